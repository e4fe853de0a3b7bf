"""The embedding encoder on the GPU, through the C ABI, against fp32 references:
each kernel against the same op in plain fp32 torch, the whole model against the CPU oracle
(oracle/bert_oracle.py).  Floating point: the tolerances are written next to each comparison; what
they cover is the fp16 storage of the matmul operands (relative 2^-11 per element), nothing else --
accumulation, softmax and LayerNorm are fp32."""
import asyncio
import math
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import bert_oracle as bo

H = 1024


@pytest.fixture(scope="module")
def sqe():
    import sqe_b200
    sqe_b200._native.load()
    rc, sms, major, _ = sqe_b200._native.device_info()
    assert rc == 1 and major == 10, sqe_b200._native.last_error()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return sqe_b200


def dev():
    return torch.device("cuda", 0)


def _gen(seed):
    return torch.Generator(device=dev()).manual_seed(seed)


def _randn(g, *shape, s=1.0, dtype=torch.float32):
    return (torch.randn(*shape, generator=g, device=dev(), dtype=torch.float32) * s).to(dtype)


@pytest.fixture(params=[1, 2, 3, 4], ids=["tiles128x64", "pairs256x256", "clusters4_multicast", "clusters8_multicast"])
def gemm_form(request, sqe):
    nat = sqe._native
    old = nat.tuning_set(nat.SQE_TUNE_ENC_GEMM_FORM, request.param)
    yield request.param
    nat.tuning_set(nat.SQE_TUNE_ENC_GEMM_FORM, old)


# ------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("m,n,k", [(128, 1024, 1024), (200, 3072, 1024), (640, 4096, 1024), (1000, 1024, 4096),
                                   (2304, 3072, 1024)])
def test_gemm_residual_fp32_epilogue(sqe, gemm_form, m, n, k):
    g = _gen(m + n + k)
    x = _randn(g, m, k, dtype=torch.float16)
    w = _randn(g, n, k, s=0.03, dtype=torch.float16)
    bias = _randn(g, n, s=0.5)
    res = _randn(g, m, n)
    out = torch.full((m, n), float("nan"), device=dev())
    sqe.encoder.gemm(x, w, bias, sqe._native.SQE_ENC_EPI_RES_F32, out, residual=res)
    want = x.float() @ w.float().T + bias + res
    # fp32 accumulation in a different order: a few ulps of the partial sums (|sum| <= ~6)
    err = float((out - want).abs().max())
    assert err < 2e-4, err


@pytest.mark.parametrize("m", [128, 333, 1536])
def test_gemm_gelu_epilogue(sqe, gemm_form, m):
    g = _gen(m)
    x = _randn(g, m, H, dtype=torch.float16)
    w = _randn(g, 4096, H, s=0.03, dtype=torch.float16)
    bias = _randn(g, 4096, s=0.5)
    out = torch.zeros((m, 4096), device=dev(), dtype=torch.float16)
    sqe.encoder.gemm(x, w, bias, sqe._native.SQE_ENC_EPI_GELU, out)
    y = x.float() @ w.float().T + bias
    want = 0.5 * y * (1.0 + torch.erf(y / math.sqrt(2.0)))
    # fp16 output: half an ulp (2^-11 relative) + the fp32 round-off of y through the gelu slope
    assert bool(((out.float() - want).abs() <= want.abs() * 6e-4 + 2e-4).all())


@pytest.mark.parametrize("m", [128, 300, 1280])
def test_gemm_qkv_split_epilogue(sqe, gemm_form, m):
    """Q (scaled by 1/8) | K row-major, V transposed -- and rows >= m are left alone."""
    g = _gen(7 * m)
    t_pad = (m + 127) // 128 * 128
    x = _randn(g, t_pad, H, dtype=torch.float16)
    w = _randn(g, 3 * H, H, s=0.03, dtype=torch.float16)
    bias = _randn(g, 3 * H, s=0.5)
    qk = torch.full((t_pad, 2 * H), 7.0, device=dev(), dtype=torch.float16)
    vt = torch.full((H, t_pad), 7.0, device=dev(), dtype=torch.float16)
    sqe.encoder.gemm(x, w, bias, sqe._native.SQE_ENC_EPI_SPLIT, qk, m=m, out1=vt, n_split=2 * H, q_cols=H,
                     q_scale=0.125)
    y = x[:m].float() @ w.float().T + bias
    y[:, :H] *= 0.125
    tol = lambda want: want.abs() * 6e-4 + 2e-4                                  # noqa: E731
    assert bool(((qk[:m].float() - y[:, :2 * H]).abs() <= tol(y[:, :2 * H])).all())
    assert bool(((vt[:, :m].float() - y[:, 2 * H:].T).abs() <= tol(y[:, 2 * H:].T)).all())
    assert bool((qk[m:] == 7.0).all()) and bool((vt[:, m:] == 7.0).all())


@pytest.mark.parametrize("m", [1, 16, 17, 100, 128])
@pytest.mark.parametrize("n,k,epi", [(3072, 1024, 0), (1024, 1024, 1), (4096, 1024, 2), (1024, 4096, 1)])
def test_gemm_small_swap_ab_split_k(sqe, m, n, k, epi):
    """The few-token form (weights as the M operand, K split over CTAs, last-CTA reduction) against
    fp32 torch, every epilogue; repeated launches are bit-identical (fixed summation order) and the
    tickets return to zero."""
    nat = sqe._native
    g = _gen(m * 7 + n + k + epi)
    x = _randn(g, 128, k, dtype=torch.float16)
    w = _randn(g, n, k, s=0.03, dtype=torch.float16)
    bias = _randn(g, n, s=0.5)
    res = _randn(g, 128, n)
    ws = torch.zeros(int(nat.load().sqe_encoder_gemm_small_workspace_bytes()), dtype=torch.uint8, device=dev())
    y = x[:m].float() @ w.float().T + bias
    outs = []
    for _ in range(2):
        if epi == 1:
            out0, out1 = torch.full((128, n), 7.0, device=dev()), None
        elif epi == 2:
            out0, out1 = torch.full((128, n), 7.0, device=dev(), dtype=torch.float16), None
        else:
            out0 = torch.full((128, 2 * H), 7.0, device=dev(), dtype=torch.float16)
            out1 = torch.full((H, 128), 7.0, device=dev(), dtype=torch.float16)
        nat.call("sqe_encoder_gemm_small", x.data_ptr(), k, w.data_ptr(), bias.data_ptr(), m, n, k, epi, out0.data_ptr(),
                 out0.stride(0), 0 if out1 is None else out1.data_ptr(), 128, 2 * H if epi == 0 else 0, H if epi == 0 else 0,
                 0.125, res.data_ptr() if epi == 1 else 0, n, 0, 0, 0, ws.data_ptr(), ws.numel(),
                 torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        outs.append((out0, out1))
    assert torch.equal(outs[0][0], outs[1][0]) and not ws[:4096].any()
    out0, out1 = outs[0]
    tol = lambda want: want.abs() * 6e-4 + 2e-4                                   # noqa: E731
    if epi == 1:
        assert float((out0[:m] - (y + res[:m])).abs().max()) < 2e-4
    elif epi == 2:
        want = 0.5 * y * (1.0 + torch.erf(y / math.sqrt(2.0)))
        assert bool(((out0[:m].float() - want).abs() <= tol(want)).all())
    else:
        y[:, :H] *= 0.125
        assert bool(((out0[:m].float() - y[:, :2 * H]).abs() <= tol(y[:, :2 * H])).all())
        assert bool(((out1[:, :m].float() - y[:, 2 * H:].T).abs() <= tol(y[:, 2 * H:].T)).all())
        assert bool((out1[:, m:] == 7.0).all())
    assert bool((out0[m:] == 7.0).all())                                          # rows >= m are left alone


def test_few_token_forward_matches_the_tile_form(sqe):
    """A query takes the swap-AB split-K products; with the knob off it takes the 128 x 64 tiles: the
    same embedding up to the summation order, and both agree with the oracle."""
    nat = sqe._native
    w, e = _pair(sqe, 25, layers=3)
    e.use_graphs = False
    g = torch.Generator().manual_seed(10)
    for lens in ([9], [3, 20], [16, 8], [32]):
        seqs = [torch.randint(0, 2000, (n,), generator=g).tolist() for n in lens]
        small = e.embed_token_ids(seqs)
        torch.cuda.synchronize()
        old = nat.tuning_set(nat.SQE_TUNE_ENC_SMALL, 1)
        try:
            tiles = e.embed_token_ids(seqs)
            torch.cuda.synchronize()
        finally:
            nat.tuning_set(nat.SQE_TUNE_ENC_SMALL, old)
        assert float((small - tiles).abs().max()) < 5e-3, lens                       # fp16 round-off of intermediates flips
        assert float((small.cpu() - bo.bert_embed(w, seqs)).abs().max()) < 3e-2, lens


def test_gemm_argument_errors(sqe):
    nat = sqe._native
    x = torch.zeros((128, H), device=dev(), dtype=torch.float16)
    w = torch.zeros((1000, H), device=dev(), dtype=torch.float16)               # n % 256 != 0
    with pytest.raises(nat.SqeError) as e:
        sqe.encoder.gemm(x, w, torch.zeros(1000, device=dev()), nat.SQE_ENC_EPI_GELU,
                         torch.zeros((128, 1000), device=dev(), dtype=torch.float16))
    assert e.value.code == -1
    with pytest.raises(RuntimeError):
        sqe.encoder.gemm(x.cpu(), w.cpu(), torch.zeros(1000), nat.SQE_ENC_EPI_GELU, torch.zeros((128, 1000)))


# ------------------------------------------------------------------------------ row kernels
def test_layernorm_matches_torch(sqe):
    g = _gen(3)
    x = _randn(g, 777, H, s=3.0) + 1.5
    x[5] = 0.25                                                                   # constant row: variance 0
    gamma, beta = 1.0 + _randn(g, H, s=0.2), _randn(g, H, s=0.2)
    o32, o16 = torch.empty_like(x), torch.empty((777, H), device=dev(), dtype=torch.float16)
    sqe.encoder.layernorm(x, gamma, beta, 1e-12, o32, o16)
    want = torch.nn.functional.layer_norm(x, (H,), gamma, beta, eps=1e-12)
    assert float((o32 - want).abs().max()) < 2e-5
    assert torch.equal(o16, o32.half())
    assert bool((o32[5] - beta).abs().max() < 1e-6)


def test_layernorm_statistics_form_is_bit_identical(sqe, gemm_form):
    """The form the forward pass uses: LayerNorm stores only its fp16 output + {mean, rstd}; the residual
    GEMM (both tile forms and the few-token form) recomputes the fp32 output from the pre-LayerNorm sum.
    Same bits as LayerNorm-then-plain-residual."""
    nat = sqe._native
    g = _gen(77 + gemm_form)
    m, n, k = 300, H, H
    pre = _randn(g, m, H, s=2.0) + 0.5                                            # the pre-LayerNorm sum
    gamma, beta = 1.0 + _randn(g, H, s=0.2), _randn(g, H, s=0.2)
    o32 = torch.empty_like(pre)
    o16 = torch.empty((m, H), device=dev(), dtype=torch.float16)
    sqe.encoder.layernorm(pre, gamma, beta, 1e-12, o32, o16)                      # plain form
    stats = torch.zeros((m, 2), device=dev())
    o16s = torch.empty_like(o16)
    sqe.encoder.layernorm(pre, gamma, beta, 1e-12, None, o16s, stats=stats)       # statistics form
    assert torch.equal(o16, o16s)
    mean = pre.mean(dim=1)
    assert float((stats[:, 0] - mean).abs().max()) < 1e-5
    x = _randn(g, m, k, dtype=torch.float16)
    w = _randn(g, n, k, s=0.03, dtype=torch.float16)
    bias = _randn(g, n, s=0.5)
    a = torch.empty((m, n), device=dev())
    b = torch.empty((m, n), device=dev())
    sqe.encoder.gemm(x, w, bias, nat.SQE_ENC_EPI_RES_F32, a, residual=o32)
    sqe.encoder.gemm(x, w, bias, nat.SQE_ENC_EPI_RES_F32, b, residual=pre, res_stats=stats, res_gamma=gamma, res_beta=beta)
    assert torch.equal(a, b)
    # few-token form (16 rows)
    ws = torch.zeros(int(nat.load().sqe_encoder_gemm_small_workspace_bytes()), dtype=torch.uint8, device=dev())
    outs = []
    for res, st in ((o32, None), (pre, stats)):
        o = torch.zeros((16, n), device=dev())
        nat.call("sqe_encoder_gemm_small", x.data_ptr(), k, w.data_ptr(), bias.data_ptr(), 16, n, k, 1, o.data_ptr(), n, 0, 0,
                 0, 0, 1.0, res.data_ptr(), H, 0 if st is None else st.data_ptr(), 0 if st is None else gamma.data_ptr(),
                 0 if st is None else beta.data_ptr(), ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
        outs.append(o)
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1])
    # pooling from the sum + statistics
    first = torch.tensor([0, 7, 299], dtype=torch.int32, device=dev())
    p0 = torch.empty((3, H), device=dev())
    p1 = torch.empty((3, H), device=dev())
    nat.call("sqe_encoder_pool", o32.data_ptr(), first.data_ptr(), 3, p0.data_ptr(), H, 0, 0, 0, torch.cuda.current_stream().cuda_stream)
    nat.call("sqe_encoder_pool", pre.data_ptr(), first.data_ptr(), 3, p1.data_ptr(), H, stats.data_ptr(), gamma.data_ptr(),
             beta.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert torch.equal(p0, p1) and torch.equal(p0, o32[first.long()])


def test_embed_ln_matches_torch(sqe):
    nat = sqe._native
    g = _gen(4)
    vocab, max_pos, rows = 999, 512, 300
    word, pos_e, type_e = _randn(g, vocab, H, s=0.5), _randn(g, max_pos, H, s=0.3), _randn(g, 2, H, s=0.1)
    gamma, beta = 1.0 + _randn(g, H, s=0.1), _randn(g, H, s=0.1)
    ids = torch.randint(0, vocab, (rows,), generator=g, device=dev(), dtype=torch.int32)
    pos = torch.randint(0, max_pos, (rows,), generator=g, device=dev(), dtype=torch.int32)
    ids[[7, 100]] = -1                                                            # padding rows
    o32 = torch.full((rows, H), 9.0, device=dev())
    o16 = torch.full((rows, H), 9.0, device=dev(), dtype=torch.float16)
    nat.call("sqe_encoder_embed_ln", ids.data_ptr(), pos.data_ptr(), word.data_ptr(), vocab, pos_e.data_ptr(),
             max_pos, type_e[0].contiguous().data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1e-12, rows,
             o32.data_ptr(), o16.data_ptr(), 0, torch.cuda.current_stream().cuda_stream)
    safe = ids.clamp(min=0).long()
    want = torch.nn.functional.layer_norm((word[safe] + type_e[0]) + pos_e[pos.long()], (H,), gamma, beta, eps=1e-12)
    want[[7, 100]] = 0.0
    assert float((o32 - want).abs().max()) < 2e-5
    assert torch.equal(o16, o32.half())
    # statistics form: out_f32 = the PRE-LayerNorm sum, stats = {mean, rstd}; the fp16 output is the same
    s32 = torch.full((rows, H), 9.0, device=dev())
    s16 = torch.full((rows, H), 9.0, device=dev(), dtype=torch.float16)
    stats = torch.full((rows, 2), 9.0, device=dev())
    nat.call("sqe_encoder_embed_ln", ids.data_ptr(), pos.data_ptr(), word.data_ptr(), vocab, pos_e.data_ptr(),
             max_pos, type_e[0].contiguous().data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1e-12, rows,
             s32.data_ptr(), s16.data_ptr(), stats.data_ptr(), torch.cuda.current_stream().cuda_stream)
    pre = (word[safe] + type_e[0]) + pos_e[pos.long()]
    pre[[7, 100]] = 0.0
    assert torch.equal(s32, pre) and torch.equal(s16, o16)
    keep = torch.ones(rows, dtype=torch.bool, device=dev())
    keep[[7, 100]] = False
    assert float((stats[keep, 0] - pre[keep].mean(dim=1)).abs().max()) < 1e-5
    recomputed = torch.addcmul(beta, (s32 - stats[:, :1]) * stats[:, 1:], gamma)
    assert float((recomputed[keep] - o32[keep]).abs().max()) < 1e-6


# ------------------------------------------------------------------------------- attention
def _attention_case(sqe, lens, seed, junk=0.0):
    g = _gen(seed)
    t_pad, pos, first, tiles = sqe.GpuEmbeddingEncoder.plan(lens)
    total = int(first[-1]) + lens[-1]
    qk = _randn(g, t_pad, 2 * H, s=1.6, dtype=torch.float16)
    qk[:, :H] *= 0.125
    v = _randn(g, t_pad, H, dtype=torch.float16)
    if junk:
        qk[total:] = junk                                                          # rows after the last sequence
        v[total:] = junk
    vt = v.T.contiguous()
    ctx = torch.full((t_pad, H), 5.0, device=dev(), dtype=torch.float16)
    tiles_d = torch.from_numpy(tiles).to(dev())
    sqe.encoder.attention(qk, vt, tiles_d, tiles.shape[0], max(lens), ctx)
    torch.cuda.synchronize()
    worst = 0.0
    for s0, n in zip(first.tolist(), lens):
        q = qk[s0:s0 + n, :H].float().view(n, 16, 64).transpose(0, 1)
        k = qk[s0:s0 + n, H:].float().view(n, 16, 64).transpose(0, 1)
        vv = v[s0:s0 + n].float().view(n, 16, 64).transpose(0, 1)
        want = (torch.softmax(q @ k.transpose(1, 2), dim=-1) @ vv).transpose(0, 1).reshape(n, H)
        worst = max(worst, float((ctx[s0:s0 + n].float() - want).abs().max()))
    assert bool((ctx[total:] == 5.0).all())                                        # nothing stored past the end
    return worst


@pytest.mark.parametrize("lens", [[1], [2], [63], [64], [65], [128], [129], [255, 257], [511], [512],
                                  [5, 130, 1, 512, 64, 300, 17], [384] * 3, [12] * 40,
                                  [512] * 28, [300] * 20, [130] * 60, [449, 77, 512, 200] * 8],
                         ids=lambda v: f"{len(v)}x{max(v)}")
def test_attention_matches_torch(sqe, lens):
    # P is rounded to fp16 (2^-11 relative per weight, weights sum to 1, |v| <~ 4) and so is the output
    assert _attention_case(sqe, lens, seed=sum(lens)) < 4e-3


def test_attention_masks_whatever_follows_a_sequence(sqe):
    assert _attention_case(sqe, [70, 3, 200], seed=1, junk=30000.0) < 4e-3


# -------------------------------------------------------------------------------- the model
def _pair(sqe, seed, layers, vocab=2000):
    w = bo.random_bert_weights(seed, layers=layers, vocab=vocab)
    gw = sqe.EncoderWeights.from_state_dict(w, device=dev())
    return w, sqe.GpuEmbeddingEncoder(gw)


def _check_vs_oracle(got, want, emulated, atol, atol_emulated):
    err = float((got - want).abs().max())
    cos = float(torch.nn.functional.cosine_similarity(got, want).min())
    err_e = float((got - emulated).abs().max())
    assert err < atol and err_e < atol_emulated and cos > 0.9999, (err, err_e, cos)


def test_two_layer_model_vs_oracle(sqe):
    w, e = _pair(sqe, 21, layers=2)
    g = torch.Generator().manual_seed(5)
    seqs = [torch.randint(0, 2000, (n,), generator=g).tolist() for n in (2, 9, 64, 65, 130, 300, 512, 1, 31)]
    got = e.embed_token_ids(seqs)
    torch.cuda.synchronize()
    # hidden states are O(1) (LayerNorm outputs); fp16 operands move them by <~ 1e-2 in two layers; against
    # the oracle WITH the same operand rounding only accumulation order and the fp16 P matrix remain
    _check_vs_oracle(got.cpu(), bo.bert_embed(w, seqs), bo.bert_embed(w, seqs, round_operands=lambda t: t.half().float()),
                     atol=2e-2, atol_emulated=6e-3)
    again = e.embed_token_ids(seqs)
    torch.cuda.synchronize()
    assert torch.equal(got, again)                                                 # deterministic
    alone = e.embed_token_ids([seqs[4]])
    torch.cuda.synchronize()
    assert float((alone[0] - got[4]).abs().max()) < 1e-3                           # batch composition: same rows


def test_full_depth_model_vs_oracle(sqe):
    """24 layers, the mxbai-embed-large geometry (random weights: the checkpoint cannot be fetched)."""
    w, e = _pair(sqe, 22, layers=24, vocab=1000)
    g = torch.Generator().manual_seed(6)
    seqs = [torch.randint(0, 1000, (n,), generator=g).tolist() for n in (7, 140, 33)]
    got = e.embed_token_ids(seqs)
    torch.cuda.synchronize()
    assert e.launches_last_forward == 2 + 24 * 7
    _check_vs_oracle(got.cpu(), bo.bert_embed(w, seqs), bo.bert_embed(w, seqs, round_operands=lambda t: t.half().float()),
                     atol=6e-2, atol_emulated=3e-2)


def test_query_shapes_replay_a_cuda_graph_with_identical_results(sqe):
    """The second batch of a shape is captured; replays serve other token ids of that shape."""
    w, e = _pair(sqe, 24, layers=2)
    plain = sqe.GpuEmbeddingEncoder(e.w, use_graphs=False)
    g = torch.Generator().manual_seed(9)
    for rep in range(5):
        seqs = [torch.randint(0, 2000, (11,), generator=g).tolist()]
        got = e.embed_token_ids(seqs)
        want = plain.embed_token_ids(seqs)
        torch.cuda.synchronize()
        assert torch.equal(got, want), rep
    assert e.graph_replays >= 3 and plain.graph_replays == 0
    pair = [torch.randint(0, 2000, (n,), generator=g).tolist() for n in (5, 40)]
    for rep in range(3):
        got, want = e.embed_token_ids(pair), plain.embed_token_ids(pair)
        torch.cuda.synchronize()
        assert torch.equal(got, want)
    assert float((got.cpu() - bo.bert_embed(w, pair)).abs().max()) < 2e-2


def test_many_tokens_split_into_batches(sqe):
    w, e = _pair(sqe, 23, layers=1)
    e.max_batch_tokens = 1024
    g = torch.Generator().manual_seed(8)
    seqs = [torch.randint(0, 2000, (int(n),), generator=g).tolist() for n in torch.randint(1, 400, (30,), generator=g)]
    got = e.embed_token_ids(seqs)
    torch.cuda.synchronize()
    want = bo.bert_embed(w, seqs)
    assert float((got.cpu() - want).abs().max()) < 1e-2


# ----------------------------------------------------------- the reference's call signatures
VOCAB = ["[PAD]", "[UNK]", "[CLS]", "[SEP]"] + [chr(c) for c in range(97, 123)] + ["##" + chr(c) for c in range(97, 123)] + \
        ["the", "cell", "##s", "protein", "bind", "##ing", "gene", "expression", "tumor", "patient", ".", ",", "?"]


def test_drop_in_coroutines_and_retrieval_end_to_end(sqe):
    """`install_encoder` patches the three embedding coroutines of a loaded reference module; their
    outputs feed `add_embeddings` / `search` (main.py:455, :499) unchanged."""
    w = bo.random_bert_weights(31, layers=2, vocab=len(VOCAB))
    vocab = {t: i for i, t in enumerate(VOCAB)}
    e = sqe.GpuEmbeddingEncoder(sqe.EncoderWeights.from_state_dict(w, device=dev()), sqe.WordPieceTokenizer(vocab))
    main = types.SimpleNamespace(embed_query=None)
    sqe.install_encoder(main, e)
    chunks = ["the cells bind the protein.", "gene expression, tumor cells", "patient tumor binding?", "  ",
              "protein " * 600]
    emb = asyncio.run(main.embed_texts_in_batches(chunks))
    assert emb.shape == (5, 1024) and emb.dtype == np.float32 and not emb[3].any()
    want = bo.bert_embed(w, [bo.encode_text(t, vocab) for t in chunks if t.strip()]).numpy()
    assert np.abs(emb[[0, 1, 2, 4]] - want).max() < 2e-2
    q = asyncio.run(main.embed_query("gene expression, tumor cells"))
    assert q.shape == (1, 1024) and np.abs(q[0] - emb[1]).max() < 5e-3          # alone: few-token products
    assert asyncio.run(main.embed_query(" ")).size == 0
    one = asyncio.run(main.ollama_embed_text("patient tumor binding?"))
    assert isinstance(one, list) and len(one) == 1024 and abs(one[0] - float(emb[2, 0])) < 5e-3
    index = sqe.GpuCorpusIndex(dtype="fp32")
    index.add_embeddings(emb, [{"doc_id": f"d{i}", "text": t} for i, t in enumerate(chunks)])
    hits = index.search(q, k=2)
    assert hits[0][0]["doc_id"] == "d1" and hits[0][1] > 0.9999
    gen = types.SimpleNamespace(bulk_index_embeddings=None)
    sqe.install_encoder(gen, e)
    assert asyncio.run(gen.embed_texts_in_batches([])).shape == (0, 1024) and not hasattr(gen, "embed_query")


def test_micro_batcher_serves_text_queries_through_the_encoder(sqe):
    """`MicroBatcher(index, encoder=...)`: a request is the query TEXT; the batch is encoded and
    searched on the device.  Same hits as embed_query + search one by one (main.py:676, :684)."""
    w = bo.random_bert_weights(33, layers=2, vocab=len(VOCAB))
    vocab = {t: i for i, t in enumerate(VOCAB)}
    e = sqe.GpuEmbeddingEncoder(sqe.EncoderWeights.from_state_dict(w, device=dev()), sqe.WordPieceTokenizer(vocab))
    rng = np.random.default_rng(4)
    words = VOCAB[-13:-3]
    chunks = [" ".join(rng.choice(words, int(rng.integers(3, 40)))) + " ." for _ in range(300)]
    chunks = list(dict.fromkeys(chunks))                                          # distinct texts
    index = sqe.GpuCorpusIndex(dtype="fp32")
    index.add_embeddings(e.embed_texts(chunks), [{"doc_id": f"d{i}", "text": t} for i, t in enumerate(chunks)])
    mb = sqe.MicroBatcher(index, max_batch=64, max_wait_s=2e-3, depth=2, encoder=e)
    try:
        picks = [int(i) for i in rng.choice(len(chunks), 90, replace=False)]
        futs = [mb.submit_text(chunks[i], 3) for i in picks]
        for i, f in zip(picks, futs):
            hits = f.result(timeout=60)
            assert hits[0][0]["doc_id"] == f"d{i}" and hits[0][1] > 0.9999, (i, hits[0])
            one = index.search(asyncio.run(e.embed_query(chunks[i])), k=3)
            assert one[0][0]["doc_id"] == f"d{i}" and abs(one[1][1] - hits[1][1]) < 2e-3     # same runner-up score
        assert mb.submit_text("   ", 3).result(timeout=10) == []
        assert mb.search_text(e.tok.encode(chunks[picks[0]]), 1)[0][0]["doc_id"] == f"d{picks[0]}"

        async def many():
            return await asyncio.gather(*[mb.asearch_text(chunks[i], 2) for i in picks[:40]])
        for i, hits in zip(picks[:40], asyncio.run(many())):
            assert hits[0][0]["doc_id"] == f"d{i}"
        assert mb.batches < mb.requests                                           # requests were coalesced
    finally:
        mb.close()


def test_a_transformers_checkpoint_loads_and_reproduces_the_library(sqe, tmp_path):
    """`EncoderWeights.load` on files written by the `transformers` library (a BertModel of the
    mxbai-embed-large geometry, 2 layers, random weights; both the torch and the safetensors format) +
    `WordPieceTokenizer.from_file` on a vocab.txt: the GPU embeddings equal what the library computes for
    the same texts with its own padded batch + attention mask (CLS token of the last hidden state)."""
    tr = pytest.importorskip("transformers")
    torch.manual_seed(123)
    cfg = tr.BertConfig(hidden_size=1024, num_hidden_layers=2, num_attention_heads=16, intermediate_size=4096,
                        vocab_size=len(VOCAB), max_position_embeddings=512)
    model = tr.BertModel(cfg).eval()                                              # with its (unused) pooler
    with torch.no_grad():
        for name, prm in model.named_parameters():
            if "query.weight" in name or "key.weight" in name:
                prm.normal_(0.0, 0.05)
            elif prm.dim() == 2 and "embeddings" not in name:
                prm.normal_(0.0, 0.03)
            elif "embeddings" in name and prm.dim() == 2:
                prm.normal_(0.0, 0.4)
    vocab_file = tmp_path / "vocab.txt"
    vocab_file.write_text("\n".join(VOCAB) + "\n", encoding="utf-8")
    tok = sqe.WordPieceTokenizer.from_file(str(vocab_file))
    texts = ["the cells bind the protein.", "gene expression, tumor cells", "patient " * 300]
    seqs = [tok.encode(t) for t in texts]
    L = max(map(len, seqs))
    ids = torch.zeros(len(seqs), L, dtype=torch.long)
    mask = torch.zeros(len(seqs), L, dtype=torch.long)
    for i, sq in enumerate(seqs):
        ids[i, : len(sq)] = torch.tensor(sq)
        mask[i, : len(sq)] = 1
    with torch.no_grad():
        want = model(input_ids=ids, attention_mask=mask).last_hidden_state[:, 0].numpy()
    pt = tmp_path / "pytorch_model.bin"
    torch.save({"bert." + k: v for k, v in model.state_dict().items()}, str(pt))          # a task-head style prefix
    paths = [str(pt)]
    try:
        from safetensors.torch import save_file
        st = tmp_path / "model.safetensors"
        save_file({k: v.contiguous() for k, v in model.state_dict().items()}, str(st))
        paths.append(str(st))
    except ImportError:
        pass
    for path in paths:
        e = sqe.GpuEmbeddingEncoder(sqe.EncoderWeights.load(path, device=dev()), tok)
        got = e.embed_texts(texts)
        assert np.abs(got - want).max() < 2e-2, path
        cos = (got * want).sum(1) / (np.linalg.norm(got, axis=1) * np.linalg.norm(want, axis=1))
        assert cos.min() > 0.9999


def test_ingest_thread_and_text_serving_share_one_encoder(sqe):
    """The reference embeds chunks from an executor thread (main.py:449-455) while requests embed
    queries (main.py:676).  Here the ingest thread runs `embed_texts` on the encoder's stream and a
    MicroBatcher runs text requests on ITS compute stream with the same encoder object (activation
    buffers are per stream): both must equal what they produce alone."""
    import threading
    w = bo.random_bert_weights(35, layers=2, vocab=len(VOCAB))
    vocab = {t: i for i, t in enumerate(VOCAB)}
    e = sqe.GpuEmbeddingEncoder(sqe.EncoderWeights.from_state_dict(w, device=dev()), sqe.WordPieceTokenizer(vocab))
    rng = np.random.default_rng(6)
    words = VOCAB[-13:-3]
    chunks = list(dict.fromkeys(" ".join(rng.choice(words, int(rng.integers(3, 30)))) + " ." for _ in range(120)))
    alone = e.embed_texts(chunks)
    index = sqe.GpuCorpusIndex(dtype="fp32")
    index.add_embeddings(alone, [{"doc_id": f"d{i}", "text": t} for i, t in enumerate(chunks)])
    mb = sqe.MicroBatcher(index, max_batch=32, max_wait_s=1e-3, depth=2, encoder=e)
    errors, ingested = [], []

    def ingest():
        try:
            for _ in range(6):
                ingested.append(e.embed_texts(chunks))
        except Exception as ex:                                                   # noqa: BLE001
            errors.append(ex)
    t = threading.Thread(target=ingest)
    t.start()
    try:
        for rep in range(6):
            futs = [mb.submit_text(chunks[i], 1) for i in range(0, len(chunks), 3)]
            for i, f in zip(range(0, len(chunks), 3), futs):
                hit = f.result(timeout=60)[0]
                assert hit[0]["doc_id"] == f"d{i}" and hit[1] > 0.9999, (rep, i, hit)
    finally:
        t.join()
        mb.close()
    assert not errors, errors
    for got in ingested:
        assert np.array_equal(got, alone)                                         # same batches, same stream: bit-identical


def test_random_length_mixes_against_the_oracle(sqe):
    """Fuzz of the packing / attention work list: 24 random batches (1..40 sequences of 1..512 tokens, some
    all-short, some all-long, token budgets that split a batch) through a one-layer model vs the oracle."""
    w, e = _pair(sqe, 41, layers=1)
    rng = np.random.default_rng(12)
    worst = 0.0
    for trial in range(24):
        n = int(rng.integers(1, 41))
        hi = int(rng.choice([8, 64, 130, 512]))
        lens = [int(x) for x in rng.integers(1, hi + 1, size=n)]
        if trial % 5 == 0:
            lens[int(rng.integers(0, n))] = 512
        e.max_batch_tokens = int(rng.choice([512, 2048, 32768]))
        seqs = [rng.integers(0, 2000, size=m).tolist() for m in lens]
        got = e.embed_token_ids(seqs)
        torch.cuda.synchronize()
        err = float((got.cpu() - bo.bert_embed(w, seqs)).abs().max())
        worst = max(worst, err)
        assert err < 1e-2, (trial, lens, err)
    assert worst > 0.0


def test_the_ollama_model_file_loads_and_runs(sqe, tmp_path):
    """The deployment's own checkpoint: a BERT GGUF file (written here by the `gguf` library the way
    llama.cpp's converter writes mxbai-embed-large: F16 matrices, F32 vectors / position table, rewritten
    vocabulary) inside an Ollama model store.  `GpuEmbeddingEncoder.from_ollama` finds the blob through
    the manifest, reads weights AND vocabulary from it; the embeddings are bit-identical to an encoder
    built from the tensors the file holds, and agree with the oracle on the ORIGINAL fp32 weights within
    the fp16 storage error.  A Q8_0 file of the same model loads too."""
    pytest.importorskip("gguf")
    import json
    from gguf_fixture import write_bert_gguf
    w = bo.random_bert_weights(77, layers=2, vocab=len(VOCAB))
    root = tmp_path / "models"
    (root / "blobs").mkdir(parents=True)
    digest = "sha256:" + "5e" * 32
    blob = root / "blobs" / digest.replace(":", "-")
    held = write_bert_gguf(str(blob), w, VOCAB, ftype="f16")
    man = root / "manifests" / "registry.ollama.ai" / "library" / "mxbai-embed-large"
    man.mkdir(parents=True)
    (man / "latest").write_text(json.dumps({"schemaVersion": 2, "layers": [
        {"mediaType": "application/vnd.ollama.image.model", "digest": digest}]}))
    e = sqe.GpuEmbeddingEncoder.from_ollama("mxbai-embed-large", models_dir=str(root), device=dev())
    assert len(e.w.layers) == 2 and e.w.vocab_size == len(VOCAB) and e.tok.cls_id == VOCAB.index("[CLS]")
    texts = ["the cells bind the protein.", "gene expression, tumor cells?", "patient " * 200, "x"]
    got = e.embed_texts(texts)
    vocab = {t: i for i, t in enumerate(VOCAB)}
    same = sqe.GpuEmbeddingEncoder(
        sqe.EncoderWeights.from_state_dict({k: torch.from_numpy(v) for k, v in held.items()}, device=dev()),
        sqe.WordPieceTokenizer(vocab))
    assert np.array_equal(got, same.embed_texts(texts))
    seqs = [bo.encode_text(t, vocab) for t in texts]
    want = bo.bert_embed(w, seqs).numpy()
    assert np.abs(got - want).max() < 2e-2
    cos = (got * want).sum(1) / (np.linalg.norm(got, axis=1) * np.linalg.norm(want, axis=1))
    assert cos.min() > 0.9999
    q8 = str(tmp_path / "q8.gguf")
    write_bert_gguf(q8, w, VOCAB, ftype="q8_0")
    got8 = sqe.GpuEmbeddingEncoder.from_gguf(q8, device=dev()).embed_texts(texts)
    cos8 = (got8 * want).sum(1) / (np.linalg.norm(got8, axis=1) * np.linalg.norm(want, axis=1))
    assert cos8.min() > 0.999

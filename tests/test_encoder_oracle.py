"""CPU tests of the embedding-encoder row: the oracle (oracle/bert_oracle.py)
against the published implementations it restates, and the product's host logic (tokeniser,
packing plan, the reference's call conventions) -- no GPU, no compute calls into the library."""
import asyncio
import json
import os

import numpy as np
import pytest
import torch

import sqe_b200
from sqe_b200 import encoder as enc
from oracle import bert_oracle as bo

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _tiny():
    z = np.load(os.path.join(GOLDEN, "bert_tiny.npz"))
    w = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w/")}
    n = sum(1 for k in z.files if k.startswith("ids/"))
    return w, [z[f"ids/{i}"].tolist() for i in range(n)], [z[f"hidden/{i}"] for i in range(n)]


def test_oracle_matches_transformers_golden():
    """fixtures = transformers.BertModel on a PADDED batch with an attention mask; the oracle runs
    every sequence alone without a mask: the same numbers (fp32 round-off)."""
    w, seqs, hidden = _tiny()
    for ids, want in zip(seqs, hidden):
        got = bo.bert_hidden_states(w, ids, heads=4).numpy()
        assert got.shape == want.shape
        np.testing.assert_allclose(got, want, rtol=0, atol=2e-5)
    cls = bo.bert_embed(w, seqs, heads=4).numpy()
    np.testing.assert_allclose(cls, np.stack([h[0] for h in hidden]), rtol=0, atol=2e-5)


def test_oracle_matches_transformers_live_at_hidden_1024():
    """One full-width layer (hidden 1024, 16 heads, FFN 4096) with the oracle's own seeded weights,
    loaded into the library's BertModel."""
    tr = pytest.importorskip("transformers")
    w = bo.random_bert_weights(3, layers=1, vocab=500)
    cfg = tr.BertConfig(hidden_size=1024, num_hidden_layers=1, num_attention_heads=16, intermediate_size=4096,
                        vocab_size=500, max_position_embeddings=512)
    m = tr.BertModel(cfg, add_pooling_layer=False).eval()
    missing, unexpected = m.load_state_dict(w, strict=False)
    assert not unexpected and all("position_ids" in k for k in missing)
    g = torch.Generator().manual_seed(1)
    seqs = [torch.randint(0, 500, (n,), generator=g).tolist() for n in (3, 70, 130)]
    L = max(map(len, seqs))
    ids = torch.zeros(len(seqs), L, dtype=torch.long)
    mask = torch.zeros(len(seqs), L, dtype=torch.long)
    for i, s in enumerate(seqs):
        ids[i, : len(s)] = torch.tensor(s)
        mask[i, : len(s)] = 1
    with torch.no_grad():
        want = m(input_ids=ids, attention_mask=mask).last_hidden_state
    for i, s in enumerate(seqs):
        got = bo.bert_hidden_states(w, s)
        assert float((got - want[i, : len(s)]).abs().max()) < 5e-5


def test_oracle_matches_transformers_live_at_full_depth():
    """All 24 layers of the mxbai-embed-large geometry (seeded weights of the oracle loaded into the
    library's BertModel), CLS embeddings of a padded, masked batch: the quantity the GPU tests compare."""
    tr = pytest.importorskip("transformers")
    w = bo.random_bert_weights(9, layers=24, vocab=300)
    cfg = tr.BertConfig(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096,
                        vocab_size=300, max_position_embeddings=512)
    m = tr.BertModel(cfg, add_pooling_layer=False).eval()
    missing, unexpected = m.load_state_dict(w, strict=False)
    assert not unexpected and all("position_ids" in k for k in missing)
    g = torch.Generator().manual_seed(4)
    seqs = [torch.randint(0, 300, (n,), generator=g).tolist() for n in (5, 40, 17)]
    L = max(map(len, seqs))
    ids = torch.zeros(len(seqs), L, dtype=torch.long)
    mask = torch.zeros(len(seqs), L, dtype=torch.long)
    for i, s in enumerate(seqs):
        ids[i, : len(s)] = torch.tensor(s)
        mask[i, : len(s)] = 1
    with torch.no_grad():
        want = m(input_ids=ids, attention_mask=mask).last_hidden_state[:, 0]
    got = bo.bert_embed(w, seqs)
    assert float((got - want).abs().max()) < 2e-4                     # fp32 round-off through 24 blocks


def test_tokenizers_agree_with_the_library_golden():
    d = json.load(open(os.path.join(GOLDEN, "bert_wordpiece.json"), encoding="utf-8"))
    vocab = {t: i for i, t in enumerate(d["vocab"])}
    tk = enc.WordPieceTokenizer(vocab)
    for c in d["cases"]:
        assert bo.wordpiece_tokenize(c["text"], vocab) == c["ids"], c["text"]
        assert tk.tokenize(c["text"]) == c["ids"], c["text"]
        assert tk.tokenize(c["text"]) == c["ids"]                       # memoised second pass
        assert tk.encode(c["text"]) == [vocab["[CLS]"]] + c["ids"] + [vocab["[SEP]"]]


def test_tokenizer_against_the_library_on_random_text():
    tok = pytest.importorskip("tokenizers")
    from tokenizers.models import WordPiece
    from tokenizers.normalizers import BertNormalizer
    from tokenizers.pre_tokenizers import BertPreTokenizer
    rng = np.random.default_rng(5)
    alphabet = list("abcdefghijklmnopqrstuvwxyz") + list("ABCDEÉéüñçß") + list(" \t\n.,;!?-()'\"$%") + ["中", "文", "　", " ", "​", "́"]
    pieces = ["[PAD]", "[UNK]", "[CLS]", "[SEP]"] + list("abcdefghijklmnopqrstuvwxyz") + ["##" + c for c in "abcdefghijklmnopqrstuvwxyz"]
    for _ in range(300):                                              # multi-letter pieces
        n = int(rng.integers(2, 5))
        s = "".join(rng.choice(list("abcdefghijklmnopqrstuvwxyz"), n))
        pieces += [s, "##" + s]
    pieces += list(".,;!?-()'\"$%") + ["中", "ss"]
    vocab = {}
    for p in pieces:
        vocab.setdefault(p, len(vocab))
    lib = tok.Tokenizer(WordPiece(vocab, unk_token="[UNK]", max_input_chars_per_word=100))
    lib.normalizer = BertNormalizer(clean_text=True, handle_chinese_chars=True, strip_accents=None, lowercase=True)
    lib.pre_tokenizer = BertPreTokenizer()
    mine = enc.WordPieceTokenizer(vocab)
    for _ in range(400):
        text = "".join(rng.choice(alphabet, int(rng.integers(0, 80))))
        want = lib.encode(text).ids
        assert mine.tokenize(text) == want, ascii(text)
        assert bo.wordpiece_tokenize(text, vocab) == want, ascii(text)


def test_encode_truncates_to_the_model_positions():
    vocab = {t: i for i, t in enumerate(["[PAD]", "[UNK]", "[CLS]", "[SEP]", "a"])}
    tk = enc.WordPieceTokenizer(vocab)
    ids = tk.encode("a " * 1000)
    assert len(ids) == 512 and ids[0] == 2 and ids[-1] == 3 and set(ids[1:-1]) == {4}
    assert bo.encode_text("a " * 1000, vocab) == ids
    assert tk.encode("") == [2, 3]


def test_packing_plan():
    """sequences start at multiples of 8 tokens (TMA boxes of V^T start on 16-byte boundaries)"""
    t_pad, pos, first, tiles = enc.GpuEmbeddingEncoder.plan([5, 130, 1, 512])
    assert t_pad == 768 and first.tolist() == [0, 8, 144, 152]
    assert pos[:8].tolist() == [0, 1, 2, 3, 4, -1, -1, -1] and pos[8] == 0 and pos[137] == 129
    assert pos[138:144].tolist() == [-1] * 6 and pos[144:152].tolist() == [0] + [-1] * 7
    assert pos[152:664].tolist() == list(range(512)) and (pos[664:] == -1).all()
    assert tiles.tolist() == [[0, 5, 0, 5], [8, 130, 0, 128], [8, 130, 128, 2], [144, 1, 0, 1],
                              [152, 512, 0, 128], [152, 512, 128, 128], [152, 512, 256, 128], [152, 512, 384, 128]]
    # an ingest batch: one entry per sequence (its K / V are fetched once), all four query tiles
    _, _, _, big = enc.GpuEmbeddingEncoder.plan([512] * 64)
    assert big.shape == (64, 4) and big[3].tolist() == [3 * 512, 512, 0, 512]
    _, _, _, mid = enc.GpuEmbeddingEncoder.plan([300] * 20)
    assert mid.shape == (40, 4) and mid[1].tolist() == [0, 300, 256, 44]
    with pytest.raises(ValueError):
        enc.GpuEmbeddingEncoder.plan([0])
    with pytest.raises(ValueError):
        enc.GpuEmbeddingEncoder.plan([513])


def test_rounded_operand_emulation_is_close_to_the_oracle():
    """The GPU path stores matmul operands in fp16; the oracle's emulation of that rounding bounds
    what the GPU tests may tolerate (and shows the tolerance is about rounding, not about a bug)."""
    w = bo.random_bert_weights(11, layers=2, vocab=300)
    g = torch.Generator().manual_seed(2)
    seqs = [torch.randint(0, 300, (n,), generator=g).tolist() for n in (4, 40)]
    exact = bo.bert_embed(w, seqs)
    half = bo.bert_embed(w, seqs, round_operands=lambda t: t.half().float())
    assert float((exact - half).abs().max()) < 2e-2
    cos = torch.nn.functional.cosine_similarity(exact, half)
    assert float(cos.min()) > 0.9999


def test_encoder_refuses_to_run_without_a_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        enc.EncoderWeights.from_state_dict(bo.random_bert_weights(0, layers=1, vocab=50), device="cpu")


def test_reference_call_conventions_for_blank_input():
    """embed_query('') -> np.array([]) (main.py:176-177); no texts -> np.array([]) in main,
    zeros((0, 1024)) in embedding_gen (embedding_gen.py:173-174).  No device work is involved."""
    e = enc.GpuEmbeddingEncoder.__new__(enc.GpuEmbeddingEncoder)
    e.blank_policy = "main"
    assert asyncio.run(e.embed_query("   ")).size == 0
    assert asyncio.run(e.embed_texts_in_batches([])).shape == (0,)
    e.blank_policy = "embedding_gen"
    assert asyncio.run(e.embed_texts_in_batches([])).shape == (0, 1024)
    assert sqe_b200.GpuEmbeddingEncoder is enc.GpuEmbeddingEncoder

"""The oracle against the reference's own outputs (tests/golden, made by
oracle/make_golden.py) and, where /root/reference is mounted, against the
reference live.  CPU only."""
import json
import os

import numpy as np
import pytest

import oracle
from oracle import numpy_oracle as no
from oracle.ref_loader import reference_available, load_reference_main


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_cosine_matches_reference_vectors(golden_dir):
    g = _load(golden_dir, "cosine.npz")
    got = np.array([oracle.cosine_similarity(a, b) for a, b in zip(g["a"], g["b"])])
    # same numpy expressions -> same bits on the same BLAS; 1e-6 leaves room for
    # another OpenBLAS kernel on another host
    np.testing.assert_allclose(got, g["out"], rtol=0, atol=1e-6)
    assert got[4] == 0.0 and got[5] == 0.0 and got[6] == 0.0       # zero-norm guard, main.py:62-63
    assert abs(got[1] - 1.0) < 1e-6 and abs(got[2] + 1.0) < 1e-6


def test_normalize_is_bit_identical_to_reference(golden_dir):
    g = _load(golden_dir, "index_search.npz")
    got = oracle.normalize_rows(g["emb"])
    assert got.dtype == np.float32
    # the reference serialises via .tolist() (exact for fp32) -> compare bits
    np.testing.assert_array_equal(got.view(np.uint32), g["stored"].view(np.uint32))
    assert not got[5].any()                                         # zero row stays zero
    qn = oracle.normalize_rows(g["q"])
    np.testing.assert_array_equal(qn.view(np.uint32), g["q_norm"].view(np.uint32))


def test_pairwise_order_is_numpys(golden_dir):
    g = _load(golden_dir, "index_search.npz")
    emb = g["emb"]
    ssq = no.pairwise_sumsq_f32(emb)
    ref = np.add.reduce(emb * emb, axis=1)
    np.testing.assert_array_equal(ssq.view(np.uint32), ref.view(np.uint32))
    np.testing.assert_array_equal(np.sqrt(ssq).view(np.uint32),
                                  np.linalg.norm(emb, axis=1).view(np.uint32))
    rng = np.random.default_rng(3)
    x = (rng.standard_normal((257, 1024)) * 10 ** rng.uniform(-6, 6, (257, 1))).astype(np.float32)
    np.testing.assert_array_equal(no.pairwise_sumsq_f32(x).view(np.uint32),
                                  np.add.reduce(x * x, axis=1).view(np.uint32))


def test_search_matches_reference_vectors(golden_dir):
    g = _load(golden_dir, "index_search.npz")
    stored = oracle.normalize_rows(g["emb"])
    qn = oracle.normalize_rows(g["q"])
    for ki, k in enumerate(g["ks"]):
        s, i = oracle.topk_cosine(stored, qn, int(k))
        np.testing.assert_array_equal(i, g["res_idx"][:, ki, :k])
        np.testing.assert_allclose(s, g["res_score"][:, ki, :k], atol=1e-6)
    # planted tie: rows 6 and 7 are identical -> 6 before 7 (main.py:84 generalised)
    _, i = oracle.topk_cosine(stored, qn[1:2], 3)
    assert list(i[0][:2]) == [6, 7] or list(i[0][:3]) == [6, 7, 40] or 40 in i[0]
    pos = {int(v): p for p, v in enumerate(i[0])}
    assert pos[6] < pos[7]
    # zero query -> every score 0.0 -> indices 0..k-1
    _, i = oracle.topk_cosine(stored, qn[2:3], 5)
    assert list(i[0]) == [0, 1, 2, 3, 4]


def test_topk_partition_equals_stable_argsort():
    rng = np.random.default_rng(11)
    s = rng.standard_normal((6, 5000)).astype(np.float32)
    s[:, 100:140] = s[:, 99:100]                 # a run of exact ties
    s[2, :] = 0.25                               # all equal
    for k in (1, 7, 64):
        gs, gi = oracle.topk_from_scores(s, k)
        for r in range(s.shape[0]):
            ref = np.argsort(-s[r], kind="stable")[:k]
            np.testing.assert_array_equal(gi[r], ref)
            np.testing.assert_array_equal(gs[r], s[r][ref])
    gs, gi = oracle.topk_from_scores(s[:, :3], 5)    # k > n
    assert (gi[:, 3:] == -1).all() and np.isneginf(gs[:, 3:]).all()


def test_topk_cosine_chunked_equals_unchunked():
    rng = np.random.default_rng(5)
    d = oracle.normalize_rows(rng.standard_normal((3000, 1024)).astype(np.float32))
    d[2500] = d[17]
    q = oracle.normalize_rows(rng.standard_normal((4, 1024)).astype(np.float32))
    q[0] = d[17]
    s0, i0 = oracle.topk_cosine(d, q, 10)
    s1, i1 = oracle.topk_cosine(d, q, 10, chunk=700)
    np.testing.assert_array_equal(i0, i1)
    np.testing.assert_array_equal(s0, s1)
    assert list(i0[0][:2]) == [17, 2500]


def test_storage_rounding():
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(2)
    x = np.concatenate([rng.standard_normal(4096).astype(np.float32) * 0.05,
                        np.array([0.0, -0.0, 1.0, 1.00390625, 1.01171875, 3.0e38, 1e-40, np.inf],
                                 dtype=np.float32)])
    bits = oracle.to_storage(x, "bf16")
    ref = torch.from_numpy(x).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    np.testing.assert_array_equal(bits, ref)
    back = oracle.from_storage(bits, "bf16")
    np.testing.assert_array_equal(back, torch.from_numpy(x).to(torch.bfloat16).float().numpy())
    h = oracle.to_storage(x, "fp16")
    np.testing.assert_array_equal(h.view(np.uint16),
                                  torch.from_numpy(x).to(torch.float16).view(torch.int16).numpy().view(np.uint16))
    np.testing.assert_array_equal(oracle.to_storage(x, "fp32"), x)


@pytest.mark.parametrize("name", ["default", "evict8", "thr095"])
def test_cache_model_replays_reference_log(golden_dir, name):
    with open(os.path.join(golden_dir, f"cache_{name}.json")) as f:
        log = json.load(f)
    vecs = _load(golden_dir, f"cache_{name}.npz")["vecs"]
    model = oracle.LfuCacheModel(max_items=log["max_items"], threshold=log["threshold"])
    for op in log["ops"]:
        v = vecs[op["vec"]][None, :]
        if op["op"] == "get":
            assert model.get(v) == op["result"], op
        else:
            model.put(v, op["response"])
    assert model.responses() == log["final_responses"]
    assert model.freqs() == log["final_freqs"]


def test_cache_lookup_edge_cases():
    rng = np.random.default_rng(9)
    q = rng.standard_normal(1024).astype(np.float32)
    assert oracle.cache_lookup(q, [], 0.96) == (-1, -1.0, False)          # empty, main.py:70-71
    idx, sim, hit = oracle.cache_lookup(q, [-q, -2 * q], 0.96)            # all sims == -1.0
    assert idx == -1 or sim <= -1.0 + 1e-6
    assert hit is False
    idx, sim, hit = oracle.cache_lookup(q, [q.copy(), q, q.copy()], 0.96)  # exact ties -> first
    assert idx == 0 and hit
    idx, sim, hit = oracle.cache_lookup(np.zeros(1024, np.float32), [q], 0.96)
    assert sim == 0.0 and not hit and idx == 0                            # 0.0 > -1.0


def test_cache_lookup_batched_equals_rowwise():
    rng = np.random.default_rng(4)
    c = oracle.normalize_rows(rng.standard_normal((300, 1024)).astype(np.float32))
    q = oracle.normalize_rows(rng.standard_normal((9, 1024)).astype(np.float32))
    q[0] = c[42]
    q[1] = oracle.normalize_rows((0.97 * c[7] + 0.1 * q[1])[None])[0]
    idx, score, hit = no.cache_lookup_batched(q, c, 0.96)
    for r in range(len(q)):
        i, s, h = oracle.cache_lookup(q[r], list(c), 0.96)
        assert idx[r] == i and bool(hit[r]) == h and abs(score[r] - s) < 1e-6
    assert hit[0] and idx[0] == 42 and hit[1] and idx[1] == 7


def test_merge_topk():
    s = np.array([[[0.9, 0.5, 0.1]], [[0.9, 0.8, -np.inf]]], dtype=np.float32)   # [lists=2,B=1,k=3]
    i = np.array([[[10, 11, 12]], [[3, 4, -1]]], dtype=np.int64)
    ms, mi = oracle.merge_topk(s, i, 4)
    assert list(mi[0]) == [3, 10, 4, 11]
    np.testing.assert_allclose(ms[0], [0.9, 0.9, 0.8, 0.5])


@pytest.mark.skipif(not reference_available(), reason="/root/reference not mounted")
def test_oracle_against_live_reference():
    m = load_reference_main()
    rng = np.random.default_rng(77)
    a = rng.standard_normal((50, 1024)).astype(np.float32)
    b = rng.standard_normal((50, 1024)).astype(np.float32)
    for x, y in zip(a, b):
        assert oracle.cosine_similarity(x, y) == m.cosine_similarity(x, y)
    assert oracle.CACHE_SIM_THRESHOLD == m.CACHE_SIM_THRESHOLD
    assert oracle.REDIS_MAX_ITEMS == m.REDIS_MAX_ITEMS
    assert oracle.EMBED_DIM == m.EMBED_DIM


@pytest.mark.skipif(not reference_available(), reason="/root/reference not mounted")
def test_reference_bulk_actions_collide_on_id_when_a_document_is_indexed_again():
    """Pins the `_id` claim behind GpuCorpusIndex's overwrite-in-place: the reference's OWN
    add_embeddings numbers `_id = f"{doc_id}_{i}"` per call (main.py:318-325) with
    `_op_type: "index"`, so indexing the same document again sends the same _ids -- which an
    OpenSearch index replaces (the stand-in bulk does what the bulk API documents)."""
    from oracle.ref_loader import FakeOpenSearch
    m = load_reference_main()
    client = FakeOpenSearch()
    idx = m.OpenSearchIndexer(client, "medical-search-index")
    rng = np.random.default_rng(9)
    e = rng.standard_normal((3, 1024)).astype(np.float32)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        idx.add_embeddings(e, [{"doc_id": "PMC1.txt", "text": f"c{i}"} for i in range(3)])
        idx.add_embeddings(e[::-1].copy(), [{"doc_id": "PMC1.txt", "text": f"d{i}"} for i in range(3)])
        idx.add_embeddings(e[:1], [{"doc_id": "PMC2.txt", "text": "x"}])
    assert [d[0] for d in client.docs] == ["PMC1.txt_0", "PMC1.txt_1", "PMC1.txt_2", "PMC2.txt_0"]
    assert [d[1]["text"] for d in client.docs] == ["d0", "d1", "d2", "x"]
    with contextlib.redirect_stdout(io.StringIO()):
        hits = idx.search(e[:1], k=1)
    assert hits[0][0]["text"] == "d2"                    # what the GPU index returns in the same scenario


# ---------------------------------------------------------------- the plain-C restatement
def _c():
    from oracle import c_oracle
    if not c_oracle.available():
        pytest.skip("no gcc and no prebuilt oracle/_build/libsqe_oracle.so")
    return c_oracle


def test_c_oracle_normalise_is_bit_identical_to_the_reference_vectors(golden_dir):
    """oracle/c_oracle.c spells out numpy's pairwise fp32 reduction in plain C; its output must be
    the reference's own stored rows bit for bit (main.py:315-316) -- an independent statement of
    the order the CUDA ingest kernel reproduces."""
    co = _c()
    g = _load(golden_dir, "index_search.npz")
    np.testing.assert_array_equal(co.normalize_rows(g["emb"]).view(np.uint32), g["stored"].view(np.uint32))
    np.testing.assert_array_equal(co.normalize_rows(g["q"]).view(np.uint32), g["q_norm"].view(np.uint32))
    rng = np.random.default_rng(8)
    x = (rng.standard_normal((300, 1024)) * 10.0 ** rng.uniform(-12, 12, (300, 1))).astype(np.float32)
    x[0] = 0.0
    x[1] = 1e-30
    np.testing.assert_array_equal(co.row_sumsq(x).view(np.uint32), np.add.reduce(x * x, axis=1).view(np.uint32))
    np.testing.assert_array_equal(co.normalize_rows(x).view(np.uint32), oracle.normalize_rows(x).view(np.uint32))


def test_c_oracle_cosine_search_and_cache_match_the_reference_vectors(golden_dir):
    co = _c()
    g = _load(golden_dir, "cosine.npz")
    got = np.array([co.cosine_similarity(a, b) for a, b in zip(g["a"], g["b"])])
    np.testing.assert_allclose(got, g["out"], rtol=0, atol=1e-6)
    assert got[4] == 0.0 and got[5] == 0.0 and got[6] == 0.0       # zero-norm guard, main.py:62-63
    g = _load(golden_dir, "index_search.npz")
    for ki, k in enumerate(g["ks"]):
        s, i = co.topk_cosine(g["stored"], g["q_norm"], int(k))
        np.testing.assert_array_equal(i, g["res_idx"][:, ki, :k])   # the reference run's own hit order
        np.testing.assert_allclose(s, g["res_score"][:, ki, :k], atol=1e-6)
    # and against the numpy restatement on fresh data: ties (7 / 50 / last) by lower row, NaN last
    rng = np.random.default_rng(12)
    d = oracle.normalize_rows(rng.standard_normal((2000, 1024)).astype(np.float32))
    d[50] = d[7]
    d[1999] = d[7]
    q = oracle.normalize_rows(rng.standard_normal((5, 1024)).astype(np.float32))
    q[0] = d[7]
    s1, i1 = co.topk_cosine(d, q, 12)
    s2, i2 = oracle.topk_cosine(d, q, 12)
    np.testing.assert_array_equal(i1, i2)
    np.testing.assert_allclose(s1, s2, atol=1e-6)
    assert list(i1[0][:3]) == [7, 50, 1999]
    sc = np.array([[0.5, np.nan, 0.5, -0.0, 0.0, 0.7]], dtype=np.float32)
    _, i = co.topk_from_scores(sc, 6)
    assert list(i[0]) == [5, 0, 2, 3, 4, 1]
    _, i = co.topk_from_scores(sc[:, :2], 4)
    assert list(i[0]) == [0, 1, -1, -1]
    # lfu_cache_get's scan (main.py:73-90): first maximum, strict '>', double threshold compare
    idx, sim, hit = co.cache_lookup(d, q[0], 0.96)
    assert (idx, hit) == (7, True) and abs(sim - 1.0) < 1e-6
    idx, sim, hit = co.cache_lookup(d, q[1], 0.96)
    wi, ws, wh = no.cache_lookup_batched(q[1:2], d, 0.96)
    assert idx == wi[0] and hit == bool(wh[0]) and abs(sim - ws[0]) < 1e-6 and hit is False
    idx, sim, hit = co.cache_lookup(-q[2:3], q[2], 0.96)                  # similarity -1 (to rounding)
    assert hit is False and (idx == -1 or sim <= -1.0 + 1e-6)             # only a sim > -1.0 replaces the start value
    assert co.cache_lookup(np.zeros((0, 1024), np.float32), q[2], 0.96) == (-1, -1.0, False)


def test_split_bf16_storage_is_exact_to_16_bits():
    """bf16x2 (new storage class): hi + lo reconstructs x to 2^-17 relative, the sum is exact in
    fp32, and zeros / signs / tiny values survive."""
    import oracle
    rng = np.random.default_rng(5)
    x = (rng.standard_normal((64, 1024)) * 10.0 ** rng.uniform(-8, 0, size=(64, 1))).astype(np.float32)
    x[0] = 0.0
    st = oracle.to_storage(x, "bf16x2")
    assert st.dtype == np.uint16 and st.shape == (64, 2048)
    back = oracle.from_storage(st, "bf16x2")
    assert back.dtype == np.float32 and back.shape == x.shape
    assert np.all(np.abs(back - x) <= np.abs(x) * 2.0 ** -16 + 1e-38)
    assert np.all(back[0] == 0.0)
    hi = oracle.from_storage(st[:, :1024], "bf16").astype(np.float64)
    lo = oracle.from_storage(st[:, 1024:], "bf16").astype(np.float64)
    assert np.array_equal((hi + lo).astype(np.float32).astype(np.float64), hi + lo)    # exact in fp32


def test_ws_session_replay_with_the_oracle_and_the_product_context_builder(golden_dir):
    """The reference's own websocket handler was recorded on a request sequence (hits, near
    misses, repeats, a blank query; tests/golden/ws_session.*).  Replaying it with the oracle for
    the retrieval path and the PRODUCT's host-side grouping (`serving.build_context_text`, the
    step right after the path, main.py:685-698) must give the same client messages and the same
    prompts, character for character."""
    import sqe_b200
    from ws_replay import load_session, replay
    meta, emb, qvec = load_session(golden_dir)
    stored = oracle.normalize_rows(emb)
    cache = oracle.LfuCacheModel()

    def os_search(query_emb, k):
        qn = oracle.normalize_rows(query_emb)
        s, i = oracle.topk_cosine(stored, qn, k)
        return [(dict(meta["docs"][r]), float(v)) for v, r in zip(s[0], i[0]) if r >= 0]
    got = replay(meta, qvec, cache.get, cache.put, os_search, sqe_b200.build_context_text)
    assert got == meta["results"]
    assert cache.responses() == meta["final_cache_responses"] and cache.freqs() == meta["final_cache_freqs"]
    # the recording exercises what it should: multi-chunk documents, cache hits, a blank query
    assert sum(r["prompt"] is None for r in got) == 3
    assert "passage 13 of PMC2002\npassage 14 of PMC2002" in got[0]["prompt"]


def _upload_fixture(golden_dir):
    with open(os.path.join(golden_dir, "upload_service.json")) as f:
        meta = json.load(f)
    arr = np.load(os.path.join(golden_dir, "upload_service.npz"))
    return meta, arr


def test_upload_service_golden_is_what_the_oracle_predicts(golden_dir):
    """tests/golden/upload_service.* = the reference's own `bulk_index_embeddings`
    (embedding_gen.py:196-257) on a call sequence.  The oracle's normalise gives the stored
    vectors bit for bit (:215-216 is the expression of main.py:315-316), `_id = f"{doc_id}_{i}"`
    with i the chunk index of THAT call (:221), an index action on a known _id replaces the
    document in place, and each user has an index of their own (:211)."""
    meta, arr = _upload_fixture(golden_dir)
    model = {}                                               # index name -> list of [_id, doc_id, text, vec]
    for ci, c in enumerate(meta["calls"]):
        emb = arr[f"emb_{ci}"]
        if emb.size == 0:                                    # :207-209
            continue
        docs = model.setdefault(f"{meta['base_index_name']}-{c['user']}", [])
        vecs = oracle.normalize_rows(emb)
        for i, (chunk, v) in enumerate(zip(c["chunks"], vecs)):
            entry = [f"{c['doc_id']}_{i}", c["doc_id"], chunk, v]
            for j, known in enumerate(docs):
                if known[0] == entry[0]:
                    docs[j] = entry
                    break
            else:
                docs.append(entry)
    assert set(model) == set(meta["indices"])
    for name, docs in model.items():
        assert [{"_id": d[0], "doc_id": d[1], "text": d[2]} for d in docs] == meta["indices"][name]
        np.testing.assert_array_equal(np.stack([d[3] for d in docs]).view(np.uint32),
                                      arr[f"stored_{name}"].view(np.uint32))
    alice = meta["indices"]["docs-alice"]
    assert len(alice) == 75 and alice[-1]["text"] == "paper v1 chunk 4" and alice[-3]["text"] == "paper v2 chunk 2"

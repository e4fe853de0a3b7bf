"""SURVEY.md 8f(3), literally: the REFERENCE'S OWN handlers on the GPU path.

When a copy of the reference's `app/` is reachable -- `SQE_REFERENCE_ROOT`, `/root/reference`, or a
copy staged under the git-ignored `oracle/_ref/reference/` by `scripts/stage_reference.sh` for a
GPU run -- this loads the reference's unmodified `app/main.py` (absent services stubbed,
oracle/ref_loader.py), calls `sqe_b200.plugin.install(main)` and then runs the reference's code:

  * `RAGModel()` + `OpenSearchIndexer.add_embeddings` + `ask_websocket_endpoint` (main.py:650-735)
    on the recorded request sequence of tests/golden/ws_session.* -- every client message and
    every prompt must equal what the reference recorded with its own retrieval path;
  * the lifespan ingest (main.py:568-580 -> `build_embeddings_from_scratch`, main.py:413-456) over
    a synthetic PMC directory of 3,027 files = 32,717 chunks (the size of the reference's own sample
    corpus), through `run_in_executor(add_embeddings)`, followed by websocket
    requests whose expected messages come from the oracle over all 32,717 chunk vectors.

Skipped when no reference copy is present (the driver's GPU box has none)."""
import asyncio
import json
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DIM = 1024


def _reference_root():
    for cand in (os.environ.get("SQE_REFERENCE_ROOT"), "/root/reference",
                 os.path.join(ROOT, "oracle", "_ref", "reference")):
        if cand and os.path.isfile(os.path.join(cand, "app", "main.py")):
            return cand
    return None


@pytest.fixture()
def ref_main():
    root = _reference_root()
    if root is None:
        pytest.skip("no copy of the reference's app/ here (scripts/stage_reference.sh stages one for a GPU run)")
    import sqe_b200
    sqe_b200._native.load()
    from oracle import ref_loader
    ref_loader.REFERENCE_ROOT = root
    m = ref_loader.load_reference_main()
    return sqe_b200, m


class FakeWebSocket:
    def __init__(self, payload):
        self.payload, self.sent, self.closed = payload, [], False

    async def accept(self):
        pass

    async def receive_text(self):
        return self.payload

    async def send_text(self, text):
        self.sent.append(text)

    async def close(self):
        self.closed = True


def _drive(m, requests, prompts):
    results = []
    for req in requests:
        ws = FakeWebSocket(json.dumps(req))
        before = len(prompts)
        asyncio.run(m.ask_websocket_endpoint(ws))
        assert ws.closed
        results.append({"sent": ws.sent, "prompt": prompts[before] if len(prompts) > before else None})
    return results


def _stub_generation(m, prompts):
    async def fake_stream(prompt, system_msg=""):                       # main.py:615-647's interface
        prompts.append(prompt)
        yield "Answer "
        yield f"#{len(prompts)}"
    m.openai_generate_text_stream = fake_stream


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_reference_websocket_handler_on_the_gpu_path_reproduces_its_recorded_session(ref_main, golden_dir, dtype):
    sqe, m = ref_main
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from ws_replay import load_session
    meta, emb, qvec = load_session(golden_dir)
    cache = sqe.plugin.install(m, dtype=dtype, strict=True)
    assert m.OpenSearchIndexer.__mro__[1] is sqe.GpuCorpusIndex
    m.rag_model = m.RAGModel()                                          # main.py:408-411 -> GpuCorpusIndex
    assert isinstance(m.rag_model.os_indexer, sqe.GpuCorpusIndex)
    m.rag_model.os_indexer.add_embeddings(emb, meta["docs"])            # main.py:309-338

    async def fake_ollama(text, model=None):                            # main.py:134-153's result
        return qvec[text].tolist()
    m.ollama_embed_text = fake_ollama
    prompts = []
    _stub_generation(m, prompts)
    results = _drive(m, meta["requests"], prompts)
    assert results == meta["results"]
    assert cache.responses() == meta["final_cache_responses"] and cache.freqs() == meta["final_cache_freqs"]


def test_reference_lifespan_ingest_of_32717_chunks_then_websocket_requests(ref_main, tmp_path):
    sqe, m = ref_main
    import oracle
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ws_replay
    rng = np.random.default_rng(2026)
    n_docs, n_chunks = 3027, 32717                                      # the reference's PMC sample
    per = np.full(n_docs, n_chunks // n_docs)
    per[: n_chunks - per.sum()] += 1
    rng.shuffle(per)
    chunk_of_text, expect_docs = {}, []
    pmc = tmp_path / "PMC"
    pmc.mkdir()
    for d in range(n_docs):
        name = f"PMC{100000 + d}.txt"
        words = []
        for c in range(int(per[d])):
            n_words = m.CHUNK_SIZE if c + 1 < per[d] else int(rng.integers(1, m.CHUNK_SIZE))
            words.append(" ".join([f"{name[:-4]}c{c}"] + ["lorem"] * (n_words - 1)))
        (pmc / name).write_text("\n".join(words))
    (pmc / "README.md").write_text("not a PMC file")                    # ignored by main.py:431
    emb = (rng.standard_normal((n_chunks, DIM)) * rng.uniform(0.3, 3.0, size=(n_chunks, 1))).astype(np.float32)

    # the order the reference will see the chunks in: os.listdir order, then chunk order
    row = 0
    for fname in os.listdir(str(pmc)):
        if fname.startswith("PMC") and fname.endswith(".txt"):
            text = m.basic_cleaning((pmc / fname).read_text())
            for chunk in m.chunk_text(text, m.CHUNK_SIZE):
                chunk_of_text[chunk] = row
                expect_docs.append({"doc_id": fname, "text": chunk})
                row += 1
    assert row == n_chunks

    unit = emb / np.linalg.norm(emb, axis=1, keepdims=True)

    def mix(rows):
        w = np.array([1.0, 0.85, 0.7, 0.58, 0.48, 0.4, 0.33, 0.27][: len(rows)], dtype=np.float32)
        return ((w[:, None] * unit[rows]).sum(axis=0) * np.float32(2.5)).astype(np.float32)

    qvec = {"first question": mix([13, 14, 15, 9000, 9001, 20000, 17, 33]),
            "second question": mix([32716, 32715, 5, 6, 7, 15000, 15001, 15002]),
            "third question": mix([100, 200, 300, 400, 500, 600, 700, 800])}
    qvec["first question again"] = qvec["first question"] * np.float32(1.7)     # cos 1 -> cache hit

    async def fake_ollama(text, model=None):                            # main.py:134-153, stubbed
        if text in qvec:
            return qvec[text]
        return emb[chunk_of_text[text]]
    m.ollama_embed_text = fake_ollama
    cache = sqe.plugin.install(m, dtype="bf16", strict=True)
    m.EMB_DIR = str(pmc)
    prompts = []
    _stub_generation(m, prompts)

    async def lifespan():                                               # main.py:568-580
        async with m.lifespan(m.app):
            pass
    asyncio.run(lifespan())
    index = m.rag_model.os_indexer
    assert isinstance(index, sqe.GpuCorpusIndex) and index.num_rows == n_chunks and index.has_any_data()
    assert index._docs == expect_docs

    requests = [{"query": "first question", "top_k": 8}, {"query": "second question"},
                {"query": "first question again", "top_k": 8}, {"query": ""},
                {"query": "third question", "top_k": 5}, {"query": "second question", "top_k": 6}]
    results = _drive(m, requests, prompts)

    # expected: the same handler flow with the oracle as the retrieval path (stored bf16 rows)
    d_st = oracle.from_storage(oracle.to_storage(oracle.normalize_rows(emb), "bf16"), "bf16")
    model = oracle.LfuCacheModel(m.REDIS_MAX_ITEMS, m.CACHE_SIM_THRESHOLD)

    def os_search(q, k):
        q_st = oracle.from_storage(oracle.to_storage(oracle.normalize_rows(q), "bf16"), "bf16")
        s, i = oracle.topk_cosine(d_st, q_st, k)
        return [(expect_docs[int(r)], float(sc)) for sc, r in zip(s[0], i[0]) if r >= 0]
    want = ws_replay.replay({"requests": requests}, qvec, model.get, model.put, os_search, sqe.build_context_text)
    assert [r["sent"] for r in results] == [w["sent"] for w in want]
    assert [r["prompt"] for r in results] == [w["prompt"] for w in want]
    assert cache.responses() == model.responses() and cache.freqs() == model.freqs()
    print(f"reference lifespan ingest: {n_chunks} chunks of {n_docs} files through run_in_executor(add_embeddings); "
          f"{len(requests)} websocket requests through main.ask_websocket_endpoint, {len(prompts)} prompts, "
          f"cache {cache.freqs()}")


def test_reference_ingest_and_requests_from_text_with_the_gpu_encoder(ref_main, tmp_path):
    """Nothing stubbed between the text and the prompt: `install_encoder(main)` + `plugin.install(main)`,
    then the reference's own `build_embeddings_from_scratch` (main.py:413-456: list the PMC directory,
    `basic_cleaning`, `chunk_text`, `embed_texts_in_batches`, `run_in_executor(add_embeddings)`) and
    `ask_websocket_endpoint` (main.py:650-735: `embed_query`, `lfu_cache_get`, `os_search`, context,
    `lfu_cache_put`) run with the embedding step AND the retrieval step on the GPU.  Expected hits: the
    CPU oracles end to end (bert_oracle embeddings of the same chunks -> numpy_oracle top-k)."""
    sqe, m = ref_main
    import oracle
    from oracle import bert_oracle as bo
    vocab_list = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"] + \
        [w for w in ("gene tumor protein cell patient dose trial cohort enzyme receptor pathway mutation tissue "
                     "serum marker assay sample therapy response control binding expression growth signal").split()] + \
        ["##s", "##ing", "##ed", ".", ","]
    vocab = {t: i for i, t in enumerate(vocab_list)}
    words = vocab_list[5:-5]
    rng = np.random.default_rng(7)
    pmc = tmp_path / "PMC"
    pmc.mkdir()
    n_files = 48
    for d in range(n_files):
        n_words = int(rng.integers(6, 60)) if d % 12 else m.CHUNK_SIZE + int(rng.integers(5, 40))   # a few two-chunk files
        text = " ".join(str(rng.choice(words)) + ("s" if rng.random() < 0.2 else "") for _ in range(n_words))
        (pmc / f"PMC{7000 + d}.txt").write_text(text + " .\n")
    (pmc / "notes.txt").write_text("ignored")
    w = bo.random_bert_weights(91, layers=2, vocab=len(vocab_list))
    dev = torch.device("cuda", 0)
    enc = sqe.GpuEmbeddingEncoder(sqe.EncoderWeights.from_state_dict(w, device=dev), sqe.WordPieceTokenizer(vocab))
    sqe.install_encoder(m, enc)
    cache = sqe.plugin.install(m, dtype="fp32", strict=True)
    m.EMB_DIR = str(pmc)
    prompts = []
    _stub_generation(m, prompts)

    async def lifespan():
        async with m.lifespan(m.app):
            pass
    asyncio.run(lifespan())
    index = m.rag_model.os_indexer
    docs = []
    for fname in os.listdir(str(pmc)):
        if fname.startswith("PMC") and fname.endswith(".txt"):
            for chunk in m.chunk_text(m.basic_cleaning((pmc / fname).read_text()), m.CHUNK_SIZE):
                docs.append({"doc_id": fname, "text": chunk})
    assert index.num_rows == len(docs) > n_files and index._docs == docs

    # the oracle's view of the same corpus and queries
    emb = bo.bert_embed(w, [bo.encode_text(d["text"], vocab) for d in docs]).numpy()
    d_n = oracle.normalize_rows(emb)
    free = "tumor gene expression in patient serum"
    f_n = oracle.normalize_rows(bo.bert_embed(w, [bo.encode_text(free, vocab)]).numpy())
    # random-weight encoders put all texts in a narrow cone (mean cosine ~0.8 here, some pairs above the
    # cache threshold 0.96): pick query chunks that are no cache hits for one another or for the free text
    picks = []
    for i in range(len(docs)):
        if all(float(d_n[i] @ d_n[j]) < 0.92 for j in picks) and float(d_n[i] @ f_n[0]) < 0.92:
            picks.append(i)
        if len(picks) == 3:
            break
    assert len(picks) == 3
    queries = [docs[i]["text"] for i in picks] + [free, "   "]
    requests = [{"query": q, "top_k": 4} for q in queries] + [{"query": queries[0], "top_k": 4}]
    results = _drive(m, requests, prompts)
    n_prompts = 0
    for req, res in zip(requests[: len(queries)], results):
        q = req["query"]
        if not q.strip():
            assert res["sent"] == ["[ERROR] Empty query."] and res["prompt"] is None
            continue
        n_prompts += 1
        q_n = oracle.normalize_rows(bo.bert_embed(w, [bo.encode_text(q, vocab)]).numpy())
        s, idx = oracle.topk_cosine(d_n, q_n, 5)
        s, idx = s[0], idx[0]
        # ranks whose oracle score is separated from the next by more than the fp16-operand error of the
        # GPU encoder (a few 1e-3 on a cosine) must come out identically; the handler's prompt holds them
        sure = 0
        while sure < 4 and s[sure] - s[sure + 1] > 2e-2:
            sure += 1
        hits = index.search(asyncio.run(m.embed_query(q)), k=4)
        assert [h[0] for h in hits[:sure]] == [docs[int(r)] for r in idx[:sure]], (q[:40], s, sure)
        assert all(abs(h[1] - float(sc)) < 1e-2 for h, sc in zip(hits, s)), (hits, s)
        assert res["prompt"] == (f"User Query:\n{q}\n\nContext:\n{sqe.build_context_text(hits)}\n"
                                 "--- End of context ---\n\nProvide your concise answer now.")
        assert res["sent"] == ["Answer ", f"#{n_prompts}"]
    for i, res in zip(picks, results):                                  # a chunk's own text finds that chunk first
        assert f"--- Document ID: {docs[i]['doc_id']} ---\n" in res["prompt"].split("Context:\n")[1][:60]
    # the repeated first query is answered from the cache (cos = 1 >= 0.96): no new prompt, freq 2
    assert results[-1] == {"sent": ["Answer #1"], "prompt": None} and len(prompts) == n_prompts
    assert cache.freqs()[-1] == 2 and sorted(cache.freqs()) == [1] * (n_prompts - 1) + [2]
    print(f"reference text ingest with the GPU encoder: {len(docs)} chunks of {n_files} files, "
          f"{len(requests)} websocket requests, {len(prompts)} prompts, cache {cache.freqs()}")


def test_reference_upload_service_from_text_with_the_gpu_encoder(tmp_path):
    """The upload micro-service, literally: the reference's own `upload_text` handler
    (embedding_gen.py:315-410: authorisation, file checks, save, `chunk_text`, `embed_texts_in_batches`,
    `bulk_index_embeddings`) with `install_encoder` + `plugin.install_embedding_gen` -- files in, per-user GPU
    indices out; only Postgres is stubbed.  Expected payload from the handler's own rules, expected vectors
    from the CPU oracles."""
    import io
    import types
    root = _reference_root()
    if root is None:
        pytest.skip("no copy of the reference's app/ here (scripts/stage_reference.sh stages one for a GPU run)")
    import oracle
    import sqe_b200 as sqe
    from oracle import bert_oracle as bo
    from oracle import ref_loader
    sqe._native.load()
    ref_loader.REFERENCE_ROOT = root
    eg = ref_loader.load_reference_embedding_gen(base_index_name="docs")
    vocab_list = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"] + \
        "gene tumor protein cell patient dose trial cohort enzyme receptor pathway mutation".split() + ["##s", ".", ","]
    vocab = {t: i for i, t in enumerate(vocab_list)}
    w = bo.random_bert_weights(19, layers=2, vocab=len(vocab_list))
    dev = torch.device("cuda", 0)
    enc = sqe.GpuEmbeddingEncoder(sqe.EncoderWeights.from_state_dict(w, device=dev), sqe.WordPieceTokenizer(vocab))
    reg = sqe.plugin.install_embedding_gen(eg, dtype="fp32", strict=True)
    sqe.install_encoder(eg, enc)
    assert enc.blank_policy == "embedding_gen"

    async def authorised(user_id):                                   # embedding_gen.py:282-309, Postgres stubbed
        return user_id != "mallory"
    eg.check_user_authorized_in_postgres = authorised
    eg.BASE_UPLOAD_DIR = str(tmp_path / "uploads")
    clock = [1700000000.7]
    eg.time = types.SimpleNamespace(time=lambda: clock[0])             # doc_id = f"{stem}_{int(time.time())}"

    class Upload:
        def __init__(self, filename, text):
            self.filename, self.file = filename, io.BytesIO(text.encode("utf-8"))

    rng = np.random.default_rng(3)
    words = vocab_list[5:-3]

    def text_of(n_words):
        return " ".join(str(rng.choice(words)) + ("s" if rng.random() < 0.25 else "") for _ in range(n_words)) + " ."
    paper, notes, bobs = text_of(eg.CHUNK_SIZE + 40), text_of(25), text_of(60)
    msg = asyncio.run(eg.upload_text(user_id="alice", files=[Upload("paper.txt", paper), Upload("notes.txt", notes)]))
    assert msg == "Uploaded 2 files & embedded documents for user='alice'."
    assert asyncio.run(eg.upload_text(user_id="bob", files=[Upload("b.txt", bobs)])).startswith("Uploaded 1 files")
    assert sorted(os.listdir(tmp_path / "uploads" / "alice")) == ["notes_1700000000.txt", "paper_1700000000.txt"]

    want_docs = [("paper_1700000000", c) for c in eg.chunk_text(paper, eg.CHUNK_SIZE)] + \
                [("notes_1700000000", c) for c in eg.chunk_text(notes, eg.CHUNK_SIZE)]
    assert len(want_docs) == 3
    idx = reg.get("alice")
    assert idx is not None and idx.index_name == "docs-alice" and idx.num_rows == 3 and reg.get("bob").num_rows == 1
    assert [idx._source(r) for r in range(3)] == [{"doc_id": d, "text": c} for d, c in want_docs]
    assert [idx.doc_id_of(r) for r in range(3)] == ["paper_1700000000_0", "paper_1700000000_1", "notes_1700000000_0"]
    # stored rows = normalised encoder outputs: against the fp32 oracles (fp16-operand error of the GPU encoder)
    want = oracle.normalize_rows(bo.bert_embed(w, [bo.encode_text(c, vocab) for _, c in want_docs]).numpy())
    got = idx.shard.float().cpu().numpy()
    assert np.abs(got - want).max() < 2e-3 and (got * want).sum(1).min() > 0.9999
    # a chunk's own text finds that chunk in its owner's index, and nothing in another user's
    q = asyncio.run(eg.embed_texts_in_batches([want_docs[1][1]]))
    hit = reg.search("alice", q, k=1)[0]
    assert hit[0] == {"doc_id": "paper_1700000000", "text": want_docs[1][1]} and hit[1] > 0.9999
    assert reg.search("bob", q, k=3)[0][0]["doc_id"] == "b_1700000000" and reg.search("carol", q, k=3) == []
    # a later re-upload is a NEW document (the timestamp is part of the doc_id): rows are appended
    clock[0] += 60.0
    asyncio.run(eg.upload_text(user_id="alice", files=[Upload("notes.txt", notes)]))
    assert idx.num_rows == 4 and idx.doc_id_of(3) == "notes_1700000060_0"
    # the handler's refusals are unchanged (embedding_gen.py:332-336, :351-355, :386-390)
    for uid, files, status in (("mallory", [Upload("a.txt", "gene")], 403), ("alice", [Upload("a.pdf", "gene")], 403),
                               ("alice", [Upload("a.txt", "   ")], 400), ("alice", [], 400)):
        with pytest.raises(eg.HTTPException) as ei:
            asyncio.run(eg.upload_text(user_id=uid, files=files))
        assert ei.value.status_code == status
    assert idx.num_rows == 4
    print(f"reference upload service with the GPU encoder: 3 uploads of 2 users, indices "
          f"{ {u: reg.get(u).num_rows for u in ('alice', 'bob')} }, 4 refusals")

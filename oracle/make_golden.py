"""Generate tests/golden/*.npz|json by running the REFERENCE's own functions.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python -m oracle.make_golden

Inputs are seeded; outputs are what `/root/reference/app/main.py`'s unmodified
`cosine_similarity`, `lfu_cache_get`, `lfu_cache_put`,
`_remove_least_frequent_item`, `OpenSearchIndexer.add_embeddings/.search`, and the websocket
handler `ask_websocket_endpoint` (a whole request sequence, `gen_ws_session`)
return (external services stubbed as in ref_loader.py).  Two module constants
are overridden for some scenarios -- `REDIS_MAX_ITEMS` (1000 -> 8, so LFU
eviction is reachable with a small fixture) and `CACHE_SIM_THRESHOLD`
(0.96 -> 0.95, BASELINE.json config 5) -- the functions read them at call time.
"""
from __future__ import annotations

import contextlib
import io
import json
import os

import numpy as np

from .ref_loader import FakeOpenSearch, load_reference_embedding_gen, load_reference_main

DIM = 1024
GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                      "tests", "golden")


def unit(v):
    return v / np.linalg.norm(v)


def with_cosine(rng, u, c):
    """A vector whose cosine with u is c (before fp32 rounding)."""
    u = unit(u.astype(np.float64))
    w = rng.standard_normal(DIM)
    w = unit(w - (w @ u) * u)
    return (c * u + np.sqrt(1.0 - c * c) * w)


def gen_cosine(m, rng):
    a = rng.standard_normal((24, DIM)).astype(np.float32)
    b = rng.standard_normal((24, DIM)).astype(np.float32)
    b[1] = a[1]                       # identical
    b[2] = -a[2]                      # antiparallel
    b[3] = a[3] * np.float32(7.25)    # scaled
    a[4] = 0.0                        # zero lhs
    b[5] = 0.0                        # zero rhs
    a[6] = 0.0
    b[6] = 0.0
    a[7] *= np.float32(1e-20)         # tiny norm (squares underflow in fp32 dot)
    b[8] *= np.float32(1e15)          # huge but finite in fp32
    for i, c in enumerate((0.94, 0.95, 0.9599, 0.9601, 0.97, 0.5, -0.3)):
        b[9 + i] = (with_cosine(rng, a[9 + i], c) * 3.0).astype(np.float32)
    out = np.array([m.cosine_similarity(a[i], b[i]) for i in range(len(a))], dtype=np.float64)
    np.savez_compressed(os.path.join(GOLDEN, "cosine.npz"), a=a, b=b, out=out)


def gen_normalize_and_search(m, rng):
    n = 192
    emb = (rng.standard_normal((n, DIM)) * rng.uniform(0.05, 20.0, size=(n, 1))).astype(np.float32)
    emb[5] = 0.0                               # zero row (embedding_gen.py:147-148 can yield one)
    emb[6] = emb[7]                            # duplicate rows -> exact tie
    emb[40] = emb[7] * np.float32(2.0)         # same direction, different norm
    emb[9] *= np.float32(1e-12)                # norm ~ 3e-11 < 1e-9: the +1e-9 matters
    emb[10] *= np.float32(1e-8)
    emb[11] *= np.float32(1e13)
    docs = [{"doc_id": f"PMC{1000 + i // 4}", "text": f"chunk {i}"} for i in range(n)]

    client = FakeOpenSearch()
    indexer = m.OpenSearchIndexer(client, "golden-index")
    assert indexer.has_any_data() is False
    indexer.add_embeddings(emb, docs)
    assert indexer.has_any_data() is True
    stored = np.asarray([s["embedding"] for _, s in client.docs], dtype=np.float32)
    ids = [i for i, _ in client.docs]

    q = rng.standard_normal((10, DIM)).astype(np.float32) * np.float32(3.0)
    q[1] = emb[7] * np.float32(0.5)            # hits the duplicate pair 6/7 (+40)
    q[2] = 0.0                                 # zero query: all scores 0 -> indices 0..k-1
    q[3] = (with_cosine(rng, emb[20], 0.9) * 4).astype(np.float32)
    ks = [1, 3, 5, 10]
    res_idx = np.full((len(q), len(ks), max(ks)), -1, dtype=np.int64)
    res_score = np.full((len(q), len(ks), max(ks)), np.nan, dtype=np.float64)
    q_norm = np.zeros_like(q)
    for qi in range(len(q)):
        for ki, k in enumerate(ks):
            hits = indexer.search(q[qi:qi + 1], k=k)
            q_norm[qi] = np.asarray(client.last_query["query"]["knn"]["embedding"]["vector"],
                                    dtype=np.float32)
            for r, (src, score) in enumerate(hits):
                res_idx[qi, ki, r] = int(src["text"].split()[1])
                res_score[qi, ki, r] = score
    empty = indexer.search(np.array([]), k=3)
    assert empty == []
    np.savez_compressed(os.path.join(GOLDEN, "index_search.npz"),
                        emb=emb, stored=stored, q=q, q_norm=q_norm, ks=np.array(ks),
                        res_idx=res_idx, res_score=res_score)
    with open(os.path.join(GOLDEN, "index_search.json"), "w") as f:
        json.dump({"ids": ids, "docs": docs}, f)


def run_cache_scenario(m, rng, max_items, threshold, n_ops, name):
    m.REDIS_MAX_ITEMS = max_items
    m.CACHE_SIM_THRESHOLD = threshold
    m.redis_client.lists.clear()
    vecs = []
    ops = []

    def add_vec(v):
        vecs.append(np.asarray(v, dtype=np.float32))
        return len(vecs) - 1

    def do_get(vi):
        r = m.lfu_cache_get(vecs[vi][None, :])
        ops.append({"op": "get", "vec": vi, "result": r})
        return r

    def do_put(vi, resp):
        m.lfu_cache_put(vecs[vi][None, :], resp)
        ops.append({"op": "put", "vec": vi, "response": resp})

    # empty cache -> None (main.py:70-71)
    v0 = add_vec(rng.standard_normal(DIM) * 2.5)
    do_get(v0)
    do_put(v0, "answer-0")
    do_get(v0)                                   # exact hit, freq 1 -> 2
    # threshold ladder around the configured threshold
    for c in (threshold - 0.02, threshold - 1e-4, threshold + 1e-4, threshold + 0.02):
        do_get(add_vec(with_cosine(rng, vecs[v0], c) * 1.7))
    # duplicate entries: newest (list index 0) must win the tie (main.py:84,128)
    v1 = add_vec(rng.standard_normal(DIM))
    do_put(v1, "dup-old")
    do_put(v1, "dup-new")
    do_get(v1)
    # zero vector query / zero vector entry
    vz = add_vec(np.zeros(DIM))
    do_get(vz)
    do_put(vz, "zero-entry")
    do_get(vz)
    do_get(v0)
    # random traffic with repeats -> exercises LFU eviction when max_items is small
    pool = [add_vec(rng.standard_normal(DIM) * rng.uniform(0.5, 3)) for _ in range(14)]
    for t in range(n_ops):
        vi = pool[int(rng.integers(len(pool)))]
        if rng.random() < 0.5:
            near = add_vec(with_cosine(rng, vecs[vi], float(rng.choice([0.90, 0.99]))))
            do_get(near)
        else:
            if do_get(vi) is None:
                do_put(vi, f"resp-{t}")
    final = [json.loads(s) for s in m.redis_client.lrange(m.REDIS_CACHE_LIST, 0, -1)]
    np.savez_compressed(os.path.join(GOLDEN, f"cache_{name}.npz"), vecs=np.stack(vecs))
    with open(os.path.join(GOLDEN, f"cache_{name}.json"), "w") as f:
        json.dump({"max_items": max_items, "threshold": threshold, "ops": ops,
                   "final_responses": [e["response"] for e in final],
                   "final_freqs": [e["freq"] for e in final]}, f)


def gen_ws_session(m, seed):
    """The reference's OWN websocket handler (`ask_websocket_endpoint`, main.py:650-735), driven
    request by request with a fake socket: embed_query (main.py:172-180, its Ollama call stubbed
    with fixed vectors) -> lfu_cache_get -> RAGModel.os_search -> the doc-id grouping and prompt
    building of main.py:685-717 -> generation (stubbed: records the prompt, streams two tokens) ->
    lfu_cache_put.  Recorded: every message sent to the client and the prompt handed to the LLM --
    what a drop-in for the retrieval path must reproduce.  Uses its own rng so the older fixtures
    stay byte-identical."""
    import asyncio
    rng = np.random.default_rng(seed)
    n, per_doc = 240, 6
    emb = (rng.standard_normal((n, DIM)) * rng.uniform(0.2, 5.0, size=(n, 1))).astype(np.float32)
    docs = [{"doc_id": f"PMC{2000 + i // per_doc}", "text": f"passage {i} of PMC{2000 + i // per_doc}"}
            for i in range(n)]
    unit = emb / np.linalg.norm(emb, axis=1, keepdims=True)

    def mix(rows):
        w = np.array([1.0, 0.85, 0.7, 0.58, 0.48, 0.4, 0.33, 0.27][: len(rows)], dtype=np.float32)
        return ((w[:, None] * unit[rows]).sum(axis=0) * np.float32(2.5)).astype(np.float32)

    qvec = {
        "what does passage thirteen say": mix([13, 14, 15, 90, 91, 200, 17, 33]),      # doc 2002 three times
        "tell me about document 2001": mix([7, 6, 8, 9, 10, 11, 100, 150]),          # six chunks of one doc
        "a different question": mix([120, 5, 121, 60, 122, 61, 62, 123]),
    }
    base = qvec["what does passage thirteen say"]
    qvec["what does passage 13 say?"] = (with_cosine(rng, base, 0.97) * 1.3).astype(np.float32)    # cache hit
    qvec["what might passage thirteen say"] = (with_cosine(rng, base, 0.95) * 0.8).astype(np.float32)  # miss
    requests = [
        {"query": "what does passage thirteen say", "top_k": 8},
        {"query": "tell me about document 2001"},                       # default top_k = 3 (main.py:667)
        {"query": "what does passage 13 say?", "top_k": 8},             # >= 0.96 to request 0: cached answer
        {"query": "   "},                                               # blank -> "[ERROR] Empty query."
        {"query": "what might passage thirteen say", "top_k": 5},       # 0.95 < 0.96: a miss, retrieval again
        {"query": "tell me about document 2001", "top_k": 6},           # exact repeat: cached answer
        {"query": "a different question", "top_k": 7},
    ]

    m.REDIS_MAX_ITEMS, m.CACHE_SIM_THRESHOLD = 1000, 0.96
    m.redis_client.lists.clear()
    m.os_client = FakeOpenSearch()
    m.rag_model = m.RAGModel()                                          # main.py:408-411
    m.rag_model.os_indexer.add_embeddings(emb, docs)                    # main.py:309-338

    async def fake_ollama(text, model=None):                            # main.py:134-153's result
        return qvec[text].tolist()
    m.ollama_embed_text = fake_ollama
    prompts = []

    async def fake_stream(prompt, system_msg=""):                       # main.py:615-647's interface
        prompts.append(prompt)
        yield "Answer "
        yield f"#{len(prompts)}"
    m.openai_generate_text_stream = fake_stream

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

    results = []
    for req in requests:
        ws = FakeWebSocket(json.dumps(req))
        before = len(prompts)
        asyncio.run(m.ask_websocket_endpoint(ws))
        assert ws.closed
        results.append({"sent": ws.sent, "prompt": prompts[before] if len(prompts) > before else None})
    cache = [json.loads(x) for x in m.redis_client.lrange(m.REDIS_CACHE_LIST, 0, -1)]
    assert [bool(r["prompt"]) for r in results] == [True, True, False, False, True, False, True]
    names = sorted(qvec)
    np.savez_compressed(os.path.join(GOLDEN, "ws_session.npz"), emb=emb,
                        qvecs=np.stack([qvec[t] for t in names]))
    with open(os.path.join(GOLDEN, "ws_session.json"), "w") as f:
        json.dump({"docs": docs, "query_texts": names, "requests": requests, "results": results,
                   "final_cache_responses": [e["response"] for e in cache],
                   "final_cache_freqs": [e["freq"] for e in cache]}, f)


def gen_upload_service(seed):
    """The reference's upload micro-service (`app/embedding_gen.py`): its own
    `bulk_index_embeddings(user_id, doc_id, embeddings, chunks)` (:196-257, normalise :215-216,
    `_id = f"{doc_id}_{i}"` :221, batches of 64 :236) and `init_user_index` (:83-122) on a call
    sequence with two users, a document longer than one bulk batch, a zero chunk embedding
    (:147-148 returns one on an embedding error), the same doc_id under two users, a re-upload
    with fewer chunks (ids _0.._2 replaced, _3 and _4 stay) and an empty call (:207-209).
    Recorded per user index: the documents in stored order and the stored vectors."""
    g = load_reference_embedding_gen("docs")
    rng = np.random.default_rng(seed)
    calls = []

    def call(user, doc_id, n_chunks, tag):
        e = (rng.standard_normal((n_chunks, DIM)) * rng.uniform(0.3, 4.0, size=(max(n_chunks, 1), 1))[:n_chunks]).astype(np.float32)
        if n_chunks > 10:
            e[9] = 0.0
        chunks = [f"{tag} chunk {i}" for i in range(n_chunks)]
        g.bulk_index_embeddings(user, doc_id, e if n_chunks else np.array([]), chunks)
        calls.append({"user": user, "doc_id": doc_id, "chunks": chunks, "emb": e})
    g.init_user_index("alice")
    g.init_user_index("alice")                                          # :92-94: already exists
    call("alice", "notes_1700000000", 70, "notes v1")
    call("alice", "paper_1700000001", 5, "paper v1")
    call("bob", "notes_1700000000", 3, "bob notes")
    call("alice", "paper_1700000001", 3, "paper v2")
    call("bob", "empty_1700000002", 0, "nothing")
    indices = {}
    stored = {}
    for name, docs in g.os_client.by_index.items():
        indices[name] = [{"_id": i, "doc_id": s["doc_id"], "text": s["text"]} for i, s in docs]
        stored[name] = np.asarray([s["embedding"] for _, s in docs], dtype=np.float32).reshape(len(docs), DIM)
    mapping = g.os_client.bodies["docs-alice"]["mappings"]["properties"]["embedding"]
    assert mapping["dimension"] == 1024 and mapping["method"]["space_type"] == "cosinesimil"
    arrays = {f"emb_{i}": c["emb"] for i, c in enumerate(calls)}
    arrays.update({f"stored_{k}": v for k, v in stored.items()})
    np.savez_compressed(os.path.join(GOLDEN, "upload_service.npz"), **arrays)
    with open(os.path.join(GOLDEN, "upload_service.json"), "w") as f:
        json.dump({"base_index_name": "docs",
                   "calls": [{"user": c["user"], "doc_id": c["doc_id"], "chunks": c["chunks"]} for c in calls],
                   "indices": indices}, f)


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    m = load_reference_main()
    assert m.CACHE_SIM_THRESHOLD == 0.96 and m.REDIS_MAX_ITEMS == 1000 and m.EMBED_DIM == 1024
    rng = np.random.default_rng(20261018)
    with contextlib.redirect_stdout(io.StringIO()):      # the reference print()s per call
        gen_cosine(m, rng)
        gen_normalize_and_search(m, rng)
        run_cache_scenario(m, rng, max_items=1000, threshold=0.96, n_ops=40, name="default")
        run_cache_scenario(m, rng, max_items=8, threshold=0.96, n_ops=60, name="evict8")
        run_cache_scenario(m, rng, max_items=8, threshold=0.95, n_ops=40, name="thr095")
        gen_ws_session(m, 20261019)
        gen_upload_service(20261020)
    print("golden vectors written to", GOLDEN)


if __name__ == "__main__":
    main()

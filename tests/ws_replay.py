"""Replay of the reference's websocket request sequence (tests/golden/ws_session.*, recorded by
running the reference's own `ask_websocket_endpoint`, main.py:650-735, see oracle/make_golden.py)
through a replacement of its retrieval path.  `handle()` below is the handler's control flow with
the retrieval calls injected; the generation step streams the same two tokens the recording's
stub did."""
import json
import os

import numpy as np


def load_session(golden_dir):
    with open(os.path.join(golden_dir, "ws_session.json")) as f:
        meta = json.load(f)
    arr = np.load(os.path.join(golden_dir, "ws_session.npz"))
    qvec = {t: arr["qvecs"][i] for i, t in enumerate(meta["query_texts"])}
    return meta, arr["emb"], qvec


def replay(meta, qvec, cache_get, cache_put, os_search, build_context_text):
    """Returns [{"sent": [...], "prompt": str | None}] for meta["requests"]."""
    out, n_prompts = [], 0
    for req in meta["requests"]:
        query = req.get("query", "")
        if not query.strip():                                        # main.py:660-663
            out.append({"sent": ["[ERROR] Empty query."], "prompt": None})
            continue
        top_k = req.get("top_k", 3)                                  # main.py:667
        query_emb = np.array([qvec[query]], dtype=np.float32)        # main.py:675 (embed_query's shape)
        cached = cache_get(query_emb)                                # main.py:676
        if cached:
            out.append({"sent": [cached], "prompt": None})           # main.py:677-681
            continue
        hits = os_search(query_emb, top_k)                           # main.py:684
        context_text = build_context_text(hits)                      # main.py:685-698
        prompt = (f"User Query:\n{query}\n\n"                        # main.py:710-715
                  f"Context:\n{context_text}\n"
                  "--- End of context ---\n\n"
                  "Provide your concise answer now.")
        n_prompts += 1
        chunks = ["Answer ", f"#{n_prompts}"]
        answer = "".join(chunks)
        if answer.strip():
            cache_put(query_emb, answer)                             # main.py:725-727
        out.append({"sent": chunks, "prompt": prompt})
    return out

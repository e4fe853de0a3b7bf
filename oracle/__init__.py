"""CPU oracle for the retrieval hot path.  TEST INFRASTRUCTURE ONLY.

This package is a numpy restatement of the reference's similarity path
(`/root/reference/app/main.py`, `app/embedding_gen.py`) plus an independent plain-C
twin of it (`c_oracle.c`, compiled with gcc by `__graft_entry__.build()`, wrapped by
`c_oracle.py`).  It exists so the CUDA path can be checked; it is never the thing
shipped or measured.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py` (its `cpu_baseline`
leg and `--impl reference` arm) may import it.  Nothing under
`semantic-query-engine_b200/` imports it, and the product path raises when
the CUDA library is missing rather than falling back to this code.

Parity pin: PINNED.  `oracle/make_golden.py` imports the reference's own
functions (with absent third-party services stubbed, see `ref_loader.py`),
runs them on seeded inputs -- single functions, a whole websocket request sequence through the
reference's handler, and the upload service's `bulk_index_embeddings` (`embedding_gen.py`) -- and
commits inputs+outputs under `tests/golden/`;
`tests/test_oracle_golden.py` checks every function here -- numpy and C -- against those
vectors.  The one leg with no reference arithmetic to pin is corpus
scoring/top-k: the reference delegates it to an external, unpinned
OpenSearch HNSW index (`app/main.py:356-361`); `numpy_oracle.topk_cosine`
restates it as the reference's own cosine (`main.py:59-64`) applied to every
row plus a stable sort (exact, ties -> lower index), as SURVEY.md §8c fixes.
"""
from .numpy_oracle import (  # noqa: F401
    EMBED_DIM,
    CACHE_SIM_THRESHOLD,
    REDIS_MAX_ITEMS,
    cosine_similarity,
    normalize_rows,
    pairwise_sumsq_f32,
    to_storage,
    from_storage,
    topk_cosine,
    topk_from_scores,
    cache_lookup,
    merge_topk,
    LfuCacheModel,
    opensearch_score,
    quantize_rows_int8,
    prefilter_bounds,
    k3_score_emulation,
)

"""Patch a loaded reference `main` module so its handlers use the GPU path.

    import main                      # the reference's app/main.py
    import sqe_b200
    sqe_b200.plugin.install(main, dtype="bf16")

After `install`, `main.OpenSearchIndexer(client, index_name)` builds a
`GpuCorpusIndex`, and `main.lfu_cache_get` / `main.lfu_cache_put` /
`main.cosine_similarity` are served by a `GpuQueryCache` / the same kernels -- the names, arguments and
return values RAGModel.ask (main.py:493,499,547) and the websocket handler
(main.py:676,684,727) already use.  See INTEGRATION.md.
"""
from __future__ import annotations

import numpy as np

from .cache import GpuQueryCache, cosine_similarity as gpu_cosine_similarity
from .corpus import GpuCorpusIndex


class GpuIndexStandIn:
    """Truthy placeholder for `main.os_client` when no OpenSearch server is running:
    `RAGModel.__init__` only builds its indexer `if os_client` (main.py:408-411), and the reason to
    install this plugin is that there is no such server any more."""

    def __repr__(self) -> str:
        return "<sqe_b200: GPU index, no OpenSearch client>"


def install(main_module, *, dtype: str = "bf16", cache_dtype: str = "fp32", device=None,
            write_through_redis: bool = False, strict: bool = False, prefilter: bool = True):
    """`write_through_redis`: mirror every cache mutation into `main.redis_client` in the
    reference's entry format and take over what that list already holds (warm start).
    `prefilter`: the handlers issue ONE query per request (main.py:499, :684), the shape the
    int8-prefiltered scan (K3p) serves: the same hits as the exact scan, bit for bit, at about half
    the HBM bytes per request, for +1 KB of HBM per stored row.  False keeps only the shard."""
    threshold = getattr(main_module, "CACHE_SIM_THRESHOLD", 0.96)
    max_items = getattr(main_module, "REDIS_MAX_ITEMS", 1000)
    list_name = getattr(main_module, "REDIS_CACHE_LIST", "query_cache_lfu")
    redis_client = getattr(main_module, "redis_client", None) if write_through_redis else None
    cache = GpuQueryCache(max_items=max_items, threshold=threshold, dtype=cache_dtype,
                          device=device, redis_client=redis_client, list_name=list_name)

    if redis_client is not None:
        try:
            n = cache.load_from_redis()
            if n:
                print(f"[sqe_b200] query cache warm-started with {n} entries from Redis list {list_name!r}")
        except Exception as e:                           # a broken entry must not stop the service
            if strict:
                raise
            print(f"[sqe_b200] could not warm-start the query cache from Redis: {e}")
    if getattr(main_module, "os_client", None) is None:
        main_module.os_client = GpuIndexStandIn()

    class OpenSearchIndexer(GpuCorpusIndex):
        def __init__(self, client=None, index_name: str = ""):
            super().__init__(client, index_name, dtype=dtype, device=device, strict=strict,
                             prefilter=prefilter)

    def lfu_cache_get(query_emb: np.ndarray):
        return cache.get(query_emb)

    def lfu_cache_put(query_emb: np.ndarray, response: str):
        cache.put(query_emb, response)

    def cosine_similarity(a: np.ndarray, b: np.ndarray) -> float:
        return gpu_cosine_similarity(a, b, device=device)

    main_module.OpenSearchIndexer = OpenSearchIndexer
    main_module.cosine_similarity = cosine_similarity
    main_module.lfu_cache_get = lfu_cache_get
    main_module.lfu_cache_put = lfu_cache_put
    main_module._sqe_b200_cache = cache
    return cache


def install_embedding_gen(module, *, dtype: str = "bf16", device=None, strict: bool = False):
    """Patch a loaded reference `embedding_gen` module (the upload micro-service): its
    `init_user_index(user_id)` and `bulk_index_embeddings(user_id, doc_id, embeddings, chunks)`
    (embedding_gen.py:83, :196) then build / fill per-user `GpuCorpusIndex` objects."""
    from .serving import UserIndexRegistry
    base = getattr(module, "BASE_OPENSEARCH_INDEX_NAME", "docs")
    reg = UserIndexRegistry(base, dtype=dtype, device=device, strict=strict)
    module.init_user_index = lambda user_id: (reg.init_user_index(user_id), None)[1]
    module.bulk_index_embeddings = reg.bulk_index_embeddings
    module._sqe_b200_registry = reg
    return reg

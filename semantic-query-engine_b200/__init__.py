"""semantic-query-engine_b200 -- B200-native retrieval hot path.

The directory name carries the reference's name (with its hyphens), so import it
through the alias module at the repository root:

    import sqe_b200
    index = sqe_b200.GpuCorpusIndex(dtype="bf16")

Contents: the host-side mirror of the reference's retrieval interface
(`GpuCorpusIndex` ~ `OpenSearchIndexer`, `GpuQueryCache` ~ `lfu_cache_get/put`),
the corpus-sharded multi-GPU mode (`ShardedCorpusIndex`), `plugin.install()` to
patch a loaded reference `main` module, and `csrc/` (the sm_100a kernels + C ABI).
"""
from . import _native
from ._native import NativeLibraryMissing, SqeError
from .cache import (CACHE_SIM_THRESHOLD, REDIS_CACHE_LIST, REDIS_MAX_ITEMS, GpuQueryCache,
                    cosine_similarity)
from .corpus import EMBED_DIM, GpuCorpusIndex
from .sharded import ShardedCorpusIndex, shard_bounds
from .serving import MicroBatcher, UserIndexRegistry, build_context_text, group_hits_by_doc
from .encoder import EncoderWeights, GpuEmbeddingEncoder, WordPieceTokenizer, install_encoder
from . import encoder, gguf_model, ops, plugin

__all__ = [
    "GpuCorpusIndex", "GpuQueryCache", "cosine_similarity", "ShardedCorpusIndex", "shard_bounds", "ops", "plugin",
    "MicroBatcher", "UserIndexRegistry", "build_context_text", "group_hits_by_doc",
    "NativeLibraryMissing", "SqeError", "EMBED_DIM", "CACHE_SIM_THRESHOLD", "REDIS_MAX_ITEMS",
    "REDIS_CACHE_LIST", "GpuEmbeddingEncoder", "EncoderWeights", "WordPieceTokenizer", "install_encoder", "encoder", "gguf_model",
]

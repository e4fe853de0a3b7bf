"""Corpus-sharded multi-GPU search (north_star subsystem 4).

One process per GPU.  Rank r owns the contiguous row block
[r*N/G, (r+1)*N/G); queries are replicated; every rank computes its local top-k with
global row numbers (idx_offset = first row of the shard), the [B,k] (score fp32, row int64) lists are exchanged over NVLink and every rank merges the
G lists.  Two exchange paths:
  * "p2p"  (default on CUDA when symmetric memory can be set up): ONE kernel per rank pushes
    its lists into every rank's peer-mapped buffer, flags, waits and merges (K4x,
    csrc/exchange.cu) -- no collective launch at all;
  * "nccl": two all-gathers + the K4 merge kernel (also what the gloo CPU test exercises).  top-k of a union is the top-k of the per-shard top-k's, so the result is
exactly the single-GPU result, ties included (global rows keep the lower-index rule).

Nothing like this exists in the reference (single process, external index); the
exchange step is the only collective on the path.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n_total: int, world: int, rank: int) -> Tuple[int, int]:
    """Rows [lo, hi) of `rank`: contiguous blocks, remainder spread over the first ranks."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


class ShardedCorpusIndex:
    """Wraps a per-rank index (normally a GpuCorpusIndex holding this rank's rows).

    `local_topk(q, k, idx_offset) -> (scores[B,k], rows[B,k])` and
    `merge(scores[G,B,k], rows[G,B,k], k) -> (scores[B,k], rows[B,k])` default to the CUDA
    kernels; the CPU tests inject checkers there to exercise the sharding arithmetic and
    the collective with the gloo backend."""

    def __init__(self, local_index=None, *, group=None,
                 local_topk: Optional[Callable] = None, merge: Optional[Callable] = None,
                 exchange: str = "auto"):
        if exchange not in ("auto", "p2p", "nccl"):
            raise ValueError("exchange must be 'auto', 'p2p' or 'nccl'")
        self.local = local_index
        self.group = group
        self._exchange_req = exchange
        self.exchange = "nccl"          # decided collectively on first use
        self._xchg = None               # (buffer, handle, peer pointers, capacity)
        self._p2p_failed = False        # symmetric memory could not be set up: stay on NCCL (reset_exchange())
        self._epoch = 0
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._local_topk = local_topk
        self._merge = merge
        self.row_offset = 0
        self.total_rows = 0
        self._gather_s = None
        self._gather_i = None
        self.fuse_small_batches = True  # b <= 2: exchange fused into the scan kernel (p2p mode only)

    # ------------------------------------------------------------------ layout
    def finalize(self, local_rows: Optional[int] = None) -> None:
        """Exchange shard sizes and fix this rank's global row offset (prefix sum)."""
        if local_rows is None:
            local_rows = self.local.num_rows
        if self.world == 1:
            self.row_offset, self.total_rows = 0, int(local_rows)
            return
        dev = self._comm_device()
        mine = torch.tensor([int(local_rows)], dtype=torch.int64, device=dev)
        counts = [torch.zeros_like(mine) for _ in range(self.world)]
        dist.all_gather(counts, mine, group=self.group)
        counts = [int(c.item()) for c in counts]
        self.row_offset = sum(counts[: self.rank])
        self.total_rows = sum(counts)
        if self.total_rows >= 0xFFFFFFFF:
            raise ValueError("the merge kernels carry global rows as 32-bit keys: at most 2^32 - 2 rows in total")

    def _comm_device(self) -> torch.device:
        if self.local is not None and hasattr(self.local, "device"):
            return self.local.device
        return torch.device("cpu")

    # ---------------------------------------------------------------- exchange
    def _setup_p2p(self, capacity: int) -> bool:
        """Allocate + rendezvous the peer-mapped gather buffers (collective).  Returns whether
        EVERY rank succeeded; otherwise all ranks stay on the NCCL path."""
        from . import ops
        dev = self._comm_device()
        ok = 1
        state = None
        try:
            import torch.distributed._symmetric_memory as symm_mem
            nbytes = ops.exchange_buffer_bytes(self.world, capacity)
            buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=dev)
            buf.zero_()
            hdl = symm_mem.rendezvous(buf, self.group if self.group is not None else dist.group.WORLD)
            state = (buf, hdl, [int(p) for p in hdl.buffer_ptrs], int(capacity))
        except Exception as e:                      # no fabric / IPC support on this box
            ok = 0
            if self._exchange_req == "p2p":
                print(f"[ShardedCorpusIndex] symmetric memory unavailable on rank {self.rank}: {e}")
        flag = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        torch.cuda.synchronize(dev)                 # everyone's zeroing has finished ...
        dist.barrier(group=self.group)              # ... before anyone pushes
        if int(flag.item()) == 1:
            self._xchg = state
            self._epoch = 0
            return True
        self._xchg = None
        return False

    def _choose_exchange(self, dev: torch.device, need: int) -> None:
        if self._exchange_req == "nccl" or dev.type != "cuda" or self._merge is not None:
            self.exchange = "nccl"
            return
        if self._p2p_failed:                        # decided collectively once; no per-search retries
            self.exchange = "nccl"
            return
        if self._xchg is not None and self._xchg[3] >= need:
            return
        if self._xchg is not None:                  # grow: nobody may still be reading the old one
            torch.cuda.synchronize(dev)
            dist.barrier(group=self.group)
        cap = max(need, 1024 * 16)
        self.exchange = "p2p" if self._setup_p2p(cap) else "nccl"
        if self.exchange == "nccl":
            self._p2p_failed = True                 # the same on every rank (all-reduced flag)
            if self._exchange_req == "p2p":
                raise RuntimeError("exchange='p2p' requested but symmetric memory could not be set up")

    def reset_exchange(self) -> None:
        """Collective: forget a failed symmetric-memory set-up so the next search tries again."""
        self._p2p_failed = False

    def _all_ranks_can_fuse(self, b: int) -> bool:
        """Every rank must take the same path (the fused form and the exchange kernel share buffers
        and epochs but not flags): the decision depends on the batch size only."""
        return self.local is not None and hasattr(self.local, "can_fuse_exchange") and \
            self.local.can_fuse_exchange(b)

    # ------------------------------------------------------------------ search
    def search_device(self, q_dev: torch.Tensor, k: int, out=None, queries_ready: bool = False
                      ) -> Tuple[torch.Tensor, torch.Tensor]:
        """q_dev [B,1024] fp32 (replicated on every rank) -> merged (scores, global rows) on
        every rank: local scan, then ONE exchange + merge."""
        if self._local_topk is not None:
            s, i = self._local_topk(q_dev, k, self.row_offset)
        elif self.world > 1 and q_dev.is_cuda and q_dev.dtype == torch.float32 and \
                self.fuse_small_batches and self._all_ranks_can_fuse(q_dev.shape[0]):
            # one or two queries: the exchange rides in the scan's last CTA -- ONE launch per rank
            # (sqe_search_gemv_sharded / sqe_search_gemv_prefiltered), no exchange kernel
            self._choose_exchange(q_dev.device, q_dev.shape[0] * k)
            if self.exchange == "p2p":
                self._epoch += 1
                return self.local.search_device(q_dev, k, idx_offset=self.row_offset, out=out,
                                                xchg=(self.rank, self._xchg[2], self._xchg[3], self._epoch),
                                                queries_ready=queries_ready)
            s, i = self.local.search_device(q_dev, k, idx_offset=self.row_offset)
        else:
            s, i = self.local.search_device(q_dev, k, idx_offset=self.row_offset,
                                            out=out if self.world == 1 else None,
                                            **({"queries_ready": True} if queries_ready else {}))
        if self.world == 1:
            return s, i
        return self.exchange_lists(s, i, k, out=out)

    def exchange_lists(self, s: torch.Tensor, i: torch.Tensor, k: int, out=None
                       ) -> Tuple[torch.Tensor, torch.Tensor]:
        """Collective: every rank passes its local [B,k] lists (global rows) and gets the merged
        top-k of all ranks."""
        b = s.shape[0]
        self._choose_exchange(s.device, b * k)
        if self.exchange == "p2p":
            from . import ops
            self._epoch += 1
            return ops.exchange_merge(s.contiguous(), i.contiguous(), k, self.rank, self._xchg[2],
                                      self._xchg[3], self._epoch, out=out)
        if self._gather_s is None or self._gather_s.shape[1:] != s.shape or self._gather_s.device != s.device:
            self._gather_s = torch.empty((self.world, b, k), dtype=s.dtype, device=s.device)
            self._gather_i = torch.empty((self.world, b, k), dtype=i.dtype, device=i.device)
        # concatenated [G*B, k] form of the output: accepted by both NCCL and gloo
        dist.all_gather_into_tensor(self._gather_s.view(self.world * b, k), s.contiguous(), group=self.group)
        dist.all_gather_into_tensor(self._gather_i.view(self.world * b, k), i.contiguous(), group=self.group)
        if self._merge is not None:
            return self._merge(self._gather_s, self._gather_i, k)
        from . import ops
        return ops.merge_topk(self._gather_s, self._gather_i, k, out=out)

    def search_batch(self, query_emb: np.ndarray, k: int = 3) -> Tuple[np.ndarray, np.ndarray]:
        """Host fp32 [B,1024] in (the same batch on every rank), host (scores, global rows) out."""
        dev = self._comm_device()
        if dev.type != "cuda" or self.local is None:
            q = torch.from_numpy(np.ascontiguousarray(query_emb, dtype=np.float32))
            s, i = self.search_device(q, k)
            return s.cpu().numpy(), i.cpu().numpy()
        from . import ops
        q = self.local._as_rows(query_emb)
        with self.local._search_lock, torch.cuda.device(dev):
            qd = self.local._stage_queries(q)                  # reusable pinned staging
            buf, s, i = ops.packed_topk_out(dev, q.shape[0], k)
            self.search_device(qd, k, out=(s, i))
            return self.local._fetch_packed(buf, q.shape[0], k)  # one D2H copy

    def search_batches(self, batches, k: int = 3, depth: int = 2):
        """Streaming form of `search_batch` (every rank passes the same sequence of batches):
        yields one `(scores, global rows)` per batch, in order, while the copies of the
        neighbouring batches overlap the scan + exchange of the current one
        (GpuCorpusIndex.search_batches)."""
        dev = self._comm_device()
        if dev.type != "cuda" or self.local is None:
            for qb in batches:
                yield self.search_batch(qb, k)
            return
        yield from self.local.search_batches(
            batches, k, depth, device_fn=lambda qd, kk, out: self.search_device(qd, kk, out=out))

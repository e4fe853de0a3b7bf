"""CPU-only checks: the C ABI library builds, loads and exports every symbol the header
declares; the product path refuses to run without a GPU (no fallback); host-side logic."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def sqe():
    import __graft_entry__
    __graft_entry__.build()
    import sqe_b200
    return sqe_b200


def test_library_exports_every_header_symbol(sqe):
    with open(os.path.join(ROOT, "include", "sqe_b200.h")) as f:
        header = f.read()
    declared = set(re.findall(r"SQE_API\s+[\w\s\*]+?\b(sqe_\w+)\s*\(", header))
    assert len(declared) >= 11
    lib = sqe._native.load()
    bound = {name for name, _, _ in sqe._native.PROTOTYPES}
    assert declared == bound, declared ^ bound
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.sqe_abi_version() == 2
    out = subprocess.run(["nm", "-D", "--defined-only", sqe._native.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (sqe_\w+)", out))
    assert declared <= exported


def test_library_contains_sm100a_code(sqe):
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.isfile(cuobjdump):
        pytest.skip("cuobjdump not installed")
    out = subprocess.run([cuobjdump, "-lelf", sqe._native.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_header_is_plain_c_and_library_is_usable_from_c(sqe, tmp_path):
    """The boundary is a C ABI: the header compiles as C99 with -pedantic and a C program can
    link the library, call it and get error codes (no exceptions, no torch)."""
    exe = str(tmp_path / "abi_smoke")
    src = os.path.join(ROOT, "tests", "c", "abi_smoke.c")
    libdir = os.path.dirname(sqe._native.LIB_PATH)
    cc = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I",
                         os.path.join(ROOT, "include"), src, "-o", exe, "-L", libdir, "-lsqe_b200",
                         "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr
    run = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0, (run.returncode, run.stdout, run.stderr)
    assert "abi 2 ok" in run.stdout


def test_no_cpu_fallback(sqe):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    x = torch.zeros((4, 1024))
    with pytest.raises(RuntimeError):
        sqe.ops.normalize_cast(x, "bf16")               # CPU tensors are refused
    rc, *_ = sqe._native.device_info()
    assert rc < 0                                        # no device -> error, not a silent path
    with pytest.raises(sqe.SqeError):
        sqe._native.call("sqe_normalize_cast", 16, 16, 4, 1024, 1, None)
    assert "CUDA" in sqe._native.last_error() or "device" in sqe._native.last_error()


def test_argument_errors_cross_the_abi_as_codes(sqe):
    lib = sqe._native.load()
    assert lib.sqe_normalize_cast(16, 16, 4, 512, 1, None) == -1          # dim != 1024
    assert b"dim" in lib.sqe_last_error()
    assert lib.sqe_normalize_cast(16, 16, 4, 1024, 7, None) == -1         # dtype
    assert lib.sqe_topk_gemv(16, 1, 10, 1024, 16, 1, 0, 16, 16, 0, 16, 1 << 20, None) == -1   # k = 0
    assert lib.sqe_topk_gemv(16, 1, 10, 1024, 16, 1, 257, 16, 16, 0, 16, 1 << 20, None) == -1
    assert lib.sqe_topk_gemv(8, 1, 10, 1024, 16, 1, 5, 16, 16, 0, 16, 1 << 20, None) == -1    # alignment
    assert lib.sqe_topk_batched(16, 0, 10, 1024, 16, 4, 5, 16, 16, 0, 16, 1 << 20, None) == -4  # fp32 shard
    assert lib.sqe_merge_topk(16, 16, 0, 1, 1, 1, 16, 16, None) == -1
    assert lib.sqe_search_gemv(16, 1, 10, 1024, 16, 1, 0, 16, 16, 0, 16, 1 << 20, None) == -1  # k = 0
    import ctypes
    peers = (ctypes.c_void_p * 2)(16, 32)
    assert lib.sqe_exchange_merge(16, 16, 4, 10, 10, 2, 2, peers, 100, 1, 3, 16, 16, None) == -1   # rank >= world
    assert lib.sqe_exchange_merge(16, 16, 4, 10, 10, 0, 2, peers, 39, 1, 3, 16, 16, None) == -1    # capacity < b*k
    assert lib.sqe_exchange_merge(16, 16, 4, 10, 10, 0, 17, peers, 100, 1, 3, 16, 16, None) == -1  # world > 16
    assert lib.sqe_exchange_buffer_bytes(8, 1024 * 10) == 256 + 2 * 8 * 10240 * 16
    assert lib.sqe_tuning_set(99, 0) == -1
    old = lib.sqe_tuning_set(0, 2)
    assert lib.sqe_tuning_set(0, old) == 2
    assert lib.sqe_topk_gemv_workspace_bytes(1, 10) > 0
    assert lib.sqe_topk_batched_workspace_bytes(1000, 8, 10) > 0
    assert lib.sqe_cache_top1_workspace_bytes(1000, 8) > 0


def test_shard_bounds_partition(sqe):
    for n, w in [(10, 3), (40, 8), (7, 8), (10_000_000, 8), (0, 2)]:
        spans = [sqe.shard_bounds(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "semantic-query-engine_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, fn)) as f:
                    src = f.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), fn


WORKER = r'''
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
import oracle
from sqe_b200.sharded import ShardedCorpusIndex, shard_bounds
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
rng = np.random.default_rng(7)
n, b, k = 1003, 5, 12
d = oracle.normalize_rows(rng.standard_normal((n, 1024)).astype(np.float32))
d[900] = d[3]                                   # tie across shards: lower global row first
q = rng.standard_normal((b, 1024)).astype(np.float32); q[0] = d[3]
lo, hi = shard_bounds(n, world, rank)
def local_topk(qt, k, off):                      # injected checker (CPU): the oracle on this shard
    s, i = oracle.topk_cosine(d[lo:hi], oracle.normalize_rows(qt.numpy()), k)
    return torch.from_numpy(s), torch.from_numpy(np.where(i >= 0, i + off, -1))
def merge(gs, gi, k):
    s, i = oracle.merge_topk(gs.numpy(), gi.numpy(), k)
    return torch.from_numpy(s), torch.from_numpy(i)
idx = ShardedCorpusIndex(None, local_topk=local_topk, merge=merge)
idx.finalize(hi - lo)
assert idx.row_offset == lo and idx.total_rows == n, (idx.row_offset, lo)
s, i = idx.search_batch(q, k)
ws, wi = oracle.topk_cosine(d, oracle.normalize_rows(q), k)
assert np.array_equal(i, wi), (rank, i, wi)
assert np.allclose(s, ws, atol=1e-6)
assert list(i[0][:2]) == [3, 900]
outs = list(idx.search_batches([q, q[:2], q[1:]], k))          # streaming form: same lists, in order
assert len(outs) == 3 and np.array_equal(outs[0][1], wi) and np.array_equal(outs[1][1], wi[:2])
assert np.array_equal(outs[2][1], wi[1:])
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_sharded_search_two_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="2")
    proc = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
         "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
        capture_output=True, text=True, env=env, timeout=240)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-2000:]
    assert proc.stdout.count("ok") == 2


def test_bench_reference_arm_runs_on_cpu():
    # torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arm must still use every core
    env = dict(os.environ, OMP_NUM_THREADS="1")
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                           "--steps", "1", "--warmup", "0", "--workload", "b1"],
                          capture_output=True, text=True, env=env, timeout=300)
    assert proc.returncode == 0, proc.stderr[-2000:]
    import json
    line = json.loads(proc.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))


def test_bench_reference_arm_under_torchrun_prints_one_line_and_uses_every_core():
    """The driver launches the reference arm like our own arm: under torchrun for N > 1.  Rank 0
    alone works and prints; torchrun's OMP_NUM_THREADS=1 must not make it single-threaded."""
    import json
    env = {k: v for k, v in os.environ.items() if k != "OMP_NUM_THREADS"}
    proc = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
         "--master-addr", "127.0.0.1", "--master-port", "29613", os.path.join(ROOT, "bench.py"),
         "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--workload", "b1"],
        capture_output=True, text=True, env=env, timeout=400)
    assert proc.returncode == 0, proc.stdout[-1500:] + proc.stderr[-1500:]
    lines = [ln for ln in proc.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, lines
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["value"] > 0
    assert line["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert "rows split over 2 rank(s)" in line["config"]["sharding"]

"""One encoder block at the benchmark shape (64 chunks x 512 tokens), kernel by kernel -- the
target of the ncu captures under profiles/ -- and, with --timers, the phase time stamps of the
attention kernel (sqe_debug_encoder_attention_timers)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import sqe_b200
from sqe_b200 import encoder as enc

nat = sqe_b200._native
dev = torch.device("cuda", 0)
n_seq, seq_len, H = 64, int(os.environ.get("ENC_SEQ", "512")), 1024
w = sqe_b200.EncoderWeights.random_init(seed=0, layers=1, device=dev)
e = sqe_b200.GpuEmbeddingEncoder(w)
t_pad, pos, first, tiles = e.plan([seq_len] * n_seq)
buf = e._buffers(t_pad)
rng = np.random.default_rng(0)
ids_np = np.where(pos >= 0, rng.integers(0, 30522, size=t_pad), -1).astype(np.int32)
ids_d = torch.from_numpy(ids_np).to(dev)
pos_d = torch.from_numpy(np.maximum(pos, 0)).to(dev)
tiles_d = torch.from_numpy(tiles).to(dev)
L = w.layers[0]


def block():
    enc.gemm(buf.h16, L["wqkv"], L["bqkv"], nat.SQE_ENC_EPI_SPLIT, buf.qk, m=t_pad, out1=buf.vt, n_split=2 * H,
             q_cols=H, q_scale=0.125)
    enc.attention(buf.qk, buf.vt, tiles_d, tiles.shape[0], seq_len, buf.ctx)
    enc.gemm(buf.ctx, L["wo"], L["bo"], nat.SQE_ENC_EPI_RES_F32, buf.sum_a, m=t_pad, residual=buf.sum_b,
                 res_stats=buf.stats_b, res_gamma=L["g2"], res_beta=L["b2"])
    enc.layernorm(buf.sum_a, L["g1"], L["b1"], 1e-12, None, buf.h16, rows=t_pad, stats=buf.stats_a)
    enc.gemm(buf.h16, L["w1"], L["bi"], nat.SQE_ENC_EPI_GELU, buf.ffn, m=t_pad)
    enc.gemm(buf.ffn, L["w2"], L["bo2"], nat.SQE_ENC_EPI_RES_F32, buf.sum_b, m=t_pad, residual=buf.sum_a,
                 res_stats=buf.stats_a, res_gamma=L["g1"], res_beta=L["b1"])


first_d = torch.from_numpy(first).to(dev)
out = torch.empty((n_seq, H), dtype=torch.float32, device=dev)
e._forward(buf, ids_d, pos_d, tiles_d, tiles.shape[0], seq_len, first_d, n_seq, out)   # real activations in the buffers
torch.cuda.synchronize()
if "--timers" in sys.argv:
    n_cta = tiles.shape[0] * 16
    dbg = torch.zeros((n_cta, 8), dtype=torch.int64, device=dev)
    nat.load().sqe_debug_encoder_attention_timers(dbg.data_ptr())
    enc.attention(buf.qk, buf.vt, tiles_d, tiles.shape[0], seq_len, buf.ctx)
    torch.cuda.synchronize()
    nat.load().sqe_debug_encoder_attention_timers(None)
    d = dbg.cpu().numpy().astype(np.float64)
    nq = d[:, 6]
    names = ["CTA total", "wait for scores (loads, QK^T)", "row maxima", "exp + P stores (+ PV overlap)",
             "wait for PV", "output"]
    for i, nm in enumerate(names):
        dt = d[:, i] / nq
        print(f"{nm:32s} per q-tile: mean {dt.mean():9.0f}  p10 {np.percentile(dt, 10):9.0f}  p90 {np.percentile(dt, 90):9.0f} cycles")
    span_us = (d[:, 7].max() - d[:, 7].min()) / 1e3
    print(f"{n_cta} CTAs x {nq.mean():.1f} q-tiles; ends span {span_us:.1f} us")
else:
    block()
    torch.cuda.synchronize()
    print("one block done")

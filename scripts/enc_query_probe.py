"""One 16-token query through the 24-layer encoder without CUDA graphs: the target of an ncu launch
list (per-kernel durations of the latency path)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import sqe_b200

dev = torch.device("cuda", 0)
w = sqe_b200.EncoderWeights.random_init(seed=0, layers=24, device=dev)
e = sqe_b200.GpuEmbeddingEncoder(w, use_graphs=False)
rng = np.random.default_rng(0)
n_tok = int(os.environ.get("ENC_QTOK", "16"))
for _ in range(3):
    q = [rng.integers(0, 30522, size=n_tok).tolist()]
    out = e.embed_token_ids(q)
    e.stream.synchronize()
print("ok", float(out.abs().sum()))

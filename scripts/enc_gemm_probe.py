"""A/B of the encoder GEMM forms (tuning knob SQE_TUNE_ENC_GEMM_FORM) on the four products of a
block at the benchmark shape: correctness against fp32 torch and time per launch.
    python scripts/enc_gemm_probe.py [forms...]      e.g.  2 3"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import sqe_b200
from sqe_b200 import encoder as enc

nat = sqe_b200._native
dev = torch.device("cuda", 0)
forms = [int(x) for x in sys.argv[1:]] or [2, 3]
M, H = int(os.environ.get("ENC_M", "32768")), 1024
g = torch.Generator(device=dev).manual_seed(0)
rnd = lambda *s, sc=1.0: torch.randn(*s, generator=g, device=dev) * sc          # noqa: E731
x = rnd(M, H).half()
x4 = rnd(M, 4 * H).half()
shapes = {"qkv": (3 * H, H, nat.SQE_ENC_EPI_SPLIT), "attn_out": (H, H, nat.SQE_ENC_EPI_RES_F32),
          "ffn1_gelu": (4 * H, H, nat.SQE_ENC_EPI_GELU), "ffn2": (H, 4 * H, nat.SQE_ENC_EPI_RES_F32)}
res = rnd(M, H)
for name, (n, k, epi) in shapes.items():
    w = rnd(n, k, sc=0.03).half()
    bias = rnd(n, sc=0.5)
    xa = x4 if k == 4 * H else x
    ref = None
    for form in forms:
        nat.tuning_set(nat.SQE_TUNE_ENC_GEMM_FORM, form)
        if epi == nat.SQE_ENC_EPI_SPLIT:
            out0 = torch.zeros((M, 2 * H), dtype=torch.float16, device=dev)
            out1 = torch.zeros((H, M), dtype=torch.float16, device=dev)
            run = lambda: enc.gemm(xa, w, bias, epi, out0, out1=out1, n_split=2 * H, q_cols=H, q_scale=0.125)   # noqa: E731
        elif epi == nat.SQE_ENC_EPI_RES_F32:
            out0 = torch.zeros((M, n), dtype=torch.float32, device=dev)
            out1 = None
            run = lambda: enc.gemm(xa, w, bias, epi, out0, residual=res)          # noqa: E731
        else:
            out0 = torch.zeros((M, n), dtype=torch.float16, device=dev)
            out1 = None
            run = lambda: enc.gemm(xa, w, bias, epi, out0)                        # noqa: E731
        try:
            run()
            torch.cuda.synchronize()
        except Exception as e:                                                    # noqa: BLE001
            print(name, "form", form, "FAILED", str(e).splitlines()[0])
            sys.exit(1)
        cur = (out0.clone(), None if out1 is None else out1.clone())
        if ref is None:
            ref = cur
            same = "reference form"
        else:
            same = "identical" if torch.equal(cur[0], ref[0]) and (cur[1] is None or torch.equal(cur[1], ref[1])) else \
                f"DIFFERS max {float((cur[0].float() - ref[0].float()).abs().max()):.3e}"
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"{name:10s} form {form}: {ms * 1e3:8.1f} us  {2.0 * M * n * k / ms / 1e9:8.1f} TFLOP/s  {same}")
        if os.environ.get("ENC_TIMERS"):
            dbg = torch.zeros((148, 8), dtype=torch.int64, device=dev)
            nat.load().sqe_debug_encoder_gemm_timers(dbg.data_ptr())
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            nat.load().sqe_debug_encoder_gemm_timers(None)
            d = dbg.cpu().numpy().astype(float)
            lead = d[d[:, 5] > 0]
            tiles = lead[:, 5].mean()
            print(f"           issuer: {lead[:, 2].mean() / tiles:8.0f} cycles per tile ({k // 64 * 512} at the issue floor), "
                  f"waiting for data {lead[:, 3].mean() / tiles:7.0f}, for the epilogue {lead[:, 4].mean() / tiles:7.0f}; "
                  f"producer waiting for free stages {d[:, 1].mean() / tiles:7.0f}")
nat.tuning_set(nat.SQE_TUNE_ENC_GEMM_FORM, 0)

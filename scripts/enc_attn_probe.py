"""Run the attention kernel on one packed batch (lengths from argv) and print the worst error
against fp32 torch -- one process per case, so a faulting case does not poison the next."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import sqe_b200
import test_gpu_encoder as t

lens = [int(x) for x in sys.argv[1:]]
try:
    print(lens, "max err", t._attention_case(sqe_b200, lens, seed=1))
except Exception as e:                                   # noqa: BLE001
    print(lens, "FAILED", str(e).splitlines()[0])

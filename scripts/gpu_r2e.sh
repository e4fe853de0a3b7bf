#!/bin/bash
# Round-2 third session, run 1: the new tests (GGUF / Ollama store loading, the reference's handlers with the
# GPU encoder from text), K2p regression on the rebuilt library, ncu --set full of the int8 scan in its final
# (resident query tile) form at 10M rows, diagnostics modes.  Outputs -> gpurun_out/r2e/
set -u
O=gpurun_out/r2e
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_encoder.py tests/test_reference_handlers_gpu.py -m gpu -q -s -k "ollama or reference" --timeout 500 2>&1 | tail -12 > $O/new_tests.log; tail -4 $O/new_tests.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "batched_prefiltered or index_with_prefilter" --timeout 500 2>&1 | tail -3 > $O/k2p_tests.log; tail -1 $O/k2p_tests.log
timeout 300 python scripts/k2p_probe.py 10000000 1024 10 bf16 > $O/k2p_probe_plain.log 2>&1; tail -1 $O/k2p_probe_plain.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:topk_batched_i8 -s 4 -c 1 -f -o $O/k2p_i8_10m python scripts/k2p_probe.py 10000000 1024 10 bf16 > $O/ncu_i8.log 2>&1; tail -1 $O/ncu_i8.log
( timeout 300 python scripts/k2p_modes.py 10000000 1024 10 ) 2>&1 | grep "rows=" | tee $O/k2p_modes.txt
ls -la $O

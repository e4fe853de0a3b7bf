# K2p operand-ring depth experiment: the shipped library (4 x 16 KB stages beside the resident query tile) against
# experimental builds with 3 and 5 stages (lib/exp/, built with -DSQE_I8_ASTAT_STAGES=...; the 5-stage build has room
# for ONE threshold-warp slot only, so only its diagnostics modes 1 / 2 -- no candidate logging -- are meaningful).
set -u
O=gpurun_out/r2e
mkdir -p $O
( for rep in 1 2 3; do for lib in lib/libsqe_b200.so lib/exp/libsqe_b200_s5.so; do
    K2P_MODES=1,2,1,2 SQE_LIB=semantic-query-engine_b200/$lib timeout 300 python scripts/k2p_modes.py 10000000 1024 10 2>&1 | grep "rows=\|rror"
  done; done ) | tee $O/k2p_ring_depth2.txt

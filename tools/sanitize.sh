#!/usr/bin/env bash
# Race / memory checking of the single-GPU kernels (run on a B200 through gpurun).
# The reference has no sanitizer usage at all; cross-GPU flag protocols cannot be
# checked by racecheck, so the multi-GPU kernels are covered by tests/gpu/comm_check.py instead.
set -euo pipefail
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
for TOOL in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $TOOL --log-file "$OUT/sanitizer_$TOOL.log" \
    python -m pytest tests/test_gpu_engine.py -m gpu -q -k "nn_kernels or fused_step_local or tcgen05" || true
  tail -5 "$OUT/sanitizer_$TOOL.log"
done
# tcgen05 / TMA conv kernels, first-layer tensor-core kernels, fused un-pool producers and the fused head
timeout 300 compute-sanitizer --tool memcheck --log-file "$OUT/sanitizer_memcheck_conv.log" \
  python tests/gpu/conv_check.py 2 || true
tail -3 "$OUT/sanitizer_memcheck_conv.log"

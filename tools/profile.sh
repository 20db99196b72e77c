#!/usr/bin/env bash
# Launch list + full ncu capture of our kernels for the MNIST-CNN step (1 GPU, via gpurun):
#   gpurun --timeout 600 -- tools/profile.sh
set -euo pipefail
OUT=gpurun_out
mkdir -p $OUT
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv \
    --log-file $OUT/launches.csv python tests/gpu/profile_step.py
for K in tfy_gemm_bf16_kernel tfy_fused_step_kernel tfy_conv3x3_c1_wgrad_kernel tfy_pool_drop_relu_bwd_kernel; do
  PROFILE_STEPS=1 ncu --set full --clock-control none --import-source on --profile-from-start off \
      -k regex:$K -c 1 -f -o $OUT/prof_$K python tests/gpu/profile_step.py
done
# read back on the CPU box with:  python tools/ncu_summary.py gpurun_out/prof_*.ncu-rep

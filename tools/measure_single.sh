#!/usr/bin/env bash
# Single-GPU measurement + evidence pass (one gpurun call):  gpurun --timeout 1700 -- tools/measure_single.sh TAG
set -uo pipefail
TAG=${1:-ms}; OUT=gpurun_out; mkdir -p $OUT
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6) > $OUT/${TAG}_pytest.log; cat $OUT/${TAG}_pytest.log
arm() { local name=$1 to=$2; shift 2; timeout "$to" "$@" > "$OUT/${TAG}_${name}.raw" 2> "$OUT/${TAG}_${name}.err"; local rc=$?
  grep "^{" "$OUT/${TAG}_${name}.raw" | tail -1 > "$OUT/${TAG}_${name}.json"
  python -c "
import json,sys
try:
    d=json.load(open('$OUT/${TAG}_${name}.json')); print('[$name] rc=$rc', round(d['value'],1), d['unit'], 'ms/step', round(d['ms_per_step'],5), 'e2e', round(d.get('e2e',{}).get('value',0),1))
except Exception as e: print('[$name] rc=$rc no result', e)"; }
arm mnist_reference 300 python bench.py --impl reference --steps 20 --warmup 5
arm mnist_ours 300 python bench.py --steps 20 --warmup 5
arm mnist_ours_2000 300 python bench.py --steps 2000 --warmup 20
arm mnist_standin_graph 300 python bench.py --impl standin --graph --steps 20 --warmup 5
arm mnist_standin 300 python bench.py --impl standin --steps 20 --warmup 5
arm resnet50_ours 300 python bench.py --config resnet50 --steps 10 --warmup 3 --repeats 3
arm resnet50_reference 300 python bench.py --config resnet50 --impl reference --steps 10 --warmup 3 --repeats 3
arm bert_ours 400 python bench.py --config bert --steps 10 --warmup 3 --repeats 3 --side-tasks
arm bert_standin 400 python bench.py --config bert --impl standin --steps 10 --warmup 3 --repeats 3
TFY_BENCH_TIMEOUT=200 arm wide_deep_ours 300 python bench.py --config wide_deep --gpus 1 --steps 200 --warmup 5
arm wide_deep_standin 300 python bench.py --config wide_deep --impl standin --gpus 1 --steps 100 --warmup 5
timeout 300 python tests/gpu/gemm_bench.py --out $OUT/${TAG}_gemm_bench.json > $OUT/${TAG}_gemm_bench.log 2>&1; tail -5 $OUT/${TAG}_gemm_bench.log | cut -c1-300
# launch list + full ncu capture of the step's kernels (never a bench number)
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv \
    --log-file $OUT/${TAG}_launches.csv python tests/gpu/profile_step.py > /dev/null 2>&1
PROFILE_STEPS=1 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"tfy_dense_bwd|tfy_gemm_bf16|tfy_fused_step|tfy_conv3x3_wgrad|tfy_dense_head" -f -o $OUT/${TAG}_prof_step python tests/gpu/profile_step.py > /dev/null 2>&1
ls -la $OUT/${TAG}_prof_step.ncu-rep 2>/dev/null
echo "measure_single done"

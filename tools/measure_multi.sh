#!/usr/bin/env bash
# Multi-GPU measurement pass (one gpurun call):  gpurun --gpus N --timeout 1500 -- tools/measure_multi.sh N TAG [full]
# Writes gpurun_out/<TAG>_*.json (one JSON line per bench arm) and logs.  `full` adds the secondary configs,
# the comm sweep and the multi-GPU correctness checks.
set -uo pipefail
N=${1:-2}; TAG=${2:-mm}; FULL=${3:-}
OUT=gpurun_out; mkdir -p $OUT
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29700
arm() {  # name timeout cmd...
  local name=$1 to=$2; shift 2
  port=$((port+1))
  timeout "$to" "$@" > "$OUT/${TAG}_${name}.raw" 2> "$OUT/${TAG}_${name}.err"
  local rc=$?
  grep "^{" "$OUT/${TAG}_${name}.raw" | tail -1 > "$OUT/${TAG}_${name}.json"
  python - "$OUT/${TAG}_${name}.json" "$name" "$rc" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    e = d.get("e2e", {})
    print(f"[{sys.argv[2]}] rc={sys.argv[3]} value={d.get('value'):.1f} {d.get('unit')} ms/step={d.get('ms_per_step'):.4f} "
          f"e2e={e.get('value', 0):.1f} exposed_ms={d.get('exposed_comm_ms')} sync={d.get('params_in_sync')} "
          f"clk={d.get('clocks', {}) and d['clocks'].get('sm_mhz')} {d.get('clocks', {}) and d['clocks'].get('reasons')}")
except Exception as exc:
    print(f"[{sys.argv[2]}] rc={sys.argv[3]} no result ({exc})")
PY
}
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > $OUT/${TAG}_clocks.csv &
SMI=$!
# ---- headline: all three arms, the driver's own step count and a long one
arm mnist_reference 300 $R --master-port $port bench.py --gpus $N --impl reference --steps 20 --warmup 5
arm mnist_ours      300 $R --master-port $port bench.py --gpus $N --steps 20 --warmup 5
arm mnist_ours_2000 300 $R --master-port $port bench.py --gpus $N --steps 2000 --warmup 20
arm mnist_standin_graph 300 $R --master-port $port bench.py --gpus $N --impl standin --graph --steps 20 --warmup 5
if [ -n "$FULL" ]; then
  arm mnist_standin 300 $R --master-port $port bench.py --gpus $N --impl standin --steps 20 --warmup 5
  arm resnet50_ours      400 $R --master-port $port bench.py --config resnet50 --gpus $N --steps 10 --warmup 3 --repeats 3
  arm resnet50_reference 400 $R --master-port $port bench.py --config resnet50 --impl reference --gpus $N --steps 10 --warmup 3 --repeats 3
  arm bert_ours    500 $R --master-port $port bench.py --config bert --gpus $N --steps 10 --warmup 3 --repeats 3 --side-tasks
  arm bert_standin 500 $R --master-port $port bench.py --config bert --impl standin --gpus $N --steps 10 --warmup 3 --repeats 3
  TFY_BENCH_TIMEOUT=240 arm wide_deep_ours 300 python bench.py --config wide_deep --gpus $N --steps 200 --warmup 5
  arm wide_deep_standin 300 python bench.py --config wide_deep --impl standin --gpus $N --steps 200 --warmup 5
  timeout 400 $R --master-port $((port+20)) tests/gpu/comm_sweep.py --max-mb 512 --out $OUT/${TAG}_comm_sweep_N$N.json > $OUT/${TAG}_comm_sweep.log 2>&1
  grep -E "barrier|graph\]|fused_adam" $OUT/${TAG}_comm_sweep.log | cut -c1-600
  timeout 300 $R --master-port $((port+21)) tests/gpu/comm_check.py --quick > $OUT/${TAG}_comm_check.log 2>&1
  grep -E "FAIL|SUMMARY" $OUT/${TAG}_comm_check.log | tail -5
  cp $OUT/comm_check_N${N}_vmm.json $OUT/${TAG}_comm_check_N$N.json 2>/dev/null
  timeout 200 $R --master-port $((port+22)) tests/gpu/ddp_check.py > $OUT/${TAG}_ddp_check.log 2>&1
  grep -E "PASS|FAIL|DDP CHECK" $OUT/${TAG}_ddp_check.log | tail -4
fi
kill $SMI 2>/dev/null
echo "measure_multi done"

#!/usr/bin/env bash
# One gpurun call that re-establishes every judged number of a round on a fresh B200 (about 6-7 GPU-minutes):
#   /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash tools/gpu_round_check.sh r02'
# Order follows the profiling recipe: each command first exits 0 WITHOUT ncu, then runs under ncu; numbers printed
# under ncu are never bench values.  Everything lands in gpurun_out/<tag>/ ; copy what is to be judged into profiles/.
set -u
TAG="${1:-rXX}"
OUT="gpurun_out/${TAG}"
mkdir -p "${OUT}"
fail=0

step() {  # step <name> <timeout-seconds> <command...>
  local name="$1" limit="$2"; shift 2
  echo "== ${name}" | tee -a "${OUT}/steps.log"
  local t0=$SECONDS
  timeout "${limit}" "$@" > "${OUT}/${name}.log" 2> "${OUT}/${name}.err"
  local rc=$?
  echo "   rc=${rc} $((SECONDS - t0)) s" | tee -a "${OUT}/steps.log"
  [ "${rc}" -eq 0 ] || fail=1
  return "${rc}"
}

nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,clocks_throttle_reasons.active --format=csv > "${OUT}/gpu.csv" 2>&1

# 1. parity gate (through the C ABI) and the smoke entry
step pytest_gpu 600 python -m pytest tests -m gpu -x -q
step smoke 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"

# 2. the bench line (reference arm first, as the driver does), default K / W
step bench_reference 600 python bench.py --impl reference
step bench 900 python bench.py

# 3. ncu launch list of the same bench command (kernel SHARES of a step) - only after the plain run exited 0
if [ -s "${OUT}/bench.log" ]; then
  step ncu_launches 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file "${OUT}/launches.csv" python bench.py --steps 2 --warmup 3 --no-cpu-baseline
  # 4. one --set full capture of the dominant kernel (fused forward step) and one of the gradient tile kernel
  step ncu_step 900 ncu --set full --clock-control none --import-source on -k regex:gp_step_kernel -s 3 -c 1 \
    -o "${OUT}/prof_step" -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline
  step grad_plain 600 python tools/grad_time.py
  step ncu_grad 900 ncu --set full --clock-control none --import-source on -k regex:gp_step_kernel -s 8 -c 1 \
    -o "${OUT}/prof_grad" -f python tools/grad_time.py
fi

# 5. selector timing (the un-remeasured __restrict__ change of round 1 touches select_update_kernel)
step selector_time 600 python tools/selector_time.py

echo "overall fail=${fail}" | tee -a "${OUT}/steps.log"
exit "${fail}"

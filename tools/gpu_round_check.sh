#!/usr/bin/env bash
# One gpurun call that re-establishes every judged number of a round on a fresh B200 (about 11 GPU-minutes):
#   /usr/local/graft/bin/gpurun --timeout 1800 -- 'bash tools/gpu_round_check.sh r02'
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
step pytest_gpu 900 python -m pytest tests -m gpu -x -q
step smoke 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"

# 2. the bench line (reference arm first, as the driver does), default K / W
step bench_reference 600 python bench.py --impl reference
step bench 900 python bench.py

# 3. ncu launch list of a short bench command (kernel SHARES of a step; incl. the widened rows: no library GEMM may show
#    up there apart from the host framework's mean network) - only after the plain run exited 0
step bench_short 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --short
if [ -s "${OUT}/bench_short.log" ]; then
  step ncu_launches 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
    --log-file "${OUT}/launches.csv" python bench.py --steps 2 --warmup 3 --no-cpu-baseline --short
  # 4. one --set full capture of the dominant kernel (fused forward step)
  step ncu_step 900 ncu --set full --clock-control none --import-source on -k regex:gp_step_kernel -s 3 -c 1 \
    -o "${OUT}/prof_step" -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --short
fi
# 5. the round-2 kernels: factorisation, GEMM, Gram tile kernel, sharded selector
step profile_plain 600 python tools/profile_round2.py
if [ -s "${OUT}/profile_plain.log" ]; then
  step ncu_round2 900 ncu --set full --clock-control none \
    -k regex:"chol_inverse_coop_kernel|dgemm_kernel|select_update_sharded_kernel|select_finalize_sharded_kernel" -c 12 \
    -o "${OUT}/prof_round2" -f python tools/profile_round2.py
fi
# 5b. kernels of the last session: guarded tensor-pipe Gram, risk reduction at N = 1e8, quarter-tile GEMM, 8-point forward
#     tiles, and a fresh capture of the gradient-mode tile kernel + SYRK
step profile2b_plain 600 python tools/profile_round2b.py
if [ -s "${OUT}/profile2b_plain.log" ]; then
  step ncu_round2b 1200 ncu --set full --clock-control none \
    -k regex:"ard_gram_mma_kernel|gram_guard_kernel|risk_kernel|dgemm_kernel|gp_step_kernel|syrk_kernel" -c 26 \
    -o "${OUT}/prof_round2b" -f python tools/profile_round2b.py
fi
step selector_time 600 python tools/selector_time.py
step factor_time 600 python tools/factor_only.py
step gram_time 600 python tools/gram_time.py
step dgemm_time 600 python tools/dgemm_time.py
step risk_time 600 python tools/risk_time.py
step small_time 600 python tools/small_time.py
step widened 900 python tools/widened_launches.py

# gpurun copies gpurun_out/ back only below 64 MiB: keep the raw-metric CSV of every capture, drop the big reports
for rep in "${OUT}"/*.ncu-rep; do
  [ -f "${rep}" ] || continue
  ncu -i "${rep}" --page raw --csv > "${rep%.ncu-rep}_raw.csv" 2>/dev/null
  size=$(stat -c %s "${rep}")
  if [ "${size}" -gt 12000000 ]; then rm -f "${rep}"; fi
done
du -sh "${OUT}" | tee -a "${OUT}/steps.log"
echo "overall fail=${fail}" | tee -a "${OUT}/steps.log"
exit "${fail}"

#!/bin/bash
# One GPU-box session: gpu test-suite, smoke, bench (+ optional ncu launch list).  Everything lands in gpurun_out/.
#   PYTEST_ARGS  extra pytest arguments (e.g. "-k trained")      RUN_NCU=0  skip the ncu launch list
#   SKIP_TESTS=1 skip pytest                                      BENCH_ARGS bench.py arguments
set -u
mkdir -p gpurun_out
OUT=gpurun_out
: > $OUT/summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc >> $OUT/gpu.txt; lscpu | grep -i "model name" >> $OUT/gpu.txt
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout ${PYTEST_TIMEOUT:-1800} python -m pytest tests -m gpu -q -rA -p no:cacheprovider ${PYTEST_ARGS:-} > $OUT/pytest.log 2>&1
  echo "pytest rc=$?" | tee -a $OUT/summary.txt
  grep -E "passed|failed|error" $OUT/pytest.log | tail -3 | tee -a $OUT/summary.txt
  grep -E "^(FAILED|ERROR)" $OUT/pytest.log | tee -a $OUT/summary.txt
fi
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1
echo "smoke rc=$?" | tee -a $OUT/summary.txt; tail -2 $OUT/smoke.log | tee -a $OUT/summary.txt
timeout 900 python bench.py ${BENCH_ARGS:---steps 10 --warmup 3} > $OUT/bench.log 2> $OUT/bench.err
rc=$?; echo "bench rc=$rc" | tee -a $OUT/summary.txt; tail -1 $OUT/bench.log | cut -c1-1500 | tee -a $OUT/summary.txt; tail -5 $OUT/bench.err
if [ $rc -eq 0 ] && [ "${RUN_NCU:-1}" = "1" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv \
      --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 > $OUT/ncu_list.log 2>&1
  echo "ncu list rc=$?" | tee -a $OUT/summary.txt
fi

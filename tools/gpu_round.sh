#!/bin/bash
# One GPU-box session: isolated MLP variants first (a protocol bug must not take the rest down),
# then the gpu test-suite, smoke, bench and the ncu launch list.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
OUT=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc >> $OUT/gpu.txt; lscpu | grep -i "model name" >> $OUT/gpu.txt
for v in 4 3 2 1; do
  timeout 400 python tests/gpu_worker.py mlp $v 1000 0 > $OUT/worker_v$v.log 2>&1
  echo "worker variant $v rc=$?" | tee -a $OUT/summary.txt
  grep RESULT $OUT/worker_v$v.log | tee -a $OUT/summary.txt
done
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider ${PYTEST_EXTRA:-} > $OUT/pytest.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -25 $OUT/pytest.log | tee -a $OUT/summary.txt
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1
echo "smoke rc=$?" | tee -a $OUT/summary.txt; tail -3 $OUT/smoke.log | tee -a $OUT/summary.txt
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/bench.log 2> $OUT/bench.err
rc=$?; echo "bench rc=$rc" | tee -a $OUT/summary.txt; tail -2 $OUT/bench.log | tee -a $OUT/summary.txt; tail -5 $OUT/bench.err
if [ $rc -eq 0 ] && [ "${RUN_NCU:-1}" = "1" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv \
      --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 > $OUT/ncu_list.log 2>&1
  echo "ncu list rc=$?" | tee -a $OUT/summary.txt
fi

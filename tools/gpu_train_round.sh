#!/bin/bash
# Training-path session: gpu train tests, the config-4 bench, and the per-kernel list with warm caches.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_train.py -q -x -p no:cacheprovider > $OUT/pytest_train.log 2>&1
echo "pytest train rc=$?"; tail -5 $OUT/pytest_train.log
timeout 300 python tools/train_bench.py --steps 20 --warmup 5 > $OUT/train_plain.log 2>&1
echo "train rc=$?"; tail -1 $OUT/train_plain.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv \
    --log-file $OUT/train_launches_warm.csv python tools/train_bench.py --steps 3 --warmup 2 > $OUT/ncu_train.log 2>&1
echo "ncu train rc=$?"

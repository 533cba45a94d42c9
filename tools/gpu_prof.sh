#!/bin/bash
# Profiling session: plain bench first (must exit 0 without ncu), then the ncu launch list of the same
# command, one `--set full` capture of the fine mlp_fused_kernel launch, and the training-step kernel
# list with warm caches (ncu's default cache flush hides L2 residency of the weight images).
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench.log 2> $OUT/bench.err || { echo "bench failed"; tail -5 $OUT/bench.err; exit 1; }
tail -1 $OUT/bench.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 > $OUT/ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_fused_kernel -s 7 -c 1 \
    -o $OUT/prof_mlp_fold -f python bench.py --steps 2 --warmup 3 > $OUT/ncu_full.log 2>&1
echo "ncu full rc=$?"
timeout 300 python tools/train_bench.py --steps 20 --warmup 5 > $OUT/train_plain.log 2>&1
echo "train rc=$?"; tail -1 $OUT/train_plain.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv \
    --log-file $OUT/train_launches_warm.csv python tools/train_bench.py --steps 3 --warmup 2 > $OUT/ncu_train.log 2>&1
echo "ncu train rc=$?"

#!/bin/bash
# Training-path profiling only: plain run, warm-cache launch list, ncu --set full of one step's big kernels.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python tools/train_bench.py --steps 4 > $OUT/train_plain.log 2>&1
echo "train rc=$?"; tail -1 $OUT/train_plain.log | cut -c1-400
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 250 -c 150 --csv \
    --log-file $OUT/train_launches_warm.csv python tools/train_bench.py --steps 4 > $OUT/ncu_train.log 2>&1
echo "ncu train list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k "regex:mlp_fused_kernel|mlp_bwd_dx_kernel|mlp_bwd_dw_kernel" \
    -s 30 -c 6 -o $OUT/prof_train -f python tools/train_bench.py --steps 4 > $OUT/ncu_train_full.log 2>&1
echo "ncu train full rc=$?"

#!/bin/bash
# Profiling session (round 2): every command first runs plain (must exit 0 without ncu), then under ncu.
#   1. launch list of the bench command (gpu__time_duration per kernel)          -> launches.csv
#   2. launch list of the training step with warm caches                         -> train_launches_warm.csv
#   3. ncu --set full of the training kernels (fwd, dX, dW, heads; one step)     -> prof_train.ncu-rep
#   4. ncu --set full of the HBM-side kernels of one frame (K1, K2, K4, dirbias) -> prof_small.ncu-rep
#   5. (FULL_MLP=1) ncu --set full of the fine mlp_fused_kernel launch           -> prof_mlp.ncu-rep
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench.log 2> $OUT/bench.err || { echo "bench failed"; tail -5 $OUT/bench.err; exit 1; }
tail -1 $OUT/bench.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 > $OUT/ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 300 python tools/train_bench.py --steps 4 > $OUT/train_plain.log 2>&1
echo "train rc=$?"; tail -1 $OUT/train_plain.log | cut -c1-400
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 250 -c 150 --csv \
    --log-file $OUT/train_launches_warm.csv python tools/train_bench.py --steps 4 > $OUT/ncu_train.log 2>&1
echo "ncu train list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k "regex:mlp_fused_kernel|mlp_bwd_dx_kernel|mlp_bwd_dw_kernel|head_rgb_kernel|head_sigma_kernel|head_dir_kernel|adam_pack_kernel|pack_fold_pair_kernel" \
    -s 70 -c 16 -o $OUT/prof_train -f python tools/train_bench.py --steps 4 > $OUT/ncu_train_full.log 2>&1
echo "ncu train full rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on \
    -k "regex:composite_fwd|sample_pdf|dirbias_kernel|coarse_z|raygen_kernel" -s 12 -c 7 \
    -o $OUT/prof_small -f python bench.py --steps 2 --warmup 3 > $OUT/ncu_small.log 2>&1
echo "ncu small rc=$?"
if [ "${FULL_MLP:-0}" = "1" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_fused_kernel -s 7 -c 1 \
      -o $OUT/prof_mlp -f python bench.py --steps 2 --warmup 3 > $OUT/ncu_full.log 2>&1
  echo "ncu mlp full rc=$?"
fi

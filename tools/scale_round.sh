#!/bin/bash
# Scaling session on one N-GPU box: bench.py at N = 1, 2, 4, 8 (as many as the box has).  Each line carries the
# one-view-per-GPU throughput (value / e2e), the single-frame strong-scaling record and the training-step record.
set -u
OUT=gpurun_out; mkdir -p $OUT
G=$(nvidia-smi -L | wc -l)
for n in 1 2 4 8; do
  [ $n -le $G ] || continue
  if [ $n -eq 1 ]; then python bench.py --gpus 1 --steps 10 --warmup 3 > $OUT/bench_n$n.log 2> $OUT/bench_n$n.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
         bench.py --gpus $n --steps 10 --warmup 3 > $OUT/bench_n$n.log 2> $OUT/bench_n$n.err; fi
  echo "bench n=$n rc=$?"; grep '^{' $OUT/bench_n$n.log | tail -1 | cut -c1-160
done

#!/bin/bash
# Scaling session on one 8-GPU box: render bench at N = 1, 2, 4, 8 (one 640x480 view per GPU + NCCL all-gather of
# the uint8 tiles) and the data-parallel training step at N = 1 and 8.  JSON lines land in gpurun_out/.
set -u
OUT=gpurun_out; mkdir -p $OUT
run() {  # n script args...
  local n=$1; shift
  if [ "$n" = "1" ]; then python "$@"; else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) "$@"; fi
}
for n in 1 2 4 8; do
  run $n bench.py --gpus $n --steps 10 --warmup 3 > $OUT/bench_n$n.log 2> $OUT/bench_n$n.err
  echo "bench n=$n rc=$?"; grep '^{' $OUT/bench_n$n.log | tail -1 | cut -c1-160
done
for n in 1 8; do
  run $n tools/train_bench.py --steps 20 --warmup 5 > $OUT/train_n$n.log 2> $OUT/train_n$n.err
  echo "train n=$n rc=$?"; grep '^{' $OUT/train_n$n.log | tail -1
done

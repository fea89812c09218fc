#!/bin/bash
# Multi-GPU session on one box: bench cfg3 (weak scaling, no collective), cfg4 (strong scaling, NCCL all-reduce inside the
# step) and cfg5 (self-play, weak) at every N given.   usage: gpu_multi.sh TAG "1 2 4 8"
TAG=${1:-x}; NS=${2:-"1 2"}; O=gpurun_out; mkdir -p $O
nvidia-smi -L | head -8
for N in $NS; do
  for WL in cfg3 cfg4 cfg5; do
    ST=100; [ $WL = cfg4 ] && ST=20
    EXTRA="--no-cpu-baseline"; [ $WL = cfg5 ] && EXTRA="--deal uniform"
    if [ $N = 1 ]; then
      python bench.py --gpus 1 --steps $ST --warmup 5 --workload $WL $EXTRA > $O/${TAG}_n${N}_${WL}.json 2> $O/${TAG}_n${N}_${WL}.err
    else
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus $N --steps $ST --warmup 5 --workload $WL $EXTRA > $O/${TAG}_n${N}_${WL}.json 2> $O/${TAG}_n${N}_${WL}.err
    fi
    echo "N=$N $WL rc=$?"; tail -1 $O/${TAG}_n${N}_${WL}.json | cut -c1-400
  done
done

#!/bin/bash
# Multi-GPU session on one box: the default bench line (cfg3 weak scaling, cfg4 strong scaling with the in-kernel peer-memory
# reduction, cfg5, ranges, sustained legs) at every N given, and the 2-GPU tests.   usage: gpu_multi.sh TAG "1 2 4 8"
TAG=${1:-x}; NS=${2:-"1 2"}; O=gpurun_out; mkdir -p $O
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_dist.py -m gpu -x -q --timeout 500 > $O/${TAG}_pytest_dist.log 2>&1; echo "pytest dist rc=$?"
for N in $NS; do
  if [ $N = 1 ]; then
    python bench.py --gpus 1 --steps 20 --warmup 5 > $O/${TAG}_n${N}.json 2> $O/${TAG}_n${N}.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
      bench.py --gpus $N --steps 20 --warmup 5 > $O/${TAG}_n${N}.json 2> $O/${TAG}_n${N}.err
  fi
  echo "N=$N rc=$?"; tail -1 $O/${TAG}_n${N}.json | cut -c1-300
done

#!/bin/bash
# Multi-GPU session: bench cfg3 (weak, no collective) and cfg4 (strong, NCCL all-reduce) at N ranks.  usage: gpu_multi.sh TAG N
TAG=${1:-x}; N=${2:-2}; O=gpurun_out; mkdir -p $O
nvidia-smi -L | head -8
for WL in cfg3 cfg4; do
  ST=100; [ $WL = cfg4 ] && ST=20
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps $ST --warmup 5 --workload $WL > $O/${TAG}_n${N}_${WL}.json 2> $O/${TAG}_n${N}_${WL}.err
  echo "rc=$?"; tail -1 $O/${TAG}_n${N}_${WL}.json
  python bench.py --gpus 1 --steps $ST --warmup 5 --workload $WL --no-cpu-baseline > $O/${TAG}_n1_${WL}.json 2>> $O/${TAG}_n${N}_${WL}.err
  tail -1 $O/${TAG}_n1_${WL}.json
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 -m pytest tests/test_dist.py -q -x -k nccl > $O/${TAG}_n${N}_pytest.log 2>&1
tail -3 $O/${TAG}_n${N}_pytest.log

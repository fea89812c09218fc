#!/bin/bash
# One GPU session: parity tests, default bench, reference arm, ncu launch list + one full capture of the top kernel.
# usage (on the GPU box, from the repo root): bash tools/gpu_round.sh TAG
TAG=${1:-x}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
tail -3 $O/${TAG}_pytest.log
timeout 300 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"
cat $O/${TAG}_bench.json
timeout 200 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_cfg4.json 2>> $O/${TAG}_bench.err
cat $O/${TAG}_bench_cfg4.json
timeout 200 python bench.py --deal reference --steps 20 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_refdeal.json 2>> $O/${TAG}_bench.err
cat $O/${TAG}_bench_refdeal.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:equity_uniform -s 4 -c 1 -o $O/${TAG}_uniform \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $O | tail -12

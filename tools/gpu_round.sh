#!/bin/bash
# One GPU session: parity tests (normal and checked build), smoke, the default bench line, the reference arm, the ncu launch
# list of a self-play step and full captures of the main kernels (summarised on the box: the .ncu-rep files stay in /tmp).
# usage (on the GPU box, from the repo root): bash tools/gpu_round.sh TAG
TAG=${1:-x}
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 900 > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
tail -3 $O/${TAG}_pytest.log
CK=$PWD/neuron_poker_b200/build/libnpk_checked.so
if [ -f $CK ]; then
  NPK_LIBRARY=$CK timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_ranges.py tests/test_holdem.py -m gpu -x -q -s \
      --timeout 900 -k "not exhaustive" > $O/${TAG}_pytest_checked.log 2>&1; echo "pytest (checked build) rc=$?"
  grep -E "checked build|passed|failed" $O/${TAG}_pytest_checked.log | tail -3
  NPK_LIBRARY=$CK timeout 600 python tools/sanitizer_workload.py > $O/${TAG}_workload_checked.log 2>&1; echo "workload (checked build) rc=$?"
  tail -2 $O/${TAG}_workload_checked.log
fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/${TAG}_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"
timeout 300 python tools/latency.py > $O/${TAG}_latency.txt 2>&1; echo "latency rc=$?"
timeout 300 python tools/latency_breakdown.py > $O/${TAG}_latency_breakdown.txt 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference.json 2>> $O/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_cfg5.csv \
    python bench.py --workload cfg5 --deal uniform --steps 3 --warmup 3 > $O/${TAG}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
cap() { NAME=$1; RX=$2; SKIP=$3; UNITS=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$RX -s $SKIP -c 1 -f -o /tmp/$NAME "$@" > $O/${TAG}_ncu_$NAME.log 2>&1
  echo "ncu $NAME rc=$?"; python tools/ncu_summary.py /tmp/$NAME.ncu-rep $UNITS > $O/${TAG}_ncu_$NAME.txt 2>&1; }
cap uniform equity_uniform 4 40960000 python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline --no-extras
cap refdeal equity_refdeal 4 40960000 python bench.py --workload cfg3 --deal reference --steps 2 --warmup 3 --no-cpu-baseline --no-extras
cap mixed equity_mixed 20 65536000 python bench.py --workload cfg5 --deal uniform --steps 3 --warmup 3
cap rank7 rank7_kernel 2 16777216 python tools/ncu_others.py rank7
cap enum enum_kernel 2 47646720 python tools/ncu_others.py enum
cap ranges_fast equity_ranges_fast 2 4096000 python tools/ncu_others.py ranges
cap ranges_generic "equity_ranges_kernel" 2 4096000 python tools/ncu_others.py ranges generic
cap holdem holdem_step 2 65536 python tools/ncu_others.py holdem
ls -la $O | tail -12

#!/bin/bash
# Collect the measured evidence of a round on ONE B200 (run through gpurun): tools/round_evidence.sh <tag>   e.g. r01
# Order follows the profiling recipe: every program first runs to completion WITHOUT ncu, then the ncu passes.
T=${1:-r01}; O=gpurun_out
python -m pytest tests -x -q -m gpu > $O/${T}_pytest_gpu.log 2>&1; tail -2 $O/${T}_pytest_gpu.log
python __graft_entry__.py smoke > $O/${T}_smoke.log 2>&1; tail -2 $O/${T}_smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err
python bench.py --dump-ops $O/${T}_ops.txt > $O/${T}_bench.json 2> $O/${T}_bench.err; cut -c1-300 $O/${T}_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-library-bar > $O/${T}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tcgen05 -s 200 -c 14 -f -o $O/${T}_conv_full \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-library-bar > $O/${T}_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused -s 32 -c 8 -f -o $O/${T}_attn_full \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-library-bar > $O/${T}_ncu_attn.log 2>&1
ls -la $O/${T}_*

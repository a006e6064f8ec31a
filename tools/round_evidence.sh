#!/bin/bash
# Collect the measured evidence of a round on ONE B200 (run through gpurun): tools/round_evidence.sh <tag>   e.g. r02
# Order follows the profiling recipe: every program first runs to completion WITHOUT ncu, then the ncu passes.
T=${1:-r02}; O=gpurun_out
python -m pytest tests -x -q -m gpu > $O/${T}_pytest_gpu.log 2>&1; tail -2 $O/${T}_pytest_gpu.log
python __graft_entry__.py smoke > $O/${T}_smoke.log 2>&1; tail -2 $O/${T}_smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err
python bench.py --dump-ops $O/${T}_ops.txt > $O/${T}_bench.json 2> $O/${T}_bench.err; cut -c1-300 $O/${T}_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-library-bar > $O/${T}_ncu_launches.log 2>&1
# single launches of the conv engine at bench shapes (tools/prof_conv.py): level-0 plain / GroupNorm-prologue 3x3, 16x16 prologue, 4x4 prologue, 1x1
# (gpurun brings back at most 64 MiB: the reports are summarised here and only the level-0 prologue conv's report is kept)
for c in plain gn0 gn small qkv; do
  python tools/prof_conv.py $c > $O/${T}_prof_$c.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:conv_tcgen05 -s 2 -c 1 -f -o /tmp/${T}_conv_$c python tools/prof_conv.py $c > $O/${T}_ncu_conv_$c.log 2>&1
  if [ $c = gn0 ]; then python tools/ncu_summary.py /tmp/${T}_conv_$c.ncu-rep $O/${T}_conv_${c}_summary.txt $O/${T}_conv_traffic.json 0 "level-0 conv3x3 128->128 @32x32 with the fused GroupNorm+SiLU prologue, batch 256 (tools/prof_conv.py gn0; algorithmic bytes 134.5 MB: 67 MB in, 67 MB out, 0.3 MB weights)"; cp /tmp/${T}_conv_$c.ncu-rep $O/
  else python tools/ncu_summary.py /tmp/${T}_conv_$c.ncu-rep $O/${T}_conv_${c}_summary.txt; fi
done
ncu --set full --clock-control none --import-source on -k regex:fused -s 32 -c 8 -f -o /tmp/${T}_attn_full \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-library-bar > $O/${T}_ncu_attn.log 2>&1
python tools/ncu_summary.py /tmp/${T}_attn_full.ncu-rep $O/${T}_attn_summary.txt
ls -la $O/${T}_*

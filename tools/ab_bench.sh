#!/bin/bash
# A/B of two builds of libdmn_b200.so on the SAME box: tools/ab_bench.sh <libA.so> [rounds]   (B = the in-tree library)
A=$1; N=${2:-2}
for i in $(seq 1 $N); do
  DMN_LIB_PATH=$A python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/ab_A$i.log 2>&1
  python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/ab_B$i.log 2>&1
done
for f in gpurun_out/ab_[AB]*.log; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    kb = d["kernel_breakdown"]
    print(sys.argv[1], "samples/s %.2f  ms/step %.4f  conv %.4f  finalize %.4f  linattn %.4f" % (
        d["value"], d["ms_per_step"], kb["conv_tcgen05"]["ms"], kb["gn_finalize"]["ms"], kb["linattn_core"]["ms"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done

#!/bin/bash
# Same-box comparison of several builds of libdmn_b200.so: tools/ab_bench.sh <rounds> <lib1.so> [lib2.so ...]  ("-" = the in-tree library)
N=$1; shift
for i in $(seq 1 $N); do
  k=0
  for lib in "$@"; do
    k=$((k+1))
    if [ "$lib" = "-" ]; then unset DMN_LIB_PATH; else export DMN_LIB_PATH=$lib; fi
    python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e --dump-ops gpurun_out/ab_ops_$k.txt > gpurun_out/ab_${k}_$i.log 2>&1
  done
done
unset DMN_LIB_PATH
k=0
for lib in "$@"; do
  k=$((k+1))
  for i in $(seq 1 $N); do python - "gpurun_out/ab_${k}_$i.log" "$lib" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    kb = d["kernel_breakdown"]
    print("%-28s samples/s %.2f  ms/step %.4f  conv %.4f  finalize %.4f  linattn %.4f" % (
        sys.argv[2], d["value"], d["ms_per_step"], kb["conv_tcgen05"]["ms"], kb["gn_finalize"]["ms"], kb.get("other", {"ms": 0})["ms"]))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
  done
done

#!/bin/bash
# usage (GPU box): tools/variants.sh "<name>:<-D flags>" ...   builds each variant into gpurun_out/ and runs the
# serial kernel-timing part of the bench with it (experiments only)
for v in "$@"; do
  name=${v%%:*}; flags=${v#*:}
  mkdir -p /tmp/var; so=/tmp/var/lib_$name.so
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -shared --cudart static $flags -o $so rbepwt_b200/csrc/rbepwt_b200.cu || exit 1
  RBEPWT_B200_LIB=$so python bench.py --steps 12 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --parity-images 0 > gpurun_out/var_$name.json 2>gpurun_out/var_$name.err
  python - <<PY
import json
d=json.load(open("gpurun_out/var_$name.json"))
k=d["kernels"]
print("%-14s value %6.0f  " % ("$name", d["value"]) + "  ".join("%s %.3f" % (n.split("+")[0], v["ms_per_step"]) for n, v in k.items()))
PY
done

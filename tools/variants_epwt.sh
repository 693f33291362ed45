#!/bin/bash
# usage (GPU box): tools/variants_epwt.sh "<name>:<-D flags>" ...   builds each variant and times EPWT config 3
for v in "$@"; do
  name=${v%%:*}; flags=${v#*:}
  mkdir -p /tmp/var; so=/tmp/var/lib_$name.so
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -shared --cudart static "$flags" -o $so rbepwt_b200/csrc/rbepwt_b200.cu || exit 1
  echo "== $name"; RBEPWT_B200_LIB=$so python tools/epwt_time.py 2>&1 | tail -2
done

#!/bin/bash
# usage (build container, then GPU box):
#   tools/ab_libs.sh build <git-rev>     here: compiles that revision's kernels into tools/scratch/lib_<rev>.so
#                                        (git-ignored, but it travels to the GPU box with the snapshot)
#   tools/ab_libs.sh run <git-rev> [B]   GPU box: heavy-tailed + uniform timing and the per-kernel bench table with that
#                                        library and with the current one, same process conditions
# A/B of two library versions inside ONE gpurun call: boxes differ by a few percent from call to call.
set -e
cmd=$1; rev=$2; B=${3:-256}
root=$(cd "$(dirname "$0")/.." && pwd)
so=$root/tools/scratch/lib_$rev.so
if [ "$cmd" = build ]; then
  mkdir -p $root/tools/scratch/src_$rev
  git -C $root archive $rev rbepwt_b200/csrc include | tar -x -C $root/tools/scratch/src_$rev
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -shared --cudart static \
       -o $so $root/tools/scratch/src_$rev/rbepwt_b200/csrc/rbepwt_b200.cu
  rm -rf $root/tools/scratch/src_$rev
  ls -la $so
  exit 0
fi
for lib in $so $root/rbepwt_b200/_lib/librbepwt_b200.so; do
  echo "== $lib"
  RBEPWT_B200_LIB=$lib python $root/tools/heavytail_timing.py $B | cut -c1-230
  RBEPWT_B200_LIB=$lib python $root/bench.py --steps 12 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels']
print('bench value %.0f  %.2f ms/step  ' % (d['value'], d['ms_per_step']) + '  '.join('%s %.2f' % (n.split('+')[0], v['ms_per_step']) for n, v in k.items()))"
done

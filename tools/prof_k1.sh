#!/bin/bash
# usage (GPU box): tools/prof_k1.sh <tag> [kernel-regex] [skip]   -- ncu --set full of one launch of a path kernel
tag=$1; k=${2:-k1_walk}; skip=${3:-3}
ARGS="--batch 512 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
python bench.py $ARGS > gpurun_out/prof_plain_$tag.json 2> gpurun_out/prof_plain_$tag.err || { echo "plain run failed"; tail -5 gpurun_out/prof_plain_$tag.err; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -o gpurun_out/prof_${k}_$tag \
    python bench.py $ARGS > gpurun_out/ncu_full_${k}_$tag.log 2>&1
cp rbepwt_b200/_lib/librbepwt_b200.so gpurun_out/lib_$tag.so
ls -la gpurun_out | grep $tag

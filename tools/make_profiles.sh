#!/bin/bash
# usage (GPU box): tools/make_profiles.sh <round-tag>
# 1. plain bench run (must exit 0), 2. ncu launch list of our kernels over one step,
# 3. ncu --set full of one launch of each hot kernel.  Everything lands in gpurun_out/.
tag=$1
ARGS="--batch 512 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
python bench.py $ARGS > gpurun_out/prof_plain_$tag.json 2> gpurun_out/prof_plain_$tag.err || { echo "plain run failed"; tail -5 gpurun_out/prof_plain_$tag.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k[0-9q_]" -c 400 --csv \
    --log-file gpurun_out/launches_$tag.csv python bench.py $ARGS > gpurun_out/ncu_list_$tag.log 2>&1
for k in k1_paths_tpr k1_bitmaps k3_dwt_level k5_idwt_level k4_threshold k0_regions_fast k0_count; do
  skip=2; [ $k = k1_paths_tpr ] && skip=3; [ $k = k3_dwt_level ] && skip=14; [ $k = k5_idwt_level ] && skip=20
  ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -o gpurun_out/prof_${k}_$tag \
      python bench.py $ARGS > gpurun_out/ncu_full_${k}_$tag.log 2>&1
done
ls -la gpurun_out | grep $tag

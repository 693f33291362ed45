#!/bin/bash
# usage (GPU box): tools/make_profiles.sh <round-tag>
# 1. plain bench run (must exit 0), 2. ncu launch list of our kernels over one step,
# 3. ncu --set full of one launch of each hot kernel.  Everything lands in gpurun_out/;
#    tools/summarize_profiles.py <tag> turns it into the committed profiles/<tag>_*.txt.
tag=$1
T=/tmp/prof_$tag; mkdir -p $T gpurun_out   # the .ncu-rep files stay on the box (gpurun_out/ is capped at 64 MiB): summaries travel
ARGS="--batch 512 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --parity-images 0"
python bench.py $ARGS > gpurun_out/prof_plain_$tag.json 2> gpurun_out/prof_plain_$tag.err || { echo "plain run failed"; tail -5 gpurun_out/prof_plain_$tag.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k[0-9q_]" -c 600 --csv \
    --log-file gpurun_out/launches_$tag.csv python bench.py $ARGS > gpurun_out/ncu_list_$tag.log 2>&1
# launch-skip: 3 warm-up steps.  Per step (512 device-resident images = 2 path groups and 2 transform sub-batches of 256):
# k1_walk launches 4 times (per group: the windowed instantiation, then the bulk one), k3/k5 14 times (per sub-batch the 7
# level launches; the tail kernel has its own name), everything else twice -- the captures land on the timed step's first launches
for spec in k1_walk:12:2 k1_bitmaps:6:1 kq_slots:6:1 k1_coop_all:6:1 k2_perm:6:1 k4_select:6:1 k3_dwt_level:42:1 k5_idwt_level:48:1 k3_dwt_tail:6:1 k0_regions_fast:6:1 k0_count:6:1; do
  IFS=: read k skip cnt <<< "$spec"
  ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c $cnt -o $T/prof_${k}_$tag \
      python bench.py $ARGS > gpurun_out/ncu_full_${k}_$tag.log 2>&1
done
PROF_GROUP=256 PROF_SRC=$T PROF_LAUNCHES=gpurun_out PROF_DST=gpurun_out/profiles_$tag python tools/summarize_profiles.py $tag
cp $T/prof_k1_walk_$tag.ncu-rep gpurun_out/ 2>/dev/null
ls -la gpurun_out/profiles_$tag

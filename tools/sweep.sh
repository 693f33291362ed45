#!/bin/bash
# usage (GPU box): tools/sweep.sh "<bench args>" ...   one short bench run per argument string
for a in "$@"; do
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --parity-images 0 --no-e2e $a > gpurun_out/sweep.json 2>gpurun_out/sweep.err || { tail -3 gpurun_out/sweep.err; continue; }
  python - "$a" <<'PY'
import json,sys
d=json.load(open("gpurun_out/sweep.json"))
k=d["kernels"]
g=lambda n: k[n]["ms_per_step"] if n in k else 0
print("%-34s value %6.0f (%.2f ms) e2e %s | serial: k0 %.2f k1 %.2f dwt %.2f sel %.2f idwt %.2f" % (sys.argv[1], d["value"], d["ms_per_step"], d["e2e"] and round(d["e2e"]["value"]), g("k0_count+k0_regions_fast+queue"), g("k1_walk"), g("k3_dwt_level"), g("k4_threshold"), g("k5_idwt_level")))
PY
done

#!/bin/bash
# usage (on the GPU box, via gpurun): tools/gpu_check.sh <tag> [bench args]
# GPU parity tests, then a short bench run whose JSON lands in gpurun_out/bench_<tag>.json
tag=$1; shift
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; prc=$?
python bench.py --steps 5 --warmup 3 "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_$tag.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_$tag.json"))
print("value %.0f img/s  %.2f ms/step  e2e %s" % (d["value"], d["ms_per_step"], d["e2e"] and round(d["e2e"]["value"])))
for k, v in d["kernels"].items():
    print("  %-28s %8.3f ms/step  share %.3f  %7.0f GB/s alg" % (k, v["ms_per_step"], v["share"], v["algorithmic_GBps"]))
PY
echo "pytest rc=$prc: $(tail -1 gpurun_out/pytest_$tag.log)"
[ $prc -ne 0 ] && tail -30 gpurun_out/pytest_$tag.log
exit 0

#!/bin/bash
# usage (on the GPU box, from the repo root): tools/gpu_check.sh TAG [bench args...]
# GPU parity tests, then one bench line into gpurun_out/bench_TAG.json
TAG=${1:-x}; shift
python -m pytest tests -m gpu -x -q 2>&1 | tail -12
python bench.py --steps 128 --warmup 16 "$@" > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 400 gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print("value %.3fM ms/step %.4f e2e %.3fM aba_ms %.4f aba_frac %.4f cpu %s" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["roofline_aba"]["avg_launch_ms"], d["roofline_aba"]["frac"], d.get("cpu_baseline",{}).get("value")))
PY

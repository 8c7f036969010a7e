#!/bin/bash
# usage (GPU box): tools/final_evidence.sh TAG -- tests, bench (+ reference arm), launch list, one ncu --set full capture
TAG=${1:-rXX}
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 256 --warmup 32 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 200 gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 64 --warmup 4 > gpurun_out/${TAG}_bench_reference_arm.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:env_step_kernel -s 20 -c 1 -o gpurun_out/${TAG}_step python tools/ncu_step.py 2 > /dev/null 2>&1
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["cpu_baseline"]["value"], d["gpu_launches"], d["roofline_aba"]["avg_launch_ms"], d["roofline_temp_states"]["frac"], d["clocks"])
PY

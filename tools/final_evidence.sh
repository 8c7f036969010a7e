#!/bin/bash
# usage (GPU box): tools/final_evidence.sh TAG -- tests, bench (+ reference arm), launch list, one ncu --set full capture of the
# step kernel (summarised into profiles/traffic.json with the git head: bench.py reads roofline.traffic / roofline_compute there)
TAG=${1:-rXX}
python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
ncu --set full --clock-control none --import-source on -k regex:env_step_kernel -s 20 -c 1 -o gpurun_out/${TAG}_step python tools/ncu_step.py 2 > /dev/null 2>&1
ncu -i gpurun_out/${TAG}_step.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_step.ncu-rep --page source --csv > gpurun_out/${TAG}_sass.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${TAG}_raw.csv ${TAG} > /dev/null && cp profiles/traffic.json gpurun_out/${TAG}_traffic.json && cp profiles/${TAG}_env_step_ncu.csv gpurun_out/
python tools/ncu_opcodes.py gpurun_out/${TAG}_sass.csv > gpurun_out/${TAG}_env_step_opcodes.txt 2>/dev/null
# per-phase table (needs the disassembly of the .so that ran):  cuobjdump -xelf all libss_b200.so; nvdisasm -g -c *.cubin > all.sass;
#   python tools/ncu_phases.py gpurun_out/${TAG}_sass.csv all.sass env_step_kernelILb0 > profiles/${TAG}_env_step_phases.txt
python bench.py --steps 256 --warmup 32 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 200 gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_reference_arm.json 2>/dev/null
python bench.py --workload mike --steps 128 --warmup 16 --no-cpu-baseline > gpurun_out/${TAG}_bench_mike.json 2>/dev/null
python bench.py --workload w3d --steps 128 --warmup 16 --no-cpu-baseline > gpurun_out/${TAG}_bench_w3d_shipped.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-walking > /dev/null 2>&1
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"].get("host_us_per_step"), d.get("walking", {}).get("ms_per_step"), d["cpu_baseline"]["value"], d["gpu_launches"], d["roofline"]["traffic"], d["roofline_aba"]["avg_launch_ms"], d["roofline_temp_states"]["frac"], d["clocks"])
r = json.load(open("gpurun_out/${TAG}_bench_reference_arm.json")); print("reference arm", r["value"], r["cpu_baseline"], r.get("reference_probe"))
PY

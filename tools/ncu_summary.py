"""ncu raw page (ncu -i X.ncu-rep --page raw --csv) of ONE env_step_kernel launch -> profiles/traffic.json: the numbers
bench.py puts on its line as roofline.traffic / roofline_compute, with the git head and the capture they come from.
usage: python tools/ncu_summary.py raw.csv TAG [envs]      (writes profiles/traffic.json and profiles/TAG_env_step_ncu.csv)"""
import csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw, tag = sys.argv[1], sys.argv[2]
envs = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
rows = list(csv.reader(open(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {}
for h, u, v in zip(hdr, units, vals):
    try:
        m[h] = (float(v.replace(",", "")), u)
    except ValueError:
        pass
def g(name, scale=1.0):
    return m[name][0] * scale if name in m else None
def bytes_of(name):
    v, u = m[name]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
rd, wr = bytes_of("dram__bytes_read.sum"), bytes_of("dram__bytes_write.sum")
try:
    head = subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True, text=True).stdout.strip()
except Exception:
    head = None
out = {
    "env_step_kernel_dram_bytes_per_launch": int(rd + wr), "dram_bytes_read": int(rd), "dram_bytes_written": int(wr),
    "algorithmic_bytes_per_launch": 884 * envs, "traffic_over_algorithmic": (rd + wr) / (884 * envs),
    "issue_active_pct": g("smsp__issue_active.avg.per_cycle_active", 100.0),
    "pipe_fma_pct_of_peak": g("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    "pipe_alu_pct_of_peak": g("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    "warp_instructions_per_env_substep": g("inst_executed", 1.0 / (envs * 4)),
    "local_load_instructions": g("sass__inst_executed_local_loads"), "local_store_instructions": g("sass__inst_executed_local_stores"),
    "registers_per_thread": g("launch__registers_per_thread"), "warps_active_pct": g("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "kernel_us_under_ncu": g("gpu__time_duration.sum", 1.0) if m.get("gpu__time_duration.sum", (0, ""))[1] in ("us", "usecond") else g("gpu__time_duration.sum"),
    "stall_per_issue": {k.split("issue_stalled_")[1].split("_per_issue")[0]: round(v[0], 3) for k, v in m.items()
                        if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and v[0] >= 0.1},
    "envs": envs, "git_head": head, "capture": "profiles/%s_env_step_ncu.csv" % tag,
    "source": "one `ncu --set full --clock-control none` launch of env_step_kernel<false> (tools/ncu_step.py: N=%d, W3D flat, collision_mode 2, step 21 after reset), L2 cold" % envs,
}
keep = [i for i, h in enumerate(hdr) if any(k in h for k in ("dram__bytes", "issue_active", "pipe_fma", "pipe_alu", "inst_executed", "local", "registers",
                                                                "warps_active", "issue_stalled", "gpu__time_duration", "Kernel Name", "shared_mem", "occupancy"))]
with open(os.path.join(ROOT, "profiles", "%s_env_step_ncu.csv" % tag), "w", newline="") as f:
    w = csv.writer(f)
    for r in rows[:3]:
        w.writerow([r[i] for i in keep])
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))

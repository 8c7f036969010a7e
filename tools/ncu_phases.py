"""Where the step kernel's issue slots and stall samples go, PHASE by phase: joins ncu's SASS page (per-instruction
executed counts and warp-stall samples, `ncu -i X.ncu-rep --page source --csv`) with `nvdisasm -g` line info and groups
the instructions by the source range (innermost inlined location) they come from.
usage: ncu_phases.py sass.csv disasm.sass kernel_name [envs*substeps]

The ranges follow the function boundaries of steppingstone_b200/csrc/ss_device.cuh at the git head named in
profiles/traffic.json; re-derive them (grep -n __device__) when the file moves."""
import csv, re, sys
from collections import defaultdict

PHASES = [  # (file, first line, last line, phase)
    ("sm_30_intrinsics.hpp", 0, 10**9, "warp shuffles (inlined everywhere)"),
    ("ss_device.cuh", 120, 225, "small helpers: rcp/sqrt, 3-vectors, transforms, tree scans"),
    ("ss_device.cuh", 226, 255, "forward kinematics"),
    ("ss_device.cuh", 256, 325, "U D U^T factorisation of the joint-space inertia"),
    ("ss_device.cuh", 326, 533, "dynamics: velocities, bias forces, composite inertias, H"),
    ("ss_device.cuh", 534, 570, "triangular solves with the stored factor"),
    ("ss_device.cuh", 571, 616, "friction basis, integration"),
    ("ss_device.cuh", 617, 681, "contact pool (merge, order, cap)"),
    ("ss_device.cuh", 682, 903, "collision detection (feet-stones, link-link, ground)"),
    ("ss_device.cuh", 904, 943, "PGS iterations"),
    ("ss_device.cuh", 944, 1126, "constraint rows: limits, Jacobians, z = U^-1 J^T, row constants"),
    ("ss_device.cuh", 1127, 1170, "Delassus rows"),
    ("ss_device.cuh", 1171, 1197, "PGS set-up and dispatch on the contact count"),
    ("ss_device.cuh", 1198, 1223, "velocity update, warm-start cache"),
    ("ss_device.cuh", 1224, 1297, "MT19937 (envs that reset)"),
    ("ss_device.cuh", 1298, 10**9, "angle helpers of the env logic"),
    ("ss_kernels.cu", 60, 110, "per-substep CTA barrier + model staging (line table puts the barrier here)"),
    ("ss_kernels.cu", 0, 10**9, "env logic: action, observation, reward, terrain, done, auto-reset, outputs"),
]

sass_csv, disasm, kname = sys.argv[1:4]
units = float(sys.argv[4]) if len(sys.argv) > 4 else 8192 * 4
seq = []; cur = None; infn = False
inl = re.compile(r'//## File "([^"]+)", line (\d+)')
for ln in open(disasm):
    if ln.startswith(".text."):
        infn = kname in ln; continue
    if not infn: continue
    m = inl.search(ln)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln): seq.append(cur)
hdr = None; data = []
for r in csv.reader(open(sass_csv)):
    if r and r[0] == "Address":
        if hdr is not None: break
        hdr = r; continue
    if hdr is not None and r and r[0].startswith("0x"): data.append(r)
ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]


def phase_of(loc):
    if loc is None: return "(no line info)"
    f, l = loc
    for pf, a, b, name in PHASES:
        if f == pf and a <= l <= b: return name
    return "other (%s)" % f


ins = defaultdict(float); smp = defaultdict(float); stall = defaultdict(lambda: defaultdict(float))
for k in range(min(len(data), len(seq))):
    p = phase_of(seq[k])
    ins[p] += float(data[k][ie]); smp[p] += float(data[k][isamp])
    for i, name in stall_cols:
        stall[p][name] += float(data[k][i])
ti, ts = sum(ins.values()), sum(smp.values())
print("kernel %s: %.3e warp instructions, %d stall samples, %d SASS instructions with line info" % (kname, ti, ts, len(seq)))
print("%-78s %9s %7s %8s  %s" % ("phase", "inst/unit", "% inst", "% sample", "top stall reasons (share of the phase's samples)"))
order = [name for _, _, _, name in PHASES] + sorted(set(ins) - {name for _, _, _, name in PHASES})
for p in order:
    if ins[p] == 0 and smp[p] == 0: continue
    top = sorted(stall[p].items(), key=lambda kv: -kv[1])[:3]
    tot = sum(stall[p].values()) or 1.0
    print("%-78s %9.0f %6.1f%% %7.1f%%  %s" % (p, ins[p] / units, 100 * ins[p] / ti, 100 * smp[p] / ts,
                                               ", ".join("%s %.0f%%" % (n, 100 * v / tot) for n, v in top)))

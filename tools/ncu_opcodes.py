"""Executed warp instructions per SASS opcode from ncu's source page (ncu -i X.ncu-rep --page source --csv > sass.csv).
usage: python tools/ncu_opcodes.py sass.csv [units]   (units = envs * substeps of the captured launch, default 8192*4)"""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]) if len(sys.argv) > 2 else 8192 * 4
hdr = None; agg = collections.Counter(); smp = collections.Counter(); tot = 0; ts = 0
for r in rows:
    if r and r[0] == "Address":
        if hdr is not None: break
        hdr = r; ie = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples"); continue
    if hdr is None or not r or not r[0].startswith("0x"): continue
    m = re.match(r"\s*(?:@!?U?P\w+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", r[1])
    op = m.group(1) if m else "?"
    if op in ("LDS", "STS", "LDL", "STL", "LDG", "STG", "LD", "ST") and m and ".128" in m.group(2): op += ".128"
    n = float(r[ie]); s = float(r[isamp])
    agg[op] += n; smp[op] += s; tot += n; ts += s
print("total %.4e warp instructions (%.0f per unit), %d stall samples" % (tot, tot / units, ts))
for op, n in agg.most_common(45):
    print("  %-10s %8.0f /unit %5.1f%% insts %5.1f%% samples" % (op, n / units, 100 * n / tot, 100 * smp[op] / max(ts, 1)))

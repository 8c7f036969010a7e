"""Executed warp instructions per source FUNCTION-ish region: joins ncu's SASS page with nvdisasm -g line info and groups
by (file, 25-line bucket).  usage: ncu_regions.py sass.csv disasm.sass kernel_name [envs*substeps]"""
import csv, re, sys
from collections import defaultdict
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
rows = list(csv.reader(open(sass_csv)))
hdr = None; data = []
for r in rows:
    if r and r[0] == "Address":
        if hdr is not None: break
        hdr = r; continue
    if hdr is not None and r and r[0].startswith("0x"): data.append(r)
ie = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples")
agg = defaultdict(float); smp = defaultdict(float); tot = 0; ts = 0
for k in range(min(len(data), len(seq))):
    f, l = seq[k] if seq[k] else ("?", 0)
    key = (f, l // 25 * 25)
    agg[key] += float(data[k][ie]); smp[key] += float(data[k][isamp]); tot += float(data[k][ie]); ts += float(data[k][isamp])
print("total %.3e insts, %d samples" % (tot, ts))
for key, v in sorted(agg.items()):
    if v / tot > 0.004:
        print("%-24s %6.0f /unit %5.1f%% insts %5.1f%% samples" % ("%s:%d" % key, v / units, 100 * v / tot, 100 * smp[key] / ts))

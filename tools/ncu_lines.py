"""Join an `ncu --page source --print-source sass --csv` dump with `nvdisasm -g` line info: per-source-line
executed warp instructions and stall samples.  usage: ncu_lines.py sass.csv all.sass kernel_mangled_prefix [top]"""
import csv, re, sys
from collections import defaultdict
sass_csv, disasm, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# disasm: sequence of (file,line) per instruction, in address order, for the chosen function
seq = []; cur = None; infn = False
inl = re.compile(r'//## File "([^"]+)", line (\d+)( inlined at "([^"]+)", line (\d+))?')
for ln in open(disasm):
    if ln.startswith(".text."):
        infn = kname in ln
        continue
    if not infn: continue
    m = inl.search(ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        seq.append(cur)
rows = list(csv.reader(open(sass_csv)))
# first kernel instance only
hdr = None; data = []
for r in rows:
    if r and r[0] == "Address": 
        if hdr is not None: break
        hdr = r; continue
    if hdr is not None and r and r[0].startswith("0x"): data.append(r)
ie = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples")
print("sass rows", len(data), "disasm insts", len(seq))
n = min(len(data), len(seq))
inst = defaultdict(float); samp = defaultdict(float)
for k in range(n):
    key = seq[k]
    inst[key] += float(data[k][ie]); samp[key] += float(data[k][isamp])
ti = sum(inst.values()); ts = sum(samp.values())
print("total warp insts %.3e  samples %d" % (ti, ts))
print("%-22s %10s %6s %8s %6s" % ("file:line", "insts", "%", "samples", "%"))
for key, v in sorted(inst.items(), key=lambda kv: -samp[kv[0]])[:top]:
    print("%-22s %10.3e %5.1f%% %8d %5.1f%%" % ("%s:%d" % key if key else "?", v, 100 * v / ti, samp[key], 100 * samp[key] / ts))

"""Where does ptxas spill?  Lists local-memory SASS instructions (LDL/STL) of one kernel by source line (needs -lineinfo).
usage: python tools/sass_spills.py lib.so [kernel-substring]   (static view; ncu's source page gives the executed counts)"""
import collections, os, re, subprocess, sys, tempfile
so = sys.argv[1]; kname = sys.argv[2] if len(sys.argv) > 2 else "env_step_kernelILb0"
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
infn = False; cur = None; n = 0
byline = collections.Counter(); ops = collections.Counter()
inl = re.compile(r'//## File "([^"]+)", line (\d+)')
for ln in dis.splitlines():
    if ln.startswith(".text."):
        infn = kname in ln; continue
    if not infn: continue
    m = inl.search(ln)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m:
        n += 1
        op = m.group(1).split(".")[0]
        ops[op] += 1
        if op in ("LDL", "STL"): byline[(cur, op)] += 1
print("kernel %s: %d SASS instructions; LDL %d STL %d" % (kname, n, ops["LDL"], ops["STL"]))
for (loc, op), c in sorted(byline.items(), key=lambda x: (x[0][0] or ("", 0))):
    print("  %-28s %s x%d" % ("%s:%d" % loc if loc else "?", op, c))
if "-ops" in sys.argv:
    for op, c in ops.most_common(40): print("  %-10s %d" % (op, c))

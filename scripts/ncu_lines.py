"""Attribute the PC-sampling stall samples and executed instructions of one kernel in an .ncu-rep to CUDA source lines.
usage: python scripts/ncu_lines.py <report.ncu-rep> <object.o> <kernel name substring> [top=30] [launch index=0]
The report must come from `ncu --set full --import-source on`; the object from the same build (compiled with -lineinfo)."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, obj, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
which = int(sys.argv[5]) if len(sys.argv) > 5 else 0
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info-inline", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
# address -> (file, line) inside the wanted function
amap, cur, infn = {}, None, False
for ln in dis:
    if ln.startswith("\t.section\t.text."):
        infn = kname in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:   # consecutive annotations = the inlining chain, innermost first: keep the outermost frame (the call site in the kernel)
        cur = (os.path.basename(m.group(1)), int(m.group(2)), m.group(1))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+\S", ln)
    if m and cur:
        amap[int(m.group(1), 16)] = cur
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, curb = [], None
for r in csv.reader(raw.splitlines()):
    if r and r[0] == "Kernel Name":
        curb = {"name": r[1], "rows": []}
        blocks.append(curb)
    elif curb is not None:
        curb["rows"].append(r)
sel = [b for b in blocks if kname in b["name"]][which]
H = sel["rows"][0]
iA, iE, iN = H.index("Address"), H.index("Instructions Executed"), H.index("# Samples")
stall = [i for i, h in enumerate(H) if h.startswith("stall_") and "Not Issued" not in h]
agg, inst, st, base = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter), None
for r in sel["rows"][1:]:
    try:
        a, n, e = int(r[iA], 16), int(r[iN]), int(r[iE])
    except Exception:
        continue
    base = a if base is None else base
    key = amap.get(a - base, ("?", 0, ""))
    agg[key] += n
    inst[key] += e
    for i in stall:
        try:
            st[key][H[i][6:]] += int(r[i])
        except Exception:
            pass
tot, toti = sum(agg.values()), sum(inst.values())
print(f"{sel['name'][:80]}: {tot} samples, {toti} warp instructions")
cache = {}
for key, n in agg.most_common(top):
    f, l, path = key
    if path and path not in cache:
        try:
            cache[path] = open(path).read().split("\n")
        except Exception:
            cache[path] = []
    text = cache.get(path, [])[l - 1].strip()[:100] if path and l - 1 < len(cache.get(path, [])) else ""
    tops = ", ".join(f"{k}:{v}" for k, v in st[key].most_common(3))
    print(f"{100 * n / tot:5.1f}%  inst {100 * inst[key] / max(toti, 1):5.1f}%  {f}:{l:4d} [{tops}] {text}")

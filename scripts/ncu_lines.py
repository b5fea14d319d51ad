"""Attribute ncu per-SASS-instruction counters to CUDA source lines.

usage: python scripts/ncu_lines.py <report.ncu-rep> <kernel mangled-name substring> [top]
Needs the .so that was profiled (same build) in sod100k_b200/: lines come from `nvdisasm --print-line-info`,
instruction order is matched against `ncu --page source --csv`.
"""
import csv, io, os, re, subprocess, sys, tempfile, collections

rep, key = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "sod100k_b200", "libcsnet_b200.so")], cwd=tmp, capture_output=True)
sass = []
for cubin in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):
    sass += subprocess.run(["nvdisasm", "--print-line-info", cubin], cwd=tmp, capture_output=True, text=True).stdout.splitlines()
lines, cur, inside = [], None, False
for l in sass:
    if l.startswith("\t.section\t.text."):
        inside = key in l
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[h]
ie, ws, src = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source")
data = [r for r in rows[h + 1:] if len(r) > ie]
print(f"sass instructions: nvdisasm {len(lines)}, ncu {len(data)}")
agg = collections.defaultdict(lambda: [0, 0])
ops = collections.Counter()
for loc, r in zip(lines, data):
    n, s = int(r[ie] or 0), int(r[ws] or 0)
    agg[loc][0] += n
    agg[loc][1] += s
    ops[r[src].split()[0 if not r[src].strip().startswith("@") else 1].split(".")[0]] += n
ti, ts = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
print(f"total warp-instructions {ti}, stall samples {ts}")
srcs = {}
for (f, ln), (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if f not in srcs:
        for d in ("sod100k_b200/csrc",):
            p = os.path.join(root, d, f)
            srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
    text = srcs[f][ln - 1].strip()[:90] if srcs[f] and ln <= len(srcs[f]) else ""
    print(f"{100 * n / ti:5.1f}% inst {100 * s / max(ts, 1):5.1f}% stall  {f}:{ln:<4} {text}")
print("opcode mix:", ", ".join(f"{k} {100 * v / ti:.1f}%" for k, v in ops.most_common(18)))

"""profiles/traffic.json entries of the il_stream ops from an ncu CSV (dram__bytes_read.sum, dram__bytes_write.sum of the
il_stream_kernel launches of ONE forward at the bench configuration, in launch order = program order of the streaming ILBlock ops).
usage: python scripts/traffic_il_stream.py <ncu.csv> [model=csnet-L-x2] [batch=256] [size=224]"""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sod100k_b200 import checkpoints, compiler

path = sys.argv[1]
model = sys.argv[2] if len(sys.argv) > 2 else "csnet-L-x2"
B = int(sys.argv[3]) if len(sys.argv) > 3 else 256
S = int(sys.argv[4]) if len(sys.argv) > 4 else 224
rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
h = rows[0]
ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
per = {}
for r in rows[1:]:
    if "il_stream_kernel" not in r[ki]:
        continue
    per.setdefault(int(r[ii]), 0)
    per[int(r[ii])] += int(r[vi].replace(",", ""))
launches = [per[k] for k in sorted(per)]
cfg, sd = checkpoints.load_npz(model)
import torch
prog = compiler.compile_csnet(cfg, {k: torch.from_numpy(v) for k, v in sd.items()}, S, S, "fp16")
# the streaming kernel takes the ILBlock ops whose width is a multiple of 16 (plan.cu make_ils)
ops = [o.name for o in prog.ops if o.kind == 3 and prog.tensors[o.dst].W % 16 == 0]
print(len(launches), "il_stream launches;", len(ops), "ops:", ops)
assert len(launches) >= len(ops), "capture at least one whole forward"
launches = launches[:len(ops)]
tj_path = os.path.join(ROOT, "profiles", "traffic.json")
tj = json.load(open(tj_path))
for name, b in zip(ops, launches[:len(ops)]):
    tj[f"{model}:{name}:bs{B}:{S}x{S}:fp16"] = b
    print(f"{name:12s} {b / 1e6:9.1f} MB")
json.dump(tj, open(tj_path, "w"), indent=1)

"""Run only the fused-ILBlock programs on a small input (debug aid for compute-sanitizer)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sod100k_b200 import checkpoints, compiler, runtime, synth
cfg, sd = checkpoints.load_npz("csnet-L-x2")
h, w = int(os.environ.get("HH", 64)), int(os.environ.get("WW", 96))
x = torch.from_numpy(synth.randn_images(2, h, w, 5)).cuda()
only = os.environ.get("ONLY")
prog = compiler.compile_csnet(cfg, sd, h, w, "fp16", fuse=({only} if only else True))
print("fused ops:", [o.name for o in prog.ops if o.kind == 3])
plan = runtime.Plan(prog, max_batch=2)
y = plan.forward(x)
torch.cuda.synchronize()
print("ok", float(y.mean()))

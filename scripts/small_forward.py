"""Tiny forward for compute-sanitizer runs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sod100k_b200 import checkpoints, synth
for tag, dt in (("csnet-L-x2", "fp32"), ("csnet-L-x1", "fp16")):
    m, cfg, sd = checkpoints.build_from_npz(tag)
    m.cuda().eval().set_precision(dt)
    with torch.no_grad():
        y = m(torch.from_numpy(synth.randn_images(2, 64, 96, 5)).cuda())
    torch.cuda.synchronize()
    print(tag, dt, float(y.mean()))

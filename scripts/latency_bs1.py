"""bs-1 latency (the reference's real caller, CSNet/test.py:71-99): engine fp16 / fp32 vs eager PyTorch (the oracle's ATen calls) on the GPU."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import csnet_oracle as O
from sod100k_b200 import checkpoints, synth

def timeit(fn, n=200):
    for _ in range(20): fn()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / n * 1e3

for bs in (1, 4):
    x = torch.from_numpy(synth.randn_images(bs, 224, 224, 1)).cuda()
    cfg, sd = checkpoints.load_npz("csnet-L-x2")
    sdc = {k: torch.from_numpy(v).cuda() for k, v in sd.items()}
    with torch.no_grad():
        eager = timeit(lambda: O.csnet_forward(cfg, sdc, x), 50)
        res = {}
        for dt in ("fp32", "fp16"):
            m, _, _ = checkpoints.build_from_npz("csnet-L-x2")
            m.cuda().eval().set_precision(dt)
            m(x); m.engine().freeze(True)
            res[dt] = timeit(lambda: m(x))
    print(f"bs {bs}: eager ATen fp32 {eager:.3f} ms | engine fp32 {res['fp32']:.3f} ms | engine fp16 {res['fp16']:.3f} ms (wall clock per forward, launches included)", flush=True)

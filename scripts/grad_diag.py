"""Per-tensor gradient error of the training kernels against the float64 oracle, next to the fp32 oracle's own error (the
conditioning yardstick): which tensors exceed 1e-3 of their scale, and is it the kernels or the conditioning?"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import csnet_oracle as O
from sod100k_b200 import train_ops as T
from tests import test_gpu_train as G

out = {}
CASES = (("csnet-L-x2", (64, 64), 2), ("csnet-L-x2", (128, 128), 4), ("csnet-L-x1", (128, 160), 4), ("csnet-L-x1", (224, 224), 4))
if len(sys.argv) > 1:
    CASES = CASES[:int(sys.argv[1])]
for tag, hw, n in CASES:
    for seed in (51, 52, 53):
        m, cfg, params, buffers, x, t = G._setup(tag, n, hw, seed)
        o = m(torch.from_numpy(x).cuda())
        loss = T.BceFn.apply(o, torch.from_numpy(t).cuda())
        loss.backward()
        _, ref, *_ = O.train_step(cfg, params, buffers, {}, torch.from_numpy(x), torch.from_numpy(t))
        g64 = G._oracle_fp64_grads(cfg, params, buffers, x, t)
        rows = []
        for name, p in m.named_parameters():
            g, r, r64 = p.grad.detach().cpu().double(), ref[name].double(), g64[name]
            scale = max(r64.abs().max().item(), 1e-6)
            err, noise = (g - r64).abs().max().item() / scale, (r - r64).abs().max().item() / scale
            if err > 5e-4:
                rows.append((name, round(err, 5), round(noise, 6)))
        key = f"{tag} {hw} n{n} seed{seed}"
        out[key] = rows
        worst = sorted(rows, key=lambda r: -r[1])[:6]
        print(key, len(rows), "tensors above 5e-4; worst:", worst, flush=True)
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "grad_diag.json"), "w"), indent=1)

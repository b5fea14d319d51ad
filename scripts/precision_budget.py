"""fp16 error budget on the structured (blob) images: max / mean |sigmoid - sigmoid_ref| against the fp32 oracle for
  (a) the reference's own ATen calls with everything cast to half on the GPU (`model.half()` equivalent: what fp16 storage costs
      an unfused implementation), (b) our generic kernels with fp16 storage (fp32 weights, fp32 accumulate), (c) + fused ILBlock
      kernels, (d) the full bench program (+ 16-bit-weight tensor-core / streaming MIX kernels, MSBlock kernel).
Writes gpurun_out/precision_budget.json.  python scripts/precision_budget.py [n_images]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import csnet_oracle as O
from sod100k_b200 import compiler, runtime, synth, checkpoints

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
out = {}
for tag in ("csnet-L-x2", "csnet-L-x1"):
    cfg, sd = checkpoints.load_npz(tag)
    sd = {k: torch.from_numpy(v) for k, v in sd.items()}
    xb, _ = synth.blob_images(n, 224, 224, 1235)
    xr = synth.randn_images(n, 224, 224, 1234)
    res = {}
    for name, x in (("blobs", xb), ("randn", xr)):
        xt = torch.from_numpy(x).cuda()
        with torch.no_grad():
            sdc = {k: v.cuda() for k, v in sd.items()}
            ref = torch.sigmoid(O.csnet_forward(cfg, sdc, xt))                       # fp32 ATen on the GPU == the CPU oracle to ~1e-6
            sdh = {k: (v.half() if v.is_floating_point() else v) for k, v in sdc.items()}
            half = torch.sigmoid(O.csnet_forward(cfg, sdh, xt.half()).float())
        r = {"reference_ops_all_half": [(half - ref).abs().max().item(), (half - ref).abs().mean().item()]}
        for label, kw in (("generic_fp16_storage", dict(fuse=False, tensor_core=False)), ("plus_fused_ilblocks", dict(fuse=True, tensor_core=False)),
                          ("full_program", dict())):
            prog = compiler.compile_csnet(cfg, sd, 224, 224, "fp16", **kw)
            p = runtime.Plan(prog, max_batch=n)
            y = torch.sigmoid(p.forward(xt))
            r[label] = [(y - ref).abs().max().item(), (y - ref).abs().mean().item()]
            p.close()
        res[name] = r
        print(tag, name, json.dumps(r), flush=True)
    out[tag] = res
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"n_images": n, "size": 224, "metric": "[max, mean] |sigmoid(y) - sigmoid(fp32 oracle)|", "results": out}, open("gpurun_out/precision_budget.json", "w"), indent=1)

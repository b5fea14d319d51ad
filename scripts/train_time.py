"""Training step timing: batch sweep, eager vs whole-step CUDA graph; optional kernel breakdown (torch.profiler, CUDA activities)."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sod100k_b200 import checkpoints, synth, train_ops
from sod100k_b200.trainer import Trainer

dev = torch.device("cuda", 0)
S = 224
batches = [int(b) for b in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["32", "128"])]
prof = len(sys.argv) > 2 and sys.argv[2] == "prof"
recompute = "recompute" in sys.argv[2:]

def timed(fn, k):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time(); e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k, (time.time() - t0) * 1e3 / k

for B in batches:
    model, cfg, _ = checkpoints.build_from_npz("csnet-L-x2")
    model.cuda(0)
    tr = Trainer(model, lr=1e-4, weight_decay=5e-3, recompute=recompute)
    x = torch.from_numpy(synth.randn_images(B, S, S, 1234)).to(dev)
    t = torch.from_numpy(synth.random_masks(B, S, S, 1236)).to(dev)
    for _ in range(3): tr.step(x, t)
    ms, wall = timed(lambda: tr.step(x, t), 5)
    mem = torch.cuda.max_memory_allocated() / 2**30
    rec = {"batch": B, "recompute": recompute, "eager_ms": ms, "eager_img_s": B / ms * 1e3, "max_mem_GiB": mem}
    if recompute:
        print(json.dumps(rec), flush=True)
        del tr, model, x, t
        torch.cuda.empty_cache(); torch.cuda.reset_peak_memory_stats()
        continue
    if prof:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as p:
            tr.step(x, t); torch.cuda.synchronize()
        ev = sorted(p.key_averages(), key=lambda e: -e.device_time_total)
        tot = sum(e.device_time_total for e in ev)
        rec["gpu_busy_ms"] = tot / 1e3
        for e in ev[:25]:
            print(f"  {e.key[:90]:90s} {e.count:5d} {e.device_time_total/1e3:9.3f} ms", file=sys.stderr)
    try:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            tr.step(x, t)
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            loss = tr.step(x, t)
        for _ in range(2): g.replay()
        gms, _ = timed(g.replay, 5)
        rec.update({"graph_ms": gms, "graph_img_s": B / gms * 1e3, "loss": float(loss)})
    except Exception as e:
        rec["graph_error"] = f"{type(e).__name__}: {e}"[:300]
    print(json.dumps(rec), flush=True)
    del tr, model, x, t
    torch.cuda.empty_cache(); torch.cuda.reset_peak_memory_stats()

"""Time forward_host / forward_host_u8 at the bench configuration (one process per CSNET_HOST_SPLIT* setting: the env is read once)."""
import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    from sod100k_b200 import synth
    from sod100k_b200.model import csnet
    from tests import fixtures
    cfg, sd = fixtures.checkpoint("csnet-L-x2")
    m = csnet.CSNet(cfg); m.load_state_dict(sd); m = m.cuda().eval(); m.set_precision("fp16")
    B, S = 256, 224
    x = torch.randn(B, 3, S, S).pin_memory(); y = torch.empty(B, 1, S, S).pin_memory()
    x8 = torch.randint(0, 256, (B, S, S, 3), dtype=torch.uint8).pin_memory(); y8 = torch.empty(B, S, S, dtype=torch.uint8).pin_memory()
    eng = m.engine()
    res = {}
    for name, fn in (("f32", lambda: eng.forward_host(x, out=y)), ("u8", lambda: eng.forward_host_u8(x8, out=y8))):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / 10
    print(json.dumps(res))
else:
    for split in ("32,32", "0,0", "16,16", "8,8", "16,0", "0,16", "24,16"):
        env = dict(os.environ, CSNET_HOST_SPLIT=split, CSNET_HOST_SPLIT_U8=split)
        out = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        print(split, out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:], flush=True)

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
        for _ in range(6): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / 20
    print(json.dumps(res))
else:
    settings = [{}, {"CSNET_HOST_SCHED": "16,32,64,112,32"}, {"CSNET_HOST_SCHED": "16,48,160,32"}, {"CSNET_HOST_SCHED": "8,16,32,64,112,24"},
                {"CSNET_HOST_SCHED": "24,72,136,24"}, {"CSNET_HOST_SCHED": "32,64,128,32"}, {"CSNET_HOST_SCHED": "16,32,64,128,16"},
                {"CSNET_HOST_SCHED_U8": "64,192"}, {"CSNET_HOST_SCHED_U8": "32,192,32"}, {"CSNET_HOST_SCHED_U8": "32,224"}, {"CSNET_HOST_SCHED_U8": "224,32"}]
    for st in settings:
        env = dict(os.environ, **st)
        out = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        print(st, out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:], flush=True)

"""Per-primitive timings of the training kernels at the shapes of csnet-L-x2 (bs 256, 224x224): forward, data gradient and weight
gradient of conv paths, depthwise, pooling prep; ms, GB/s of compulsory traffic and GFMA/s."""
import os, sys, json, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sod100k_b200 import train_ops as T

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda", 0)

def timed(fn, k=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k

def conv_case(cin, cout, h, w, k=1, dil=1):
    x = torch.randn(N, cin, h, w, device=dev); wt = torch.randn(cin, k * k, cout, device=dev) * 0.1
    y = torch.empty(N, cout, h, w, device=dev); dx = torch.empty_like(x); dw = torch.empty_like(wt)
    st = torch.cuda.current_stream().cuda_stream
    p = T.TrainPath(x.data_ptr(), wt.data_ptr(), cin, h, w, 0, cin, 0, 1, k, dil, 1, dil * (k // 2), 1, 0, cout)
    arr = (T.TrainPath * 1)(p)
    L = T.lib()
    f = timed(lambda: L.csnet_train_mix_fwd(y.data_ptr(), N, cout, h, w, arr, 1, st))
    d = timed(lambda: L.csnet_train_mix_dgrad(y.data_ptr(), N, cout, h, w, C.byref(p), dx.data_ptr(), st))
    g = timed(lambda: L.csnet_train_mix_wgrad(y.data_ptr(), N, cout, h, w, C.byref(p), dw.data_ptr(), st))
    by = (cin + cout) * N * h * w * 4 / 1e9
    fma = cin * cout * k * k * N * h * w / 1e9
    print(json.dumps({"conv": f"{cin}->{cout} k{k} d{dil} @{h}x{w}", "fwd_ms": round(f, 3), "dgrad_ms": round(d, 3), "wgrad_ms": round(g, 3),
                      "GB": round(by, 2), "fwd_GBs": round(by / f * 1e3), "wgrad_GBs": round(by / g * 1e3), "fwd_GFMAs": round(fma / f * 1e3)}), flush=True)

def dw_case(c, h, w):
    x = torch.randn(N, c, h, w, device=dev); wt = torch.randn(c, 9, device=dev); y = torch.empty_like(x); dw = torch.empty_like(wt)
    st = torch.cuda.current_stream().cuda_stream
    L = T.lib()
    f = timed(lambda: L.csnet_train_dw_conv(x.data_ptr(), wt.data_ptr(), y.data_ptr(), N, c, h, w, 100.0, 0, st))
    g = timed(lambda: L.csnet_train_dw_wgrad(x.data_ptr(), y.data_ptr(), dw.data_ptr(), N, c, h, w, 100.0, st))
    by = 2 * c * N * h * w * 4 / 1e9
    print(json.dumps({"dw": f"{c} @{h}x{w}", "fwd_ms": round(f, 3), "wgrad_ms": round(g, 3), "fwd_GBs": round(by / f * 1e3), "wgrad_GBs": round(by / g * 1e3)}), flush=True)

def pool_case(c, h, w, pool=2):
    x = torch.randn(N, c, h, w, device=dev); y = torch.empty(N, c, h // pool, w // pool, device=dev); idx = torch.empty(y.shape, dtype=torch.uint8, device=dev)
    dx = torch.empty_like(x)
    st = torch.cuda.current_stream().cuda_stream
    L = T.lib()
    f = timed(lambda: L.csnet_train_pool_fwd(x.data_ptr(), N, c, 0, c, h, w, 0, pool, y.data_ptr(), idx.data_ptr(), st))
    b = timed(lambda: L.csnet_train_pool_bwd(y.data_ptr(), idx.data_ptr(), N, c, h, w, 0, pool, dx.data_ptr(), st))
    by = c * N * h * w * 4 / 1e9
    print(json.dumps({"pool": f"{c} @{h}x{w} /{pool}", "fwd_ms": round(f, 3), "bwd_ms": round(b, 3), "fwd_GBs": round(1.25 * by / f * 1e3), "bwd_GBs": round(1.25 * by / b * 1e3)}), flush=True)

def rs_case(c, h, w, up):
    x = torch.randn(N, c, h, w, device=dev); y = torch.empty(N, c, h * up, w * up, device=dev); dx = torch.empty_like(x)
    st = torch.cuda.current_stream().cuda_stream
    p = T.TrainPath(x.data_ptr(), None, c, h, w, 0, c, 0, 1, 0, 1, 1, 0, up, 0, c)
    arr = (T.TrainPath * 1)(p)
    L = T.lib()
    f = timed(lambda: L.csnet_train_mix_fwd(y.data_ptr(), N, c, h * up, w * up, arr, 1, st))
    d = timed(lambda: L.csnet_train_mix_dgrad(y.data_ptr(), N, c, h * up, w * up, C.byref(p), dx.data_ptr(), st))
    by = c * N * h * w * 4 * (1 + up * up) / 1e9
    print(json.dumps({"resample": f"{c} @{h}x{w} x{up}", "fwd_ms": round(f, 3), "dgrad_ms": round(d, 3), "fwd_GBs": round(by / f * 1e3), "dgrad_GBs": round(by / d * 1e3)}), flush=True)

conv_case(18, 18, 224, 224)
conv_case(13, 18, 224, 224)
conv_case(18, 13, 112, 112)
conv_case(34, 31, 112, 112)
conv_case(51, 23, 56, 56)
conv_case(38, 64, 28, 28)
conv_case(64, 64, 14, 14)
conv_case(18, 28, 112, 112, 3)
conv_case(51, 23, 56, 56, 3)
conv_case(64, 64, 28, 28, 3)
conv_case(17, 2, 112, 112, 3, 4)
conv_case(17, 2, 112, 112, 3, 16)
conv_case(38, 6, 56, 56, 3, 8)
dw_case(18, 224, 224)
dw_case(34, 112, 112)
dw_case(64, 14, 14)
pool_case(18, 224, 224)
rs_case(13, 112, 112, 2)
rs_case(17, 28, 28, 4)

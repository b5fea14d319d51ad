"""Drop-in boundary checks that need no GPU: module tree / state_dict contract, initialisation parity,
C-ABI library exports, loud failure without a device."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch
import torch.nn as nn

from sod100k_b200 import runtime
from sod100k_b200.model import csnet
from tests import fixtures

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from sod100k_b200 import build

    build.build()
    return runtime.load_library()


def test_header_symbols_are_exported(lib):
    header = open(os.path.join(ROOT, "include", "csnet_b200.h")).read()
    declared = set(re.findall(r"\b(csnet_[a-z0-9_]+)\s*\(", header))
    assert declared == set(runtime.SYMBOLS)
    for s in declared:
        assert hasattr(lib, s), s
    assert lib.csnet_abi_version() == runtime.ABI_VERSION


def test_ctypes_structs_match_header_layout():
    from sod100k_b200 import ir

    assert ctypes.sizeof(ir.TensorDesc) == 32
    assert ctypes.sizeof(ir.PathDesc) == 56
    assert ctypes.sizeof(ir.OpDesc) == 32 + 8 * 56 + 24 * 8


def test_invalid_program_is_rejected_before_touching_the_device(lib):
    from sod100k_b200 import ir

    t = (ir.TensorDesc * 2)(ir.TensorDesc(3, 16, 16, 0, 0, 0, 0), ir.TensorDesc(4, 16, 16, 0, -1, 0, 0))
    op = ir.OpDesc()
    op.kind, op.dst, op.n_paths, op.bias_off, op.slope_off = ir.OP_MIX, 1, 1, -1, -1
    op.paths[0] = ir.PathDesc(0, 0, 3, 0, 1, 1, 1, 1, 0, 1, 0, 4, 1000)      # weights outside the blob
    h = ctypes.c_void_p()
    rc = lib.csnet_plan_create(ctypes.byref(h), t, 2, (ir.OpDesc * 1)(op), 1, 16, 1, 0)
    assert rc == -1 and b"weights outside blob" in lib.csnet_last_error()


def test_no_gpu_fails_loudly():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cfg, sd = fixtures.checkpoint("csnet-L-x1")
    m = csnet.CSNet(cfg)
    m.load_state_dict(sd)
    m.eval()
    with pytest.raises(runtime.EngineError):
        m(torch.zeros(1, 3, 32, 32))


@pytest.mark.parametrize("tag", ["csnet-L-x2", "csnet-L-x1"])
def test_checkpoints_load_strict(tag):
    cfg, sd = fixtures.checkpoint(tag)
    m = csnet.CSNet(cfg)
    res = m.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert list(m.state_dict().keys()) == list(sd.keys())           # same ORDER as the reference too
    _, meta = fixtures.forward_golden()
    assert sum(p.numel() for p in m.parameters()) == meta[tag]["params"]


def test_same_seed_same_initial_parameters_as_reference():
    z, _ = fixtures.forward_golden()
    torch.manual_seed(0)
    m = csnet.build_model(basic_split=[0.5, 0.5], expand=2.0)
    sd = m.state_dict()
    n = 0
    for k in z.files:
        if k.startswith("seed0-init/"):
            assert np.array_equal(sd[k.split("/", 1)[1]].numpy().reshape(-1)[:256], z[k]), k
            n += 1
    assert n == 6


def test_module_tree_supports_reference_callers():
    m = csnet.build_model(basic_split=[0.5, 0.5], expand=2.0)
    # train.py:101-105 picks BN gammas by name
    picked = [n for n, _ in m.named_parameters()
              if "stage" in n and ("conv1x1.bns" in n or "conv3x3_1.bns" in n) and "weight" in n]
    assert len(picked) == 15 * 4 + 3 * 2
    # finetune / foo_bns walk isinstance targets
    assert sum(isinstance(x, csnet.ILBlock) for x in m.modules()) == 18
    assert sum(isinstance(x, csnet.gOctaveCBR) for x in m.modules()) == 20
    assert sum(isinstance(x, nn.BatchNorm2d) for x in m.modules()) == 106
    # bookkeeping surface
    m.flops_hook(expandflop=1.0)
    m.set_batchsize(24)
    m.clear_flops()
    assert m.get_flops() == 0
    assert not any(x._forward_hooks for x in m.modules())


def test_layer_config_pickle_roundtrip(tmp_path):
    cfg = csnet.init_layers(40, [0.5, 0.5])
    csnet.save_layer_config(cfg, str(tmp_path), 3, latest=True)
    back = csnet.load_layer_config(str(tmp_path / "layer_config_latest.bin"))
    m = csnet.build_model(predefine=str(tmp_path / "layer_config_3.bin"))
    assert back[-1] == cfg[-1] and len(m.state_dict()) > 700


def test_simplesum_reproduces_the_reference_counters():
    """the reference's own counter gives 0.0936 M / 0.4354 G (x1) and 0.1409 M / 0.7167 G (x2)."""
    import contextlib
    import io

    from sod100k_b200 import checkpoints
    from sod100k_b200.model.utils.simplesum_octconv import simplesum

    for tag, params, gflops in (("csnet-L-x1", 93647, 0.4354), ("csnet-L-x2", 140894, 0.7167)):
        m, _, _ = checkpoints.build_from_npz(tag)
        with contextlib.redirect_stdout(io.StringIO()) as out:
            p, f = simplesum(m, inputsize=(3, 224, 224), device=-1)
        assert p == params and abs(f / 1e9 - gflops) < 5e-5
        assert "Number of params" in out.getvalue() and "Number of FLOPs" in out.getvalue()
        assert not any(x._forward_hooks for x in m.modules())


def test_lowering_choices_by_dtype():
    """16-bit programs use the fused forms (ILBlock kernels incl. the stem, cls_layer folded into fuse1x1, pooled /
    up-sampled branches materialised once); fp32 programs — the 1e-3 parity path — keep the reference's op order."""
    from sod100k_b200 import compiler, ir
    from tests import fixtures
    cfg, sd = fixtures.checkpoint("csnet-L-x2")
    p16 = compiler.compile_csnet(cfg, sd, 224, 224, "fp16")
    kinds = [o.kind for o in p16.ops]
    names = [o.name for o in p16.ops]
    assert kinds.count(ir.OP_ILBLOCK) == 12 and names[0] == "stage0.0" and p16.ops[0].paths[0].ksize == 3      # stem form first
    assert kinds.count(ir.OP_MIXPROJ) == 1 and "cls_layer" not in names
    assert sum(".pool" in n for n in names) >= 4 and sum(".maxpool" in n for n in names) >= 5 and sum(".up" in n for n in names) == 2
    computed = [o.name for o in p16.ops if o.kind != ir.OP_ILBLOCK for q in o.paths if q.ksize > 0 and (q.pre_avg or q.pool > 1 or q.up > 1)]
    assert computed == [], "every conv path of a 16-bit program reads a plain tensor"
    assert len(p16.ops) == 74
    p32 = compiler.compile_csnet(cfg, sd, 224, 224, "fp32")
    assert all(o.kind in (ir.OP_MIX, ir.OP_DW) for o in p32.ops) and "cls_layer" in [o.name for o in p32.ops]
    assert not any("pool" in o.name or ".up" in o.name for o in p32.ops)
    # gating by name: only the requested module changes
    p = compiler.compile_csnet(cfg, sd, 224, 224, "fp16", fuse={"stage2.0.conv1x1"}, tensor_core=False)
    pooled = [o.name for o in p.ops if "pool" in o.name]
    assert len(pooled) == 3 and all(n.startswith("stage2.0.conv1x1.") for n in pooled)


def test_module_copies_and_pickles_without_its_engine():
    """The reference module supports copy.deepcopy (EMA) and torch.save(model); the engine's ctypes handles must not travel."""
    import copy
    import ctypes
    import io
    import pickle

    cfg, sd = fixtures.checkpoint("csnet-L-x1")
    m = csnet.CSNet(cfg)
    m.load_state_dict(sd)

    class _FakeEngine:                                   # what a used engine looks like to pickle: a raw pointer inside
        def __init__(self):
            self.h = ctypes.c_void_p(1234)

        def invalidate(self):
            pass

    object.__setattr__(m, "_engine", _FakeEngine())
    m2 = copy.deepcopy(m)
    assert m2._engine is None and m._engine is not None
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))
    buf = io.BytesIO()
    torch.save(m, buf)
    buf.seek(0)
    m3 = torch.load(buf, weights_only=False)
    assert m3._engine is None and list(m3.state_dict().keys()) == list(m.state_dict().keys())
    pickle.dumps(m)


def test_frozen_parameters_take_no_optimizer_step():
    from sod100k_b200 import trainer

    cfg, sd = fixtures.checkpoint("csnet-L-x1")
    m = csnet.CSNet(cfg)
    n_all = sum(1 for _ in m.parameters())
    m.cls_layer.weight.requires_grad_(False)
    normal, picked = trainer.reference_param_groups(m)
    assert len(normal) + len(picked) == n_all - 1
    assert all(p.requires_grad for p in normal + picked)


def test_program_signature_tracks_structure_not_values():
    from sod100k_b200 import compiler

    cfg, sd = fixtures.checkpoint("csnet-L-x1")
    a = compiler.compile_csnet(cfg, sd, 64, 64, "fp16")
    sd2 = {k: (v * 1.5 if k.endswith("prelus.0.weight") else v) for k, v in sd.items()}
    b = compiler.compile_csnet(cfg, sd2, 64, 64, "fp16")
    assert a.signature() == b.signature() and not np.array_equal(a.blob, b.blob)
    c = compiler.compile_csnet(cfg, sd, 64, 64, "fp16", tensor_core=False)       # the kernel veto flags differ
    assert c.signature() != a.signature()

"""Host-side check of compiler + generic kernel bodies (emulated on the CPU) against the oracle."""
import numpy as np
import pytest
import torch

from oracle import csnet_oracle as O
from sod100k_b200 import compiler, synth
from tests import emu, fixtures


def _oracle(cfg, sd, x, taps=None):
    with torch.no_grad():
        return O.csnet_forward(cfg, sd, torch.from_numpy(x), taps=taps).numpy()


@pytest.mark.parametrize("tag,hw", [("csnet-L-x2", (64, 96)), ("csnet-L-x1", (64, 64))])
def test_emulated_program_matches_oracle_checkpoints(tag, hw):
    cfg, sd = fixtures.checkpoint(tag)
    x = synth.randn_images(2, hw[0], hw[1], 11)
    prog = compiler.compile_csnet(cfg, sd, hw[0], hw[1], "fp32", reuse_arena=False)
    names = ["stage1.0/0", "stage1.0/1", "stage2.0/1", "stage2.3/0", "stage4.3/0", "oct_fuse.ms/2", "oct_fuse.fuse1x1/0"]
    y, got = emu.run(prog, x, names)
    taps = {}
    ref = _oracle(cfg, sd, x, taps)
    for n in names:
        blk, b = n.rsplit("/", 1)
        r = taps[blk][int(b)].numpy()
        assert np.abs(got[n] - r).max() <= 1e-4 * max(1.0, np.abs(r).max()), n
    assert np.abs(y - ref).max() <= 1e-4
    # arena reuse must not change the result
    prog2 = compiler.compile_csnet(cfg, sd, hw[0], hw[1], "fp32", reuse_arena=True)
    assert prog2.arena_bytes_per_image < prog.arena_bytes_per_image
    y2, _ = emu.run(prog2, x)
    assert np.array_equal(y, y2)


@pytest.mark.parametrize("tag", ["init-std", "init-3br"])
def test_emulated_program_unpruned_architectures(tag):
    cfg, sd, m = fixtures.synthetic_model(tag)
    h, w = m["hw"]
    x = synth.randn_images(1, h, w, 1240 + m["seed"])
    prog = compiler.compile_csnet(cfg, sd, h, w, "fp32")
    y, _ = emu.run(prog, x)
    z, _ = fixtures.forward_golden()
    ref = z[f"{tag}/randn"]
    assert np.abs(y - ref).max() <= 2e-4 * max(1.0, np.abs(ref).max())


def test_size_must_be_multiple_of_16():
    cfg, sd = fixtures.checkpoint("csnet-L-x1")
    with pytest.raises(ValueError):
        compiler.compile_csnet(cfg, sd, 100, 100)


def test_emulated_input_side_upsample_paths():
    """1x1 up-paths lowered as `conv(up(x))` (what 16-bit programs use) equal the reference's `up(conv(x))`."""
    cfg, sd = fixtures.checkpoint("csnet-L-x2")
    x = synth.randn_images(1, 64, 64, 12)
    prog = compiler.compile_csnet(cfg, sd, 64, 64, "fp32", upsample_inputs=True)
    assert any(q.ksize == 1 and q.up > 1 for o in prog.ops for q in o.paths)
    y, _ = emu.run(prog, x)
    assert np.abs(y - _oracle(cfg, sd, x)).max() <= 1e-4


def test_emulated_resample_paths_pool_upsample_and_copy():
    """A `ksize = 0` path adds its source avg-pooled (pre_avg), max-pooled (pool), bilinearly up-sampled (up) or as it is:
    the single-path MIX ops the compiler emits for materialised branches (generic_ops.cuh: mix_finish).  fp32 program on
    the host emulation against torch."""
    import torch
    import torch.nn.functional as F
    from sod100k_b200 import ir
    b = ir.Builder()
    C_, H, W = 3, 16, 24
    x = b.tensor(C_, H, W, ir.F32, external=0, name="x")
    outs = {
        "avg": (b.tensor(C_, H // 2, W // 2, ir.F32, external=1), ir.Path(x, C_, C_, ksize=0, pre_avg=1)),
        "max": (b.tensor(C_, H // 2, W // 2, ir.F32, external=2), ir.Path(x, C_, C_, ksize=0, pool=2)),
        "max4": (b.tensor(C_, H // 4, W // 4, ir.F32, external=3), ir.Path(x, C_, C_, ksize=0, pool=4)),
        "up2": (b.tensor(C_, H * 2, W * 2, ir.F32, external=4), ir.Path(x, C_, C_, ksize=0, up=2)),
        "copy": (b.tensor(C_, H, W, ir.F32, external=5), ir.Path(x, C_, C_, ksize=0)),
    }
    for name, (t, path) in outs.items():
        b.op(ir.OP_MIX, t, [path], name=name)
    prog = b.finish(reuse=False)
    rng = np.random.default_rng(5)
    xv = rng.standard_normal((2, C_, H, W)).astype(np.float32)
    arrays = [xv] + [np.zeros((2, prog.tensors[t].C, prog.tensors[t].H, prog.tensors[t].W), np.float32) for t, _ in outs.values()]
    emu.run_ext(prog, arrays, 2)
    xt = torch.from_numpy(xv)
    want = [F.avg_pool2d(xt, 2), F.max_pool2d(xt, 2), F.max_pool2d(xt, 4),
            F.interpolate(xt, scale_factor=2, mode="bilinear", align_corners=False), xt]
    for (name, _), got, ref in zip(outs.items(), arrays[1:], want):
        assert np.abs(got - ref.numpy()).max() <= 1e-6, name

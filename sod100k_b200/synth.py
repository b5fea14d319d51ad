"""Seeded synthetic inputs shared by bench.py, the tests and the golden-fixture script.

SURVEY.md §8(d): images are ImageNet-normalised-range noise (`randn`) for throughput and a
structured "blobs" set for F-measure (pure noise gives an all-background prediction).  Everything is
generated with numpy's PCG64 `Generator`, whose stream is stable across numpy versions, so fixtures
store the *seed*, not the tensor.
"""
from __future__ import annotations

import numpy as np

IMAGENET_MEAN = (0.485, 0.456, 0.406)   # /root/reference/CSNet/test.py:68
IMAGENET_STD = (0.229, 0.224, 0.225)    # /root/reference/CSNet/test.py:69


def randn_images(n: int, h: int, w: int, seed: int = 1234) -> np.ndarray:
    """float32 [n,3,h,w] standard-normal images."""
    rng = np.random.default_rng(seed)
    return rng.standard_normal((n, 3, h, w), dtype=np.float32)


def blob_images(n: int, h: int, w: int, seed: int = 1235):
    """Bright Gaussian blobs on a dark smooth background.

    Returns (images float32 [n,3,h,w] already mean/std normalised, masks float32 [n,1,h,w] in {0,1}).
    The mask is the blob support (intensity above half maximum), used as synthetic ground truth.
    """
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.arange(h, dtype=np.float32), np.arange(w, dtype=np.float32), indexing="ij")
    imgs = np.empty((n, 3, h, w), np.float32)
    masks = np.zeros((n, 1, h, w), np.float32)
    mean = np.asarray(IMAGENET_MEAN, np.float32)[:, None, None]
    std = np.asarray(IMAGENET_STD, np.float32)[:, None, None]
    for i in range(n):
        # smooth dark background: low-frequency cosine texture in [0.05, 0.25]
        fx, fy = rng.uniform(0.5, 2.0, 2)
        ph = rng.uniform(0, 2 * np.pi, 2)
        bg = 0.15 + 0.10 * np.cos(2 * np.pi * fx * xx / w + ph[0]) * np.cos(2 * np.pi * fy * yy / h + ph[1])
        rgb = np.stack([bg * s for s in rng.uniform(0.7, 1.0, 3)]).astype(np.float32)
        for _ in range(int(rng.integers(1, 3))):
            cy, cx = rng.uniform(0.25, 0.75) * h, rng.uniform(0.25, 0.75) * w
            sy, sx = rng.uniform(0.08, 0.2) * h, rng.uniform(0.08, 0.2) * w
            g = np.exp(-0.5 * (((yy - cy) / sy) ** 2 + ((xx - cx) / sx) ** 2)).astype(np.float32)
            col = rng.uniform(0.6, 1.0, 3).astype(np.float32)[:, None, None]
            rgb = np.maximum(rgb, g[None] * col)
            masks[i, 0] = np.maximum(masks[i, 0], (g > 0.5).astype(np.float32))
        imgs[i] = (np.clip(rgb, 0.0, 1.0) - mean) / std
    return imgs, masks


def random_masks(n: int, h: int, w: int, seed: int = 1236) -> np.ndarray:
    """float32 [n,1,h,w] Bernoulli(0.5) masks (throughput-only training targets)."""
    rng = np.random.default_rng(seed)
    return (rng.random((n, 1, h, w), dtype=np.float32) > 0.5).astype(np.float32)


def synth_state(shapes: dict, seed: int = 0) -> dict:
    """Seeded synthetic parameters for a CSNet of arbitrary layer_config.

    `shapes` maps state_dict key -> shape.  Conv weights ~ U(-b, b) with b = 1/sqrt(fan_in)
    (divided by 100 for Conv2dX100 weights, which the forward multiplies by 100); BN gamma ~ U(0.5,1.5),
    beta ~ N(0,0.1), running_mean ~ N(0,0.1), running_var ~ U(0.5,1.5); PReLU slope ~ U(0.1,0.4).
    Returns numpy arrays (float32; int64 for num_batches_tracked).
    """
    rng = np.random.default_rng(seed)
    out = {}
    for key in sorted(shapes):
        shp = tuple(shapes[key])
        if key.endswith("num_batches_tracked"):
            out[key] = np.zeros(shp, np.int64)
        elif ".bns." in key or ".bn." in key:
            if key.endswith("running_var") or key.endswith(".weight"):
                out[key] = rng.uniform(0.5, 1.5, shp).astype(np.float32)
            else:
                out[key] = (0.1 * rng.standard_normal(shp)).astype(np.float32)
        elif ".prelus." in key or ".prelu." in key:
            out[key] = rng.uniform(0.1, 0.4, shp).astype(np.float32)
        elif key == "cls_layer.bias":
            out[key] = (0.1 * rng.standard_normal(shp)).astype(np.float32)
        else:  # conv weight [cout, cin/groups, k, k]
            fan_in = int(np.prod(shp[1:]))
            b = 1.0 / np.sqrt(fan_in)
            w = rng.uniform(-b, b, shp).astype(np.float32)
            x100 = (".convs." in key) or (".msconv." in key)   # depthwise / dilated Conv2dX100 weights
            out[key] = (w / 100.0).astype(np.float32) if x100 else w
    return out


def synth_state_r(shapes: dict, seed: int = 0) -> dict:
    """Seeded synthetic parameters for CSF+Res2Net (config 5): the reference ships no weights for it.  Conv weights
    ~ N(0, 1/fan_in), norm gammas ~ U(0.5, 1.5) (the last BN of every residual branch ~ U(0.1, 0.3) so activations stay
    O(1) through 16 residual blocks), betas / means ~ N(0, 0.1), variances ~ U(0.5, 1.5), PReLU slopes ~ U(0.1, 0.4)."""
    rng = np.random.default_rng(seed)
    out = {}
    for key in sorted(shapes):
        shp = tuple(shapes[key])
        if key.endswith("num_batches_tracked"):
            out[key] = np.zeros(shp, np.int64)
        elif key.endswith("running_var"):
            out[key] = rng.uniform(0.5, 1.5, shp).astype(np.float32)
        elif key.endswith("running_mean"):
            out[key] = (0.1 * rng.standard_normal(shp)).astype(np.float32)
        elif ".prelu" in key:
            out[key] = rng.uniform(0.1, 0.4, shp).astype(np.float32)
        elif len(shp) == 1 and key.endswith(".weight"):
            lo, hi = (0.1, 0.3) if ".bn3." in key else (0.5, 1.5)
            out[key] = rng.uniform(lo, hi, shp).astype(np.float32)
        elif len(shp) == 1:
            out[key] = (0.1 * rng.standard_normal(shp)).astype(np.float32)
        else:
            fan_in = int(np.prod(shp[1:]))
            out[key] = (rng.standard_normal(shp) / np.sqrt(fan_in)).astype(np.float32)
    return out

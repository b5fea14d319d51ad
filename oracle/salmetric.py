"""TEST INFRASTRUCTURE — numpy restatement of the reference's SalMetric (F-measure / MAE).

Follows /root/reference/CSNet_training/SalMetric/src/sal_metric.cpp:
  * compute_mae (:86-97): mean |sal - gt| / 255 over the 8-bit maps;
  * compute_precision_and_recall (:99-120): for th in 0..255, a = sal > th, b = gt > 128,
    pre = (|a&b| + 1e-4) / (|a| + 1e-4), rec = (|a&b| + 1e-4) / (|b| + 1e-4), accumulated PER IMAGE;
  * do_evaluation (:164-185): per-threshold means over images, F = 1.3*P*R / (0.3*P + R),
    report max-F (and its argmax P/R), mean-F, mean P/R, MAE.
Constants: sal_metric.hpp:50-52 (THRESHOLDS 256, EPSILON 1e-4, BETA 0.3).
PINNED against the reference's own code: oracle/build_ref.py compiles the UNMODIFIED sal_metric.cpp (from where it lies under
/root/reference, with oracle/cvshim standing in for the three OpenCV headers: cv::Mat + a PGM imread) into oracle/_ref/salmetric;
tests/golden/salmetric_ref.json holds its reports on seeded maps (tests/golden/make_salmetric_golden.py) and
tests/test_salmetric.py compares this restatement with them (and with the live binary when present), next to the loop-for-loop
restatement and the hand-computed cases.
Also provides the reference's png quantisation: (sigmoid * 255).astype(uint8) (CSNet/test.py:94-96).
"""
from __future__ import annotations

import numpy as np

THRESHOLDS = 256
EPSILON = np.float32(1e-4)
BETA = np.float32(0.3)


def quantise(prob: np.ndarray) -> np.ndarray:
    """CSNet/test.py:94-96 — float saliency in [0,1] -> uint8 (truncation, as numpy astype does)."""
    return (np.asarray(prob, np.float64) * 255).astype(np.uint8)


def precision_recall(sal_u8: np.ndarray, gt_u8: np.ndarray):
    """Per-image precision[256], recall[256] (sal_metric.cpp:99-120)."""
    sal = sal_u8.reshape(-1).astype(np.int64)
    b = gt_u8.reshape(-1) > THRESHOLDS // 2
    hist_all = np.bincount(sal, minlength=256)
    hist_pos = np.bincount(sal[b], minlength=256)
    # a = sal > th  ->  count of values strictly above th = suffix sum from th+1
    a_sum = (hist_all[::-1].cumsum()[::-1] - hist_all).astype(np.float32)
    ab = (hist_pos[::-1].cumsum()[::-1] - hist_pos).astype(np.float32)
    b_sum = np.float32(b.sum())
    return (ab + EPSILON) / (a_sum + EPSILON), (ab + EPSILON) / (b_sum + EPSILON)


def mae(sal_u8: np.ndarray, gt_u8: np.ndarray) -> float:
    """sal_metric.cpp:86-97."""
    return float(np.mean(np.abs(sal_u8.astype(np.float32) - gt_u8.astype(np.float32)) / 255.0))


def evaluate(sal_maps, gt_maps) -> dict:
    """sal_metric.cpp:122-197 over lists of uint8 maps."""
    n = len(sal_maps)
    p = np.zeros(THRESHOLDS, np.float64)
    r = np.zeros(THRESHOLDS, np.float64)
    m = 0.0
    for s, g in zip(sal_maps, gt_maps):
        pi, ri = precision_recall(s, g)
        p += pi
        r += ri
        m += mae(s, g)
    p, r, m = p / n, r / n, m / n
    f = ((1 + BETA) * p * r) / (BETA * p + r)
    k = int(np.argmax(f))
    return dict(max_f=float(f[k]), mean_f=float(f.mean()), precision=float(p[k]), recall=float(r[k]),
                mean_precision=float(p.mean()), mean_recall=float(r.mean()), mae=float(m), argmax=k)

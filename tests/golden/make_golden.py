"""Regenerates tests/golden/*.npz from the reference's own known-answer tests and from the compiled reference
(oracle/_ref/libccv_ref.so).  Run in the build container (needs /root/reference built through `make -C oracle`):

    python tests/golden/make_golden.py

Two kinds of fixtures:
  literal_*   inputs + expected outputs typed from the reference's unit tests (test/unit/nnc/gemm.tests.c etc.); these
              pin the oracle itself.
  cpuref_*    seeded inputs + the outputs CCV_NNC_BACKEND_CPU_REF produced for them here; these travel to the GPU box
              so parity can be checked even without the compiled reference.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ccv_b200 import nnc  # noqa: E402
from oracle import ref  # noqa: E402
from tests.util import seeded  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def literal_gemm():
    """test/unit/nnc/gemm.tests.c:13-200 (forward cases, hand-computed expectations)."""
    a = np.array([[1, 2], [3, 4], [5, 6], [7, 8]], np.float32)
    b = np.array([[7, 8, 9], [10, 11, 12]], np.float32)
    c = a.astype(np.float64) @ b.astype(np.float64)
    bias = np.array([22, 23, 24], np.float32)
    return dict(a=a, b=b, c=c.astype(np.float32), bt=np.ascontiguousarray(b.T), at=np.ascontiguousarray(a.T)[None], bias=bias,
                c_bias=(c + bias).astype(np.float32))


def cpuref_cases():
    cases = {}
    # GEMM forward NT with bias (dense layer form) and backward
    a, w, bias = seeded((12, 20), 1, -1, 1), seeded((8, 20), 2, -1, 1), seeded((8,), 3)
    b = np.zeros((12, 8), np.float32)
    assert ref.run(nnc.CMD_GEMM_FORWARD((0, 0), (0, 1)), None, 0, [a, w, bias], [b]) == 0
    g = seeded((12, 8), 4, -1, 1)
    h, dw, db = np.zeros_like(a), np.zeros_like(w), np.zeros_like(bias)
    assert ref.run(nnc.CMD_GEMM_BACKWARD((0, 0), (0, 1)), None, 0, [g, a, w], [h, dw, db]) == 0
    cases["gemm"] = dict(a=a, w=w, bias=bias, b=b, g=g, h=h, dw=dw, db=db)
    # convolution 3x3 stride 2 pad 1 with bias, forward + backward (NHWC)
    x, wt, cb = seeded((2, 9, 11, 8), 5), seeded((12, 3, 3, 8), 6) / 72, seeded((12,), 7)
    y = np.zeros((2, 5, 6, 12), np.float32)
    hint = nnc.hint((2, 2), (1, 1))
    assert ref.run(nnc.CMD_CONVOLUTION_FORWARD(1, 12, 3, 3, 8), hint, 0, [x, wt, cb], [y]) == 0
    gy = seeded(y.shape, 8)
    gx, gw, gb = np.zeros_like(x), np.zeros_like(wt), np.zeros_like(cb)
    assert ref.run(nnc.CMD_CONVOLUTION_BACKWARD(1, 12, 3, 3, 8), hint, 0, [gy, x, wt], [gx, gw, gb]) == 0
    cases["conv"] = dict(x=x, w=wt, bias=cb, y=y, gy=gy, gx=gx, gw=gw, gb=gb)
    # batch norm training forward
    bx = seeded((4, 5, 6, 16), 9, -1, 1)
    scale, bb = seeded((1, 1, 1, 16), 10), seeded((1, 1, 1, 16), 11)
    mean, var = seeded((1, 1, 1, 16), 12), seeded((1, 1, 1, 16), 13)
    mean0, var0 = mean.copy(), var.copy()
    by, sm, sis = np.zeros_like(bx), np.zeros_like(mean), np.zeros_like(mean)
    assert ref.run(nnc.CMD_BATCH_NORM_FORWARD(1e-4, 0, 0.9), None, 0, [bx, scale, bb, mean, var], [by, mean, var, sm, sis]) == 0
    cases["bnorm"] = dict(x=bx, scale=scale, bias=bb, mean0=mean0, var0=var0, y=by, mean=mean, var=var, saved_mean=sm, saved_inv_std=sis)
    # float -> half conversion bit patterns (truncating, lib/ccv_util.c:1434-1440)
    f = np.concatenate([seeded((4096,), 14, -70000, 70000), seeded((4096,), 15, -1e-4, 1e-4), np.array([0.0, -0.0, 65504.0, 65520.0, 1e-8, 6.1e-5, 5.9e-8, np.inf, -np.inf], np.float32)]).astype(np.float32)
    cases["f2h"] = dict(f=f, h=ref.float_to_half(f))
    return cases


if __name__ == "__main__":
    np.savez(os.path.join(OUT, "literal_gemm.npz"), **literal_gemm())
    for name, d in cpuref_cases().items():
        np.savez(os.path.join(OUT, "cpuref_%s.npz" % name), **d)
    print("wrote", sorted(os.listdir(OUT)))

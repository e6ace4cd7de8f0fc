"""Pins the oracle (CPU only): the plain-C restatement oracle/nnc_port.c and the compiled reference oracle/_ref are
checked against the reference's literal known-answer vectors (tests/golden/literal_*.npz, typed from
test/unit/nnc/gemm.tests.c), against each other on seeded inputs, and against the committed CPU_REF fixtures."""
import os

import numpy as np
import pytest

from ccv_b200 import nnc
from tests.util import assert_close, ref_exec, seeded

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def port():
    from oracle import port as p
    if not p.available():
        pytest.skip("oracle/_ref/libnnc_port.so not built (make -C oracle port)")
    return p


def test_port_matches_reference_literal_gemm_vectors(port):
    g = np.load(os.path.join(GOLDEN, "literal_gemm.npz"))
    assert np.array_equal(port.gemm_forw(g["a"], g["b"], None, 4, 3, 2, 0, 0), g["c"])
    assert np.array_equal(port.gemm_forw(g["a"], g["bt"], None, 4, 3, 2, 0, 1), g["c"])
    assert np.array_equal(port.gemm_forw(g["at"][0], g["bt"], None, 4, 3, 2, 1, 1), g["c"])
    assert np.array_equal(port.gemm_forw(g["a"], g["b"], g["bias"], 4, 3, 2, 0, 0), g["c_bias"])


@pytest.mark.ref
def test_compiled_reference_matches_its_own_literal_gemm_vectors(ref):
    g = np.load(os.path.join(GOLDEN, "literal_gemm.npz"))
    c = np.zeros((4, 3), np.float32)
    assert ref_exec(ref, nnc.CMD_GEMM_FORWARD(), None, 0, [g["a"], g["b"]], [c])[0] == 0
    assert np.array_equal(c, g["c"])
    assert ref_exec(ref, nnc.CMD_GEMM_FORWARD((0, 0), (0, 1)), None, 0, [g["a"], g["bt"], g["bias"]], [c])[0] == 0
    assert np.array_equal(c, g["c_bias"])


def test_port_matches_committed_cpu_ref_fixtures(port):
    g = np.load(os.path.join(GOLDEN, "cpuref_gemm.npz"))
    assert_close(port.gemm_forw(g["a"], g["w"], g["bias"], 12, 8, 20, 0, 1), g["b"], 1e-6, "gemm forward")
    h, dw, db = port.gemm_back(g["g"], g["a"], g["w"], 12, 8, 20, 0, 1)
    for a, n in ((h, "h"), (dw, "dw"), (db, "db")):
        assert_close(a, g[n], 1e-6, n)
    c = np.load(os.path.join(GOLDEN, "cpuref_conv.npz"))
    d = port.conv_desc(2, 9, 11, 8, 12, 3, 3, 2, 1)
    assert_close(port.conv_forw(d, c["x"], c["w"], c["bias"]), c["y"], 1e-6, "conv forward")
    gx, gw, gb = port.conv_back(d, c["gy"], c["x"], c["w"])
    assert_close(gx, c["gx"], 1e-6, "dgrad"), assert_close(gw, c["gw"], 1e-6, "wgrad"), assert_close(gb, c["gb"], 1e-6, "dbias")
    b = np.load(os.path.join(GOLDEN, "cpuref_bnorm.npz"))
    mean, var = b["mean0"].reshape(-1).copy(), b["var0"].reshape(-1).copy()
    y, sm, sis = port.bnorm_forw_train(b["x"], b["scale"].reshape(-1), b["bias"].reshape(-1), mean, var, 1e-4, 0.9)
    assert_close(y, b["y"], 1e-5, "bn y"), assert_close(mean, b["mean"].reshape(-1), 1e-6, "bn mean"), assert_close(var, b["var"].reshape(-1), 1e-5, "bn var")
    assert_close(sis, b["saved_inv_std"].reshape(-1), 1e-5, "bn inv std")
    f = np.load(os.path.join(GOLDEN, "cpuref_f2h.npz"))
    assert np.array_equal(port.float_to_half(f["f"]), f["h"])


@pytest.mark.ref
def test_port_matches_compiled_reference_on_seeded_inputs(port, ref):
    # GEMM NN with transposed a
    a, w = seeded((48, 33), 1, -1, 1), seeded((48, 21), 2, -1, 1)
    b = np.zeros((33, 21), np.float32)
    assert ref_exec(ref, nnc.CMD_GEMM_FORWARD((0, 1), (0, 0)), None, 0, [a, w], [b])[0] == 0
    assert_close(port.gemm_forw(a, w, None, 33, 21, 48, 1, 0), b, 1e-6, "gemm TN")
    # dilated convolution
    x, wt = seeded((1, 15, 15, 6), 3), seeded((10, 3, 3, 6), 4) / 54
    d = port.conv_desc(1, 15, 15, 6, 10, 3, 3, 1, 2, 2)
    y = np.zeros((1, 15, 15, 10), np.float32)
    hint = nnc.hint((1, 1), (2, 2))
    assert ref_exec(ref, nnc.CMD_CONVOLUTION_FORWARD(1, 10, 3, 3, 6, (2, 2)), hint, 0, [x, wt], [y])[0] == 0
    assert_close(port.conv_forw(d, x, wt, None), y, 1e-6, "dilated conv")
    # pooling (one image: CPU_REF ignores the rest of a batch)
    img = seeded((9, 9, 5), 5, -1, 1)
    for kind, fwd in (("max", nnc.CMD_MAX_POOL_FORWARD(3, 3)), ("avg", nnc.CMD_AVERAGE_POOL_FORWARD(3, 3))):
        out = np.zeros((5, 5, 5), np.float32)
        assert ref_exec(ref, fwd, nnc.hint((2, 2), (1, 1)), 0, [img], [out])[0] == 0
        p = port.pool_desc(1, 9, 9, 5, 3, 2, 1)
        got = (port.max_pool_forw if kind == "max" else port.avg_pool_forw)(p, img[None])[0]
        assert_close(got, out, 1e-6, kind + " pool")
    # softmax + sgd
    s = seeded((7, 100), 6, -3, 3)
    o = np.zeros_like(s)
    assert ref_exec(ref, nnc.CMD_SOFTMAX_FORWARD(), None, 0, [s], [o])[0] == 0
    assert_close(port.softmax_forw(s), o, 1e-6, "softmax")
    g, a2, m = seeded((5, 6), 7, -1, 1), seeded((5, 6), 8, -1, 1), seeded((5, 6), 9, -1, 1)
    bo, no = np.zeros_like(a2), np.zeros_like(a2)
    assert ref_exec(ref, nnc.CMD_SGD_FORWARD(0, 0.1, 0.5, 0.01, 0.9, 0.2), None, 0, [g, a2, m], [bo, no])[0] == 0
    pb, pn = port.sgd(g, a2, m, 0, 0.1, 0.5, 0.01, 0.9, 0.2)
    assert_close(pb, bo, 1e-6, "sgd b"), assert_close(pn, no, 1e-6, "sgd n")
    # half conversion: bit-exact against the reference's table method
    f = np.concatenate([seeded((20000,), 10, -70000, 70000), seeded((20000,), 11, -1e-4, 1e-4)]).astype(np.float32)
    assert np.array_equal(port.float_to_half(f), ref.float_to_half(f))


@pytest.mark.ref
def test_committed_fixtures_are_what_the_reference_produces(ref):
    """tests/golden/cpuref_*.npz must be reproducible from the compiled reference (guards against stale fixtures)."""
    g = np.load(os.path.join(GOLDEN, "cpuref_gemm.npz"))
    b = np.zeros_like(g["b"])
    assert ref_exec(ref, nnc.CMD_GEMM_FORWARD((0, 0), (0, 1)), None, 0, [g["a"], g["w"], g["bias"]], [b])[0] == 0
    assert np.array_equal(b, g["b"])

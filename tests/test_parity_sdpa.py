"""SCALED_DOT_PRODUCT_ATTENTION forward / backward vs the reference's CPU_REF.  Shape trials follow
test/int/nnc/cublas.tests.c:2752-2953 (GQA Hq/Hk = 8/2, head dims 40/64/128, causal, inputs i / count ramps, tolerance
3e-3 absolute there; relative 2e-3 of max|ref| here) plus an additive mask case from test/unit/nnc/attention.tests.c."""
import numpy as np
import pytest

from ccv_b200 import abi, nnc as _nnc
from tests.util import assert_close, gpu_exec, ref_exec, seeded

pytestmark = [pytest.mark.gpu, pytest.mark.ref]


def _cmd(cmd_id, scale, causal):
    c = _nnc._simple(cmd_id)
    s = c.info.scaled_dot_product_attention
    s.scale, s.is_causal = scale, causal
    return c


TRIALS = [
    # B, Sq, Sk, Hq, Hk, D, causal
    (2, 32, 32, 4, 4, 64, 0), (2, 32, 48, 8, 2, 64, 0), (1, 40, 40, 8, 8, 40, 1), (2, 24, 36, 4, 2, 128, 1), (1, 128, 128, 2, 2, 128, 0),
]


@pytest.mark.parametrize("B,Sq,Sk,Hq,Hk,D,causal", TRIALS)
def test_sdpa_forward_backward_vs_cpu_ref(gpu, ref, B, Sq, Sk, Hq, Hk, D, causal):
    nnc = gpu
    scale = 1.0 / np.sqrt(D)
    q, k, v = seeded((B, Sq, Hq, D), 1, -1, 1), seeded((B, Sk, Hk, D), 2, -1, 1), seeded((B, Sk, Hk, D), 3, -1, 1)
    fwd = _cmd(abi.CCV_NNC_SCALED_DOT_PRODUCT_ATTENTION_FORWARD, scale, causal)
    mk = lambda: [np.zeros((B, Sq, Hq, D), np.float32), None]
    st_r, (o_r, _) = ref_exec(ref, fwd, None, 0, [q, k, v], mk())
    st_g, (o_g, _) = gpu_exec(nnc, fwd, None, 0, [q, k, v], mk())
    assert st_r == 0 and st_g == 0
    assert_close(o_g, o_r, 2e-3, "attention output")
    g = seeded((B, Sq, Hq, D), 4, -1, 1)
    bwd = _cmd(abi.CCV_NNC_SCALED_DOT_PRODUCT_ATTENTION_BACKWARD, scale, causal)
    ins = [g, None, None, q, k, v]
    mk = lambda: [np.zeros_like(q), np.zeros_like(k), np.zeros_like(v)]
    st_r, (dq_r, dk_r, dv_r) = ref_exec(ref, bwd, None, 0, ins, mk())
    st_g, (dq_g, dk_g, dv_g) = gpu_exec(nnc, bwd, None, 0, ins, mk())
    assert st_r == 0 and st_g == 0
    assert_close(dq_g, dq_r, 3e-3, "dq"), assert_close(dk_g, dk_r, 3e-3, "dk"), assert_close(dv_g, dv_r, 3e-3, "dv")


def test_sdpa_with_additive_mask(gpu, ref):
    nnc = gpu
    B, S, H, D = 2, 24, 4, 32
    q, k, v = seeded((B, S, H, D), 1, -1, 1), seeded((B, S, H, D), 2, -1, 1), seeded((B, S, H, D), 3, -1, 1)
    mask = np.where(np.random.RandomState(5).rand(1, 1, S, S) < 0.3, -1e9, 0.0).astype(np.float32)
    mask[..., np.arange(S), np.arange(S)] = 0  # never mask a whole row
    fwd = _cmd(abi.CCV_NNC_SCALED_DOT_PRODUCT_ATTENTION_FORWARD, 1.0 / np.sqrt(D), 0)
    st_r, (o_r, _) = ref_exec(ref, fwd, None, 0, [q, k, v, mask], [np.zeros_like(q), None])
    st_g, (o_g, _) = gpu_exec(nnc, fwd, None, 0, [q, k, v, mask], [np.zeros_like(q), None])
    assert st_r == 0 and st_g == 0
    assert_close(o_g, o_r, 2e-3, "masked attention")

"""Prints the error of the fused 16-bit SDPA backward (dq, dk, dv separately) against CPU_REF for a few shapes: a diagnosis aid for
ccv_b200/csrc/sm100_fmha_bwd.cu (the parity test proper is tests/test_parity_sdpa.py::test_attention_16bit_backward_fused)."""
import sys
import numpy as np
sys.path.insert(0, ".")
sys.path.insert(0, "oracle")
from ccv_b200 import abi, nnc
from oracle import ref as oref
from tests.util import ref_exec, seeded, to_bf16, from_bf16


def run(B, Sq, Sk, Hq, Hk, causal, saved):
    D = 128
    scale = 1.0 / np.sqrt(D)
    arrs = [seeded((B, Sq, Hq, D), 4, -1, 1), None, None, seeded((B, Sq, Hq, D), 1, -1, 1), seeded((B, Sk, Hk, D), 2, -1, 1), seeded((B, Sk, Hk, D), 3, -1, 1)]
    bits = [None if a is None else to_bf16(a) for a in arrs]
    vals = [None if b is None else from_bf16(b) for b in bits]
    bwd = nnc._simple(abi.CCV_NNC_SCALED_DOT_PRODUCT_ATTENTION_BACKWARD)
    bwd.info.scaled_dot_product_attention.scale, bwd.info.scaled_dot_product_attention.is_causal = scale, causal
    st, wants = ref_exec(oref, bwd, None, 0, vals, [np.zeros_like(vals[3]), np.zeros_like(vals[4]), np.zeros_like(vals[5])])
    stream = nnc.Stream(0)
    ins = [None if b is None else nnc.gpu_tensor(list(b.shape), datatype=abi.CCV_16BF).upload(b) for b in bits]
    if saved:
        to, tl = nnc.gpu_tensor([B, Sq, Hq, D], datatype=abi.CCV_16BF), nnc.gpu_tensor([B, Hq, Sq])
        fwd = nnc._simple(abi.CCV_NNC_SCALED_DOT_PRODUCT_ATTENTION_FORWARD)
        fwd.info.scaled_dot_product_attention.scale, fwd.info.scaled_dot_product_attention.is_causal = scale, causal
        assert nnc.cmd_exec(fwd, None, 0, [ins[3], ins[4], ins[5]], [to, tl], stream) == 0
        ins = ins + [None, None, None, to, tl]
    outs = [nnc.gpu_tensor(list(vals[i].shape), datatype=abi.CCV_16BF) for i in (3, 4, 5)]
    n0 = nnc.launch_count()
    rc = nnc.cmd_exec(bwd, None, 0, ins, outs, stream)
    stream.wait()
    line = "B%d Sq%d Sk%d H%d/%d causal%d saved%d rc=%d launches=%d" % (B, Sq, Sk, Hq, Hk, causal, saved, rc, nnc.launch_count() - n0)
    for t, want, name in zip(outs, wants, ("dq", "dk", "dv")):
        got = from_bf16(t.download())
        err = np.abs(got - want)
        worst = np.unravel_index(np.argmax(err), err.shape)
        line += " | %s err %.3e of max %.3e at %s nan=%d" % (name, err.max(), np.abs(want).max(), worst, int(np.isnan(got).sum()))
    print(line, flush=True)


if __name__ == "__main__":
    nnc.init()
    oref.ref()
    for cfg in [(1, 128, 128, 1, 1, 0, 1), (1, 128, 128, 1, 1, 0, 0), (1, 64, 64, 1, 1, 0, 1), (1, 256, 256, 1, 1, 0, 1), (1, 256, 256, 2, 2, 1, 1), (1, 128, 128, 4, 2, 0, 1), (2, 200, 328, 4, 1, 0, 1), (1, 328, 200, 2, 2, 1, 0)]:
        run(*cfg)

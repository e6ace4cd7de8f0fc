"""oracle/port.py -- TEST INFRASTRUCTURE ONLY: numpy-facing wrapper of oracle/_ref/libnnc_port.so (oracle/nnc_port.c)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_PATH = os.path.join(_HERE, "_ref", "libnnc_port.so")
_lib = None


class Conv(C.Structure):
    _fields_ = [(n, C.c_int) for n in "N H W C K R S P Q stride_h stride_w pad_h pad_w dil_h dil_w groups".split()]


class Pool(C.Structure):
    _fields_ = [(n, C.c_int) for n in "N H W C P Q R S stride_h stride_w pad_h pad_w".split()]


def available():
    return os.path.exists(PORT_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(PORT_PATH)
        _lib.port_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def gemm_forw(a, w, bias, M, N, K, ta, tb):
    b = np.zeros((M, N), np.float32)
    lib().port_gemm_forw(_p(a), _p(w), _p(bias), _p(b), M, N, K, ta, tb)
    return b


def gemm_back(g, a, w, M, N, K, ta, tb, accumulate=0, h=None, dw=None, dbias=None):
    h = np.zeros_like(a) if h is None else h
    dw = np.zeros_like(w) if dw is None else dw
    dbias = np.zeros((N,), np.float32) if dbias is None else dbias
    lib().port_gemm_back(_p(g), _p(a), _p(w), _p(h), _p(dw), _p(dbias), M, N, K, ta, tb, accumulate)
    return h, dw, dbias


def conv_desc(N, H, W, Cc, K, R, S, st, pad, dil=1, groups=1):
    P = (H + 2 * pad - ((R - 1) * dil + 1)) // st + 1
    Q = (W + 2 * pad - ((S - 1) * dil + 1)) // st + 1
    return Conv(N, H, W, Cc, K, R, S, P, Q, st, st, pad, pad, dil, dil, groups)


def conv_forw(c, a, w, bias):
    b = np.zeros((c.N, c.P, c.Q, c.K), np.float32)
    lib().port_conv_forw(C.byref(c), _p(a), _p(w), _p(bias), _p(b))
    return b


def conv_back(c, g, a, w, accumulate=0):
    h, dw, db = np.zeros_like(a), np.zeros_like(w), np.zeros((c.K,), np.float32)
    lib().port_conv_back(C.byref(c), _p(g), _p(a), _p(w), _p(h), _p(dw), _p(db), accumulate)
    return h, dw, db


def bnorm_forw_train(x, scale, bias, mean, var, eps, momentum):
    Cc = x.shape[-1]
    rows = x.size // Cc
    y, sm, sis = np.zeros_like(x), np.zeros((Cc,), np.float32), np.zeros((Cc,), np.float32)
    lib().port_bnorm_forw_train(_p(x), _p(scale), _p(bias), _p(mean), _p(var), _p(y), _p(sm), _p(sis), C.c_size_t(rows), Cc, C.c_float(eps), C.c_float(momentum))
    return y, sm, sis


def bnorm_back(g, x, scale, sm, sis):
    Cc = x.shape[-1]
    rows = x.size // Cc
    h, ds, db = np.zeros_like(x), np.zeros((Cc,), np.float32), np.zeros((Cc,), np.float32)
    lib().port_bnorm_back(_p(g), _p(x), _p(scale), _p(sm), _p(sis), _p(h), _p(ds), _p(db), C.c_size_t(rows), Cc)
    return h, ds, db


def float_to_half(f):
    f = f32(f)
    h = np.zeros(f.shape, np.uint16)
    lib().port_float_to_half(_p(f), _p(h), C.c_size_t(f.size))
    return h


def softmax_forw(a):
    b = np.zeros_like(a)
    lib().port_softmax_forw(_p(a), _p(b), a.shape[0], a.shape[1])
    return b


def sgd(g, a, m, nesterov, rate, scale, decay, momentum, dampening):
    b, n = np.zeros_like(a), np.zeros_like(a)
    lib().port_sgd(_p(g), _p(a), _p(m), _p(b), _p(n), C.c_size_t(a.size), nesterov, C.c_float(rate), C.c_float(scale), C.c_float(decay), C.c_float(momentum), C.c_float(dampening))
    return b, n


def pool_desc(N, H, W, Cc, R, st, pad):
    P, Q = (H + 2 * pad - R) // st + 1, (W + 2 * pad - R) // st + 1
    return Pool(N, H, W, Cc, P, Q, R, R, st, st, pad, pad)


def max_pool_forw(p, a):
    b = np.zeros((p.N, p.P, p.Q, p.C), np.float32)
    lib().port_max_pool_forw(C.byref(p), _p(a), _p(b))
    return b


def avg_pool_forw(p, a):
    b = np.zeros((p.N, p.P, p.Q, p.C), np.float32)
    lib().port_avg_pool_forw(C.byref(p), _p(a), _p(b))
    return b


def sdpa_forw(q, k, v, scale, is_causal, mask=None):
    """q [B, Sq, H, D], k [B, Sk, Hk, D], v [B, Sk, Hk, Dv] -> o [B, Sq, H, Dv]; mask [Sq, Sk] additive or None"""
    q, k, v = f32(q), f32(k), f32(v)
    B, Sq, H, D = q.shape
    Sk, Hk, Dv = k.shape[1], k.shape[2], v.shape[3]
    o = np.zeros((B, Sq, H, Dv), np.float32)
    m = None if mask is None else f32(mask)
    lib().port_sdpa_forw(_p(q), _p(k), _p(v), _p(m) if m is not None else None, _p(o), B, Sq, Sk, H, Hk, D, Dv, C.c_float(scale), int(is_causal))
    return o


def row_norm_forw(x, scale, bias, inner, eps, rms=0):
    """layer norm / rms norm over the trailing `inner` elements; returns (y, saved_mean, saved_inv_std)"""
    x = f32(x)
    rows = x.size // inner
    y, sm, sis = np.zeros_like(x), np.zeros((rows,), np.float32), np.zeros((rows,), np.float32)
    s = None if scale is None else f32(scale)
    b = None if bias is None else f32(bias)
    lib().port_row_norm_forw(_p(x), _p(s) if s is not None else None, _p(b) if b is not None else None, _p(y), _p(sm), _p(sis), rows, inner, C.c_float(eps), int(rms))
    return y, sm, sis

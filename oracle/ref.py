"""oracle/ref.py -- TEST INFRASTRUCTURE ONLY.

ctypes driver for the two CPU checkers built by oracle/Makefile:
  * oracle/_ref/libccv_ref.so   the unmodified reference (CCV_NNC_BACKEND_CPU_REF) + oracle/ref_shim.c
  * oracle/_ref/libnnc_port.so  our plain-C restatement (oracle/nnc_port.c)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module; the
product (ccv_b200/) never does.  Struct layouts come from ccv_b200.abi, which mirrors the same reference headers."""
import ctypes as C
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from ccv_b200 import abi  # noqa: E402

REF_PATH = os.path.join(_HERE, "_ref", "libccv_ref.so")
PORT_PATH = os.path.join(_HERE, "_ref", "libnnc_port.so")
_ref = None
_port = None

NP_TO_CCV = {np.dtype(np.float32): abi.CCV_32F, np.dtype(np.int32): abi.CCV_32S, np.dtype(np.float64): abi.CCV_64F, np.dtype(np.float16): abi.CCV_16F, np.dtype(np.uint8): abi.CCV_8U}


def available():
    return os.path.exists(REF_PATH)


def ref():
    global _ref
    if _ref is None:
        l = C.CDLL(REF_PATH)
        vp, i32, u32 = C.c_void_p, C.c_int, C.c_uint32
        l.ref_tensor_new.restype = vp
        l.ref_tensor_new.argtypes = [vp, C.POINTER(abi.TensorParam)]
        l.ref_tensor_free.argtypes = [vp]
        l.ref_tensor_view_new.restype = vp
        l.ref_tensor_view_new.argtypes = [vp, C.POINTER(abi.TensorParam), C.POINTER(i32), C.POINTER(i32)]
        l.ref_tensor_view_free.argtypes = [vp]
        l.ref_cmd_exec.restype = i32
        l.ref_cmd_exec.argtypes = [u32, u32, i32, C.POINTER(abi.CmdParam), C.POINTER(abi.Hint), i32, C.POINTER(vp), i32, C.POINTER(vp), i32]
        l.ref_cmd_time.restype = C.c_double
        l.ref_cmd_time.argtypes = [u32, u32, C.POINTER(abi.CmdParam), C.POINTER(abi.Hint), i32, C.POINTER(vp), i32, C.POINTER(vp), i32, i32]
        l.ref_hint_auto.argtypes = [C.POINTER(abi.CmdParam), C.POINTER(abi.TensorParam), C.POINTER(abi.TensorParam), C.POINTER(abi.Hint)]
        l.ref_float_to_half.argtypes = [vp, vp, C.c_size_t]
        l.ref_half_to_float.argtypes = [vp, vp, C.c_size_t]
        l.ref_num_threads.restype = i32
        l.ref_set_num_threads.argtypes = [i32]
        l.ref_nnc_init()
        _ref = l
    return _ref


class RefTensor(object):
    """A ccv_nnc_tensor_t of the reference library wrapping a numpy array's memory (no copy)."""

    def __init__(self, array, fmt=abi.CCV_TENSOR_FORMAT_NHWC, dims=None):
        self.array = array
        d = list(array.shape) if dims is None else list(dims)
        self.params = abi.tensor_param(abi.CCV_TENSOR_CPU_MEMORY, fmt, NP_TO_CCV[array.dtype], d)
        self.ptr = ref().ref_tensor_new(array.ctypes.data, C.byref(self.params))
        self.is_view = False

    def view(self, dims, ofs, stride):
        v = RefTensor.__new__(RefTensor)
        v.array = self.array
        v.params = abi.tensor_param(abi.CCV_TENSOR_CPU_MEMORY, self.params.format, self.params.datatype, dims)
        o = (C.c_int * abi.MAX_DIM_ALLOC)(*list(ofs))
        s = (C.c_int * abi.MAX_DIM_ALLOC)(*list(stride))
        v.ptr = ref().ref_tensor_view_new(self.ptr, C.byref(v.params), o, s)
        v.is_view = True
        v._owner = self
        return v

    def free(self):
        if self.ptr:
            (ref().ref_tensor_view_free if self.is_view else ref().ref_tensor_free)(self.ptr)
            self.ptr = None


def _ptrs(tensors):
    arr = (C.c_void_p * max(len(tensors), 1))()
    for i, t in enumerate(tensors):
        arr[i] = t.ptr if t is not None else None
    return arr


def cmd_exec(cmd, hint, flags, inputs, outputs, backend=abi.CCV_NNC_BACKEND_CPU_REF):
    """ccv_nnc_cmd_exec of the REFERENCE with cmd.backend = CPU_REF; inputs/outputs are RefTensor or None."""
    return ref().ref_cmd_exec(cmd.cmd, backend, 0, C.byref(cmd.info), C.byref(hint if hint is not None else abi.NO_HINT), flags, _ptrs(inputs), len(inputs), _ptrs(outputs), len(outputs))


def cmd_time(cmd, hint, flags, inputs, outputs, reps=3, backend=abi.CCV_NNC_BACKEND_CPU_REF):
    """best-of-`reps` wall-clock seconds of the reference's own implementation (for cpu_baseline)."""
    return ref().ref_cmd_time(cmd.cmd, backend, C.byref(cmd.info), C.byref(hint if hint is not None else abi.NO_HINT), flags, _ptrs(inputs), len(inputs), _ptrs(outputs), len(outputs), reps)


def run(cmd, hint, flags, in_arrays, out_arrays, fmt=abi.CCV_TENSOR_FORMAT_NHWC, in_fmts=None, out_fmts=None):
    """Convenience: wrap numpy arrays (None allowed), run CPU_REF, return status. Outputs are written in place."""
    ins = [None if a is None else RefTensor(a, (in_fmts[i] if in_fmts else fmt)) for i, a in enumerate(in_arrays)]
    outs = [None if a is None else RefTensor(a, (out_fmts[i] if out_fmts else fmt)) for i, a in enumerate(out_arrays)]
    # in-place aliases (e.g. batch norm running stats) must be the same tensor memory: callers pass the same array
    status = cmd_exec(cmd, hint, flags, ins, outs)
    for t in ins + outs:
        if t is not None:
            t.free()
    return status


def num_threads():
    return int(ref().ref_num_threads())


def physical_cores():
    """Distinct (package, core) pairs among the CPUs this process may run on; falls back to the affinity count."""
    try:
        allowed = os.sched_getaffinity(0)
    except AttributeError:
        allowed = set(range(os.cpu_count() or 1))
    cores = set()
    for cpu in allowed:
        try:
            base = "/sys/devices/system/cpu/cpu%d/topology/" % cpu
            cores.add((open(base + "physical_package_id").read().strip(), open(base + "core_id").read().strip()))
        except OSError:
            return max(1, len(allowed))
    return max(1, len(cores))


def set_num_threads(n):
    """OpenMP threads of the compiled reference for the timed CPU_REF runs (overrides an inherited OMP_NUM_THREADS)."""
    ref().ref_set_num_threads(int(n))
    return num_threads()


def float_to_half(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    out = np.empty(a.shape, dtype=np.uint16)
    ref().ref_float_to_half(a.ctypes.data, out.ctypes.data, a.size)
    return out

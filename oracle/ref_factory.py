"""oracle/ref_factory.py -- TEST INFRASTRUCTURE ONLY: lets ccv_b200.resnet50.Net build its command list over host
tensors of the REFERENCE library so that the identical list runs on CCV_NNC_BACKEND_CPU_REF (whole-model parity test,
bench.py cpu_baseline / --impl reference)."""
import numpy as np

from ccv_b200 import abi
from oracle import ref

NP = {abi.CCV_32F: np.float32, abi.CCV_32S: np.int32}


class HostTensor(ref.RefTensor):
    def upload(self, a, stream=None):
        self.array.reshape(-1)[:] = np.asarray(a, dtype=self.array.dtype).reshape(-1)
        return self

    def download(self, stream=None):
        return self.array.copy()

    @property
    def nbytes(self):
        return self.array.nbytes


class RefFactory(object):
    def alloc(self, shape, datatype=abi.CCV_32F):
        return HostTensor(np.zeros(tuple(shape), NP[datatype]))

    def alias(self, base, elem_offset, shape, datatype=abi.CCV_32F):
        n = int(np.prod(shape))
        return HostTensor(base.array.reshape(-1)[elem_offset:elem_offset + n].reshape(tuple(shape)))


POOLS = (abi.CCV_NNC_MAX_POOL_FORWARD, abi.CCV_NNC_MAX_POOL_BACKWARD, abi.CCV_NNC_AVERAGE_POOL_FORWARD, abi.CCV_NNC_AVERAGE_POOL_BACKWARD)


def run_node(node):
    """One ccv_nnc_cmd_exec on CCV_NNC_BACKEND_CPU_REF.  CPU_REF's pooling kernels only walk image 0 of a batch
    (pool/ccv_nnc_max_pool_cpu_ref.c:37-59), so pooling commands on a batch are issued once per image
    on [H, W, C] slices."""
    cmd, hint, flags, ins, outs = node
    batch = next((t.array.shape[0] for t in list(ins) + list(outs) if t is not None and t.array.ndim == 4), 1)
    if cmd.cmd in POOLS and batch > 1:
        for n in range(batch):
            sub_i = [None if t is None else HostTensor(t.array[n]) for t in ins]
            sub_o = [None if t is None else HostTensor(t.array[n]) for t in outs]
            st = ref.cmd_exec(cmd, hint, flags, sub_i, sub_o)
            for t in sub_i + sub_o:
                if t is not None:
                    t.free()
            if st != 0:
                raise RuntimeError("CPU_REF returned %d for command 0x%08x" % (st, cmd.cmd))
        return
    st = ref.cmd_exec(cmd, hint, flags, ins, outs)
    if st != 0:
        raise RuntimeError("CPU_REF returned %d for command 0x%08x" % (st, cmd.cmd))


def run_nodes(nodes, before=None, after=None):
    """ccv_nnc_graph_run's sync path on the reference: one ccv_nnc_cmd_exec (backend CPU_REF) per node, in order.
    before(i, node) / after(i, node) let a test snapshot the operands of every node (teacher-forced parity)."""
    for i, node in enumerate(nodes):
        if before:
            before(i, node)
        run_node(node)
        if after:
            after(i, node)

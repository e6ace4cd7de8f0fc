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


def run_nodes(nodes):
    """ccv_nnc_graph_run's sync path on the reference: one ccv_nnc_cmd_exec (backend CPU_REF) per node, in order."""
    for cmd, hint, flags, ins, outs in nodes:
        st = ref.cmd_exec(cmd, hint, flags, ins, outs)
        if st != 0:
            raise RuntimeError("CPU_REF returned %d for command 0x%08x" % (st, cmd.cmd))

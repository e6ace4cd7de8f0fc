"""Batch-sharded data parallelism for a command-list model: one process per GPU, a full replica each, and
ONE sum-allreduce over the flat gradient buffer between the backward commands and the SGD commands.  The reference does the
same exchange with one CMD_COMM_ALLREDUCE per parameter inside a single process
(lib/nnc/ccv_nnc_symbolic_graph_parallel.c:546-575, lib/nnc/cmd/comm/gpu/ccv_nnc_comm_gpu_nccl.cu:12-50); on GPUs
the exchange is ONE CCV_NNC_COMM_ALLREDUCE_FORWARD command of this backend (ccv_b200/csrc/sm100_comm.cu, NCCL resolved by
the library itself) over the flat gradient tensor; torch.distributed only carries the 128-byte communicator id between the
ranks.  FlatAllreduce (torch.distributed all_reduce) remains for host tensors: the gloo tests of the sharding logic."""
import numpy as np


class FlatAllreduce(object):
    """Sums net.g_flat across ranks in place. GPU tensors are wrapped zero-copy through __cuda_array_interface__ and the
    collective is enqueued on the net's own CUDA stream; host tensors (oracle-side factory, gloo) through numpy."""

    def __init__(self, net, dist, stream=None, device=0):
        import torch
        self.dist, self.torch, self.stream = dist, torch, stream
        flat = net.g_flat
        if hasattr(flat, "array"):  # host tensor of the oracle-side factory
            self.tensor = torch.from_numpy(flat.array.reshape(-1))
            self.ext = None
        else:
            class _Flat(object):
                __cuda_array_interface__ = {"shape": (net.flat_count,), "typestr": "<f4", "data": (flat.data_ptr, False), "version": 2}
            self.tensor = torch.as_tensor(_Flat(), device=torch.device("cuda", device))
            self.ext = torch.cuda.ExternalStream(stream.cuda_stream, device=torch.device("cuda", device))

    def __call__(self):
        if self.ext is None:
            self.dist.all_reduce(self.tensor)
        else:
            with self.torch.cuda.stream(self.ext):
                self.dist.all_reduce(self.tensor)


class CommandAllreduce(object):
    """The gradient exchange as the backend's own command: CMD_COMM_ALLREDUCE_FORWARD(g_flat) -> g_flat on `stream`."""
    _bound = None

    def __init__(self, net, dist, stream, rank, world):
        from ccv_b200 import nnc
        self.nnc, self.net, self.stream = nnc, net, stream
        if CommandAllreduce._bound != (world, rank):  # one communicator per process: later models of the same job reuse it
            box = [nnc.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            nnc.comm_init_rank(box[0], world, rank)
            CommandAllreduce._bound = (world, rank)
        self.cmd = nnc.CMD_COMM_ALLREDUCE_FORWARD()

    def __call__(self):
        flats = getattr(self.net, "grad_flats", None) or [self.net.g_flat]  # a 16-bit model has a 16-bit and a small fp32 buffer: one NCCL group
        st = self.nnc.cmd_exec(self.cmd, None, 0, flats, flats, self.stream)
        if st != 0:
            raise RuntimeError("COMM_ALLREDUCE returned %d: %s" % (st, self.nnc.lib().ccv_nnc_sm100_last_error()))


def shard(array, rank, world):
    """contiguous batch shard `rank` of `world` (lib/nnc/ccv_cnnp_model.c:1436-1441: P consecutive tensor lists)"""
    n = array.shape[0] // world
    return np.ascontiguousarray(array[rank * n:(rank + 1) * n])

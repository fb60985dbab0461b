"""Raw NCCL communicator for the data-parallel gradient all-reduce (csrc/dp.cu).

torch.distributed (backend "nccl") stays the plumbing — rendezvous, barriers, the broadcast of the NCCL unique id —
but the all-reduce itself is a plain stream-ordered ncclAllReduce on a communicator this module creates, because that
call can be captured into a CUDA graph together with the kernels around it (ProcessGroupNCCL work could not on this
stack: watchdog thread / work objects).  One process per GPU; call after torch.cuda.set_device().
"""
import ctypes as C

import torch
import torch.distributed as dist

from ._lib import NcclId, check, fn


class Communicator:
    def __init__(self, rank, world_size, group=None):
        if not fn["uwr_nccl_available"]():
            raise RuntimeError("libnccl.so.2 is not loadable in this process")
        uid = NcclId()
        if rank == 0:
            check(fn["uwr_nccl_unique_id"](C.byref(uid)), "uwr_nccl_unique_id")
        payload = [C.string_at(C.byref(uid), 128) if rank == 0 else None]
        dist.broadcast_object_list(payload, src=0, group=group)
        C.memmove(C.byref(uid), payload[0], 128)
        self._comm = C.c_void_p()
        check(fn["uwr_nccl_comm_init"](C.byref(self._comm), world_size, C.byref(uid), rank), "uwr_nccl_comm_init")
        self.rank, self.world = rank, world_size

    def all_reduce_sum_(self, t, stream=None):
        """in-place sum over the ranks of a contiguous float32 CUDA tensor, enqueued on `stream` (default: current)"""
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise TypeError("all_reduce_sum_ needs a contiguous CUDA float32 tensor")
        s = (stream or torch.cuda.current_stream()).cuda_stream
        check(fn["uwr_nccl_allreduce_sum_f32"](self._comm, t.data_ptr(), t.numel(), s), "uwr_nccl_allreduce_sum_f32")
        return t

    def close(self):
        if self._comm:
            fn["uwr_nccl_comm_destroy"](self._comm)
            self._comm = C.c_void_p()

"""torch.distributed plumbing for the multi-rank build.

`init_nccl(ctx)`: the normal path -- libvi_b200 owns its NCCL communicator (include/vi_b200.h vi_comm_init); torch is
only used to hand rank 0's 128-byte communicator id to the other ranks.
`Collectives`: the callback path (vi_set_collective) for hosts that bring their own transport; the same functions
work on host pointers with the gloo backend, which is how the CPU tests exercise them.

The reference has no counterpart (it is single-process); this is the host half of north_star's
"per-level range statistics are combined with a single NCCL all-reduce ... partitioning stays local".
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch
import torch.distributed as dist


class _DevicePtr:
    """Exposes a raw device pointer to torch through __cuda_array_interface__."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


def _as_tensor(ptr: int, nbytes: int, device: torch.device) -> torch.Tensor:
    if nbytes == 0:
        return torch.empty(0, dtype=torch.uint8, device=device)
    if device.type == "cuda":
        return torch.as_tensor(_DevicePtr(ptr, nbytes), device=device)
    buf = (ctypes.c_uint8 * nbytes).from_address(ptr)
    return torch.from_numpy(np.frombuffer(buf, dtype=np.uint8))


class Collectives:
    """allreduce / alltoallv over a torch.distributed process group on raw pointers."""

    def __init__(self, device: torch.device, group=None):
        self.device = torch.device(device)
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.calls = {"allreduce": 0, "alltoallv": 0, "allreduce_bytes": 0, "alltoallv_bytes": 0}

    def allreduce(self, ptr: int, count: int) -> int:
        """sum of `count` uint64 words in place (int64 view: two's-complement sums are the same bits)"""
        t = _as_tensor(ptr, count * 8, self.device).view(torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        if self.device.type == "cuda":
            torch.cuda.current_stream(self.device).synchronize()
        self.calls["allreduce"] += 1
        self.calls["allreduce_bytes"] += count * 8
        return 0

    def alltoallv(self, send_ptr: int, send_bytes, recv_ptr: int, recv_bytes) -> int:
        send = _as_tensor(send_ptr, int(sum(send_bytes)), self.device)
        recv = _as_tensor(recv_ptr, int(sum(recv_bytes)), self.device)
        if self.device.type == "cuda":
            dist.all_to_all_single(recv, send, output_split_sizes=[int(b) for b in recv_bytes],
                                   input_split_sizes=[int(b) for b in send_bytes], group=self.group)
            torch.cuda.current_stream(self.device).synchronize()
        else:
            # gloo has no all-to-all: point-to-point exchange (the CPU tests only)
            outs = list(torch.split(recv, [int(b) for b in recv_bytes]))
            ins = list(torch.split(send, [int(b) for b in send_bytes]))
            outs[self.rank].copy_(ins[self.rank])
            reqs = []
            for peer in range(self.world):
                if peer == self.rank:
                    continue
                if ins[peer].numel():
                    reqs.append(dist.isend(ins[peer].clone(), peer, group=self.group))
                if outs[peer].numel():
                    reqs.append(dist.irecv(outs[peer], peer, group=self.group))
            for r in reqs:
                r.wait()
        self.calls["alltoallv"] += 1
        self.calls["alltoallv_bytes"] += int(sum(send_bytes))
        return 0

    def attach(self, ctx) -> None:
        ctx.set_collective(self.rank, self.world, self.allreduce, self.alltoallv)


def init_nccl(ctx, device=None, group=None) -> None:
    """Collective: gives `ctx` a library-owned NCCL communicator spanning the ranks of `group`."""
    import vectorindex as vi
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = torch.device(device) if device is not None else torch.device("cpu")
    if dist.get_backend(group) == "nccl" and dev.type != "cuda":
        dev = torch.device("cuda", torch.cuda.current_device())
    uid = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(vi.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, src=0, group=group)
    ctx.comm_init(bytes(uid.cpu().numpy().tobytes()), rank, world)

"""Host-side plumbing for the multi-GPU path (one process per GPU, torch.distributed).

The path has exactly one exchange step per precompute and one per solve:
collocation / boundary points are sharded by rows, every sum over points (Gram matrix, boundary
Gram, boundary column sums, point counts, projected right-hand sides) is an all-reduce.  Nothing
here touches CUDA directly, so the logic is covered on CPU with the gloo backend
(tests/test_distributed_cpu.py); on the GPU box the same calls run over NCCL / NVLink.
"""
from __future__ import annotations

import torch

FP = 704


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous block sharding of n rows: the first n % world ranks get one extra row."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class PackedNormal:
    """The single all-reduce payload of precompute: [G (FP*FP) | G_bc (FP*FP) | s (FP) | N | N_bc], float64.

    G = DH^T DH (solver.py:167), G_bc = H_bc^T H_bc and s = H_bc^T 1 (solver.py:176-178, :196); the
    global point counts ride in the same buffer so that one collective carries everything
    precompute_LHS_RHS (solver.py:187-199) needs.
    """

    SIZE = 2 * FP * FP + FP + 2

    def __init__(self, device):
        self.buf = torch.zeros(self.SIZE, dtype=torch.float64, device=device)

    @property
    def G(self):
        return self.buf[: FP * FP].view(FP, FP)

    @property
    def Gbc(self):
        return self.buf[FP * FP: 2 * FP * FP].view(FP, FP)

    @property
    def s(self):
        return self.buf[2 * FP * FP: 2 * FP * FP + FP]

    def set_counts(self, n_pde: int, n_bc: int):
        self.buf[-2] = float(n_pde)
        self.buf[-1] = float(n_bc)

    def counts(self):
        c = self.buf[-2:].cpu()
        return int(round(c[0].item())), int(round(c[1].item()))

    def allreduce(self, group=None):
        import torch.distributed as dist

        dist.all_reduce(self.buf, op=dist.ReduceOp.SUM, group=group)
        return self


def allreduce_sum(t: torch.Tensor, group=None) -> torch.Tensor:
    import torch.distributed as dist

    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


class Communicator:
    """The sum all-reduce of the path.

    On CUDA it is the library's own NCCL communicator behind the C ABI (csrc/comm.cu: fps_comm_init /
    fps_allreduce_f64, enqueued on the caller's stream): rank 0 creates the 128-byte ncclUniqueId and torch.distributed
    merely carries it to the other ranks - the one thing a non-torch host would do by file or MPI.  ONE communicator per
    (device, process group) serves every Solver of the process and lives until the process ends: ncclCommDestroy is a
    collective, so it must not depend on when a garbage collector happens to run on each rank.  FPS_COMM=torch (or a
    failing NCCL initialisation) keeps everything on torch.distributed; CPU tensors (the gloo tests) always go there."""

    _cache = {}

    def __init__(self, lib, dev_index, group=None):
        import os

        import torch.distributed as dist

        self.group, self.lib, self.handle = group, lib, None
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if os.environ.get("FPS_COMM", "abi") != "abi":
            return
        key = (dev_index, id(group) if group is not None else None)
        if key in Communicator._cache:
            self.handle = Communicator._cache[key]
            return
        import ctypes as C

        dev = torch.device("cuda", dev_index)
        ident = [None]
        if self.rank == 0:
            buf = (C.c_char * 128)()
            if lib.fps_comm_unique_id(buf) == 0:
                ident[0] = bytes(buf)
        dist.broadcast_object_list(ident, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group,
                                   device=dev)
        handle, ok = C.c_void_p(), 0
        if ident[0] is not None:
            ok = 1 if lib.fps_comm_init(C.byref(handle), dev_index, ident[0], self.rank, self.world) == 0 else 0
        flag = torch.tensor([float(ok)], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)      # every rank must take the same path
        if flag.item() > 0.5:
            self.handle = handle
        elif ok:
            lib.fps_comm_abort(handle)
        Communicator._cache[key] = self.handle

    @property
    def native(self):
        return self.handle is not None

    def allreduce(self, t: torch.Tensor, stream=None) -> torch.Tensor:
        if self.handle is not None and t.is_cuda and t.dtype == torch.float64 and t.is_contiguous():
            from . import _lib

            st = stream if stream is not None else torch.cuda.current_stream(t.device).cuda_stream
            _lib.check(self.lib.fps_allreduce_f64(self.handle, t.data_ptr(), t.numel(), st), "fps_allreduce_f64")
            return t
        return allreduce_sum(t, self.group)

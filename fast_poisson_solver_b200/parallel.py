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

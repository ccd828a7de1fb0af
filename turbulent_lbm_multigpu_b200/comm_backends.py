"""Transport for the halo exchange -- what replaces MPI_Isend/Irecv/Waitall in
CController::syncAlpha/syncBeta (reference src/CController.hpp:299-311,361-373).

* ``TorchDistributedBackend``: one process per sub-domain/GPU (torchrun); NCCL send/recv
  on device buffers over NVLink, or gloo on host buffers (CPU tests, host-staged mode).

Sub-domains that share a process need no backend at all: ``controller.InProcessSimulation`` moves
their halos with the library's peer-copy / push-pull kernels directly.
"""
from __future__ import annotations


class TorchDistributedBackend:
    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def exchange(self, pairs):
        """pairs: list of (dst_rank, send_tensor, recv_tensor); one grouped launch, returns works."""
        dist = self.dist
        ops = []
        for dst, send, recv in pairs:
            ops.append(dist.P2POp(dist.isend, send, dst, self.group))
            ops.append(dist.P2POp(dist.irecv, recv, dst, self.group))
        return dist.batch_isend_irecv(ops) if ops else []

    def max_over_ranks(self, value, device=None):
        import torch
        t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
        return float(t.item())

    def barrier(self):
        self.dist.barrier(self.group)

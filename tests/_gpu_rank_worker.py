"""Worker for tests/test_gpu_multiprocess.py: one process per sub-domain under torchrun, the product's
CManager/CController with the CUDA solver and the selected halo transport.  One rank per GPU when the
box has enough of them (NCCL + gloo rendezvous); otherwise the ranks share the GPUs round robin and
rendezvous over gloo only (NCCL refuses two ranks on one device) -- the p2p (CUDA IPC) and host
transports do not need NCCL, so the multi-process path is exercised on a single-GPU box too."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from turbulent_lbm_multigpu_b200 import capi
    from turbulent_lbm_multigpu_b200.comm_backends import TorchDistributedBackend
    from turbulent_lbm_multigpu_b200.configuration import CConfiguration
    from turbulent_lbm_multigpu_b200.controller import CManager
    from turbulent_lbm_multigpu_b200.domain import CDomain

    D = tuple(int(v) for v in os.environ["LBM_TEST_DOMAIN"].split(","))
    nums = tuple(int(v) for v in os.environ["LBM_TEST_NUMS"].split(","))
    steps = int(os.environ["LBM_TEST_STEPS"])
    sync = os.environ["LBM_TEST_SYNC"]
    out = os.environ["LBM_TEST_OUT"]
    axis_order = os.environ.get("LBM_TEST_AXIS_ORDER") or "xyz"     # the oracle run it is compared with
    order = {"linear": capi.LBM_BETA_ORDER_LINEAR, "shipped": capi.LBM_BETA_ORDER_SHIPPED}[
        os.environ.get("LBM_TEST_BETA_ORDER") or "linear"]
    cs = float(os.environ.get("LBM_TEST_CS") or 0.0)
    dtype = np.float64 if os.environ.get("LBM_TEST_DTYPE") == "f64" else np.float32
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    ndev = torch.cuda.device_count()
    shared = world > ndev
    local = local % ndev
    torch.cuda.set_device(local)
    if shared:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("cpu:gloo,cuda:nccl", device_id=torch.device("cuda", local))
    rank = dist.get_rank()
    cfg = CConfiguration()
    cfg.loops = steps
    cfg.domain_size, cfg.subdomain_num = D, nums
    cfg.debug_mode = True                      # STORE_VELOCITY / STORE_DENSITY
    cfg.smagorinsky_constant = cs
    compute_stream, comm_stream = torch.cuda.Stream(), torch.cuda.Stream(priority=-1)
    mgr = CManager(CDomain(-1, D, (0, 0, 0), (0.1, 0.1, 0.1)), nums, backend=TorchDistributedBackend(),
                   device=local, sync_mode=sync, config=cfg, dtype=dtype,
                   beta_order=order, axis_order=axis_order, block_size=int(os.environ.get("LBM_TEST_BLOCK") or 0),
                   compute_stream=compute_stream.cuda_stream, comm_stream=comm_stream.cuda_stream)
    mgr.initSimulation(rank)
    ctrl = mgr.getController()
    if os.environ.get("LBM_TEST_GRAPH") == "1":
        # what `bench.py --graph` does: capture the beta+alpha cycle once, replay it.  The halo sequence
        # numbers are device-resident, so every replay must synchronise with the neighbour like an eager step.
        assert steps % 2 == 0 and steps >= 4
        ctrl.computeNextStep()
        ctrl.computeNextStep()
        torch.cuda.synchronize()
        dist.barrier()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=compute_stream, capture_error_mode="thread_local"):
            ctrl.computeNextStep()
            ctrl.computeNextStep()
        # the capture launched nothing; the step counter of the library advanced by two: rewind it
        sv = ctrl.getSolver()
        sv.simulation_step_counter = sv.simulation_step_counter - 2
        with torch.cuda.stream(compute_stream):
            for _ in range((steps - 2) // 2):
                graph.replay()
        torch.cuda.synchronize()
        sv.simulation_step_counter = steps
    else:
        for _ in range(steps):
            ctrl.computeNextStep()
    s = ctrl.getSolver()
    s.wait()
    torch.cuda.synchronize()
    np.savez(os.path.join(out, "rank%d.npz" % rank), dd=s.storeDensityDistribution(), velocity=s.storeVelocity(),
             flags=s.storeFlags())
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

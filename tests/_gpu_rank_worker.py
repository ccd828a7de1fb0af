"""Worker for tests/test_gpu_multiprocess.py: one process per GPU under torchrun (NCCL rendezvous),
the product's CManager/CController with the CUDA solver and the selected halo transport."""
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
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("cpu:gloo,cuda:nccl", device_id=torch.device("cuda", local))
    rank = dist.get_rank()
    cfg = CConfiguration()
    cfg.loops = steps
    cfg.domain_size, cfg.subdomain_num = D, nums
    cfg.debug_mode = True                      # STORE_VELOCITY / STORE_DENSITY
    compute_stream, comm_stream = torch.cuda.Stream(), torch.cuda.Stream(priority=-1)
    mgr = CManager(CDomain(-1, D, (0, 0, 0), (0.1, 0.1, 0.1)), nums, backend=TorchDistributedBackend(),
                   device=local, sync_mode=sync, config=cfg, dtype=np.float32,
                   beta_order=capi.LBM_BETA_ORDER_LINEAR, axis_order=axis_order,
                   compute_stream=compute_stream.cuda_stream, comm_stream=comm_stream.cuda_stream)
    mgr.initSimulation(rank)
    ctrl = mgr.getController()
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

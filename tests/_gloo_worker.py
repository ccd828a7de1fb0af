"""Worker for tests/test_multiprocess_gloo.py: one process per sub-domain, gloo backend.

The PRODUCT's host logic (CManager decomposition, CComm tables, CController time loop and the
reference-style host-staged syncAlpha/syncBeta over torch.distributed) is driven with an
oracle-backed solver injected from here (tests may use the oracle; the product never does)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    from oracle import port
    from turbulent_lbm_multigpu_b200.comm_backends import TorchDistributedBackend
    from turbulent_lbm_multigpu_b200.configuration import CConfiguration
    from turbulent_lbm_multigpu_b200.controller import CManager
    from turbulent_lbm_multigpu_b200.domain import CDomain
    from turbulent_lbm_multigpu_b200.skeleton import compute_parameters

    D = tuple(int(v) for v in os.environ["LBM_TEST_DOMAIN"].split(","))
    nums = tuple(int(v) for v in os.environ["LBM_TEST_NUMS"].split(","))
    steps = int(os.environ["LBM_TEST_STEPS"])
    out = os.environ["LBM_TEST_OUT"]
    dist.init_process_group("gloo")
    rank = dist.get_rank()
    cfg = CConfiguration()
    cfg.loops = steps
    cfg.domain_size, cfg.subdomain_num = D, nums

    class Engine(port.OracleSolver):          # the oracle behind the CLbmSolver surface
        def wait(self):
            pass

    def factory(uid, domain, BC, cfg):
        p = compute_parameters(domain.getSize(), domain.getLength(), cfg.gravitation, cfg.viscosity,
                               cfg.drivenCavityVelocity, dtype=np.float32)
        bc6 = [BC[a][s] for a in range(3) for s in range(2)]
        return Engine(domain.getSize(), bc6, p.inv_tau, p.gravitation, p.u_lid, variant=port.NOSHM,
                                 tau=p.tau)

    mgr = CManager(CDomain(-1, D, (0, 0, 0), (0.1, 0.1, 0.1)), nums, backend=TorchDistributedBackend(),
                   sync_mode="host", solver_factory=factory, config=cfg, dtype=np.float32)
    mgr.initSimulation(rank)
    mgr.startSimulation(quiet=True)
    s = mgr.getController().getSolver()
    inner = tuple(v - 2 for v in mgr.getSubdomainSize())
    np.savez(os.path.join(out, "rank%d.npz" % rank), velocity=s.storeVelocity((1, 1, 1), inner),
             flags=s.storeFlags((1, 1, 1), inner), dd=s.dd,
             comms=np.array([c.as_tuple()[0] for c in mgr.getController().getComms()]))
    dist.barrier()
    dist.destroy_process_group()
    del torch


if __name__ == "__main__":
    main()

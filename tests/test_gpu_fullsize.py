"""Parity at BASELINE.json's full sizes through size-independent properties (the CPU oracle
cannot run 256^3 x many steps inside a test budget; tests/test_gpu_parity.py covers the sizes
it can).  Every property is bit-exact:

  * the production kernel configuration (2 cells per thread, 128-thread blocks, concurrent
    wrapping kernel) equals the scalar 1-cell-per-thread kernels, which test_gpu_parity.py pins
    against the oracle;
  * the reference's validate criterion (src/main.cpp:309-408) at full size: a (1,1,2)-decomposed
    256x256x256 run equals the single 256x256x254 domain on the interior blocks;
  * the device checksum reduction agrees with the reference's serial host-order sum;
  * the 64^3 / 100-step known answers of SURVEY.md A.2 (generated from the reference kernels).
"""
import numpy as np
import pytest

from helpers import bits_equal, make_cuda, make_oracle, params_for, set_lid
from turbulent_lbm_multigpu_b200 import capi
from turbulent_lbm_multigpu_b200.configuration import CConfiguration
from turbulent_lbm_multigpu_b200.controller import (InProcessSimulation, validation_domain_size,
                                                    validation_sub_origin)
from turbulent_lbm_multigpu_b200.domain import CDomain
from turbulent_lbm_multigpu_b200.solver import CLbmSolver

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("size,dtype,cs,steps", [
    ((256, 256, 256), np.float32, 0.0, 6),      # BASELINE configs[1] as the reference runs it (BGK)
    ((256, 256, 256), np.float32, 0.1, 6),      # ... and as bench.py runs it (Smagorinsky C_s = 0.1)
    ((192, 96, 96), np.float64, 0.0, 6),        # fp64 production kernels (no work-group quirk: 128 % 192 != 0)
    ((192, 96, 96), np.float64, 0.1, 6),
    ((384, 384, 16), np.float64, 0.1, 4),       # row length / plane size of configs[4]
])
def test_production_kernels_equal_oracle_and_reference_at_full_size(size, dtype, cs, steps):
    """The DEFAULT launch configuration (what bench.py times: 2 cells per thread, 128-thread
    blocks, shipped summation order, concurrent wrapping kernel) against the CPU oracle and --
    for BGK, which is all the reference has -- against the reference's own kernels (oracle/_ref,
    strict IEEE build, shipped shared-memory beta path).  Every population and every flag of
    every cell, bit for bit, after each of the first steps."""
    from oracle import ref
    c = make_cuda(size, dtype, cs=cs, store=False)
    cfgd = c.config()
    assert cfgd["vector_width"] == 2 and cfgd["block_size"] == 128 and cfgd["wg_quirk"] == 0
    o = make_oracle(size, dtype, cs=cs)
    r = None
    if cs == 0.0 and ref.available():
        p = params_for(size, dtype)
        try:
            r = ref.RefSolver(size, [1] * 6, p.inv_tau, p.gravitation, p.u_lid, dtype=dtype, variant=ref.SHM)
            set_lid(r, size)
        except KeyError:
            r = None
    if cs == 0.0 and size[0] != 384:
        assert r is not None, "oracle/_ref has no instance for %s %s" % (size, np.dtype(dtype).name)
    for i in range(steps):
        c.simulationStep()
        o.simulationStep()
        got = c.storeDensityDistribution()
        assert bits_equal(got, o.dd), ("oracle", i)
        if r is not None:
            r.simulationStep()
            assert bits_equal(got, r.dd), ("reference kernels", i)
    assert bits_equal(c.storeFlags(), o.flags)
    if r is not None:
        assert bits_equal(c.storeFlags(), r.flags)
    c.close()


@pytest.mark.parametrize("size,dtype,cs,steps", [
    ((256, 256, 256), np.float32, 0.0, 6),      # BASELINE configs[1] without / with Smagorinsky
    ((256, 256, 256), np.float32, 0.1, 6),
    ((192, 192, 192), np.float64, 0.1, 5),      # the fp64 path (configs[4] is 384^3: same kernels)
])
def test_vectorised_kernels_equal_scalar_kernels_at_full_size(size, dtype, cs, steps):
    fast = make_cuda(size, dtype, cs=cs, store=False)                       # library defaults
    slow = make_cuda(size, dtype, cs=cs, store=False, vector_width=1, block_size=256)
    assert fast.config()["vector_width"] == 2 and slow.config()["vector_width"] == 1
    fast.simulationSteps(steps)
    slow.simulationSteps(steps)
    a = fast.storeDensityDistribution()
    b = slow.storeDensityDistribution()
    assert bits_equal(a, b)
    fast.close()
    slow.close()


def test_validate_criterion_at_full_size():
    D, nums, steps, L = (256, 256, 256), (1, 1, 2), 10, (0.1, 0.1, 0.1)
    cfg = CConfiguration()
    cfg.debug_mode = True
    cfg.smagorinsky_constant = 0.1
    sim = InProcessSimulation(CDomain(-1, D, (0, 0, 0), L), nums, transport="p2p", overlap=True, config=cfg,
                              dtype=np.float32)
    sim.run(steps)
    p = sim.controllers[0].getSolver().params
    V = validation_domain_size(D, nums)
    single = CLbmSolver(0, 0, [[1, 1]] * 3, CDomain(0, V, (0, 0, 0), L), dtype=np.float32, store_velocity=True,
                        store_density=False, smagorinsky_cs=0.1, params=p)
    rect = (V[0] - 2, 1, V[2] - 2)
    single.setFlags(np.full(rect[0] * rect[2], 4, np.int32), (1, V[1] - 2, 1), rect)
    single.simulationSteps(steps)
    inner = tuple(s - 2 for s in sim.sub_size)
    for r, ctrl in enumerate(sim.controllers):
        o = validation_sub_origin(r, nums, inner)
        got = ctrl.getSolver().storeVelocity(origin=(1, 1, 1), size=inner)
        exp = single.storeVelocity(origin=o, size=inner)
        assert bits_equal(got, exp), r
        assert np.abs(got).max() > 0


def test_device_checksum_matches_host_order_at_full_size():
    s = make_cuda((256, 256, 256), np.float32, cs=0.1, store=True)
    s.simulationSteps(20)
    host = s.getVelocityChecksum(host_order=True)       # float accumulator, index order (reference)
    dev = s.getVelocityChecksum(host_order=False)       # warp-shuffle reduction in double
    vel = s.storeVelocity().reshape(3, -1).astype(np.float64)
    fl = s.storeFlags()
    exact = float((vel[0] + vel[1] + vel[2])[fl == 2].sum())
    assert abs(dev - exact) <= 1e-9 * max(1.0, abs(exact))
    assert abs(host - exact) <= 2e-2 * max(1.0, abs(exact))   # fp32 serial accumulation over 16.5M cells
    s.close()

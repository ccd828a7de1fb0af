"""Multi-sub-domain parity on the GPU: the reference's own validate criterion
(src/main.cpp:309-408) -- an n-way decomposed run must equal the single domain of size
D - 2(n-1) bit for bit on the interior blocks -- plus equality with the oracle's decomposed
run on every population of every sub-domain (ghost layers included)."""
import numpy as np
import pytest

from helpers import bits_equal
from oracle import multi as omulti
from turbulent_lbm_multigpu_b200 import capi
from turbulent_lbm_multigpu_b200.configuration import CConfiguration
from turbulent_lbm_multigpu_b200.controller import (InProcessSimulation, validation_domain_size,
                                                    validation_sub_origin)
from turbulent_lbm_multigpu_b200.domain import CDomain
from turbulent_lbm_multigpu_b200.solver import CLbmSolver

pytestmark = pytest.mark.gpu


def _cfg():
    c = CConfiguration()
    c.debug_mode = True      # STORE_VELOCITY / STORE_DENSITY on, like a DEBUG build of the reference
    return c


@pytest.mark.parametrize("D,nums,steps", [
    ((32, 16, 16), (2, 1, 1), 40),
    ((36, 16, 16), (3, 1, 1), 41),
    ((16, 16, 24), (1, 1, 2), 60),
    ((16, 24, 16), (1, 2, 1), 61),
    ((24, 24, 12), (2, 2, 1), 60),
    ((24, 24, 24), (2, 2, 2), 81),
    ((64, 32, 48), (2, 1, 3), 30),
])
@pytest.mark.parametrize("slots,transport,overlap", [
    (capi.LBM_HALO_SLOTS_MINIMAL, "copy", False),
    (capi.LBM_HALO_SLOTS_REFERENCE, "copy", False),
    (capi.LBM_HALO_SLOTS_MINIMAL, "p2p", False),      # push / flag / pull over peer memory
    (capi.LBM_HALO_SLOTS_MINIMAL, "p2p", True),       # ... with the shell/interior split
    (capi.LBM_HALO_SLOTS_REFERENCE, "p2p", True),
])
def test_decomposed_equals_single_domain(D, nums, steps, slots, transport, overlap):
    L = (0.1, 0.1, 0.1)
    sim = InProcessSimulation(CDomain(-1, D, (0, 0, 0), L), nums, slots=slots, transport=transport,
                              overlap=overlap, config=_cfg(),
                              dtype=np.float32, beta_order=capi.LBM_BETA_ORDER_LINEAR)
    sim.run(steps)
    p = sim.controllers[0].getSolver().params
    V = validation_domain_size(D, nums)
    single = CLbmSolver(0, 0, [[1, 1]] * 3, CDomain(0, V, (0, 0, 0), L), dtype=np.float32, store_velocity=True,
                        store_density=True, beta_order=capi.LBM_BETA_ORDER_LINEAR, params=p)
    rect = (V[0] - 2, 1, V[2] - 2)
    single.setFlags(np.full(rect[0] * rect[2], 4, np.int32), (1, V[1] - 2, 1), rect)
    single.simulationSteps(steps)
    inner = tuple(s - 2 for s in sim.sub_size)
    for r, ctrl in enumerate(sim.controllers):
        o = validation_sub_origin(r, nums, inner)
        s = ctrl.getSolver()
        assert bits_equal(s.storeVelocity(origin=(1, 1, 1), size=inner), single.storeVelocity(origin=o, size=inner)), (r, "velocity")
        assert bits_equal(s.storeFlags(origin=(1, 1, 1), size=inner), single.storeFlags(origin=o, size=inner)), (r, "flags")
        fl = s.storeFlags(origin=(1, 1, 1), size=inner)
        m = np.isin(fl, (2, 4))
        a, b = s.storeDensity(origin=(1, 1, 1), size=inner), single.storeDensity(origin=o, size=inner)
        assert bits_equal(a[m], b[m]), (r, "density")

    # and against the oracle's decomposed run: every population of every rank
    make, po = omulti.make_oracle_factory(D, nums, L, dtype=np.float32, variant=1)
    md = omulti.MultiDomain(D, nums, make, slots="minimal" if slots == capi.LBM_HALO_SLOTS_MINIMAL else "reference")
    md.run(steps)
    for r, ctrl in enumerate(sim.controllers):
        assert bits_equal(ctrl.getSolver().storeDensityDistribution(), md.ranks[r]["solver"].dd), (r, "dd vs oracle")


@pytest.mark.parametrize("D,nums,steps", [
    ((32, 16, 16), (2, 1, 1), 40),
    ((36, 16, 16), (3, 1, 1), 41),
    ((24, 24, 12), (2, 2, 1), 60),
    ((24, 24, 24), (2, 2, 2), 81),
    ((64, 32, 48), (2, 1, 3), 30),
    ((16, 16, 24), (1, 1, 2), 31),       # no x neighbour: z,y,x order with nothing left to do after the interior
])
@pytest.mark.parametrize("overlap", [False, True])
def test_zyx_axis_order_x_faces_after_the_interior(D, nums, steps, overlap):
    """LBM_AXIS_ORDER_ZYX (include/lbm_b200.h): z and y faces under the interior kernel, x faces
    exchanged after it without an x shell.  Same acceptance as the default order: the reference's
    validate criterion against the single domain, plus every population of every sub-domain
    (ghost layers included) against the oracle run with the same phase order."""
    L = (0.1, 0.1, 0.1)
    sim = InProcessSimulation(CDomain(-1, D, (0, 0, 0), L), nums, transport="p2p", overlap=overlap, config=_cfg(),
                              dtype=np.float32, beta_order=capi.LBM_BETA_ORDER_LINEAR, axis_order="zyx")
    assert all(c.getSolver().commAxisOrder() == capi.LBM_AXIS_ORDER_ZYX for c in sim.controllers)
    sim.run(steps)
    p = sim.controllers[0].getSolver().params
    V = validation_domain_size(D, nums)
    single = CLbmSolver(0, 0, [[1, 1]] * 3, CDomain(0, V, (0, 0, 0), L), dtype=np.float32, store_velocity=True,
                        store_density=True, beta_order=capi.LBM_BETA_ORDER_LINEAR, params=p)
    rect = (V[0] - 2, 1, V[2] - 2)
    single.setFlags(np.full(rect[0] * rect[2], 4, np.int32), (1, V[1] - 2, 1), rect)
    single.simulationSteps(steps)
    inner = tuple(s - 2 for s in sim.sub_size)
    make, po = omulti.make_oracle_factory(D, nums, L, dtype=np.float32, variant=1)
    md = omulti.MultiDomain(D, nums, make, slots="minimal", axis_order=(2, 1, 0))
    md.run(steps)
    for r, ctrl in enumerate(sim.controllers):
        o = validation_sub_origin(r, nums, inner)
        s = ctrl.getSolver()
        assert bits_equal(s.storeVelocity(origin=(1, 1, 1), size=inner), single.storeVelocity(origin=o, size=inner)), (r, "velocity")
        assert bits_equal(s.storeDensityDistribution(), md.ranks[r]["solver"].dd), (r, "dd vs oracle")


@pytest.mark.parametrize("D,nums,steps", [
    ((40, 24, 32), (1, 1, 2), 31),       # z-slabs: what bench.py / the driver's scaling run uses
    ((40, 24, 48), (1, 1, 3), 30),       # a rank with two faces
    ((48, 40, 24), (2, 2, 2), 41),       # blocks (x cut: z,y,x order chosen like bench.py does)
    ((80, 24, 16), (2, 1, 1), 30),       # the reference's x split
    ((512, 12, 10), (2, 1, 1), 21),      # rows of 256 cells = whole blocks: the block-uniform lane test of the fused x exchange
    ((512, 24, 12), (2, 2, 1), 20),
    ((1024, 10, 8), (2, 1, 1), 11),      # two blocks per row
    ((120, 24, 16), (3, 1, 1), 21),      # a rank with two x faces
    ((600, 12, 10), (2, 1, 1), 15),      # rows of 300 cells: no block size divides a row, separate x kernels
    ((768, 12, 10), (2, 1, 1), 15),      # rows of 384 cells: the fused launches run with 96-thread blocks (192 groups = 2 x 96)
    ((900, 10, 8), (3, 1, 1), 11),       # rows of 300 cells, two x faces
])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("cs", [0.0, 0.1])
def test_bench_defaults_decomposed_equal_oracle(D, nums, steps, dtype, cs):
    """The configuration bench.py and the driver's scaling run actually use -- SHIPPED summation
    order, Smagorinsky C_s = 0.1 (and BGK), fp32 and fp64, p2p transport with the overlapped step,
    production (vectorised) kernels: sub-domain rows are no divisor of 128, so the work-group quirk
    is off -- against the oracle's decomposed run with the same options, every population of every
    rank, plus the reference's validate criterion against the single domain."""
    L = (0.1, 0.1, 0.1)
    cfg = _cfg()
    cfg.smagorinsky_constant = cs
    zyx = nums[0] > 1
    sim = InProcessSimulation(CDomain(-1, D, (0, 0, 0), L), nums, transport="p2p", overlap=True, config=cfg,
                              dtype=dtype, axis_order="zyx" if zyx else "xyz")
    for c in sim.controllers:
        k = c.getSolver().config()
        assert k["wg_quirk"] == 0 and k["vector_width"] == 2
    sim.run(steps)
    make, po = omulti.make_oracle_factory(D, nums, L, dtype=dtype, variant=0, smagorinsky_cs=cs)
    md = omulti.MultiDomain(D, nums, make, slots="minimal", axis_order=(2, 1, 0) if zyx else (0, 1, 2))
    md.run(steps)
    for r, ctrl in enumerate(sim.controllers):
        assert bits_equal(ctrl.getSolver().storeDensityDistribution(), md.ranks[r]["solver"].dd), (r, "dd vs oracle")
        assert bits_equal(ctrl.getSolver().storeFlags(), md.ranks[r]["solver"].flags), (r, "flags vs oracle")
    p = sim.controllers[0].getSolver().params
    V = validation_domain_size(D, nums)
    single = CLbmSolver(0, 0, [[1, 1]] * 3, CDomain(0, V, (0, 0, 0), L), dtype=dtype, store_velocity=True,
                        store_density=True, smagorinsky_cs=cs, params=p)
    rect = (V[0] - 2, 1, V[2] - 2)
    single.setFlags(np.full(rect[0] * rect[2], 4, np.int32), (1, V[1] - 2, 1), rect)
    single.simulationSteps(steps)
    inner = tuple(s - 2 for s in sim.sub_size)
    for r, ctrl in enumerate(sim.controllers):
        o = validation_sub_origin(r, nums, inner)
        s = ctrl.getSolver()
        assert bits_equal(s.storeVelocity(origin=(1, 1, 1), size=inner), single.storeVelocity(origin=o, size=inner)), (r, "velocity")


@pytest.mark.parametrize("D,nums,steps", [
    ((384, 24, 16), (2, 1, 1), 21),      # rows of 192 cells = three 32-thread blocks, no work-group quirk: vectorised kernels
    ((576, 24, 16), (3, 1, 1), 21),      # a rank with two x faces
    ((384, 40, 24), (2, 2, 2), 31),      # blocks: rim lines forwarded from the y/z phases
    ((128, 24, 16), (2, 1, 1), 31),      # rows of 64 cells = ONE block (both lanes in every block); 128 % 64 == 0, so the
    ((128, 40, 24), (2, 2, 2), 41),      # work-group quirk is live and the scalar kernel does the exchange by location
    ((256, 40, 12), (2, 2, 1), 40),      # rows of 128 cells, quirk live
])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_fused_x_exchange_with_small_blocks(D, nums, steps, dtype):
    """The fused x exchange (step kernels push and pull the x faces themselves) needs rows that are whole
    thread blocks; with 32-thread blocks that is 64 cells, small enough for the oracle."""
    L = (0.1, 0.1, 0.1)
    cfg = _cfg()
    cfg.smagorinsky_constant = 0.1
    sim = InProcessSimulation(CDomain(-1, D, (0, 0, 0), L), nums, transport="p2p", overlap=True, config=cfg,
                              dtype=dtype, axis_order="zyx", block_size=32)
    for c in sim.controllers:
        k = c.getSolver().config()
        assert k["vector_width"] == 2 and k["block_size"] == 32
        assert (k["wg_quirk"] == 0) == (128 % sim.sub_size[0] != 0)
    sim.run(steps)
    make, po = omulti.make_oracle_factory(D, nums, L, dtype=dtype, variant=0, smagorinsky_cs=0.1)
    md = omulti.MultiDomain(D, nums, make, slots="minimal", axis_order=(2, 1, 0))
    md.run(steps)
    for r, ctrl in enumerate(sim.controllers):
        assert bits_equal(ctrl.getSolver().storeDensityDistribution(), md.ranks[r]["solver"].dd), (r, "dd vs oracle")


@pytest.mark.parametrize("D,nums,steps", [
    ((96, 32, 32), (3, 1, 1), 21),       # rows of 32 cells in one 128-thread block, work-group quirk live, two x faces
    ((64, 32, 32), (2, 2, 1), 20),       # ... with a y neighbour: rim lines
    ((200, 12, 10), (2, 1, 1), 15),      # rows of 100 cells: 50 groups in one block of 64
    ((82, 12, 10), (2, 1, 1), 15),       # rows of 41 cells: one cell per thread, the lanes are threads 1 and 39
])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_fused_x_exchange_any_row_length(D, nums, steps, dtype, monkeypatch):
    """Default block size, rows that no block size divides: on request (LBM_B200_XFUSE=2; slower than the
    separate x kernels there, so not the default) the fused launches pick the block size that leaves the
    fewest idle threads and the last block of a row is partly empty."""
    monkeypatch.setenv("LBM_B200_XFUSE", "2")
    L = (0.1, 0.1, 0.1)
    cfg = _cfg()
    cfg.smagorinsky_constant = 0.1
    sim = InProcessSimulation(CDomain(-1, D, (0, 0, 0), L), nums, transport="p2p", overlap=True, config=cfg,
                              dtype=dtype, axis_order="zyx")
    sim.run(steps)
    make, po = omulti.make_oracle_factory(D, nums, L, dtype=dtype, variant=0, smagorinsky_cs=0.1)
    md = omulti.MultiDomain(D, nums, make, slots="minimal", axis_order=(2, 1, 0))
    md.run(steps)
    for r, ctrl in enumerate(sim.controllers):
        assert bits_equal(ctrl.getSolver().storeDensityDistribution(), md.ranks[r]["solver"].dd), (r, "dd vs oracle")


@pytest.mark.parametrize("D,nums,steps", [((512, 12, 10), (2, 1, 1), 15), ((768, 12, 10), (3, 1, 1), 11)])
@pytest.mark.parametrize("dtype,vw", [(np.float32, 1), (np.float32, 4), (np.float64, 1)])
def test_fused_x_exchange_other_vector_widths(D, nums, steps, dtype, vw):
    """1 and 4 cells per thread (the shipped default is 2): the lanes sit in other threads / elements
    (XLane), rows of 256 cells are 2 blocks of 128 one-cell threads or one block of 64 four-cell threads."""
    L = (0.1, 0.1, 0.1)
    cfg = _cfg()
    cfg.smagorinsky_constant = 0.1
    sim = InProcessSimulation(CDomain(-1, D, (0, 0, 0), L), nums, transport="p2p", overlap=True, config=cfg,
                              dtype=dtype, axis_order="zyx", vector_width=vw)
    for c in sim.controllers:
        k = c.getSolver().config()
        assert k["vector_width"] == vw and k["wg_quirk"] == 0
    sim.run(steps)
    make, po = omulti.make_oracle_factory(D, nums, L, dtype=dtype, variant=0, smagorinsky_cs=0.1)
    md = omulti.MultiDomain(D, nums, make, slots="minimal", axis_order=(2, 1, 0))
    md.run(steps)
    for r, ctrl in enumerate(sim.controllers):
        assert bits_equal(ctrl.getSolver().storeDensityDistribution(), md.ranks[r]["solver"].dd), (r, "dd vs oracle")


@pytest.mark.parametrize("D,nums", [((384, 24, 16), (2, 1, 1)), ((384, 40, 24), (2, 2, 2)), ((128, 24, 16), (2, 1, 1))])
def test_fused_x_exchange_materialises_on_host_access(D, nums):
    """With the z,y,x order the x faces are received into their blocks and read from there by the next
    step kernel; a host read of the populations in between (any step parity) must still see them in dd, and
    the run must continue unharmed."""
    L = (0.1, 0.1, 0.1)
    cfg = _cfg()
    cfg.smagorinsky_constant = 0.1
    sim = InProcessSimulation(CDomain(-1, D, (0, 0, 0), L), nums, transport="p2p", overlap=True, config=cfg,
                              dtype=np.float32, axis_order="zyx", block_size=32)
    make, po = omulti.make_oracle_factory(D, nums, L, dtype=np.float32, variant=0, smagorinsky_cs=0.1)
    md = omulti.MultiDomain(D, nums, make, slots="minimal", axis_order=(2, 1, 0))
    for chunk in (7, 1, 2, 5):
        sim.run(chunk)
        md.run(chunk)
        for r, ctrl in enumerate(sim.controllers):
            assert bits_equal(ctrl.getSolver().storeDensityDistribution(), md.ranks[r]["solver"].dd), (r, chunk)


def _device_count():
    import ctypes
    n = ctypes.c_int(0)
    capi.load().lbmGetDeviceCount(ctypes.byref(n))
    return n.value


@pytest.mark.parametrize("transport,overlap", [("copy", False), ("p2p", True)])
def test_decomposed_with_padded_slot_stride(transport, overlap, monkeypatch):
    """halo pack / push / unpack / peer copy with a padded dd slot stride"""
    monkeypatch.setenv("LBM_B200_SLOT_PAD_BYTES", "4352")
    D, nums, steps, L = (24, 24, 24), (2, 2, 2), 21, (0.1, 0.1, 0.1)
    sim = InProcessSimulation(CDomain(-1, D, (0, 0, 0), L), nums, transport=transport, overlap=overlap,
                              config=_cfg(), dtype=np.float32, beta_order=capi.LBM_BETA_ORDER_LINEAR)
    sim.run(steps)
    make, po = omulti.make_oracle_factory(D, nums, L, dtype=np.float32, variant=1)
    md = omulti.MultiDomain(D, nums, make, slots="minimal")
    md.run(steps)
    for r, ctrl in enumerate(sim.controllers):
        assert bits_equal(ctrl.getSolver().storeDensityDistribution(), md.ranks[r]["solver"].dd), (r, "dd vs oracle")


@pytest.mark.parametrize("transport,overlap", [("copy", False), ("p2p", False), ("p2p", True)])
def test_decomposed_across_devices(transport, overlap):
    """The same validate criterion with the sub-domains on DIFFERENT GPUs of the box (peer
    stores over NVLink); skipped on a single-GPU box."""
    ndev = _device_count()
    if ndev < 2:
        pytest.skip("needs >= 2 GPUs")
    D, nums, steps, L = (32, 32, 48), (1, 2, 2), 41, (0.1, 0.1, 0.1)
    devices = list(range(min(ndev, 4)))
    sim = InProcessSimulation(CDomain(-1, D, (0, 0, 0), L), nums, devices=devices, transport=transport,
                              overlap=overlap, config=_cfg(), dtype=np.float32,
                              beta_order=capi.LBM_BETA_ORDER_LINEAR)
    sim.run(steps)
    make, po = omulti.make_oracle_factory(D, nums, L, dtype=np.float32, variant=1)
    md = omulti.MultiDomain(D, nums, make, slots="minimal")
    md.run(steps)
    for r, ctrl in enumerate(sim.controllers):
        assert bits_equal(ctrl.getSolver().storeDensityDistribution(), md.ranks[r]["solver"].dd), (r, "dd vs oracle")

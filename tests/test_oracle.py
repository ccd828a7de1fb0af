"""CPU tests of the oracle itself: the plain-C restatement against (1) the committed golden
vectors produced by the reference's own kernels, (2) the known answers of SURVEY.md A.2, and
(3) -- when oracle/_ref is built (this container) -- the reference kernels run side by side.
All comparisons are bit-exact."""
import glob
import os

import numpy as np
import pytest

from helpers import bits_equal, set_lid
from oracle import multi, port, ref

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _oracle_from_fixture(z):
    size = tuple(int(v) for v in z["size"])
    dtype = z["dd_%d" % int(z["steps"][0])].dtype
    s = port.OracleSolver(size, [int(b) for b in z["bc"]], z["inv_tau"], z["gravitation"], z["u_lid"], dtype=dtype,
                          variant=int(z["variant"]))
    set_lid(s, size)
    return s


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "f*.npz"))), ids=os.path.basename)
def test_port_reproduces_reference_golden_vectors(path):
    z = np.load(path)
    s = _oracle_from_fixture(z)
    done = 0
    for k in [int(v) for v in z["steps"]]:
        while done < k:
            s.simulationStep()
            done += 1
        assert bits_equal(s.dd, z["dd_%d" % k]), (path, k, "dd")
        assert bits_equal(s.velocity, z["velocity_%d" % k]), (path, k, "velocity")
        assert bits_equal(s.density, z["density_%d" % k]), (path, k, "density")
    assert bits_equal(s.flags, z["flags"])


def test_known_answers_64cubed():
    """Values produced by the reference kernels and stored by
    tests/golden/make_golden.py: centre velocity, centre density buffer, checksum."""
    kat = np.load(os.path.join(GOLDEN, "kat_64.npz"))
    # literal pins from SURVEY.md A.2
    assert np.float32(kat["f32_64_shm_100"][0]) == np.float32(-0.000472238287)
    assert np.float32(kat["f32_64_noshm_100"][0]) == np.float32(-0.000472255051)
    assert np.float32(kat["f32_64_shm_101"][3]) == np.float32(1.0000093)
    from turbulent_lbm_multigpu_b200.skeleton import compute_parameters
    size = (64, 64, 64)
    p = compute_parameters(size, (0.1,) * 3)
    for variant, name in ((0, "shm"), (1, "noshm")):
        s = port.OracleSolver(size, [1] * 6, p.inv_tau, p.gravitation, p.u_lid, variant=variant)
        set_lid(s, size)
        n = s.n
        g = 32 + 32 * 64 + 32 * 64 * 64
        for k in (100, 101):
            while s.simulation_step_counter < k:
                s.simulationStep()
            got = np.array([s.velocity[g], s.velocity[n + g], s.velocity[2 * n + g], s.density[g],
                            s.getVelocityChecksum()], dtype=np.float64)
            assert np.array_equal(got, kat["f32_64_%s_%d" % (name, k)]), (name, k, got)


def test_density_after_alpha_is_signed_zero():
    """`#define tmp rho` (lbm_alpha.cl:173): with g_x = 0 the density stored by an alpha step is +-0."""
    s = port.OracleSolver((16, 8, 8), [1] * 6, 1.9, (0.0, -1e-4, 0.0), 0.02)
    set_lid(s, (16, 8, 8))
    s.simulationStep(); s.simulationStep()
    fluid = s.flags == 2
    assert np.all(s.density[fluid] == 0.0)


def test_skeleton_c_matches_python_mirror_and_known_values():
    from turbulent_lbm_multigpu_b200.skeleton import compute_parameters
    for dtype in (np.float32, np.float64):
        for sx in (16, 64, 96, 128, 256, 512):
            c = port.skeleton(sx, 0.1, dtype=dtype)
            p = compute_parameters((sx, sx, sx), (0.1, 0.1, 0.1), dtype=dtype)
            assert c["tau"] == p.tau and c["inv_tau"] == p.inv_tau and c["d_timestep"] == p.d_timestep
            assert c["gravitation"] == tuple(p.gravitation)
            assert c["drivenCavityVelocity"] == tuple(p.drivenCavityVelocity)
            assert c["d_reynolds"] == p.d_reynolds
    p = compute_parameters((64, 64, 64), (0.1,) * 3)
    assert p.inv_tau == np.float32(1.42278874) and p.u_lid == np.float32(0.0126204686)
    assert p.gravitation[1] == np.float32(-9.99999975e-05)
    p = compute_parameters((16, 16, 16), (0.1,) * 3)
    assert p.tau == np.float32(0.525355637) and p.inv_tau == np.float32(1.90347254)
    assert p.u_lid == np.float32(0.0252409372)


def test_copy_rect_restatement_matches_slicing():
    s = port.OracleSolver((24, 8, 8), [1] * 6, 1.5, (0, -1e-4, 0), 0.01)
    s.dd[:] = np.arange(s.dd.size, dtype=np.float32)
    origin, size = (3, 2, 1), (5, 4, 3)
    out = np.zeros(19 * 60, np.float32)
    for f in range(19):
        s.copy_rect(s.dd, f * s.n, origin, s.size, out, f * 60, (0, 0, 0), size, size)
    assert bits_equal(out, s.storeDensityDistribution(origin, size))


@pytest.mark.parametrize("D,nums,steps", [
    ((32, 16, 16), (2, 1, 1), 40), ((36, 16, 16), (3, 1, 1), 41), ((16, 16, 24), (1, 1, 2), 60),
    ((16, 24, 16), (1, 2, 1), 61), ((24, 24, 12), (2, 2, 1), 60), ((24, 24, 24), (2, 2, 2), 81),
])
@pytest.mark.parametrize("slots", ["reference", "minimal"])
def test_decomposed_run_equals_single_domain(D, nums, steps, slots):
    """The reference's validate criterion (src/main.cpp:309-408) on the oracle; also proves the
    minimal 5-slot payload is equivalent to shipping all 19 slots."""
    L = (0.1, 0.1, 0.1)
    make, p = multi.make_oracle_factory(D, nums, L, variant=1)
    md = multi.MultiDomain(D, nums, make, slots=slots)
    md.run(steps)
    V = multi.validation_domain(D, nums)
    single = port.OracleSolver(V, [1] * 6, p["inv_tau"], p["gravitation"], p["drivenCavityVelocity"][0], variant=1,
                               tau=p["tau"])
    multi.set_lid_geometry(single, V)
    for _ in range(steps):
        single.simulationStep()
    inner = tuple(v - 2 for v in md.sub_size)
    for r in range(md.nranks):
        o = multi.validation_origin(r, nums, md.sub_size)
        assert bits_equal(md.interior(r, "velocity"), single.storeVelocity(o, inner)), r
        assert bits_equal(md.interior(r, "flags"), single.storeFlags(o, inner)), r


@pytest.mark.parametrize("D,nums,steps", [
    ((32, 16, 16), (2, 1, 1), 40), ((24, 24, 12), (2, 2, 1), 41), ((24, 24, 24), (2, 2, 2), 61),
])
@pytest.mark.parametrize("slots", ["reference", "minimal"])
def test_axis_phase_order_does_not_change_the_result(D, nums, steps, slots):
    """The halo a sync delivers does not depend on the order of its three axis phases (every face
    spans the full extent of the other axes): z,y,x -- the product's LBM_AXIS_ORDER_ZYX -- meets
    the reference's validate criterion like x,y,z does, and every non-ghost population equals the
    x,y,z run bit for bit; only leftovers in ghost cells differ."""
    L = (0.1, 0.1, 0.1)
    make, p = multi.make_oracle_factory(D, nums, L, variant=1)
    ref_order = multi.MultiDomain(D, nums, make, slots=slots)
    zyx = multi.MultiDomain(D, nums, make, slots=slots, axis_order=(2, 1, 0))
    ref_order.run(steps)
    zyx.run(steps)
    V = multi.validation_domain(D, nums)
    single = port.OracleSolver(V, [1] * 6, p["inv_tau"], p["gravitation"], p["drivenCavityVelocity"][0], variant=1,
                               tau=p["tau"])
    multi.set_lid_geometry(single, V)
    for _ in range(steps):
        single.simulationStep()
    inner = tuple(v - 2 for v in zyx.sub_size)
    for r in range(zyx.nranks):
        o = multi.validation_origin(r, nums, zyx.sub_size)
        assert bits_equal(zyx.interior(r, "velocity"), single.storeVelocity(o, inner)), r
        assert bits_equal(zyx.interior(r, "dd"), ref_order.interior(r, "dd")), r


# ------------------------------------------------------------------ side by side with oracle/_ref
needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")


@needs_ref
@pytest.mark.parametrize("size,dtype,variant,bc,steps", [
    ((16, 16, 16), np.float32, 0, (1,) * 6, 21),
    ((16, 16, 16), np.float32, 1, (1,) * 6, 21),
    ((40, 24, 16), np.float32, 0, (1,) * 6, 11),
    ((40, 24, 16), np.float32, 1, (1,) * 6, 11),
    ((32, 32, 32), np.float32, 0, (1,) * 6, 8),
    ((16, 16, 16), np.float32, 0, (8,) * 6, 12),
    ((16, 16, 16), np.float32, 0, (8, 1, 1, 8, 8, 1), 12),
    ((136, 8, 8), np.float32, 0, (8, 8, 1, 1, 1, 8), 7),
    ((256, 8, 4), np.float32, 0, (8, 8, 8, 8, 1, 1), 7),
    ((12, 12, 12), np.float32, 1, (1,) * 6, 10),
    ((16, 16, 16), np.float64, 0, (1,) * 6, 9),
    ((40, 24, 16), np.float64, 0, (1,) * 6, 6),
    ((40, 24, 16), np.float64, 1, (1,) * 6, 6),
])
def test_port_equals_reference_kernels_step_by_step(size, dtype, variant, bc, steps):
    from turbulent_lbm_multigpu_b200.skeleton import compute_parameters
    p = compute_parameters(size, (0.1,) * 3, dtype=dtype)
    a = ref.RefSolver(size, list(bc), p.inv_tau, p.gravitation, p.u_lid, dtype=dtype, variant=variant)
    b = port.OracleSolver(size, list(bc), p.inv_tau, p.gravitation, p.u_lid, dtype=dtype, variant=variant)
    for s in (a, b):
        set_lid(s, size)
    for i in range(steps):
        a.simulationStep()
        b.simulationStep()
        for name in ("dd", "velocity", "density", "flags"):
            assert bits_equal(getattr(a, name), getattr(b, name)), (size, variant, i, name)
    assert a.getVelocityChecksum() == b.getVelocityChecksum()


@needs_ref
@pytest.mark.parametrize("dtype,variant", [(np.float32, 0), (np.float32, 1), (np.float64, 0)])
def test_port_equals_reference_kernels_with_obstacles_inside(dtype, variant):
    """An obstacle block, scattered obstacle cells, ghost cells inside the domain and a few
    velocity-injection cells away from the lid plane: every flag branch of both kernels
    (lbm_alpha.cl:177-343, lbm_beta.cl:495-656) next to fluid, bit for bit against the reference."""
    from turbulent_lbm_multigpu_b200.skeleton import compute_parameters
    size = (40, 24, 16)
    p = compute_parameters(size, (0.1,) * 3, dtype=dtype)
    a = ref.RefSolver(size, [1, 8, 1, 1, 8, 1], p.inv_tau, p.gravitation, p.u_lid, dtype=dtype, variant=variant)
    b = port.OracleSolver(size, [1, 8, 1, 1, 8, 1], p.inv_tau, p.gravitation, p.u_lid, dtype=dtype, variant=variant)
    rng = np.random.default_rng(11)
    speckle = np.where(rng.random(10 * 8 * 6) < 0.15, 1, 2).astype(np.int32)
    for s in (a, b):
        set_lid(s, size)
        s.setFlags(np.ones(6 * 5 * 4, np.int32), (9, 7, 5), (6, 5, 4))
        s.setFlags(speckle, (24, 4, 3), (10, 8, 6))
        s.setFlags(np.full(3, 4, np.int32), (20, 10, 8), (3, 1, 1))
        s.setFlags(np.full(4, 8, np.int32), (30, 15, 10), (2, 2, 1))
    for i in range(14):
        a.simulationStep()
        b.simulationStep()
        for name in ("dd", "velocity", "density", "flags"):
            assert bits_equal(getattr(a, name), getattr(b, name)), (variant, i, name)


@needs_ref
def test_reference_rect_kernels_equal_port_slicing():
    from turbulent_lbm_multigpu_b200.skeleton import compute_parameters
    size = (24, 20, 12)
    p = compute_parameters(size, (0.1,) * 3)
    a = ref.RefSolver(size, [1] * 6, p.inv_tau, p.gravitation, p.u_lid)
    b = port.OracleSolver(size, [1] * 6, p.inv_tau, p.gravitation, p.u_lid)
    for s in (a, b):
        set_lid(s, size)
        for _ in range(3):
            s.simulationStep()
    origin, rect = (3, 2, 1), (7, 5, 4)
    assert bits_equal(a.storeDensityDistribution(origin, rect), b.storeDensityDistribution(origin, rect))
    assert bits_equal(a.storeVelocity(origin, rect), b.storeVelocity(origin, rect))
    assert bits_equal(a.storeFlags(origin, rect), b.storeFlags(origin, rect))
    src = np.random.default_rng(1).random(19 * 140).astype(np.float32)
    for norm in (None, (1, 0, 0), (0, 0, -1)):
        a.setDensityDistribution(src, origin, rect, norm)
        b.setDensityDistribution(src, origin, rect, norm)
        assert bits_equal(a.dd, b.dd)

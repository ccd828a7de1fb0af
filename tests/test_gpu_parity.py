"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle, bit for bit.

Tolerance: NONE.  The kernels are compiled with --fmad=false in the reference's operation
order, the oracle with -ffp-contract=off, so fp32 and fp64 populations, velocities, densities
and flags must be bit-identical after every step (north-star tolerance 1e-5 fp32 / 1e-12 fp64
is met with zero error).
"""
import numpy as np
import pytest

from helpers import WALLS, assert_state_equal, bits_equal, make_cuda, make_oracle

pytestmark = pytest.mark.gpu

GHOSTS = (8, 8, 8, 8, 8, 8)


@pytest.mark.parametrize("size,dtype,order,steps", [
    ((16, 16, 16), np.float32, 0, 21),      # wg % sx == 0: the work-group x-shift quirk is live
    ((16, 16, 16), np.float32, 1, 21),
    ((32, 32, 32), np.float32, 0, 12),
    ((40, 24, 16), np.float32, 0, 11),      # non power of two, no quirk
    ((40, 24, 16), np.float32, 1, 11),
    ((24, 20, 12), np.float32, 0, 10),
    ((64, 64, 64), np.float32, 0, 10),
    ((136, 8, 8), np.float32, 0, 7),
    ((15, 9, 7), np.float32, 1, 9),         # odd sizes: scalar kernels, N % 128 != 0
    ((18, 10, 6), np.float32, 1, 9),        # vector width 2
    ((16, 16, 16), np.float64, 0, 9),
    ((40, 24, 16), np.float64, 0, 8),
    ((40, 24, 16), np.float64, 1, 8),
    ((15, 9, 7), np.float64, 1, 7),
])
def test_step_by_step_bit_exact(size, dtype, order, steps):
    c = make_cuda(size, dtype, order=order)
    o = make_oracle(size, dtype, order=order)
    assert_state_equal(c, o, ctx="after init")
    for i in range(steps):
        c.simulationStep()
        o.simulationStep()
        assert_state_equal(c, o, ctx="%s %s order %d step %d" % (size, np.dtype(dtype).name, order, i))
    assert c.simulation_step_counter == steps


@pytest.mark.parametrize("vw", [1, 2, 4])
@pytest.mark.parametrize("block", [64, 128, 256])
def test_vector_widths_and_block_sizes(vw, block):
    size = (48, 20, 12)
    c = make_cuda(size, np.float32, vector_width=vw, block_size=block)
    assert c.config()["vector_width"] == vw
    o = make_oracle(size, np.float32)
    for i in range(8):
        c.simulationStep()
        o.simulationStep()
    assert_state_equal(c, o, ctx="vw %d block %d" % (vw, block))


@pytest.mark.parametrize("bc", [GHOSTS, (8, 1, 1, 8, 8, 1), (1, 8, 8, 1, 1, 8)])
@pytest.mark.parametrize("size", [(16, 16, 16), (40, 24, 16)])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_ghost_layer_faces(bc, size, dtype):
    c = make_cuda(size, dtype, bc=bc)
    o = make_oracle(size, dtype, bc=bc)
    for i in range(10):
        c.simulationStep()
        o.simulationStep()
        assert_state_equal(c, o, ctx="bc %s step %d" % (bc, i))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("cs", [0.1, 0.17])
def test_smagorinsky_matches_its_specification(dtype, cs):
    """No reference counterpart: the oracle's restatement is the specification."""
    size = (40, 24, 16)
    c = make_cuda(size, dtype, cs=cs)
    o = make_oracle(size, dtype, cs=cs)
    for i in range(12):
        c.simulationStep()
        o.simulationStep()
        assert_state_equal(c, o, ctx="smagorinsky cs=%g step %d" % (cs, i))


def test_smagorinsky_zero_is_bgk():
    size = (40, 24, 16)
    a = make_cuda(size, np.float32, cs=0.0)
    o = make_oracle(size, np.float32, cs=0.0)
    for i in range(6):
        a.simulationStep()
        o.simulationStep()
    assert_state_equal(a, o)


def test_store_flags_off_leaves_populations_identical():
    size = (40, 24, 16)
    a = make_cuda(size, np.float32, store=False)
    o = make_oracle(size, np.float32)
    for i in range(9):
        a.simulationStep()
        o.simulationStep()
    assert_state_equal(a, o, what=("dd", "flags"))


def test_obstacles_inside_the_domain():
    size = (40, 24, 16)
    rng = np.random.default_rng(7)
    blk = (6, 5, 4)
    obst = np.ones(blk[0] * blk[1] * blk[2], np.int32)
    c = make_cuda(size, np.float32)
    o = make_oracle(size, np.float32)
    for s in (c, o):
        s.setFlags(obst, (9, 7, 5), blk)
        s.setFlags(np.full(3, 4, np.int32), (20, 10, 8), (3, 1, 1))
    for i in range(15):
        c.simulationStep()
        o.simulationStep()
    assert_state_equal(c, o)
    del rng


def test_rect_access_round_trips():
    size = (24, 20, 12)
    c = make_cuda(size, np.float32)
    o = make_oracle(size, np.float32)
    for i in range(5):
        c.simulationStep()
        o.simulationStep()
    origin, rect = (3, 2, 1), (7, 5, 4)
    assert bits_equal(c.storeDensityDistribution(origin=origin, size=rect), o.storeDensityDistribution(origin, rect))
    assert bits_equal(c.storeVelocity(origin=origin, size=rect), o.storeVelocity(origin, rect))
    assert bits_equal(c.storeDensity(origin=origin, size=rect), o.storeDensity(origin, rect))
    assert bits_equal(c.storeFlags(origin=origin, size=rect), o.storeFlags(origin, rect))
    rng = np.random.default_rng(3)
    src = rng.random(19 * 7 * 5 * 4).astype(np.float32)
    for norm in (None, (1, 0, 0), (0, -1, 0), (0, 0, 1)):
        c.setDensityDistribution(src, origin, rect, norm)
        o.setDensityDistribution(src, origin, rect, norm)
        assert bits_equal(c.storeDensityDistribution(), o.dd)
    v = rng.random(3 * 7 * 5 * 4).astype(np.float32)
    c.setVelocity(v, origin, rect)
    assert bits_equal(c.storeVelocity(origin=origin, size=rect), v)
    r = rng.random(7 * 5 * 4).astype(np.float32)
    c.setDensity(r, origin, rect)
    assert bits_equal(c.storeDensity(origin=origin, size=rect), r)


def test_checksum_host_order_and_device_reduction():
    size = (32, 32, 32)
    c = make_cuda(size, np.float32)
    o = make_oracle(size, np.float32)
    for i in range(20):
        c.simulationStep()
        o.simulationStep()
    exact = o.getVelocityChecksum()
    assert c.getVelocityChecksum(host_order=True) == exact
    dev = c.getVelocityChecksum(host_order=False)
    n = o.n
    fluid = o.flags == 2
    ref64 = float(np.sum(((o.velocity[:n] + o.velocity[n:2 * n]) + o.velocity[2 * n:])[fluid].astype(np.float64)))
    assert abs(dev - ref64) <= 1e-9 * max(1.0, abs(ref64))


def test_reset_and_driven_cavity_value():
    size = (16, 16, 16)
    c = make_cuda(size, np.float32)
    o = make_oracle(size, np.float32)
    for i in range(4):
        c.simulationStep()
    c.reset()
    from helpers import set_lid
    set_lid(c, size)
    assert c.simulation_step_counter == 0
    for i in range(4):
        c.simulationStep()
        o.simulationStep()
    assert_state_equal(c, o)


def test_known_answer_64cubed_100_steps():
    """SURVEY.md A.2 / BASELINE.md §4: values produced by the reference kernels as shipped."""
    c = make_cuda((64, 64, 64), np.float32)
    c.simulationSteps(100)
    n = 64 ** 3
    g = 32 + 32 * 64 + 32 * 64 * 64
    v = c.storeVelocity()
    assert np.float32(v[g]) == np.float32(-0.000472238287)
    assert np.float32(v[n + g]) == np.float32(-9.76771116e-06)
    assert np.float32(v[2 * n + g]) == np.float32(-3.35276127e-08)
    c.simulationStep()
    v = c.storeVelocity()
    assert np.float32(v[g]) == np.float32(-0.000478317961)
    assert np.float32(c.storeDensity()[g]) == np.float32(1.0000093)


@pytest.mark.parametrize("size,dtype,order", [((40, 24, 16), np.float32, 0), ((16, 16, 16), np.float32, 0),
                                               ((40, 24, 16), np.float64, 1)])
def test_padded_slot_stride_is_invisible(size, dtype, order, monkeypatch):
    """The library pads the slot stride of dd for the big power-of-two domains (L2-slice
    aliasing); forced on here: steps, whole-array and rect access, halo pack/unpack must be
    unaffected (host buffers stay dense [19][cells])."""
    monkeypatch.setenv("LBM_B200_SLOT_PAD_BYTES", "4352")
    c = make_cuda(size, dtype, order=order)
    import ctypes
    stride = ctypes.c_size_t(0)
    assert c._lib.lbmGetSlotStride(c.handle, ctypes.byref(stride)) == 0
    assert stride.value == int(np.prod(size)) + 4352 // np.dtype(dtype).itemsize
    o = make_oracle(size, dtype, order=order)
    for i in range(7):
        c.simulationStep()
        o.simulationStep()
        assert_state_equal(c, o, ctx="padded step %d" % i)
    origin, rect = (2, 1, 3), (5, 4, 6)
    block = c.storeDensityDistribution(origin=origin, size=rect).reshape(19, rect[2], rect[1], rect[0])
    full = o.dd.reshape(19, size[2], size[1], size[0])
    assert bits_equal(block, full[:, 3:9, 1:5, 2:7])
    new = np.arange(19 * 5 * 4 * 6, dtype=dtype) * dtype(1e-3)
    c.setDensityDistribution(new, origin, rect)
    got = c.storeDensityDistribution().reshape(19, size[2], size[1], size[0])
    full2 = full.copy()
    full2[:, 3:9, 1:5, 2:7] = new.reshape(19, 6, 4, 5)
    assert bits_equal(got, full2)
    c.setDensityDistribution(o.dd)          # whole-array upload through the pitched copy
    assert bits_equal(c.storeDensityDistribution(), o.dd)


def test_debug_print_dumps_the_device_arrays():
    """reference src/CLbmSolver.hpp:1032-1101: debug_print / debugDD show what the store* calls return."""
    import io
    from turbulent_lbm_multigpu_b200 import debug
    c = make_cuda((16, 16, 16), np.float32)
    for _ in range(3):
        c.simulationStep()
    buf = io.StringIO()
    c.debug_print(file=buf)
    assert buf.getvalue() == debug.debug_print(c.storeDensityDistribution(), c.storeVelocity(), c.storeDensity(),
                                               c.storeFlags())
    assert buf.getvalue().count("\n0: ") == 4 and "FLAGS:\n0: 1 0 0 0 1 0 0 0 " in buf.getvalue()
    buf = io.StringIO()
    c.debugDD(18, 16, 64, file=buf)
    rest = c.storeDensityDistribution().reshape(19, -1)[18]
    assert buf.getvalue().split()[1] == "%.4f" % float(rest[0]) and buf.getvalue().startswith("%d: " % (18 * 4096 // 16))
    c.close()

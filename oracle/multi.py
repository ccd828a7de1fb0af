"""numpy restatement of the reference's decomposition and halo exchange -- TEST INFRASTRUCTURE ONLY.

Follows, with file:line, the reference's
  * CManager::setSubdomainNums / initSimulation (src/CManager.hpp:49-204): sub-domain size
    D/n (ghost layers INSIDE that size), rank id = nx + ny*NX + nz*NX*NY, BC table, the up
    to six CComm descriptors in the fixed order x-,x+,y-,y+,z-,z+, lid geometry on the
    top-y ranks;
  * CController::setGeometry (src/CController.hpp:531-546);
  * CController::computeNextStep / syncAlpha / syncBeta (src/CController.hpp:265-391):
    after a beta step the GHOST layer is sent back and only slots with direction.e_f > 0 are
    written into the receiver's outermost real layer; after an alpha step the outermost real
    layer is copied (all 19 slots) into the neighbour's ghost layer;
  * the validate-mode mapping of src/main.cpp:332-337,358-387.

The reference runs one MPI rank per sub-domain and walks its CComm list sequentially with
blocking pairwise exchanges.  Within one axis the layers read ({1,S-2} alpha / {0,S-1} beta)
and written ({0,S-1} / {1,S-2}) are disjoint, while later axes read the rims written by
earlier ones, so that schedule is equivalent to "for axis in x,y,z: every rank packs the
faces of that axis; then every rank unpacks" -- which is what ``sync`` does.

``slots='reference'`` ships all 19 slots on alpha sync like the reference;
``slots='minimal'`` ships only the 5 slots a neighbour ever consumes (alpha: e_f.dir < 0,
beta: e_f.dir > 0) -- the payload the product uses; tests prove both give bit-identical
real cells.
"""
from __future__ import annotations

import numpy as np

from .port import LBM_UNITS, OracleSolver

FLAG_OBSTACLE, FLAG_FLUID, FLAG_VELOCITY_INJECTION, FLAG_GHOST_LAYER = 1, 2, 4, 8


class Comm:
    """CComm<T> (src/CComm.hpp:8-79)."""

    def __init__(self, dst, send_size, recv_size, send_origin, recv_origin, direction):
        self.dst = dst
        self.send_size, self.recv_size = tuple(send_size), tuple(recv_size)
        self.send_origin, self.recv_origin = tuple(send_origin), tuple(recv_origin)
        self.direction = tuple(direction)

    def as_tuple(self):
        return (self.dst, self.send_size, self.recv_size, self.send_origin, self.recv_origin, self.direction)


def decompose(domain_size, subdomain_nums):
    """CManager::setSubdomainNums (src/CManager.hpp:49-63)."""
    for d, n in zip(domain_size, subdomain_nums):
        if d % n != 0:
            raise ValueError("Number of subdomains does not match with the grid size!")
    return tuple(d // n for d, n in zip(domain_size, subdomain_nums))


def rank_layout(rank, subdomain_nums, sub_size):
    """CManager::initSimulation (src/CManager.hpp:78-199): coords, BC[3][2], CComm list."""
    NX, NY, NZ = subdomain_nums
    S = sub_size
    rid = max(rank, 0)
    nx = rid % NX
    ny = (rid // NX) % NY
    nz = rid // (NX * NY)
    coords = (nx, ny, nz)
    bc = [[FLAG_GHOST_LAYER, FLAG_GHOST_LAYER] for _ in range(3)]
    for a, (c, n) in enumerate(zip(coords, subdomain_nums)):
        if c == 0:
            bc[a][0] = FLAG_OBSTACLE
        if c == n - 1:
            bc[a][1] = FLAG_OBSTACLE
    stride = (1, NX, NX * NY)
    comms = []
    for a in range(3):
        face = [S[0], S[1], S[2]]
        face[a] = 1
        if bc[a][0] == FLAG_GHOST_LAYER:
            so, ro, d = [0, 0, 0], [0, 0, 0], [0, 0, 0]
            so[a], ro[a], d[a] = 1, 0, 1
            comms.append(Comm(rid - stride[a], face, face, so, ro, d))
        if bc[a][1] == FLAG_GHOST_LAYER:
            so, ro, d = [0, 0, 0], [0, 0, 0], [0, 0, 0]
            so[a], ro[a], d[a] = S[a] - 2, S[a] - 1, -1
            comms.append(Comm(rid + stride[a], face, face, so, ro, d))
    origin = tuple(c * s for c, s in zip(coords, S))
    return coords, bc, comms, origin


def _dot(a, b):
    return sum(int(x) * int(y) for x, y in zip(a, b))


class MultiDomain:
    """All ranks of a decomposed run inside one process (each rank = one solver object)."""

    def __init__(self, domain_size, subdomain_nums, make_solver, slots="reference", axis_order=(0, 1, 2)):
        self.domain_size = tuple(domain_size)
        self.nums = tuple(subdomain_nums)
        self.sub_size = decompose(domain_size, subdomain_nums)
        self.nranks = self.nums[0] * self.nums[1] * self.nums[2]
        self.slots = slots
        # phase order of a sync: (0, 1, 2) is the reference's CComm walk (src/CManager.hpp:122-199);
        # (2, 1, 0) is the product's LBM_AXIS_ORDER_ZYX option -- same halo, different leftovers in ghost cells
        self.axis_order = tuple(axis_order)
        self.ranks = []
        for r in range(self.nranks):
            coords, bc, comms, origin = rank_layout(r, self.nums, self.sub_size)
            bc6 = [bc[a][s] for a in range(3) for s in range(2)]
            solver = make_solver(r, self.sub_size, bc6)
            if coords[1] == self.nums[1] - 1:  # CManager.hpp:200-202 -> setGeometry
                set_lid_geometry(solver, self.sub_size)
            self.ranks.append(dict(coords=coords, bc=bc, comms=comms, origin=origin, solver=solver))

    def sync(self, beta):
        S = self.sub_size
        for axis in self.axis_order:
            staged = []
            for r, rk in enumerate(self.ranks):
                for c in rk["comms"]:
                    if c.direction[axis] == 0:
                        continue
                    if beta:   # CController::syncBeta :337-341 send/recv roles swapped
                        s_origin, s_size = c.recv_origin, c.recv_size
                    else:
                        s_origin, s_size = c.send_origin, c.send_size
                    buf = rk["solver"].storeDensityDistribution(s_origin, s_size)
                    staged.append((r, c, buf))
            for r, c, buf in staged:
                dst = self.ranks[c.dst]
                back = next(k for k in dst["comms"] if k.dst == r and k.direction[axis] == -c.direction[axis])
                if beta:
                    r_origin, r_size, norm = back.send_origin, back.send_size, back.direction
                    if self.slots == "minimal":
                        buf = _mask(buf, r_size, lambda f: _dot(norm, LBM_UNITS[f]) > 0, dst["solver"], r_origin)
                        dst["solver"].setDensityDistribution(buf, r_origin, r_size)
                    else:
                        dst["solver"].setDensityDistribution(buf, r_origin, r_size, norm)
                else:
                    r_origin, r_size = back.recv_origin, back.recv_size
                    if self.slots == "minimal":
                        norm = back.direction
                        buf = _mask(buf, r_size, lambda f: _dot(norm, LBM_UNITS[f]) < 0, dst["solver"], r_origin)
                    dst["solver"].setDensityDistribution(buf, r_origin, r_size)

    def step(self):
        """CController::computeNextStep (src/CController.hpp:385-391) on every rank."""
        for rk in self.ranks:
            rk["solver"].simulationStep()
        counter = self.ranks[0]["solver"].simulation_step_counter
        self.sync(beta=bool(counter & 1))

    def run(self, loops):
        for _ in range(loops):
            self.step()

    def interior(self, rank, what="velocity"):
        """validate mode: block origin (1,1,1), size S-2 (src/main.cpp:332-337)."""
        s = self.ranks[rank]["solver"]
        size = tuple(v - 2 for v in self.sub_size)
        fn = {"velocity": s.storeVelocity, "density": s.storeDensity, "flags": s.storeFlags,
              "dd": s.storeDensityDistribution}[what]
        return fn((1, 1, 1), size)


def _mask(buf, size, keep, solver, origin):
    """Replace the slots a minimal exchange would not ship by the receiver's current values."""
    cur = solver.storeDensityDistribution(origin, size).reshape(19, -1)
    new = np.asarray(buf).reshape(19, -1).copy()
    for f in range(19):
        if not keep(f):
            new[f] = cur[f]
    return new.reshape(-1)


def set_lid_geometry(solver, size):
    """CController::setGeometry (src/CController.hpp:531-546)."""
    origin = (1, size[1] - 2, 1)
    rect = (size[0] - 2, 1, size[2] - 2)
    solver.setFlags(np.full(rect[0] * rect[2], FLAG_VELOCITY_INJECTION, np.int32), origin, rect)


def validation_domain(domain_size, subdomain_nums):
    """Single domain equivalent to the decomposed run (src/main.cpp:358-361)."""
    return tuple(d - 2 * (n - 1) for d, n in zip(domain_size, subdomain_nums))


def validation_origin(rank, subdomain_nums, sub_size):
    """Where rank's interior block sits in the validation domain (src/main.cpp:374-386)."""
    NX, NY, _ = subdomain_nums
    nx, ny, nz = rank % NX, (rank // NX) % NY, rank // (NX * NY)
    return (1 + nx * (sub_size[0] - 2), 1 + ny * (sub_size[1] - 2), 1 + nz * (sub_size[2] - 2))


def make_oracle_factory(domain_size, subdomain_nums, domain_length, dtype=np.float32, variant=1,
                        smagorinsky_cs=0.0, solver_cls=OracleSolver, **phys):
    """Sub-domain solvers parametrised like CManager/CLbmSolver do: length = L/n per axis
    (src/CManager.hpp:66-69) and cell length = length_x / S_x (src/CLbmSkeleton.hpp:157)."""
    from .port import skeleton
    sub = decompose(domain_size, subdomain_nums)
    sub_len_x = np.dtype(dtype).type(domain_length[0]) / np.dtype(dtype).type(subdomain_nums[0])
    p = skeleton(sub[0], sub_len_x, dtype=dtype, **phys)
    assert not p["error"], "tau outside [0.51, 2.5]"

    def make(rank, size, bc6):
        kw = {}
        if solver_cls is OracleSolver:
            kw = dict(tau=p["tau"], smagorinsky_cs=smagorinsky_cs)
        return solver_cls(size, bc6, p["inv_tau"], p["gravitation"], p["drivenCavityVelocity"][0],
                          dtype=dtype, variant=variant, **kw)
    return make, p

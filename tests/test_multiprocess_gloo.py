"""N > 1 host path on CPU: world_size 2 and 4 `gloo` runs of the product's CManager/CController
(host-staged reference sync over torch.distributed) against the single-domain oracle run -- the
reference's validate criterion (src/main.cpp:309-408), bit-exact."""
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import bits_equal
from oracle import multi, port

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("D,nums,steps", [
    ((32, 16, 16), (2, 1, 1), 21),
    ((16, 16, 24), (1, 1, 2), 20),
    ((24, 24, 12), (2, 2, 1), 21),
])
def test_gloo_ranks_equal_single_domain(tmp_path, D, nums, steps):
    world = nums[0] * nums[1] * nums[2]
    env = dict(os.environ, LBM_TEST_DOMAIN=",".join(map(str, D)), LBM_TEST_NUMS=",".join(map(str, nums)),
               LBM_TEST_STEPS=str(steps), LBM_TEST_OUT=str(tmp_path), OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(HERE, "_gloo_worker.py")]
    r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]

    L = (0.1, 0.1, 0.1)
    make, p = multi.make_oracle_factory(D, nums, L, variant=1)
    V = multi.validation_domain(D, nums)
    single = port.OracleSolver(V, [1] * 6, p["inv_tau"], p["gravitation"], p["drivenCavityVelocity"][0], variant=1,
                               tau=p["tau"])
    multi.set_lid_geometry(single, V)
    for _ in range(steps):
        single.simulationStep()
    # the oracle's own in-process decomposed run: every population must agree too
    md = multi.MultiDomain(D, nums, make, slots="reference")
    md.run(steps)
    sub = multi.decompose(D, nums)
    inner = tuple(v - 2 for v in sub)
    for r_ in range(world):
        z = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r_))
        o = multi.validation_origin(r_, nums, sub)
        assert bits_equal(z["velocity"], single.storeVelocity(o, inner)), r_
        assert bits_equal(z["flags"], single.storeFlags(o, inner)), r_
        assert bits_equal(z["dd"], md.ranks[r_]["solver"].dd), r_
        assert list(z["comms"]) == [c.dst for c in md.ranks[r_]["comms"]]

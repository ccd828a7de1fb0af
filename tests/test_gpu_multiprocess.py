"""One process per GPU (torchrun), the layout bench.py measures: every halo transport of the
product -- one-sided NVLink peer stores over CUDA IPC ("p2p"), NCCL send/recv with and without
the shell/interior overlap ("overlap", "device"), and the reference's host-staged algorithm
("host") -- must reproduce the oracle's decomposed run bit for bit on EVERY population of every
rank (ghost layers included).  Needs >= 2 GPUs on the box; skipped otherwise."""
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import bits_equal
from oracle import multi as omulti

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _gpus():
    import ctypes
    from turbulent_lbm_multigpu_b200 import capi
    n = ctypes.c_int(0)
    capi.load().lbmGetDeviceCount(ctypes.byref(n))
    return n.value


def _free_port():
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run_ranks(tmp_path, D, nums, steps, sync, axis_order="", order="linear", cs=0.0, dtype="f32", share=False,
               graph=False, block=0):
    world = nums[0] * nums[1] * nums[2]
    if _gpus() < world and not share:
        pytest.skip("needs %d GPUs" % world)
    env = dict(os.environ, LBM_TEST_DOMAIN=",".join(map(str, D)), LBM_TEST_NUMS=",".join(map(str, nums)),
               LBM_TEST_STEPS=str(steps), LBM_TEST_SYNC=sync, LBM_TEST_OUT=str(tmp_path),
               LBM_TEST_AXIS_ORDER=axis_order, LBM_TEST_BETA_ORDER=order, LBM_TEST_CS=repr(cs), LBM_TEST_DTYPE=dtype, LBM_TEST_GRAPH="1" if graph else "0",
               LBM_TEST_BLOCK=str(block))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(HERE, "_gpu_rank_worker.py")]
    r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:]
    npdt = np.float64 if dtype == "f64" else np.float32
    make, po = omulti.make_oracle_factory(D, nums, (0.1, 0.1, 0.1), dtype=npdt, variant=0 if order == "shipped" else 1,
                                          smagorinsky_cs=cs)
    md = omulti.MultiDomain(D, nums, make, slots="reference" if sync == "host" else "minimal",
                            axis_order=(2, 1, 0) if axis_order == "zyx" else (0, 1, 2))
    md.run(steps)
    for rank in range(world):
        z = np.load(os.path.join(str(tmp_path), "rank%d.npz" % rank))
        o = md.ranks[rank]["solver"]
        assert bits_equal(z["flags"], o.flags), rank
        assert bits_equal(z["dd"], o.dd), (rank, sync)


@pytest.mark.parametrize("D,nums,steps,axis_order,dtype", [
    ((40, 24, 8 * 12), (1, 1, 8), 21, "", "f32"),        # the driver's 8-GPU scaling layout (z-slabs)
    ((48, 40, 24), (2, 2, 2), 21, "zyx", "f32"),         # BASELINE configs[3] block layout
    ((48, 40, 48), (1, 2, 4), 20, "", "f64"),            # pencils, fp64 (configs[4] dtype)
])
def test_eight_ranks_bench_defaults_equal_oracle(tmp_path, D, nums, steps, axis_order, dtype):
    """8 processes / 8 GPUs with what bench.py runs: p2p transport, SHIPPED order, C_s = 0.1."""
    _run_ranks(tmp_path, D, nums, steps, "p2p", axis_order, order="shipped", cs=0.1, dtype=dtype)


@pytest.mark.parametrize("D,nums,steps,sync,axis_order", [
    ((40, 24, 32), (1, 1, 2), 11, "p2p", ""),
    ((80, 24, 16), (2, 1, 1), 10, "p2p", "zyx"),          # rows no multiple of a block: separate x push / pull kernels
    ((384, 24, 16), (2, 1, 1), 11, "p2p", "zyx:32"),      # fused x exchange (rows of 192 cells, 32-thread blocks)
    ((40, 24, 32), (1, 1, 2), 6, "host", ""),
])
def test_two_processes_sharing_one_gpu(tmp_path, D, nums, steps, sync, axis_order):
    """The multi-process path (CUDA IPC mapped receive blocks, device-side flags) on whatever the box
    has -- two processes time-slice one GPU when there is only one.  Bench defaults otherwise."""
    axis_order, _, block = axis_order.partition(":")
    _run_ranks(tmp_path, D, nums, steps, sync, axis_order, order="shipped", cs=0.1, share=True, block=int(block or 0))


@pytest.mark.parametrize("D,nums,steps,axis_order", [
    ((40, 24, 32), (1, 1, 2), 104, ""),          # 51 replays of the captured beta+alpha cycle, z-slabs
    ((384, 24, 16), (2, 1, 1), 104, "zyx"),      # x faces exchanged from inside the step kernels (32-thread blocks)
])
def test_cuda_graph_replay_keeps_synchronising(tmp_path, D, nums, steps, axis_order):
    """`bench.py --graph` with the p2p transport: the captured 2-step cycle replayed 51 times must equal the
    oracle's decomposed run -- the halo sequence numbers are counted on the device, not frozen in the graph."""
    _run_ranks(tmp_path, D, nums, steps, "p2p", axis_order, order="shipped", cs=0.1, share=True, graph=True,
               block=32 if axis_order == "zyx" else 0)


@pytest.mark.parametrize("D,nums,steps,sync", [
    ((32, 32, 48), (1, 1, 2), 21, "p2p"), ((32, 32, 48), (1, 1, 2), 21, "overlap"),
    ((32, 32, 48), (1, 1, 2), 20, "device"), ((32, 32, 48), (1, 1, 2), 21, "host"),
    ((96, 32, 32), (2, 1, 1), 20, "p2p"), ((96, 32, 32), (2, 1, 1), 21, "overlap"),
    ((32, 32, 64), (1, 2, 2), 15, "p2p"),
    ((96, 32, 32), (2, 1, 1), 21, "p2p:zyx"),      # x faces exchanged after the interior kernel
    ((64, 32, 64), (2, 1, 2), 15, "p2p:zyx"),
])
def test_torchrun_ranks_equal_oracle(tmp_path, D, nums, steps, sync):
    sync, _, axis_order = sync.partition(":")
    _run_ranks(tmp_path, D, nums, steps, sync, axis_order)

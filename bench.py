#!/usr/bin/env python3
"""bench.py -- MLUPS of the D3Q19 alpha/beta time step on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one lattice-Boltzmann time step (one alpha or beta pass over every cell of the
sub-domain, plus the halo exchange when N > 1).  Workload (BASELINE.json configs[1]):
lid-driven cavity, 256^3 cells per GPU, D3Q19, fp32, Smagorinsky LES (C_s = 0.1), conf.xml
physics; for N > 1 the domain is z-slab decomposed (1,1,N) with 256^3 cells per GPU
(weak scaling), ghost layers inside the sub-domain size as in the reference.
MLUPS = steps * prod(global domain size) / seconds / 1e6 (reference src/CController.hpp:451-452).

Prints ONE JSON line (rank 0).  `--impl reference` times the reference's own kernels on the
host CPU cores (oracle/_ref, built from /root/reference by oracle/build_ref.py) instead.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SIZE = 256
CS = 0.1
BYTES_PER_LUP_F32 = 2 * 19 * 4 + 4      # 156 B fp32, 308 B fp64


PRESETS = {
    "1": dict(size=256, scaling="weak", decomp="slab", dtype="f32"),
    "2": dict(size=512, scaling="strong", decomp="slab", dtype="f32"),
    "3-slab": dict(size=512, scaling="weak", decomp="slab", dtype="f32"),
    "3-block": dict(size=512, scaling="weak", decomp="block", dtype="f32"),
    "3-pencil": dict(size=512, scaling="weak", decomp="pencil", dtype="f32"),
    "4": dict(size=384, scaling="strong", decomp="slab", dtype="f64"),
    # /root/reference/benchmark.py:62-83 (weak: x = 1024 n, y = 1024, z = 32, -X n) and :107-129 (strong 1024x1024x32)
    "recipe-weak": dict(size=1024, scaling="weak", decomp="slab-x", dtype="f32", shape=(1024, 1024, 32)),
    "recipe-strong": dict(size=1024, scaling="strong", decomp="slab-x", dtype="f32", shape=(1024, 1024, 32)),
}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle reasons DURING the timed region.  NVML is polled from a thread every
    few ms (the timed region of the default run is < 0.1 s, too short for `nvidia-smi -lms`);
    falls back to one nvidia-smi query if the NVML binding is unavailable."""
    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.samples = []          # (t, sm_mhz, reasons_bitmask, power_w)
        self.mark_t = None
        self.stop_flag = False
        self.thread = None
        self.sm_max = None
        self.backend = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.gpu
            if vis:
                try:
                    idx = int(vis.split(",")[self.gpu])
                except Exception:
                    idx = self.gpu
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nv = pynvml
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.backend = "nvml"
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception:
            self.backend = "nvidia-smi"

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                try:
                    pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                except Exception:
                    pw = None
                self.samples.append((time.perf_counter(), sm, rs, pw))
            except Exception:
                pass
            time.sleep(0.004)

    def mark(self):
        """start of the timed region (samples before it belong to the warm-up)"""
        self.mark_t = time.perf_counter()

    def stop(self):
        end_t = time.perf_counter()
        self.stop_flag = True
        if self.backend == "nvml":
            self.thread.join(timeout=1.0)
            timed = [x for x in self.samples if self.mark_t is None or self.mark_t <= x[0] <= end_t]
            use = timed if len(timed) >= 3 else self.samples
            reasons = set()
            for _, _, rs, _ in use:
                for name, bit in self.BAD.items():
                    if rs & bit:
                        reasons.add(name)
            pw = [x[3] for x in use if x[3] is not None]
            return {"sm_mhz": float(np.median([x[1] for x in use])) if use else None, "sm_max_mhz": self.sm_max,
                    "samples": len(use), "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons),
                    "source": "nvml polled every 4 ms during the timed region"}
        try:
            q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=10).stdout.strip().split(",")
            reasons = [n for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), out[2:6])
                       if v.strip().lower().startswith("active")]
            return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "samples": 1, "reasons": reasons,
                    "source": "nvidia-smi after the timed region"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": []}


def decomposition(n_gpus, decomp):
    """(1,1,n) z-slabs (contiguous faces), the reference's x split, y-z pencils
    (faces made of whole x rows: no strided x faces) or 3-D blocks."""
    if decomp == "slab-x":
        return (n_gpus, 1, 1)
    if decomp == "pencil":
        return {1: (1, 1, 1), 2: (1, 1, 2), 4: (1, 2, 2), 8: (1, 2, 4)}[n_gpus]
    if decomp == "block":
        return {1: (1, 1, 1), 2: (1, 1, 2), 4: (1, 2, 2), 8: (2, 2, 2)}[n_gpus]
    return (1, 1, n_gpus)


def scenario(n_gpus, size=SIZE, scaling="weak", decomp="slab", cs=CS, shape=None):
    from turbulent_lbm_multigpu_b200.configuration import CConfiguration
    cfg = CConfiguration()
    nums = decomposition(n_gpus, decomp)
    base = tuple(shape) if shape else (size, size, size)
    if scaling == "weak":
        cfg.domain_size = tuple(b * k for b, k in zip(base, nums))
        # weak scaling keeps the cell length (hence tau, u_lid) constant: benchmark.py:72
        cfg.domain_length = tuple(0.1 * k for k in nums)
    else:
        cfg.domain_size = base
        cfg.domain_length = (0.1, 0.1, 0.1)
    cfg.subdomain_num = nums
    cfg.smagorinsky_constant = cs
    cfg.loops = 0
    return cfg


# ====================================================================== reference arm
def cpu_reference(size, budget_s=12.0, threads=None, dtype=np.float32, steps=None, warmup=2):
    """The reference's kernels (oracle/_ref) on the host cores: bounded sample of the workload.
    steps=None: as many steps as fit in budget_s (the cpu_baseline leg of the GPU arm);
    steps=K: exactly K timed steps after `warmup` untimed ones (the --impl reference arm)."""
    from oracle import ref
    from turbulent_lbm_multigpu_b200.skeleton import compute_parameters
    p = compute_parameters((size,) * 3, (0.1,) * 3, dtype=np.float32)
    fast = ref.available(fast=True)
    kind = "reference"
    if not ref.available(fast=fast):
        raise RuntimeError("oracle/_ref is not built")
    lib = ref.load(fast=fast)
    if threads:
        lib.ref_set_threads(threads)
    cores = lib.ref_get_max_threads()
    s = ref.RefSolver((size,) * 3, [1] * 6, p.inv_tau, p.gravitation, p.u_lid, variant=ref.NOSHM, fast=fast)
    rect = (size - 2, 1, size - 2)
    s.setFlags(np.full(rect[0] * rect[2], 4, np.int32), (1, size - 2, 1), rect)
    for _ in range(max(2, warmup)):                 # warm-up (whole beta/alpha cycles first)
        s.simulationStep()
    if steps is None:
        t0 = time.perf_counter()
        s.simulationStep(); s.simulationStep()
        per = (time.perf_counter() - t0) / 2
        steps = max(2, int(budget_s / max(per, 1e-6)) // 2 * 2)
        steps = min(steps, 200)
    t0 = time.perf_counter()
    for _ in range(steps):
        s.simulationStep()
    dt = time.perf_counter() - t0
    mlups = size ** 3 * steps / dt / 1e6
    sample = ("%d steps of the %d^3 fp32 cavity, reference kernels lbm_alpha.cl/lbm_beta.cl (BGK -- the "
              "reference has no Smagorinsky; beta via its own USE_SHARED_MEMORY 0 path) compiled as C++ "
              "(%s), OpenMP over work-items" % (steps, size, "-O3 -march=x86-64-v3" if fast else "-O2 strict"))
    return dict(value=mlups, unit="MLUPS", cores=cores, kind=kind, sample=sample), dt / steps * 1e3, steps


def reference_sample_size(steps, warmup, threads, limit_s=150.0):
    """Edge of the cubic cavity one reference-arm step runs on: the 256^3 workload itself when
    `warmup + steps` of it fit in limit_s on these host cores, else the largest of 192/128/96/64
    that does (MLUPS is a rate, the sample only bounds the wall clock).  Calibrated with two
    64^3 steps: cost per step scales with the cell count."""
    from oracle import ref
    from turbulent_lbm_multigpu_b200.skeleton import compute_parameters
    fast = ref.available(fast=True)
    lib = ref.load(fast=fast)
    if threads:
        lib.ref_set_threads(threads)
    p = compute_parameters((64,) * 3, (0.1,) * 3, dtype=np.float32)
    s = ref.RefSolver((64,) * 3, [1] * 6, p.inv_tau, p.gravitation, p.u_lid, variant=ref.NOSHM, fast=fast)
    s.simulationStep(); s.simulationStep()
    t0 = time.perf_counter()
    for _ in range(4):
        s.simulationStep()
    per_cell = (time.perf_counter() - t0) / 4 / 64 ** 3 * 1.5        # in-cache calibration: be pessimistic
    for size in (256, 192, 128, 96, 64):
        if per_cell * size ** 3 * (steps + max(2, warmup)) <= limit_s:
            return size
    return 64


RESULT_OUT = sys.stdout     # main() swaps in a private copy of the original stdout


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # torchrun exports OMP_NUM_THREADS=1: the reference arm uses every host core it may run on
    threads = len(os.sched_getaffinity(0))
    K, W = max(1, args.steps), max(0, args.warmup)
    size = reference_sample_size(K, W, threads)
    size = min(size, args.size)
    cb, ms, steps = cpu_reference(size, threads=threads, steps=K, warmup=W)
    line = {
        "impl": "reference", "metric": "MLUPS", "value": cb["value"], "unit": "MLUPS", "n_gpus": args.gpus,
        "steps": steps, "warmup": max(2, W), "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=RESULT_OUT, flush=True)
    return 0


def workload_config(args, n):
    cfg = scenario(n, args.size, args.scaling, args.decomp, args.cs, getattr(args, "shape", None))
    sub = [d // k for d, k in zip(cfg.domain_size, cfg.subdomain_num)]
    elem = 4 if args.dtype == "f32" else 8
    ws = (19 * elem + 4) * sub[0] * sub[1] * sub[2] / 1e9
    tag = {(256, "weak", "f32"): " (BASELINE configs[1])", (512, "strong", "f32"): " (BASELINE configs[2])",
           (512, "weak", "f32"): " (BASELINE configs[3])", (384, "strong", "f64"): " (BASELINE configs[4])"}
    shape = getattr(args, "shape", None)
    return {"workload": "lid-driven cavity %s %s D3Q19 %s %s%s" % (
                ("%dx%dx%d" % tuple(shape)) if shape else ("%d^3" % args.size),
                "per GPU" if args.scaling == "weak" else "global", "fp32" if args.dtype == "f32" else "fp64",
                ("Smagorinsky C_s=%g" % args.cs) if args.cs else "BGK",
                " (reference benchmark.py recipe)" if shape else tag.get((args.size, args.scaling, args.dtype), "")),
            "global_domain": list(cfg.domain_size), "subdomain_num": list(cfg.subdomain_num),
            "subdomain_size": sub,
            "parallelism": ("%s domain decomposition, 1 process per GPU" % args.decomp) if n > 1 else "single GPU",
            "l2": "working set %.2f GB per GPU >> 126 MB L2 (no flush needed)" % ws}


# ====================================================================== parity leg
def verify_parity(args, rank, world, local, dist, backend, axis_order, np_dtype):
    """A small run of the SAME decomposition, transport, phase order, precision, Smagorinsky constant and
    kernel instantiations as the timed workload (sub-domains of 40 x 24 x 16 cells: rows no divisor of 128, so
    the production kernels run), 12 steps, every population and flag of every rank compared bit for bit
    with the CPU oracle's decomposed run (oracle/multi.py; the oracle is only the checker here).  The verdict
    travels in the JSON line, so a multi-GPU scaling job carries its own parity evidence."""
    import torch
    from turbulent_lbm_multigpu_b200.configuration import CConfiguration
    from turbulent_lbm_multigpu_b200.controller import CManager
    from turbulent_lbm_multigpu_b200.domain import CDomain
    steps, sub = 12, (40, 24, 16)
    nums = decomposition(world, args.decomp)
    D = tuple(s * k for s, k in zip(sub, nums))
    L = (0.1, 0.1, 0.1)
    cfg = CConfiguration()
    cfg.domain_size, cfg.subdomain_num, cfg.smagorinsky_constant = D, nums, args.cs
    cs_, ms_ = torch.cuda.Stream(), torch.cuda.Stream(priority=-1)
    mgr = CManager(CDomain(-1, D, (0, 0, 0), L), nums, backend=backend, device=local, config=cfg,
                   sync_mode=args.sync if world > 1 else "host", dtype=np_dtype, axis_order=axis_order,
                   store_velocity=False, store_density=False,
                   compute_stream=cs_.cuda_stream, comm_stream=ms_.cuda_stream)
    mgr.initSimulation(rank)
    ctrl = mgr.getController()
    for _ in range(steps):
        ctrl.computeNextStep()
    s = ctrl.getSolver()
    s.wait()
    mine = (s.storeDensityDistribution(), s.storeFlags(), s.config())
    s.close()
    gathered = [mine]
    if dist is not None:
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
    if rank != 0:
        return None
    out = {"ok": False, "steps": steps, "subdomain_size": list(sub), "subdomain_num": list(nums), "ranks": world,
           "compared": "19 populations + flags of every cell (ghost layers included) of every rank, bit-exact, "
                       "against the CPU oracle's decomposed run (oracle/multi.py)"}
    try:
        from oracle import multi as omulti
        make, _ = omulti.make_oracle_factory(D, nums, L, dtype=np_dtype, variant=0, smagorinsky_cs=args.cs)
        md = omulti.MultiDomain(D, nums, make, slots="reference" if (world > 1 and args.sync == "host") else "minimal",
                                axis_order=(2, 1, 0) if axis_order == "zyx" else (0, 1, 2))
        md.run(steps)
        bad = []
        for r in range(world):
            dd, fl, kc = gathered[r]
            o = md.ranks[r]["solver"]
            if not (np.array_equal(dd.view(np.uint8), o.dd.view(np.uint8)) and np.array_equal(fl, o.flags)):
                bad.append(r)
            if kc["wg_quirk"] != 0:
                bad.append(("quirk", r))
        out["ok"] = not bad
        if bad:
            out["differing_ranks"] = bad
    except Exception as e:         # the checker being absent is reported, not hidden
        out["ok"] = None
        out["error"] = "oracle unavailable: %s" % e
    return out


# ====================================================================== GPU arm
def run_gpu(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    n = args.gpus
    if world != n:
        if world == 1 and n > 1:
            raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (n, n))
        n = world
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from turbulent_lbm_multigpu_b200.comm_backends import TorchDistributedBackend
    from turbulent_lbm_multigpu_b200.controller import CManager
    from turbulent_lbm_multigpu_b200.domain import CDomain

    np_dtype = np.float32 if args.dtype == "f32" else np.float64
    elem = np.dtype(np_dtype).itemsize
    bytes_per_lup = 2 * 19 * elem + 4           # 156 B fp32, 308 B fp64
    if args.store:
        bytes_per_lup += 4 * elem               # STORE_VELOCITY + STORE_DENSITY: 3 + 1 more values written per cell
    cfg = scenario(n, args.size, args.scaling, args.decomp, args.cs, getattr(args, "shape", None))
    domain = CDomain(-1, cfg.domain_size, (0, 0, 0), cfg.domain_length)
    backend = TorchDistributedBackend() if world > 1 else None
    # the library launches on torch-owned streams so that NCCL (torch.distributed) and a
    # CUDA-graph capture of the two-step cycle see the same stream order
    compute_stream, comm_stream = torch.cuda.Stream(), torch.cuda.Stream(priority=-1)
    axis_order = args.axis_order
    if axis_order == "auto":
        axis_order = "zyx" if cfg.subdomain_num[0] > 1 else "xyz"
    mgr = CManager(domain, cfg.subdomain_num, backend=backend, device=local, config=cfg,
                   sync_mode=args.sync if world > 1 else "host", dtype=np_dtype, axis_order=axis_order,
                   store_velocity=args.store, store_density=args.store,
                   compute_stream=compute_stream.cuda_stream, comm_stream=comm_stream.cuda_stream)
    mgr.initSimulation(rank)
    ctrl = mgr.getController()
    s = ctrl.getSolver()
    cells_global = int(np.prod(cfg.domain_size))
    sub = mgr.getSubdomainSize()
    cells_local = int(np.prod(sub))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if dist is None:
            return float(v)
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    parity = None
    if not args.no_verify:
        parity = verify_parity(args, rank, world, local, dist, backend, axis_order, np_dtype)
        barrier()

    W, K = max(3, args.warmup), args.steps
    W += W & 1                              # whole beta/alpha cycles
    K += K & 1
    if args.min_seconds > 0:                # sustained run: enough steps for the requested wall time
        for _ in range(4):
            ctrl.computeNextStep()
        barrier()
        s.timerStart()
        for _ in range(10):
            ctrl.computeNextStep()
        est = max_over_ranks(s.timerStop()) / 10.0
        K = max(K, int(args.min_seconds * 1e3 / max(est, 1e-3)) + 1)
        K += K & 1
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                     # sampled from the warm-up on: the timed region is short
    for _ in range(W):
        ctrl.computeNextStep()
    barrier()
    # launch-bound inner loop: capture the beta+alpha cycle (kernels, stream edges, NCCL
    # send/recv) into one CUDA graph and replay it
    graph, per_cycle = None, 0
    if args.graph:
        try:
            l0 = s.launchCount()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=compute_stream, capture_error_mode="thread_local"):
                ctrl.computeNextStep()
                ctrl.computeNextStep()
            per_cycle = s.launchCount() - l0
            with torch.cuda.stream(compute_stream):
                graph.replay()              # warm the instantiated graph
            barrier()
        except Exception as e:              # keep the eager path usable
            if rank == 0:
                print("cuda graph capture unavailable (%s): eager launches" % e, file=sys.stderr)
            graph = None
            barrier()
    if rank == 0:
        sampler.mark()
    launches0 = s.launchCount()
    s.timerStart()
    if graph is not None:
        with torch.cuda.stream(compute_stream):
            for _ in range(K // 2):
                graph.replay()
    else:
        for _ in range(K):
            ctrl.computeNextStep()
    ms = s.timerStop()                     # CUDA events on the launching (compute) stream
    barrier()
    launches = (K // 2) * per_cycle if graph is not None else s.launchCount() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms = max_over_ranks(ms)
    value = cells_global * K / (ms * 1e-3) / 1e6

    # ---- halo exposed vs hidden (N > 1): the same steps without any exchange
    halo = None
    if world > 1:
        for _ in range(4):
            s.simulationStep()
        barrier()
        s.timerStart()
        for _ in range(K):
            s.simulationStep()
        ms_nc = max_over_ranks(s.timerStop())
        barrier()
        timeline = None
        if args.sync == "p2p":
            # device timestamps of single overlapped steps: fork -> shell done / interior done /
            # exchange done / join (mean over 10 beta and 10 alpha steps, max over ranks)
            acc = np.zeros((2, 4))
            for i in range(20):
                acc[i & 1] += np.array(s.commStepTimed())
            acc /= 10.0
            timeline = {}
            for k, name in enumerate(("beta_step", "alpha_step")):
                v = [max_over_ranks(float(x)) for x in acc[k]]
                timeline[name] = {"shell_done": v[0], "interior_done": v[1], "exchange_done": v[2], "join": v[3]}
            barrier()
        face_bytes = 0
        for c in ctrl.getComms():
            face_bytes += 5 * int(np.prod(c.getSendSize())) * elem
        halo = {"ms_per_step_no_exchange": ms_nc / K, "ms_per_step": ms / K,
                "exposed_frac": max(0.0, 1.0 - ms_nc / ms), "hidden_frac": min(1.0, ms_nc / ms),
                "bytes_sent_per_step_per_gpu": face_bytes, "timeline_ms": timeline,
                "note": "exposed = 1 - t_step(step kernels only) / t_step(with halo exchange), max over ranks"}

    # ---- per-kernel timing (alpha / beta alone), single GPU only: explains the roofline
    kern = {}
    if world == 1:
        for name, fn in (("lbm_alpha_kernel", s.simulationStepAlpha), ("lbm_beta_kernel", s.simulationStepBeta)):
            for _ in range(3):
                fn()
            s.wait()
            s.timerStart()
            reps = 20
            for _ in range(reps):
                fn()
            kms = s.timerStop() / reps
            kern[name] = {"ms": kms, "GBs": bytes_per_lup * cells_local / (kms * 1e-3) / 1e9}

    # ---- end to end through the public host API with HOST buffers every step: each rank
    # uploads the flags of its lid plane (y = Sy-2; pinned host memory -> device) before the
    # step and reads the 19 populations of its centre line back after it
    rect = (sub[0] - 2, 1, sub[2] - 2)
    ro = (1, sub[1] - 2, 1)
    lid = torch.empty(rect[0] * rect[2], dtype=torch.int32).pin_memory()
    lid_np = lid.numpy()
    s.storeFlags(lid_np, ro, rect)          # re-uploading the current flags leaves the run unchanged
    probe = torch.empty(19 * sub[1], dtype=torch.float32 if elem == 4 else torch.float64).pin_memory()
    probe_np = probe.numpy()
    po, ps = (sub[0] // 2, 0, sub[2] // 2), (1, sub[1], 1)
    for _ in range(4):
        s.setFlags(lid_np, ro, rect); ctrl.computeNextStep(); s.storeDensityDistribution(probe_np, po, ps)
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        s.setFlags(lid_np, ro, rect)                        # H2D: boundary input of the step
        ctrl.computeNextStep()
        s.storeDensityDistribution(probe_np, po, ps)        # D2H: populations on the centre line
    barrier()
    dt = max_over_ranks(time.perf_counter() - t0)
    e2e = {"value": cells_global * K / dt / 1e6, "unit": "MLUPS",
           "h2d_bytes_per_step": int(lid_np.nbytes) * n, "d2h_bytes_per_step": int(probe_np.nbytes) * n,
           "api": "CLbmSolver.setFlags + CController.computeNextStep + storeDensityDistribution (pinned host buffers)",
           "note": "the step's host input is one boundary plane of flags, its host output one line of populations: the "
                   "copies are small by nature of the workload (the lattice stays resident in HBM); e2e < value is "
                   "the per-step host round trip (two blocking copies), not bandwidth"}

    if rank == 0:
        peak, peak_src = measured_peak()
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f)
            if tj.get("cells_per_launch") == cells_local and tj.get("dtype", "f32") == args.dtype:
                traffic = tj.get("dram_bytes_per_launch")
        except Exception:
            pass
        achieved = bytes_per_lup * cells_local * K / (ms * 1e-3) / 1e9      # per GPU
        line = {
            "metric": "MLUPS", "value": value, "unit": "MLUPS", "n_gpus": n, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic", "config": workload_config(args, n),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bytes_per_lup * cells_local,
                         "kernels": kern,
                         "note": "per GPU; step kernels lbm_alpha_kernel/lbm_beta_kernel alternate; %d B per "
                                 "lattice-site update x cells per launch / CUDA-event time" % bytes_per_lup},
            "halo": halo,
            "parity": parity,
            "sync_mode": args.sync if world > 1 else None, "axis_order": axis_order if world > 1 else None,
            "cuda_graph": graph is not None,
            "kernel_config": dict(s.config(), store_velocity_density=bool(args.store)),
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                cb, _, _ = cpu_reference(min(args.size, 256), budget_s=args.cpu_budget,
                                         threads=len(os.sched_getaffinity(0)))
                line["cpu_baseline"] = cb
            except Exception as e:    # the checker being absent must not hide the GPU result
                line["cpu_baseline"] = {"value": None, "unit": "MLUPS", "cores": os.cpu_count(), "kind": "reference",
                                        "sample": "unavailable: %s" % e}
        print(json.dumps(line), file=RESULT_OUT, flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=SIZE, help="cells per axis (per GPU for weak scaling)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--decomp", default="slab", choices=["slab", "slab-x", "pencil", "block"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--cs", type=float, default=CS, help="Smagorinsky constant (0 = plain BGK, the reference)")
    ap.add_argument("--sync", default="p2p", choices=["host", "device", "overlap", "p2p"],
                    help="halo transport for N > 1 (p2p: one-sided NVLink peer stores; overlap/device: NCCL)")
    ap.add_argument("--cpu-budget", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--axis-order", default="auto", choices=["auto", "xyz", "zyx"],
                    help="phase order of the halo sync (p2p): xyz = the reference's; zyx = x faces exchanged after "
                         "the interior kernel, no x shell; auto = zyx when the decomposition cuts x")
    ap.add_argument("--graph", action="store_true", help="capture the 2-step cycle in a CUDA graph")
    ap.add_argument("--config", default=None, choices=sorted(PRESETS),
                    help="BASELINE.json configs[] presets: 1 = 256^3/GPU weak z-slabs fp32 (the default), 2 = 512^3 strong, "
                         "3-slab / 3-block / 3-pencil = 512^3/GPU weak, 4 = 384^3 fp64 strong, "
                         "recipe-weak / recipe-strong = the reference's benchmark.py shapes (1024 n x 1024 x 32, -X n)")
    ap.add_argument("--min-seconds", type=float, default=0.0,
                    help="raise --steps until the timed region lasts at least this long (sustained-clock record)")
    ap.add_argument("--no-verify", action="store_true",
                    help="skip the parity leg (a small run of the same decomposition / transport / kernels checked "
                         "bit for bit against the CPU oracle; its verdict is the `parity` key of the JSON line)")
    ap.add_argument("--store", action="store_true", help="STORE_VELOCITY/STORE_DENSITY instantiations (visualisation/validate builds)")
    args = ap.parse_args()
    if args.config:     # a preset fills in what the command line left at its default
        for k, v in PRESETS[args.config].items():
            if not hasattr(args, k) or getattr(args, k) == ap.get_default(k):
                setattr(args, k, v)
    # The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner
    # under NCCL_DEBUG=VERSION, OpenMP/torch warnings): keep the real stdout for the result line and
    # point file descriptor 1 at stderr for everything else.
    global RESULT_OUT
    sys.stdout.flush()
    RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())

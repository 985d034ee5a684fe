#!/usr/bin/env python
"""bench.py -- self-play hot path throughput on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2]/[3]): 1024 concurrent self-play games per GPU with
device-resident PUCT trees, options.def.yml settings (2x64 residual tower, 1024-node budget per
move, cpuct 1.5, bootstrap 20 %), random-init weights, synthetic positions (the games themselves).
A "step" is one pass of the hot path over one batch: every tree selects a leaf (terminals
absorbed, moves made at the node budget), leaves are encoded, the tower + heads run, results are
expanded and backed up  ==  1024 NN evaluations per GPU.

Prints ONE JSON line (rank 0).  `value` = whole-job NN evals/s with everything resident in HBM;
`e2e` = the same loop through the reference-shaped host-buffer API (kb_pool_step_hostio:
observations and policy/value rows cross PCIe every step, like kami::NN::infer / MCTS::expand).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

TREES_PER_GPU = 1024
FILTERS, RESIDUALS = 64, 2          # options.def.yml:29,53
SELFPLAY_NODES = 1024               # options.def.yml selfplay_nodes
NODE_CAPACITY = 1 << 19  # 2 x 10 MB per tree: the copying collector runs about once per 40 moves
TOWER64_DRAM_BYTES_PER_LAUNCH = 12852992  # dram__bytes_read.sum + dram__bytes_write.sum of one k_tower64<16> launch, profiles/r02_ncu_summary.txt
# State preparation (untimed).  The step gets slower as the synthetic games leave the opening (more legal moves per
# position, terminal leaves to absorb: 84 us/step after 5 k steps from fresh trees, 118-135 us from 20 k steps on,
# tools/steady_state.py), so the timed region must not start from young games.  The games are aged quickly with a small
# node budget (a move every ~30 steps: each tree plays ~200 plies, i.e. the pool holds games in every phase), then the
# trees are rebuilt under the full budget until every tree has moved a few times at 1024 nodes.
AGE_STEPS, AGE_NODES = 6000, 32
PREROLL_STEPS = 3500
TERMINAL_CAP = 1                    # kb_pool_set_terminal_cap: see include/kami_b200.h
METRIC = "selfplay_nn_evals_per_sec"
UNIT = "evals/s"


def peaks():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), float(d["bf16_tflops"]), float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "measured"
    except Exception:
        return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod

        torch.cuda.set_device(local)
        dist_mod.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
        dist = dist_mod
    return rank, world, local, dist


def barrier(dist, local):
    if dist is not None:
        import torch

        dist.barrier(device_ids=[local])
        torch.cuda.synchronize()


def reduce_max(dist, local, x):
    from kami_b200.parallel import Reducer

    return Reducer(dist, "cuda:%d" % local).max(x)


def reduce_sum(dist, local, x):
    from kami_b200.parallel import Reducer

    return Reducer(dist, "cuda:%d" % local).sum(x)


# ---------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation (oracle/_ref, compiled from
# /root/reference) on the host cores -- selfplay.cpp:113-200 with the reference's objects
# ---------------------------------------------------------------------------------------------
def reference_loop(seconds_budget, threads, ibatch, steps=None, iters_per_step=4):
    """Runs `threads` inference threads x `ibatch` reference MCTS trees sharing one reference NN
    (LibTorch CPU, force_cpu), exactly the structure of Selfplay::inference_main.  Returns
    (evals, moves, seconds, kind)."""
    import harness as H

    if H.ref_core() is None or H.ref_nn_lib() is None:
        return _port_loop(seconds_budget, ibatch, steps, iters_per_step)
    cfg = H.default_cfg(noise_weight=0.05, **H.DEF_YML)
    nn = H.RefNN(FILTERS, RESIDUALS, seed=1, force_cpu=True)
    H.ref_nn_lib().ref_nn_set_threads(max(1, os.cpu_count() or 1))
    counts = [[0, 0] for _ in range(threads)]
    stop = threading.Event()
    limit = None if steps is None else steps * iters_per_step

    def worker(tid):
        trees = [H.RefMcts(cfg) for _ in range(ibatch)]
        batch = np.zeros((ibatch, H.OBSIZE), np.float32)
        it = 0
        while not stop.is_set() and (limit is None or it < limit):
            for i, t in enumerate(trees):
                while True:
                    ok = False
                    while t.n() < SELFPLAY_NODES:
                        ok, obs = t.select()
                        if ok:
                            break
                    if ok:
                        batch[i] = obs
                        break
                    ply = t.env.ply()  # selfplay.cpp:153-156 with options.def.yml's schedule
                    a = t.pick(0.95 ** ply if ply < 20 else 0.5)
                    t.push(a)
                    counts[tid][1] += 1
                    if t.env.terminal()[0]:
                        t.reset()
            pol, val = nn.infer(batch)
            for i, t in enumerate(trees):
                t.expand(pol[i], float(val[i]))
            counts[tid][0] += ibatch
            it += 1

    ths = [threading.Thread(target=worker, args=(i,)) for i in range(threads)]
    t0 = time.time()
    for t in ths:
        t.start()
    if limit is None:
        time.sleep(seconds_budget)
        stop.set()
    for t in ths:
        t.join()
    dt = time.time() - t0
    return sum(c[0] for c in counts), sum(c[1] for c in counts), dt, "reference"


def _port_loop(seconds_budget, ibatch, steps, iters_per_step):
    """Fallback when oracle/_ref is absent: the oracle port (C restatement + numpy network)."""
    import harness as H
    import nn_oracle as NO

    params = NO.init_params(FILTERS, RESIDUALS, seed=1)
    cfg = H.default_cfg(noise_weight=0.0, **H.DEF_YML)
    trees = [H.OracleMcts(cfg) for _ in range(ibatch)]
    evals = moves = it = 0
    limit = None if steps is None else steps * iters_per_step
    t0 = time.time()
    while (limit is None and time.time() - t0 < seconds_budget) or (limit is not None and it < limit):
        batch = np.zeros((ibatch, H.OBSIZE), np.float32)
        for i, t in enumerate(trees):
            while True:
                ok = False
                while t.n() < SELFPLAY_NODES:
                    ok, obs = t.select()
                    if ok:
                        break
                if ok:
                    batch[i] = obs
                    break
                t.push(t.pick(0.0))
                moves += 1
                if t.env.terminal()[0]:
                    t.reset()
        pol, val = NO.infer(params, batch)
        for i, t in enumerate(trees):
            t.expand(pol[i], float(val[i]))
        evals += ibatch
        it += 1
    return evals, moves, time.time() - t0, "port"


_REAL_STDOUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries loaded along the way print on their own (the compiled reference's
    static initialiser writes "Initialized neocortex lookup tables" to std::cout, NCCL and torchrun print banners), so fd 1
    is pointed at stderr for the whole run and the JSON line goes to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(obj):
    sys.stdout.flush()
    f = _REAL_STDOUT or sys.stdout
    f.write(json.dumps(obj) + "\n")
    f.flush()


def run_reference(args, rank):
    if rank != 0:
        return
    threads = 3   # options.def.yml inference_threads
    ibatch = 16   # options.def.yml selfplay_batch
    iters_per_step = 4
    # warmup
    reference_loop(0, threads, ibatch, steps=max(1, args.warmup), iters_per_step=1)
    evals, moves, dt, kind = reference_loop(0, threads, ibatch, steps=args.steps, iters_per_step=iters_per_step)
    v = evals / dt
    cores = os.cpu_count() or 1
    sample = "%d steps x %d iterations of %d threads x %d trees (selfplay.cpp:113-200), LibTorch CPU fp32, %d torch threads" % (
        args.steps, iters_per_step, threads, ibatch, cores)
    out = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / max(1, args.steps) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                         # (a move needs the 1 024-node budget: a bounded sample from fresh trees holds none, so the rate is
                         # derived from the budget instead of printing 0)
                         "positions_per_sec": moves / dt if moves else v / SELFPLAY_NODES,
                         "positions_note": "measured" if moves else "evals/s / node budget (no move inside the bounded sample)"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(out)


def workload_config():
    return {"workload": "1024 concurrent self-play games per GPU, device-resident PUCT trees, options.def.yml "
                        "(2x64 tower, 1024-node budget/move, cpuct 1.5, bootstrap 20%, noise 0.05), random-init weights",
            "trees_per_gpu": TREES_PER_GPU, "filters": FILTERS, "residuals": RESIDUALS, "selfplay_nodes": SELFPLAY_NODES,
            "evals_per_step_per_gpu": TREES_PER_GPU,
            "state": "games aged to steady state before the timed region: %d steps at a %d-node budget, then %d steps at the full budget" % (
                AGE_STEPS, AGE_NODES, PREROLL_STEPS),
            "terminal_cap": "a tree that absorbed %d terminal visits in one step sits the step out (kb_pool_set_terminal_cap); value counts real "
                            "evaluations only" % TERMINAL_CAP,
            "reference_arm": "--impl reference runs the reference's stock CPU shape of this workload: 3 inference threads x 16 "
                             "trees (options.def.yml inference_threads / selfplay_batch) sharing one LibTorch-CPU network, from fresh trees",
            "l2": "inputs larger than L2: live node pools ~0.6 MB x 1024 trees per GPU >> 126 MB, no flush"}


def random_blob(F, R, seed):
    """Random-init weights in kb_net_load_blob order with LibTorch's default distributions (U(+-1/sqrt(fan_in)) for
    conv / linear weights and biases; BatchNorm gamma 1, beta 0, running mean 0, running var 1), numpy only."""
    rng = np.random.RandomState(seed)
    parts = []

    def conv(o, c, k):
        b = 1.0 / np.sqrt(c * k * k)
        parts.append(rng.uniform(-b, b, o * c * k * k))
        parts.append(rng.uniform(-b, b, o))

    def bn(c):
        parts.extend([np.ones(c), np.zeros(c), np.zeros(c), np.ones(c)])

    conv(F, 30, 3)
    bn(F)
    for _ in range(R):
        conv(F, F, 3)
        bn(F)
        conv(F, F, 3)
        bn(F)
    conv(128, F, 1)
    bn(128)
    conv(73, 128, 1)
    conv(1, F, 1)
    bn(1)
    parts.append(rng.uniform(-0.125, 0.125, 256 * 64))
    parts.append(rng.uniform(-0.125, 0.125, 256))
    return np.concatenate(parts).astype(np.float32)


# ---------------------------------------------------------------------------------------------
# training leg (BASELINE config 5)
# ---------------------------------------------------------------------------------------------
def train_leg(api, L, dist, local, world, pool, F=256, R=20, B=1024, steps=5):
    """Synthetic replay batch from the product's own path: the pool's current leaf positions (1024 distinct
    mid-game positions), their input planes (kb_encode_planes) and a random sparse visit distribution over
    their legal actions (kb_legal_actions) -- the shape of ReplayBuffer::select_batch (replaybuffer.h:61-84)."""
    import kami_b200
    from kami_b200.parallel import data_parallel_step

    rank = dist.get_rank() if dist is not None else 0
    pool.select()
    pos = pool.leaf_positions()[:B]
    obs = api.encode_planes(pos)
    acts, cnt = api.legal_actions(pos)
    rng = np.random.RandomState(17 + rank)
    pi = np.zeros((len(pos), 4672), np.float32)
    for i in range(len(pos)):
        w = rng.randint(1, 40, size=int(cnt[i])).astype(np.float32)
        pi[i, acts[i, :cnt[i]]] = w / w.sum()
    z = rng.choice(np.array([-1.0, 0.5, 1.0], np.float32), size=len(pos)).astype(np.float32)
    rep = (B + len(pos) - 1) // len(pos)
    arrs = [np.ascontiguousarray(np.tile(a, (rep,) + (1,) * (a.ndim - 1))[:B], np.float32) for a in (obs, pi, z)]
    tr = kami_b200.Trainer(F, R, B)
    tr.load_blob(random_blob(F, R, seed=1 + 0 * rank))  # every replica starts from the same weights
    devp = []
    for a in arrs:
        p = C.c_void_p()
        api._ck(L.kb_dev_alloc(C.byref(p), a.nbytes))
        api._ck(L.kb_dev_upload(p, a.ctypes.data_as(C.c_void_p), a.nbytes))
        devp.append(p)
    view = None
    for _ in range(3):
        view = data_parallel_step(tr, dist, "cuda:%d" % local, L, devp[0], devp[1], devp[2], B, 0.002, view)
    ms = C.c_float()
    barrier(dist, local)
    L.kb_dev_sync()
    L.kb_timer_start()
    for _ in range(steps):
        view = data_parallel_step(tr, dist, "cuda:%d" % local, L, devp[0], devp[1], devp[2], B, 0.002, view)
    L.kb_timer_stop(C.byref(ms))
    barrier(dist, local)
    t = reduce_max(dist, local, ms.value) / steps
    loss = tr.forward_backward_dev(devp[0], devp[1], devp[2], B)
    for p in devp:
        L.kb_dev_free(p)
    tower_f = 2.0 * 64 * (9 * 30) * F + R * 2.0 * (2.0 * 64 * 9 * F * F)
    heads_f = 2.0 * 64 * F * 128 + 2.0 * 64 * 128 * 73
    hbm, tfp, tfs, kind = peaks()
    tfl = 3.0 * (tower_f + heads_f) * B * world / (t * 1e-3) / 1e12
    return {"batch_per_gpu": B, "n_gpus": world, "ms_per_step": t, "samples_per_sec": B * world / (t * 1e-3),
            "tflops_3x_forward_convention": tfl, "frac_of_bf16_peak_per_gpu": tfl / world / tfp,
            "grad_bucket_mb": tr.n * 4 / 1e6, "collective": "none (1 GPU)" if world == 1 else "NCCL all-reduce(sum), one fp32 bucket per step",
            "loss_after": loss, "peak_kind": kind}


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def bind_to_gpu_numa(local):
    """Best effort: run this rank (and first-touch its pinned host buffers) on the NUMA node the GPU hangs off, so the
    host-buffer path does not cross sockets.  Returns a short description for the JSON line."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bdf = bus.lower()[-12:]  # 0000:xx:yy.z
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return "numa node unknown"
        cpus = []
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.extend(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return "numa node %d has no allowed cpu" % node
        os.sched_setaffinity(0, cpus)
        return "node %d (%d cpus)" % (node, len(cpus))
    except Exception as e:  # never lose the bench to this
        return "not bound (%s)" % type(e).__name__


def encoder_leg(api, L, pool, hbm_peak, peak_kind, n=131072, reps=10):
    """north_star kernel 1 on its own (encoder roofline): n mid-game positions (the pool's current leaves,
    tiled) -> input planes, positions and planes resident in HBM.  Algorithmic bytes per position: 64 B of position
    read + the planes written -- 7 680 B as fp32 [64][30] (Env::observe's layout), 4 096 B as the tower's bf16 input
    (30 channels padded to 32).  The 1 GB / 1.5 GB outputs are larger than L2, no flush needed."""
    leaves = pool.leaf_positions()
    pos = np.tile(leaves, (n + len(leaves) - 1) // len(leaves))[:n].copy()
    dpos, dobs, dtall = C.c_void_p(), C.c_void_p(), C.c_void_p()
    api._ck(L.kb_dev_alloc(C.byref(dpos), pos.nbytes))
    api._ck(L.kb_dev_alloc(C.byref(dobs), n * 1920 * 4))
    api._ck(L.kb_dev_alloc(C.byref(dtall), L.kb_net_planes_bytes(n)))
    api._ck(L.kb_dev_upload(dpos, pos.ctypes.data_as(C.c_void_p), pos.nbytes))
    out = {"positions": n, "bound": "hbm", "peak": hbm_peak, "unit": "GB/s", "peak_kind": peak_kind}
    ms = C.c_float()
    # traffic: dram__bytes_read.sum + dram__bytes_write.sum of one launch at n = 131072 (ncu --set full, profiles/r01_ncu_summary_v3.txt)
    ncu_traffic = {"k_encode_f32": 10495232 + 951133184, "k_encode_tall": 10511872 + 510364672}
    for name, fn, nbytes in (("k_encode_f32", lambda: L.kb_encode_planes_dev(dpos, n, dobs), 64 + 7680),
                             ("k_encode_tall", lambda: L.kb_encode_planes_bf16_dev(dpos, n, dtall), 64 + 4096)):
        for _ in range(3):
            api._ck(fn())
        L.kb_dev_sync()
        L.kb_timer_start()
        for _ in range(reps):
            api._ck(fn())
        L.kb_timer_stop(C.byref(ms))
        t = ms.value * 1e-3 / reps
        out[name] = {"us_per_launch": t * 1e6, "positions_per_sec": n / t, "algorithmic_bytes_per_position": nbytes,
                     "achieved": n * nbytes / t / 1e9, "frac": n * nbytes / t / 1e9 / hbm_peak,
                     "traffic": ncu_traffic[name] if n == 131072 else None}
    for p in (dpos, dobs, dtall):
        L.kb_dev_free(p)
    return out


def infer256_leg(api, L, pool, net, reps=20):
    """BASELINE config 2 (the reference's test/nncuda.cpp / test/nn.cpp loop at batch 256): 256 synthetic positions ->
    planes -> NN::infer through the host-buffer call (H2D observations, D2H policy [256][4672] + value inside the timed
    region).  Ours with pageable buffers (what an unmodified caller passes) and with pinned ones (kb_host_register),
    next to the UNMODIFIED reference NN::infer on the same input: its own CUDA path (LibTorch CUDA on this GPU,
    test/nncuda.cpp's configuration: fp32 weights, cuDNN with LibTorch's default TF32 convolutions) and its CPU path."""
    pos = pool.leaf_positions()[:256]
    obs = api.encode_planes(pos)
    out = {"batch": 256, "filters": FILTERS, "residuals": RESIDUALS}
    net.infer(obs)
    t0 = time.time()
    for _ in range(reps):
        pol, val = net.infer(obs)
    dt = (time.time() - t0) / reps
    out["ours"] = {"ms_per_call": dt * 1e3, "pred_per_sec": 256 / dt, "api": "kb_net_infer, pageable host buffers"}
    pobs, ppol, pval = obs.copy(), np.zeros((256, 4672), np.float32), np.zeros(256, np.float32)
    for a in (pobs, ppol, pval):
        api._ck(L.kb_host_register(a.ctypes.data_as(C.c_void_p), a.nbytes))
    try:
        f = lambda: api._ck(L.kb_net_infer(net.h, api._fp(pobs), 256, api._fp(ppol), api._fp(pval)))
        f()
        t0 = time.time()
        for _ in range(reps):
            f()
        dp = (time.time() - t0) / reps
        out["ours_pinned"] = {"ms_per_call": dp * 1e3, "pred_per_sec": 256 / dp, "api": "kb_net_infer, caller buffers pinned with kb_host_register",
                              "same_bits_as_pageable": bool(np.array_equal(ppol, pol) and np.array_equal(pval, val))}
    finally:
        for a in (pobs, ppol, pval):
            L.kb_host_unregister(a.ctypes.data_as(C.c_void_p))
    import harness as H  # baselines only: the compiled reference under oracle/_ref

    try:  # in a process of its own: LibTorch's CUDA backend stays out of this one
        import tempfile

        with tempfile.NamedTemporaryFile(suffix=".npy", delete=False) as f:
            np.save(f, obs)
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ref_cuda_infer.py"), f.name, str(reps)], capture_output=True, timeout=180)
        os.unlink(f.name)
        lines = [ln for ln in r.stdout.decode().splitlines() if ln.startswith("{")]
        out["reference_cuda"] = json.loads(lines[-1]) if lines else {"error": "exit code %d: %s" % (r.returncode, r.stderr.decode()[-200:])}
    except Exception as e:
        out["reference_cuda"] = {"error": str(e)[:200]}
    if H.ref_nn_lib() is not None:
        nn = H.RefNN(FILTERS, RESIDUALS, seed=1, force_cpu=True)
        H.ref_nn_lib().ref_nn_set_threads(max(1, os.cpu_count() or 1))
        nn.infer(obs)
        t0 = time.time()
        for _ in range(3):
            nn.infer(obs)
        dr = (time.time() - t0) / 3
        out["reference_cpu"] = {"ms_per_call": dr * 1e3, "pred_per_sec": 256 / dr, "cores": os.cpu_count() or 1}
    return out


def selfplay_20x256_leg(api, L, dist, local, rank, kw, per_rank_seed, steps=24):
    """SURVEY 8(d) configs 3-4 "F=64,R=2 AND F=256,R=20": the same 1024-tree self-play step with BASELINE config 5's
    network (per-layer tcgen05 convs, legal-move softmax), resident."""
    import kami_b200

    big = kami_b200.NN(256, 20)
    big.load_blob(random_blob(256, 20, seed=1))
    pool = kami_b200.TreePool(TREES_PER_GPU, 1 << 17, api.tree_cfg(seed=per_rank_seed(7000, rank), **kw))
    pool.step(big, 6)
    pool.reset_stats()
    ms = C.c_float()
    barrier(dist, local)
    L.kb_dev_sync()
    L.kb_timer_start()
    pool.step(big, steps)
    L.kb_timer_stop(C.byref(ms))
    barrier(dist, local)
    t = reduce_max(dist, local, ms.value)
    ev = reduce_sum(dist, local, float(pool.stats()["evals"]))
    t_f, h_f = big.flops()
    hbm, tfp, tfs, kind = peaks()
    world = dist.get_world_size() if dist is not None else 1
    v = ev / (t * 1e-3)
    return {"filters": 256, "residuals": 20, "trees_per_gpu": TREES_PER_GPU, "steps": steps, "ms_per_step": t / steps, "evals_per_sec": v,
            "n_gpus": world, "tflops_per_gpu": v / world * (t_f + h_f) / 1e12, "frac_of_bf16_peak_per_gpu": v / world * (t_f + h_f) / 1e12 / tfp,
            "peak_kind": kind, "note": "trees a few moves old (6 warm-up steps): the step is the network"}


def selfplay_cpp_leg(devices, seconds=5.0):
    """The e2e the C++ product serves: kami::Selfplay (kami/selfplay.h) start()/stop() from kami/tests/selfplay_e2e.cpp --
    inference threads on device pools, finished games drained into the host ReplayBuffer; one host process, one
    inference thread per GPU."""
    exe = os.path.join(ROOT, "kami", "_dropin", "selfplay_e2e")
    if not os.path.exists(exe):
        return {"unavailable": "kami/_dropin/selfplay_e2e not built"}
    out = subprocess.run([exe, str(seconds), str(devices)], capture_output=True, timeout=120)
    for line in out.stdout.decode().splitlines():
        if line.startswith("{"):
            return json.loads(line)
    return {"error": (out.stderr.decode() or out.stdout.decode())[-300:]}


def run_ours(args, rank, world, local, dist):
    numa = bind_to_gpu_numa(local)
    import kami_b200
    from kami_b200 import api
    from kami_b200.parallel import per_rank_seed

    api.init(local)
    L = kami_b200.lib()
    hbm_peak, tf_peak, tf_sustained, peak_kind = peaks()

    net = kami_b200.NN(FILTERS, RESIDUALS)
    net.load_blob(random_blob(FILTERS, RESIDUALS, seed=1))
    kw = dict(noise_weight=0.05, selfplay_nodes=SELFPLAY_NODES, alpha_initial=1.0, alpha_decay=0.95, alpha_final=0.5,
              alpha_cutoff=20, draw_value_pct=50, **kami_b200.DEF_YML)
    pool = kami_b200.TreePool(TREES_PER_GPU, NODE_CAPACITY, api.tree_cfg(seed=per_rank_seed(1000, rank), **kw))

    def timed_steps(fn, k):
        ms = C.c_float()
        barrier(dist, local)
        L.kb_dev_sync()
        L.kb_timer_start()
        fn(k)
        L.kb_timer_stop(C.byref(ms))
        barrier(dist, local)
        return reduce_max(dist, local, ms.value)

    # state preparation (untimed): age the synthetic games to steady state, then W warm-up steps
    pool.set_terminal_cap(TERMINAL_CAP)
    if args.preroll > 0:
        pool.set_selfplay_nodes(AGE_NODES)
        pool.step(net, AGE_STEPS)
        pool.set_selfplay_nodes(SELFPLAY_NODES)
        pool.step(net, args.preroll)
    pool.reset_stats()
    # nvidia-smi needs a few hundred ms before its first line, a 20-step timed region is 1.4 ms: the sampler starts with the
    # warm-up (the same step, ~150 ms of it) and, if the region was too short for three samples, keeps watching an untimed
    # continuation of the same step -- every sample is taken under this workload, and the window is named in the JSON
    sampler = ClockSampler(local)
    sampler.start()
    pool.step(net, max(3, args.warmup) + 2048)  # (+ 2048 full-budget steps whose evals / moves ratio is kept for positions/s)
    st_pre = pool.stats()
    pool.reset_stats()
    ms = timed_steps(lambda k: pool.step(net, k), args.steps)
    st = pool.stats()
    for _ in range(12):
        if len(sampler.rows) >= 3:
            break
        pool.step(net, 1024)
        L.kb_dev_sync()
    clocks = sampler.stop()
    clocks["window"] = "warm-up + timed region (+ untimed continuation of the same step until 3 samples)"
    evals_local = float(st["evals"])
    evals_per_move = (st_pre["evals"] + st["evals"]) / max(1.0, float(st_pre["moves"] + st["moves"]))
    # phase times (and the roofline's kernel duration): a separate short PROFILED call -- the timed run above records
    # no events and fuses expand + select everywhere (kb_pool_set_profiling is opt-in)
    pool.set_profiling(True)
    pool.step(net, 64)
    ph = pool.phase_ms()
    pool.set_profiling(False)
    evals = reduce_sum(dist, local, evals_local)
    moves = reduce_sum(dist, local, float(st["moves"]))
    value = evals / (ms * 1e-3)

    # roofline of the dominant kernel of the step (timed live with CUDA events inside kb_pool_step)
    t_f, h_f = net.flops()
    phases = {"select+encode": ph["select"], "tower+heads": ph["tower"], "expand+backup": ph["expand"]}
    # dominant kernel of the step: k_tower64<16> (60 % of the step in the ncu launch list, profiles/r02_ncu_summary.txt)
    achieved = (t_f + h_f) * TREES_PER_GPU / (ph["tower"] * 1e-3) / 1e12
    roof = {"kernel": "k_tower64 (one fused tcgen05 launch: %dx%d tower + policy/value heads + legal-move softmax)" % (RESIDUALS, FILTERS),
            "bound": "tensor", "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s", "frac": achieved / tf_peak,
            # dram__bytes_read.sum + dram__bytes_write.sum of one k_tower64<16> launch at 1024 boards, from the
            # ncu --set full capture summarised in profiles/r02_ncu_summary.txt (12.85 MB read: 147 input
            # slabs of 80 KB + weights)
            "traffic": TOWER64_DRAM_BYTES_PER_LAUNCH if TREES_PER_GPU == 1024 else None,
            "traffic_unit": "bytes/launch", "peak_kind": peak_kind + " burst",
            "algorithmic": "%.3f MFLOP/position x %d positions per launch" % ((t_f + h_f) / 1e6, TREES_PER_GPU)}
    # the tree kernels (latency-bound, one warp per tree): algorithmic bytes counted by the kernels themselves
    per_step = (12.0 * st["children_scanned"] + 16.0 * st["path_nodes"] + 16.0 * st["children_created"]) / max(1, args.steps)
    per_step += 3904.0 * TREES_PER_GPU
    dur = (ph["select"] + ph["expand"]) * 1e-3
    roof_tree = {"kernel": "k_pool_select + k_pool_expand", "bound": "hbm", "achieved": per_step / dur / 1e9, "peak": hbm_peak,
                 "unit": "GB/s", "frac": per_step / dur / 1e9 / hbm_peak, "traffic": 3878656,
                 "traffic_unit": "bytes/step (ncu, k_pool_expand_select, profiles/r02_ncu_summary.txt)", "peak_kind": peak_kind,
                 "algorithmic": "12 B x children scanned + 16 B x path nodes + 16 B x children created + 3904 B planes per leaf"}

    # end to end through the reference-shaped host-buffer API (pinned host memory)
    n = TREES_PER_GPU
    bufs = []

    def pinned(shape):
        nbytes = int(np.prod(shape)) * 4
        p = C.c_void_p()
        api._ck(L.kb_host_alloc_pinned(C.byref(p), nbytes))
        bufs.append(p)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(int(np.prod(shape)),)).reshape(shape)

    obs_h, pol_h, val_h = pinned((n, 1920)), pinned((n, 4672)), pinned((n,))
    e2e_steps = max(3, min(args.steps, 400))
    pool.step_hostio(net, 3, obs_h, pol_h, val_h)
    pool.reset_stats()
    ms_e2e = timed_steps(lambda k: pool.step_hostio(net, k, obs_h, pol_h, val_h), e2e_steps)
    evals_e2e = reduce_sum(dist, local, float(pool.stats()["evals"]))
    e2e_value = evals_e2e / (ms_e2e * 1e-3)
    # host -> device: the observations into NN::infer and the values into MCTS::expand are copied; the dense policy array
    # (pinned) is read in place by the expand kernels -- only the legal moves' entries cross the link, as 32-byte sectors
    # (KB_HOSTIO_ZEROCOPY=0 copies all of it back: + n * 4672 * 4 bytes, 0.77 instead of 0.60 ms per step)
    st_e2e = pool.stats()
    legal_per_eval = float(st_e2e["children_created"]) / max(1.0, float(st_e2e["evals"]))
    h2d_copied = n * 1920 * 4 + n * 4
    h2d_in_place = int(n * legal_per_eval * 32)
    h2d = h2d_copied + h2d_in_place
    d2h = n * 1920 * 4 + n * 4672 * 4 + n * 4   # leaf observations out, policy + value out (all copied)

    extras = {}
    cpu = None
    # the same host-buffer loop with the compact forms on the wire (kb_pool_step_hostio_compact): 80-byte leaf positions
    # instead of [1920] fp32 rows, [128] legal-move priors instead of [4672] policy rows
    try:
        from kami_b200 import api as _api

        leaves_h = np.zeros(n, _api.POSITION_DTYPE)
        prior_h = pinned((n, 128))
        api._ck(L.kb_host_register(leaves_h.ctypes.data_as(C.c_void_p), leaves_h.nbytes))
        try:
            pool.step_hostio_compact(net, 3, leaves_h, prior_h, val_h)
            pool.reset_stats()
            ck = max(3, min(args.steps, 1000))
            ms_c = timed_steps(lambda k: pool.step_hostio_compact(net, k, leaves_h, prior_h, val_h), ck)
            ev_c = reduce_sum(dist, local, float(pool.stats()["evals"]))
        finally:
            L.kb_host_unregister(leaves_h.ctypes.data_as(C.c_void_p))
        if rank == 0:
            extras["e2e_compact"] = {"value": ev_c / (ms_c * 1e-3), "unit": UNIT, "steps": ck, "ms_per_step": ms_c / ck,
                                     "h2d_bytes_per_step": n * (80 + 512 + 4), "d2h_bytes_per_step": n * (80 + 512 + 4),
                                     "api": "kb_pool_step_hostio_compact: leaf positions D2H -> H2D -> encode -> tower -> legal priors + value D2H -> H2D -> expand"}
    except Exception as e:
        if rank == 0:
            extras["e2e_compact"] = {"error": str(e)[:200]}
    if rank == 0 and not args.no_extras:
        # 20x256 tower (BASELINE config 5's network) forward only: % of dense BF16 peak at batch 1024
        try:
            big = kami_b200.NN(256, 20)
            big.load_blob(random_blob(256, 20, seed=1))
            B = 1024
            planes, pol, val = C.c_void_p(), C.c_void_p(), C.c_void_p()
            api._ck(L.kb_dev_alloc(C.byref(planes), L.kb_net_planes_bytes(B)))
            api._ck(L.kb_dev_alloc(C.byref(pol), B * 4672 * 4))
            api._ck(L.kb_dev_alloc(C.byref(val), B * 256 * 4))
            for _ in range(3):
                api._ck(L.kb_net_forward_dev(big.h, planes, B, pol, val))
            msb = C.c_float()
            L.kb_dev_sync()
            L.kb_timer_start()
            reps = 5
            for _ in range(reps):
                api._ck(L.kb_net_forward_dev(big.h, planes, B, pol, val))
            L.kb_timer_stop(C.byref(msb))
            tb, hb = big.flops()
            tfl = tb * B * reps / (msb.value * 1e-3) / 1e12
            extras["tower_20x256"] = {"batch": B, "ms_per_forward": msb.value / reps, "tower_tflops": tfl,
                                      "frac_of_bf16_peak": tfl / tf_peak, "frac_of_sustained": tfl / tf_sustained,
                                      "evals_per_sec": B * reps / (msb.value * 1e-3), "peak_kind": peak_kind}
            for p in (planes, pol, val):
                L.kb_dev_free(p)
            del big
        except Exception as e:  # never lose the bench line to the extra measurement
            extras["tower_20x256"] = {"error": str(e)}
        try:
            extras["roofline_encoder"] = encoder_leg(api, L, pool, hbm_peak, peak_kind)
        except Exception as e:
            extras["roofline_encoder"] = {"error": str(e)}
        if world == 1:
            try:
                extras["config2_infer_256"] = infer256_leg(api, L, pool, net)
            except Exception as e:
                extras["config2_infer_256"] = {"error": str(e)}
            ev, mv, dt, kind = reference_loop(12.0, 3, 16)
            cores = os.cpu_count() or 1
            cpu = {"value": ev / dt, "unit": UNIT, "cores": cores, "kind": kind,
                   "positions_per_sec": mv / dt if mv else ev / dt / SELFPLAY_NODES,
                   "sample": "%.0f s of 3 inference threads x 16 trees (options.def.yml) on %d host cores, %s" % (
                       dt, cores, "unmodified reference Env/MCTS + LibTorch CPU fp32 NN" if kind == "reference" else "oracle port")}

    if not args.no_extras:
        # BASELINE config 5: data-parallel NN::train mini-batches of the 20x256 tower, bf16 tcgen05 forward /
        # dgrad / wgrad, ONE NCCL all-reduce of the flat fp32 gradient bucket (95 MB) per step over NVLink
        try:
            tt = train_leg(api, L, dist, local, world, pool)
            if rank == 0:
                extras["train_20x256"] = tt
        except Exception as e:
            if rank == 0:
                extras["train_20x256"] = {"error": str(e)}
    if not args.no_extras:
        try:
            sp = selfplay_20x256_leg(api, L, dist, local, rank, kw, per_rank_seed)
            if rank == 0:
                extras["selfplay_20x256"] = sp
        except Exception as e:
            if rank == 0:
                extras["selfplay_20x256"] = {"error": str(e)[:200]}
    if not args.no_extras and world > 1 and 8192 % world == 0:
        # BASELINE config 4 exactly: 8192 concurrent games in total, sharded by game over the N GPUs (8192 / N trees per
        # GPU, no collective), same network and budget.  The headline line above keeps 1024 trees per GPU (weak scaling).
        try:
            per = 8192 // world
            pool4 = kami_b200.TreePool(per, NODE_CAPACITY, api.tree_cfg(seed=per_rank_seed(5000, rank), **kw))
            pool4.step(net, args.preroll if args.preroll > 0 else 64)
            pool4.reset_stats()
            k4 = max(3, min(args.steps, 400))
            ms4 = timed_steps(lambda k: pool4.step(net, k), k4)
            ev4 = reduce_sum(dist, local, float(pool4.stats()["evals"]))
            if rank == 0:
                extras["config4_8192_games"] = {"trees_per_gpu": per, "n_gpus": world, "steps": k4, "ms_per_step": ms4 / k4,
                                                "evals_per_sec": ev4 / (ms4 * 1e-3), "scaling": "strong (8192 games in total)"}
            del pool4
        except Exception as e:
            if rank == 0:
                extras["config4_8192_games"] = {"error": str(e)}
    for p in bufs:
        L.kb_host_free_pinned(p)
    del pool
    if not args.no_extras:
        # one C++ host process driving all N GPUs through kami::Selfplay.  The other ranks wait on a HOST-side (gloo)
        # barrier: an NCCL barrier would spin a kernel on their GPUs, which the C++ process is using
        host_group = dist.new_group(backend="gloo") if dist is not None else None
        barrier(dist, local)
        if rank == 0:
            try:
                extras["e2e_selfplay_cpp"] = selfplay_cpp_leg(world)
            except Exception as e:
                extras["e2e_selfplay_cpp"] = {"error": str(e)[:200]}
        if dist is not None:
            dist.barrier(group=host_group)
    if rank != 0:
        return
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic", "config": workload_config(),
        # moves come in bursts (a tree moves every ~870 steps), so a short timed region can hold none: positions/s is
        # value / (evals per move over the full-budget part of the state preparation + the timed region)
        "positions_per_sec": value / max(1.0, evals_per_move),
        "positions_in_timed_region": moves,
        "evals_per_move": evals_per_move,
        "skipped_leaves_in_timed_region": int(st["skipped_leaves"]),
        "phase_ms_profiled_call_mean_of_32_steps": phases,
        "roofline": roof,
        "roofline_tree_kernels": roof_tree,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "ms_per_step": ms_e2e / e2e_steps, "h2d_copied_bytes_per_step": h2d_copied,
                "h2d_read_in_place_bytes_per_step": h2d_in_place,
                "api": "kb_pool_step_hostio: leaf observations D2H -> H2D -> tower -> dense policy + value D2H -> value H2D, "
                       "expand reads policy[action] of the legal moves from the caller's pinned policy array"},
        "gpu_launches": int(st["kernel_launches"]),
        "host_numa_binding": numa,
        "clocks": clocks,
    }
    if cpu:
        out["cpu_baseline"] = cpu
    out.update(extras)
    emit(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--preroll", type=int, default=PREROLL_STEPS, help="untimed state-preparation steps (profiling runs shorten it)")
    ap.add_argument("--no-extras", action="store_true", help="skip the 20x256 tower and CPU baseline legs (profiling runs)")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        if args.steps == 2000 and args.warmup == 200:  # defaults sized for the GPU arm; keep the CPU arm to ~1 min
            args.steps, args.warmup = 40, 3
        run_reference(args, rank)
        return
    rank, world, local, dist = dist_setup(args.gpus)
    try:
        run_ours(args, rank, world, local, dist)
    finally:
        if dist is not None:
            dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- H*v throughput of the dmft-ed Lanczos hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload cfg4|cfg3|cfg2]

A "step" is one sector Hamiltonian-vector product y = H x on device-resident real fp64 vectors of the named
workload (default: BASELINE config 4 -- Norb=2, Nbath=7, Ns=16, half-filling sector (8,8), Dim=165,636,900).
`value` = whole-job H*v per second (matvec/s), timed with CUDA events on the launching stream, max over ranks.
`e2e`   = the same metric through the reference-facing C-ABI with HOST buffers: one sp_lanc_tridiag-equivalent
          call (ED_GF_NORMAL.f90:187-192) per step -- upload of the complex(8) start vector from pinned host
          memory, nlanc Lanczos steps on the device, alpha/beta back to the host -- counted as nlanc matvecs.
`roofline` = algorithmic bytes (2 * Dim * 8 B per H*v) / average H*v duration vs the measured HBM copy
          bandwidth (MEASURED_PEAKS.json).
`cpu_baseline` / `--impl reference` = the CPU oracle (literal C restatement of directMatVec_cc; the Fortran
          reference cannot be built in this image) on all host cores over a bounded row sample.
`parity`  = one H*v of the Philox start vector compared with the CPU oracle on whole reference rows that touch every
          (down-block, up-block) tile (oracle/parity_check.py), 1e-12 * |y|_inf; a mismatch fails the run.
With N>1 (torchrun) the default is ONE sector vector sharded over the ranks -> "scaling": "strong".  Round 2 shards by
CONSERVED OCCUPATION PAIRS (`--mode pairs`, edgpu_sector_build_shard): H is block diagonal over (down-block, up-block)
pairs, every rank owns whole pairs, H*v needs no exchange; the Lanczos scalars of the e2e leg are summed with NCCL inside
the C-ABI (edgpu_comm_init).  The line carries `parity_max_rel_err` of the sharded product against the oracle, a `cfg5`
sub-record (Ns=18: every GPU alone vs sharded over N) and `independent_chains` (one whole vector per GPU, weak scaling).
`--mode strips` is the round-1 scheme (up-spin column strips, NVLink exchange per product) kept for comparison.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (Norb, Nbath, nup, ndw, description)
    "cfg2": (1, 9, 5, 5, "ed_hm_bethe Norb=1 Nbath=9 Ns=10 sector (5,5) dim 63,504"),
    "cfg3": (2, 6, 7, 7, "ed_hm_2bands_bethe Norb=2 Nbath=6 Ns=14 sector (7,7) dim 11,778,624"),
    "cfg4": (2, 7, 8, 8, "ed_hm_bethe_Nbands Norb=2 Nbath=7 Ns=16 sector (8,8) dim 165,636,900"),
    "cfg5": (3, 5, 9, 9, "ed_hm_bethe_Nbands Norb=3 Nbath=5 Ns=18 sector (9,9) dim 2,363,904,400"),
}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
        except Exception:
            pass
    return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------
def cpu_reference_sample(workload, seconds_target=15.0, threads=None, steps=1, warmup=0):
    """Times the oracle's gather-form restatement of directMatVec_cc on all host cores over a bounded block of
    rows of the same sector (full map, full complex(8) input vector, recursive binary search).  Returns
    (matvec_per_s, cores, sample_description, per-step seconds)."""
    import numpy as np
    from oracle import ed_oracle as O
    O.build()
    Norb, Nbath, nup, ndw, _ = WORKLOADS[workload]
    p = O.Params(Norb=Norb, Nbath=Nbath, uloc=tuple([2.0] * Norb), lanc_method="lanczos", lanc_nstates_sector=1)
    bath = O.init_bath(p)
    model = O.Model(p, bath)
    smap = O.build_sector(p.Ns, nup, ndw)
    dim = smap.size
    rng = np.random.default_rng(20240607)
    vin = np.empty(dim, dtype=np.complex128)
    vin.real = rng.standard_normal(dim)
    vin.imag = 0.0
    L = O.lib()
    P = threads or L.ora_num_threads()
    import ctypes as C
    dp, u64p = C.POINTER(C.c_double), C.POINTER(C.c_uint64)
    # calibrate on a small block, then size the sample for ~seconds_target
    n0 = min(dim, 200_000 * P)
    hv = np.zeros(min(dim, max(n0, 1)), dtype=np.complex128)
    hv_full = np.zeros(dim, dtype=np.complex128) if dim <= 20_000_000 else None

    def run(nrows):
        out = hv_full if hv_full is not None else np.zeros(dim if nrows > hv.size else hv.size, dtype=np.complex128)
        return L.ora_gather_hxv_mt(model.h, smap.ctypes.data_as(u64p), dim, vin.view(np.float64).ctypes.data_as(dp),
                                   out.view(np.float64).ctypes.data_as(dp), 0, nrows, P)

    t0 = run(n0)
    rate = n0 / t0
    nrows = int(min(dim, max(n0, rate * seconds_target)))
    if hv_full is None and nrows > hv.size:
        hv_full = np.zeros(dim, dtype=np.complex128)
    times = []
    for i in range(warmup + steps):
        t = run(nrows)
        if i >= warmup:
            times.append(t)
    tavg = sum(times) / len(times)
    mv_per_s = (nrows / dim) / tavg
    sample = (f"rows [0,{nrows}) of {dim} ({100.0 * nrows / dim:.1f}% of one H*v) per step, gather form of "
              f"directMatVec_cc, complex(8) vectors, full map + recursive binary search, {P} pthreads, "
              f"{tavg:.2f} s per step")
    return mv_per_s, P, sample, tavg


def dim_of(workload):
    import math
    Norb, Nbath, nup, ndw, _ = WORKLOADS[workload]
    Ns = Norb * (Nbath + 1)
    return math.comb(Ns, nup) * math.comb(Ns, ndw)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # every step is a bounded sample; the whole --steps K --warmup W run is sized to ~2.5 minutes of CPU work
    nrun = max(1, args.steps) + min(args.warmup, 1)
    mv, P, sample, tavg = cpu_reference_sample(args.workload, seconds_target=min(args.cpu_seconds, 150.0 / nrun), steps=max(1, args.steps),
                                               warmup=min(args.warmup, 1))
    Norb, Nbath, nup, ndw, desc = WORKLOADS[args.workload]
    line = {
        "impl": "reference", "metric": "hxv_matvecs_per_s", "value": mv, "unit": "matvec/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tavg, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(args.workload, dim_of(args.workload), 2.0 * dim_of(args.workload) * 8.0 / max(1, args.gpus),
                               "1 GPU" if args.gpus == 1 else f"sector vector sharded by conserved occupation pairs over {args.gpus} ranks (LPT over "
                               "pair sizes); H*v has no exchange; Lanczos scalars by ncclAllReduce"),
        "note": "the Fortran reference cannot be built here (no Fortran compiler, SciFortran absent): this is the oracle port of its algorithm on the host cores",
        "cpu_baseline": {"value": mv, "unit": "matvec/s", "cores": P, "kind": "port", "sample": sample},
        "e2e": {"value": mv, "unit": "matvec/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def star_blocks(norb, nbath, n):
    """kernel launches of one spin pass for n particles: one per star-occupation block of >= 256 configurations plus one
    fringe launch for all smaller blocks (hxv_star.cu)"""
    import itertools
    import math
    sizes = [math.prod(math.comb(nbath + 1, m) for m in t) for t in itertools.product(range(nbath + 2), repeat=norb) if sum(t) == n]
    return sum(1 for z in sizes if z >= 256) + (1 if any(z < 256 for z in sizes) else 0)


def run_sharded(args, edb, world, rank, local):
    """N>1: ONE sector vector sharded by up-spin column blocks (north star / SURVEY 8e.2).  Down term local, up term
    through two NCCL all-to-all transposes per H*v.  value = H*v per second of the single sharded vector."""
    import ctypes as C
    import numpy as np
    import torch
    import torch.distributed as dist
    sharded = importlib.import_module("dmft-ed_b200.sharded")
    Norb, Nbath, nup, ndw, desc = WORKLOADS[args.workload]
    stream = torch.cuda.current_stream().cuda_stream
    ctx = edb.Context(Norb, Nbath, 1, True, device=local, stream=stream, layout=2, hxv_kernel=2)
    inp = edb.default_input(Norb=Norb, Nbath=Nbath, uloc=[2.0] * Norb)
    bath = np.zeros(edb.lib().ed_get_bath_dimension(inp))
    tmp = C.c_void_p()
    edb.lib().ed_init_solver(C.byref(inp), local, C.c_void_p(stream), bath.ctypes.data_as(edb.dp), bath.size, None, C.byref(tmp))
    edb.lib().ed_finalize_solver(tmp)
    ctx.set_hamiltonian(bath, [2.0] * Norb)
    s = ctx.sector(nup, ndw)
    g = torch.Generator(device="cuda").manual_seed(20240607 + rank)
    if args.exchange == "auto":
        args.exchange = "peer" if world <= 4 else "nccl"
    if args.exchange == "peer":
        sh = sharded.PeerShardedHxv(edb, s, rank, world, nvec=3)
        plan = sh.plan
        sh.vec(0)[:, :plan.ncols[rank]] = torch.randn(plan.dim_dw, plan.ncols[rank], dtype=torch.float64, device="cuda", generator=g)
        apply = lambda: sh.apply(0, 1)
        how = ("up term by a copy-engine kernel that reads x from and writes into the owners' shards over NVLink "
               "(CUDA IPC peer memory), 2 one-element all-reduce barriers per H*v")
    else:
        sh = sharded.make_gpu_shard(edb, s, rank, world, nchunks=args.chunks)
        plan = sh.plan
        x = sh.zeros()
        x[:, :plan.ncols[rank]] = torch.randn(plan.dim_dw, plan.ncols[rank], dtype=torch.float64, device="cuda", generator=g)
        y = sh.zeros()
        apply = lambda: sh.apply(x, y)
        how = f"up term via 2 NCCL all-to-all transposes per H*v, pipelined in {plan.nchunks} row groups"

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        apply()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sh.bytes_alltoall = sh.bytes_nvlink = 0
    barrier()
    ev0.record()
    for _ in range(args.steps):
        apply()
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    dim = s.dim
    alg_bytes = 2.0 * dim * 8.0
    peak, peak_src = measured_peaks()
    achieved = alg_bytes / (ms_step * 1e-3) / 1e9
    nvl = (sh.bytes_nvlink if args.exchange == "peer" else sh.bytes_alltoall) / args.steps
    # end to end: the column shard of the start vector comes from pinned host memory, nlanc sharded Lanczos steps
    # (H*v + all-reduced scalars + vector updates), alpha/beta go back to the host
    e2e = None
    if not args.no_e2e:
        nlanc = min(args.nlanc, 50)
        hostv = torch.empty(plan.dim_dw, plan.ldc[rank], dtype=torch.float64).pin_memory()
        hostv.zero_()
        hostv[:, :plan.ncols[rank]] = 1.0 / np.sqrt(dim)
        times = []
        for i in range(2):
            barrier()
            t0 = time.perf_counter()
            if args.exchange == "peer":
                sh.vec(0).copy_(hostv, non_blocking=True)
                al, be = sh.lanczos_tridiag(nlanc)
            else:
                x.copy_(hostv, non_blocking=True)
                al, be = sh.lanczos_tridiag(x, nlanc)
            barrier()
            times.append(time.perf_counter() - t0)
        te = torch.tensor([times[-1]], device="cuda", dtype=torch.float64)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        te = float(te.item())
        e2e = {"value": nlanc / te, "unit": "matvec/s", "h2d_bytes_per_step": int(world * hostv.numel() * 8),
               "d2h_bytes_per_step": int(2 * nlanc * 8),
               "call": f"pinned host column shards -> device, {nlanc} steps of the sharded Lanczos recurrence "
                       "(dmft-ed_b200/sharded.py: sharded_lanczos), alpha/beta -> host", "s_per_call": te,
               "alpha0": float(al[0])}
    # the other level of parallelism of the north star, for comparison in the same run: independent Lanczos chains, one
    # whole sector vector per GPU, no data-path collective -> aggregate H*v per second over all ranks
    chains = None
    try:
        xf, yf = s.vec(), s.vec()
        xf.fill_normal(20240607 + rank)
        for _ in range(args.warmup):
            s.hxv(xf, yf)
        barrier()
        ev0.record()
        for _ in range(args.steps):
            s.hxv(xf, yf)
        ev1.record()
        barrier()
        tc = torch.tensor([ev0.elapsed_time(ev1)], device="cuda", dtype=torch.float64)
        dist.all_reduce(tc, op=dist.ReduceOp.MAX)
        chains = {"value": world * args.steps / (float(tc.item()) * 1e-3), "unit": "matvec/s", "scaling": "weak",
                  "what": "independent chains: one whole sector vector per GPU, no collective (bench.py --mode chains)"}
        xf.free(); yf.free()
    except Exception as e:                   # e.g. the whole vector does not fit next to the shards
        chains = {"error": str(e)[:200]}
    if rank == 0:
        line = {
            "metric": "hxv_matvecs_per_s", "value": 1e3 / ms_step, "unit": "matvec/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {desc}", "bath": "init_dmft_bath noise=0 hwband=2", "uloc": 2.0,
                       "vector": "N(0,1) per rank", "dim": dim,
                       "l2": f"inputs exceed L2: {alg_bytes / world / 1e9:.3f} GB touched per rank and step",
                       "parallelism": f"sector vector sharded by up-spin column blocks over {world} ranks; down term local, " + how},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak * world, "unit": "GB/s", "frac": achieved / (peak * world),
                         "traffic": None, "peak_source": peak_src + f" x {world} GPUs", "algorithmic_bytes_per_launch_set": alg_bytes,
                         "nvlink_bytes_sent_per_rank_per_hxv": nvl,
                         "nvlink_floor_ms": nvl / 770e9 * 1e3, "nvlink_peak_source": "770 GB/s per direction (B200_PROFILING.md peer copy)"},
            "e2e": e2e, "cpu_baseline": None, "independent_chains": chains,
            "gpu_launches": int(args.steps * (star_blocks(Norb, Nbath, nup) + star_blocks(Norb, Nbath, ndw))), "clocks": clocks,
        }
        print(json.dumps(line))
    dist.destroy_process_group()
    return 0


def make_model_ctx(edb, workload, local, stream, layout=0, kernel=0, flags=0):
    """context + the synthetic Hamiltonian of SURVEY 8d (init_dmft_bath, Uloc=2, Ust=Jh=0, xmu=0, HFMODE=T)"""
    import ctypes as C
    import numpy as np
    Norb, Nbath, nup, ndw, desc = WORKLOADS[workload]
    ctx = edb.Context(Norb, Nbath, 1, True, device=local, stream=stream, layout=layout, hxv_kernel=kernel, debug_flags=flags)
    inp = edb.default_input(Norb=Norb, Nbath=Nbath, uloc=[2.0] * Norb)
    bath = np.zeros(edb.lib().ed_get_bath_dimension(inp))
    tmp = C.c_void_p()
    edb.lib().ed_init_solver(C.byref(inp), local, C.c_void_p(stream), bath.ctypes.data_as(edb.dp), bath.size, None, C.byref(tmp))
    edb.lib().ed_finalize_solver(tmp)
    ctx.set_hamiltonian(bath, [2.0] * Norb)
    return ctx, bath


PARITY_SEED = 20240607
PARITY_TOL = 1e-12


def oracle_model(workload, bath):
    from oracle import ed_oracle as O
    O.build()
    Norb, Nbath, nup, ndw, _ = WORKLOADS[workload]
    p = O.Params(Norb=Norb, Nbath=Nbath, uloc=tuple([2.0] * Norb), lanc_method="lanczos", lanc_nstates_sector=1)
    return O, O.Model(p, bath)


def parity_rows(workload, max_rows=24):
    from oracle import ed_oracle as O
    from oracle import parity_check as PC
    O.build()
    Norb, Nbath, nup, ndw, _ = WORKLOADS[workload]
    return PC.pick_rows(O, Norb, Nbath, nup, ndw, per_block=3, max_rows=max_rows)


def parity_check(workload, bath, rows, dim_up, fetch_rows):
    """oracle as the CHECKER of one GPU product (x = Philox uniforms, seed PARITY_SEED)"""
    from oracle import parity_check as PC
    O, model = oracle_model(workload, bath)
    Norb, Nbath, nup, ndw, _ = WORKLOADS[workload]
    res = PC.check_rows(O, model, nup, ndw, PARITY_SEED, rows, dim_up, fetch_rows)
    res["tol"] = PARITY_TOL
    res["pass"] = bool(res["max_rel_err"] < PARITY_TOL)
    res["what"] = ("y = H x, x = Philox uniforms: whole reference rows touching every (down-block, up-block) tile vs the "
                   "window oracle (direct/HxV*.f90 row rule)")
    return res


def run_pairs(args, edb, world, rank, local):
    """N>1: ONE sector vector sharded by conserved occupation pairs (edgpu_sector_build_shard): no exchange in H*v."""
    import ctypes as C
    import numpy as np
    import torch
    import torch.distributed as dist
    Norb, Nbath, nup, ndw, desc = WORKLOADS[args.workload]
    stream = torch.cuda.current_stream().cuda_stream
    ctx, bath = make_model_ctx(edb, args.workload, local, stream, flags=args.flags)
    # NCCL communicator of the C-ABI: rank 0 makes the id, the host broadcasts it (here through torch.distributed)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(ctx.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    ctx.comm_init(bytes(uid.cpu().numpy().tobytes()), rank, world)
    s = ctx.sector_shard(nup, ndw, rank, world)
    info = s.info()
    dim = s.dim
    x, y = s.vec(), s.vec()
    x.fill_uniform(PARITY_SEED)

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    def allmax(v):
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(args.warmup):
        s.hxv(x, y)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        s.hxv(x, y)
    ev1.record()
    barrier()
    ms_mine = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    ms_total = allmax(ms_mine)
    ms_step = ms_total / args.steps
    tmin = -allmax(-ms_mine) / args.steps
    # ---- parity of the sharded product: every rank contributes its elements of the sampled rows (zeros elsewhere) ----
    parity = None
    if not args.no_cpu:
        rows, dim_up, dim_dw = parity_rows(args.workload)
        got = {}
        for rd in rows:
            t = torch.from_numpy(y.download_rows(rd, rd + 1)).cuda()
            dist.all_reduce(t)
            got[rd] = t.cpu().numpy()
        if rank == 0:
            parity = parity_check(args.workload, bath, rows, dim_up, lambda a, b: got[a])
    # ---- end to end: complex(8) start vector in pinned host memory -> sharded chain through the C-ABI -> alpha/beta ----
    e2e = None
    if not args.no_e2e:
        nlanc = args.nlanc
        host = torch.empty(2 * dim, dtype=torch.float64).pin_memory()
        hv = host.numpy()
        hv[0::2] = 1.0 / np.sqrt(dim)
        hv[1::2] = 0.0
        a, b, nu = np.zeros(nlanc), np.zeros(nlanc), C.c_int32()
        times = []
        for i in range(1 + args.e2e_steps):
            barrier()
            t0 = time.perf_counter()
            ctx.check(edb.lib().edgpu_vec_upload(x.h, hv.ctypes.data, 1))
            t1 = time.perf_counter()
            ctx.check(edb.lib().edgpu_lanczos_tridiag(s.h, x.h, nlanc, 1e-13, a.ctypes.data_as(edb.dp), b.ctypes.data_as(edb.dp), C.byref(nu)))
            barrier()
            if i > 0:
                times.append((time.perf_counter() - t0, t1 - t0))
        te, tu = allmax(max(t[0] for t in times)), allmax(max(t[1] for t in times))
        e2e = {"value": nlanc / te, "unit": "matvec/s", "h2d_bytes_per_step": int(world * dim * 16), "d2h_bytes_per_step": int(2 * nlanc * 8),
               "call": f"edgpu_vec_upload(complex(8) host, every rank keeps its pairs) + edgpu_lanczos_tridiag(nlanc={nlanc}) on the "
                       "pair-sharded sector, Lanczos scalars summed by ncclAllReduce inside the C-ABI", "s_per_call": te, "upload_s": tu,
               "ms_per_lanczos_step": (te - tu) / nlanc * 1e3, "alpha0": float(a[0]), "beta1": float(b[1]) if nlanc > 1 else None}
        del host
    x.free(); y.free(); s.free()
    # ---- independent chains (weak scaling) and the cfg5 sub-record (the config the 70 % target names) ----
    chains = None
    try:
        sf = ctx.sector(nup, ndw)
        xf, yf = sf.vec().fill_normal(PARITY_SEED + rank), sf.vec()
        for _ in range(args.warmup):
            sf.hxv(xf, yf)
        barrier()
        ev0.record()
        for _ in range(args.steps):
            sf.hxv(xf, yf)
        ev1.record()
        barrier()
        tc = allmax(ev0.elapsed_time(ev1))
        chains = {"value": world * args.steps / (tc * 1e-3), "unit": "matvec/s", "scaling": "weak", "ms_per_hxv_1gpu": tc / args.steps,
                  "what": "independent chains: one whole sector vector per GPU, no collective"}
        xf.free(); yf.free(); sf.free()
    except Exception as e:
        chains = {"error": str(e)[:200]}
    sub5 = None
    if args.cfg5 and args.workload != "cfg5":
        try:
            sub5 = cfg5_record(args, edb, world, rank, local, stream, barrier, allmax)
        except Exception as e:
            sub5 = {"error": str(e)[:300]}
    alg_bytes = 2.0 * dim * 8.0
    peak, peak_src = measured_peaks()
    achieved = alg_bytes / (ms_step * 1e-3) / 1e9
    if rank == 0:
        eff1 = (chains["ms_per_hxv_1gpu"] / (world * ms_step)) if chains and "ms_per_hxv_1gpu" in chains else None
        line = {
            "metric": "hxv_matvecs_per_s", "value": 1e3 / ms_step, "unit": "matvec/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": bench_config(args.workload, dim, alg_bytes / world,
                                   f"sector vector sharded by conserved occupation pairs over {world} ranks (LPT over pair sizes); H*v has no "
                                   "exchange; Lanczos scalars by ncclAllReduce"),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak * world, "unit": "GB/s", "frac": achieved / (peak * world),
                         "traffic": None, "peak_source": peak_src + f" x {world} GPUs", "algorithmic_bytes_per_launch_set": alg_bytes,
                         "nvlink_bytes_per_hxv": 0, "ms_per_hxv_fastest_rank": tmin,
                         "load_balance": tmin / ms_step, "local_doubles_rank0": info["nalloc"]},
            "parity": parity, "parity_max_rel_err": parity["max_rel_err"] if parity else None,
            "e2e": e2e, "cpu_baseline": None, "independent_chains": chains,
            "efficiency_vs_1gpu_same_run": eff1, "cfg5": sub5,
            "gpu_launches": int(args.steps * 4), "clocks": clocks,
        }
        print(json.dumps(line))
        if parity is not None and not parity["pass"]:
            dist.destroy_process_group()
            return 1
    dist.destroy_process_group()
    return 0


def cfg5_record(args, edb, world, rank, local, stream, barrier, allmax):
    """Ns=18 (2.36e9 states, 18.9 GB per vector): every GPU alone on the whole vector, then the vector sharded by pairs over
    the ranks; efficiency = t(1 GPU) / (N * t(N GPUs)), device times, max over ranks."""
    import torch
    Norb, Nbath, nup, ndw, desc = WORKLOADS["cfg5"]
    ctx, bath = make_model_ctx(edb, "cfg5", local, stream)
    steps = 3
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(sec):
        x, y = sec.vec().fill_uniform(PARITY_SEED), sec.vec()
        sec.hxv(x, y); sec.hxv(x, y)
        barrier()
        ev0.record()
        for _ in range(steps):
            sec.hxv(x, y)
        ev1.record()
        barrier()
        t = allmax(ev0.elapsed_time(ev1)) / steps
        x.free(); y.free()
        return t

    s1 = ctx.sector(nup, ndw)
    t1 = timed(s1)
    dim = s1.dim
    s1.free()
    sn = ctx.sector_shard(nup, ndw, rank, world)
    tn = timed(sn)
    sn.free()
    ctx.close()
    return {"workload": f"cfg5: {desc}", "ms_per_hxv_1gpu": t1, "ms_per_hxv_sharded": tn, "n_gpus": world,
            "efficiency": t1 / (world * tn), "matvec_per_s_sharded": 1e3 / tn, "dim": dim,
            "definition": "t(1 GPU, whole vector, max over the N GPUs each running it alone) / (N * t(vector sharded by pairs over N GPUs, max over ranks)), CUDA events"}


def bench_config(workload, dim, bytes_per_rank, parallelism):
    desc = WORKLOADS[workload][4]
    return {"workload": f"{workload}: {desc}", "bath": "init_dmft_bath noise=0 hwband=2", "uloc": 2.0, "dim": dim,
            "vector": "Philox (counter = reference index)",
            "l2": (f"inputs exceed L2: {bytes_per_rank / 1e9:.3f} GB touched per rank and step" if bytes_per_rank > 3e8
                   else "vector fits in L2 (launch-bound case)"),
            "parallelism": parallelism}


# ------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="cfg4", choices=list(WORKLOADS))
    ap.add_argument("--nlanc", type=int, default=200, help="Lanczos steps per e2e call (reference lanc_ngfiter)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-solve", action="store_true", help="skip the ed_solve wall-time section")
    ap.add_argument("--no-solve-cfg3", action="store_true", help="skip the full ed_solve of BASELINE config 3 (Ns=14, 225 sectors)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--mode", default="auto", choices=["auto", "pairs", "strips", "chains"],
                    help="N>1: 'pairs' (default) = one sector vector sharded by conserved occupation pairs, no exchange; 'strips' = "
                         "round-1 scheme (up-spin column strips + NVLink exchange); 'chains' = independent H*v streams per rank")
    ap.add_argument("--cfg5", type=int, default=1, help="N>1: add the Ns=18 sub-record (1 GPU vs sharded)")
    ap.add_argument("--exchange", default="auto", choices=["auto", "peer", "nccl"],
                    help="sharded mode: exchange fused into the up kernel over peer memory (CUDA IPC), or NCCL all-to-all "
                         "transposes; auto = what measured faster on B200 x8 (peer up to 4 ranks, nccl at 8)")
    ap.add_argument("--chunks", type=int, default=1, help="row groups of the pipelined all-to-all exchange (sharded mode)")
    ap.add_argument("--flags", type=int, default=0, help="edgpu_params.reserved[0] test hooks (A/B runs of kernel variants)")
    ap.add_argument("--layout", type=int, default=0)
    ap.add_argument("--kernel", type=int, default=0)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl != "reference":
        args.warmup = max(args.warmup, 3) if os.environ.get("BENCH_ALLOW_SHORT_WARMUP") is None else args.warmup

    if args.impl == "reference":
        return run_reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    edb = importlib.import_module("dmft-ed_b200")
    mode = args.mode
    if mode == "auto":
        mode = "pairs" if world > 1 else "chains"
    if world > 1 and mode == "pairs":
        return run_pairs(args, edb, world, rank, local)
    if world > 1 and mode == "strips":
        return run_sharded(args, edb, world, rank, local)
    Norb, Nbath, nup, ndw, desc = WORKLOADS[args.workload]
    stream = torch.cuda.current_stream().cuda_stream
    ctx = edb.Context(Norb, Nbath, 1, True, device=local, stream=stream, layout=args.layout, hxv_kernel=args.kernel,
                      debug_flags=args.flags)
    # synthetic inputs: deterministic bath of init_dmft_bath, Uloc=2, Ust=Jh=0, xmu=0, HFMODE=T
    inp = edb.default_input(Norb=Norb, Nbath=Nbath, uloc=[2.0] * Norb)
    sol_bath = np.zeros(edb.lib().ed_get_bath_dimension(inp))
    # init_dmft_bath through the host mirror without creating a second context
    import ctypes as C
    tmp_solver = C.c_void_p()
    edb.lib().ed_init_solver(C.byref(inp), local, C.c_void_p(stream), sol_bath.ctypes.data_as(edb.dp), sol_bath.size, None,
                             C.byref(tmp_solver))
    edb.lib().ed_finalize_solver(tmp_solver)
    ctx.set_hamiltonian(sol_bath, [2.0] * Norb)
    s = ctx.sector(nup, ndw)
    dim = s.dim
    x, y = s.vec(), s.vec()
    x.fill_uniform(PARITY_SEED + rank)      # Philox uniforms in (-1,1), counter = reference index
    ctx.sync()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up, then K timed steps -------------------------------------------------------------------------
    for _ in range(args.warmup):
        s.hxv(x, y)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        s.hxv(x, y)
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * args.steps / (ms_total * 1e-3)

    # per-launch duration of the dominant kernel measured with CUDA events inside the library (same stream)
    ms_kernel, launches = s.bench_hxv(x, y, max(5, args.steps), flush_l2=False)
    alg_bytes = 2.0 * dim * 8.0
    peak, peak_src = measured_peaks()
    achieved = alg_bytes / (ms_kernel * 1e-3) / 1e9
    traffic = None
    tj = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tj):
        try:
            tjd = json.load(open(tj))
            # per-H*v DRAM bytes from an ncu capture; only valid for the kernel version it was taken at
            traffic = tjd.get(args.workload) if tjd.get("edgpu_version") == edb.lib().edgpu_version() else None
        except Exception:
            traffic = None

    # ---- end to end through the C-ABI with host buffers --------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        nlanc = args.nlanc
        host = torch.empty(2 * dim, dtype=torch.float64).pin_memory()      # complex(8) start vector, pinned
        hv = host.numpy()
        hv[0::2] = 1.0 / np.sqrt(dim)
        hv[1::2] = 0.0
        a = np.zeros(nlanc)
        b = np.zeros(nlanc)
        nu = C.c_int32()
        times, tup = [], []
        for i in range(1 + args.e2e_steps):
            barrier()
            t0 = time.perf_counter()
            ctx.check(edb.lib().edgpu_vec_upload(x.h, hv.ctypes.data, 1))
            t1 = time.perf_counter()                                   # (upload returns after the copy has completed)
            ctx.check(edb.lib().edgpu_lanczos_tridiag(s.h, x.h, nlanc, 1e-13, a.ctypes.data_as(edb.dp), b.ctypes.data_as(edb.dp),
                                                      C.byref(nu)))
            barrier()
            if i > 0:
                times.append(time.perf_counter() - t0)
                tup.append(t1 - t0)
        te = max(times)
        if world > 1:
            t = torch.tensor([te], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            te = float(t.item())
        e2e = {"value": world * nlanc / te, "unit": "matvec/s", "h2d_bytes_per_step": int(dim * 16),
               "d2h_bytes_per_step": int(2 * nlanc * 8),
               "call": f"edgpu_vec_upload(complex(8) host) + edgpu_lanczos_tridiag(nlanc={nlanc}) = sp_lanc_tridiag at "
                       "ED_GF_NORMAL.f90:187-192", "s_per_call": te, "upload_s": max(tup),
               "ms_per_lanczos_step": (te - max(tup)) / nlanc * 1e3}

    # ---- ed_solve wall time (BASELINE metric, second half): full sector scan + GF + Sigma + observables ------------
    # cfg1 / cfg2 / cfg3 on the GPU; beside them the CPU oracle's ed_solve (rank 0, same inputs): the full scan for cfg1, and
    # for cfg2 the half-filling window of 7 sectors (ED_SECTORS) on BOTH sides -- the oracle's full Ns=10 scan takes
    # ~4.5 minutes, which does not fit a bench run.
    solve = None
    if rank == 0 and world == 1 and not args.no_solve:
        solve = {}
        win2 = [(5, 5), (4, 5), (5, 4), (6, 5), (5, 6), (4, 4), (6, 6)]
        cases = [("cfg1", dict(Norb=1, Nbath=4), None), ("cfg2", dict(Norb=1, Nbath=9), None), ("cfg2_window", dict(Norb=1, Nbath=9), win2)]
        if not args.no_solve_cfg3:
            cases.append(("cfg3", dict(Norb=2, Nbath=6, uloc=[2.0, 2.0]), None))
        for name, kw, secs in cases:
            si = edb.default_input(lanc_method="lanczos", lanc_nstates_sector=1, ed_sparse_H=0, Lmats=1024, Lreal=1024, **kw)
            so = edb.Solver(si, device=local, stream=stream)
            if secs:
                so.set_sectors(secs)
            so.solve()                                    # warm-up (table builds, allocator)
            t0 = time.perf_counter()
            so.solve()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            st, zeta, egs = so.states()
            nsec = len(secs) if secs else (si.Nbath * si.Norb + si.Norb + 1) ** 2
            solve[name] = {"wall_s": dt, "phases_s": so.timings(), "egs": egs, "n_gs": len(st), "sectors": nsec,
                           "scan": "lanc_method=lanczos, direct H*v" + ("" if not secs else ", half-filling window")}
            so.close()
            if not args.no_cpu and name in ("cfg1", "cfg2_window"):
                from oracle import ed_oracle as O
                O.build()
                po = O.Params(lanc_method="lanczos", lanc_nstates_sector=1, Lmats=1024, Lreal=1024,
                              **{k: (tuple(v) if isinstance(v, list) else v) for k, v in kw.items()})
                t0 = time.perf_counter()
                ro = O.ed_solve(po, O.init_bath(po), sectors=secs)
                solve[name]["cpu_oracle_wall_s"] = time.perf_counter() - t0
                solve[name]["cpu_oracle_egs"] = ro.egs
                solve[name]["cpu_kind"] = "port, 1 thread (numpy + C oracle)"

    # ---- parity of the timed product against the oracle (rank 0, N=1): sampled reference rows -------------------------
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu:
        x.fill_uniform(PARITY_SEED)
        s.hxv(x, y)
        rows, dim_up, _ = parity_rows(args.workload)
        parity = parity_check(args.workload, sol_bath, rows, dim_up, y.download_rows)

    # ---- CPU baseline beside it (rank 0, N=1 only) ----------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        x.free(); y.free()
        mv, P, sample, _ = cpu_reference_sample(args.workload, seconds_target=args.cpu_seconds)
        cpu = {"value": mv, "unit": "matvec/s", "cores": P, "kind": "port", "sample": sample}

    if rank == 0:
        line = {
            "metric": "hxv_matvecs_per_s", "value": value, "unit": "matvec/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if world == 1 else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": bench_config(args.workload, dim, alg_bytes, "1 GPU" if world == 1 else f"{world} ranks, independent H*v streams (chain-level)"),
            "parity": parity, "parity_max_rel_err": parity["max_rel_err"] if parity else None,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch_set": alg_bytes,
                         "ms_per_hxv_events": ms_kernel},
            "e2e": e2e, "cpu_baseline": cpu, "ed_solve": solve, "gpu_launches": int(launches * args.steps / max(5, args.steps)),
            "clocks": clocks,
        }
        print(json.dumps(line))
        if parity is not None and not parity["pass"]:
            return 1
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

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
With N>1 (torchrun) the default is ONE sector vector sharded by up-spin column blocks over the ranks (BASELINE config 4:
"direct H*v sharded over 1/2/4/8 GPUs") -> "scaling": "strong".  The down term is local; the up term exchanges rows either
inside the copy-engine up kernel over CUDA-IPC peer memory (`--exchange peer`, default up to 4 ranks) or through two NCCL
all-to-all transposes (`--exchange nccl`, default at 8 ranks).  The same line carries `independent_chains` (the other
level of parallelism of the north star measured in the same run: one whole vector per GPU, no collective, weak scaling;
`--mode chains` makes it the primary value) and an `e2e` leg (pinned host shards -> device, sharded Lanczos steps,
alpha/beta -> host).
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


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    mv, P, sample, tavg = cpu_reference_sample(args.workload, seconds_target=args.cpu_seconds, steps=max(1, args.steps),
                                               warmup=min(args.warmup, 1))
    Norb, Nbath, nup, ndw, desc = WORKLOADS[args.workload]
    line = {
        "impl": "reference", "metric": "hxv_matvecs_per_s", "value": mv, "unit": "matvec/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tavg, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}", "bath": "init_dmft_bath noise=0 hwband=2", "uloc": 2.0,
                   "note": "the Fortran reference cannot be built here (no Fortran compiler, SciFortran absent): this "
                           "is the oracle port of its algorithm"},
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
    ap.add_argument("--solve-cfg3", action="store_true", help="also time a full ed_solve of BASELINE config 3 (Ns=14, 225 sectors)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--mode", default="auto", choices=["auto", "shard", "chains"],
                    help="N>1: 'shard' = one sector vector sharded by up-spin column blocks with all-to-all transposes "
                         "(strong scaling); 'chains' = independent H*v streams per rank (weak scaling)")
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
        mode = "shard" if world > 1 else "chains"
    if world > 1 and mode == "shard":
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
    x.fill_normal(20240607 + rank)          # Philox N(0,1), SURVEY 8d
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
            traffic = json.load(open(tj)).get(args.workload)
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
    solve = None
    if rank == 0 and world == 1 and not args.no_solve:
        solve = {}
        cases = [("cfg1", dict(Norb=1, Nbath=4)), ("cfg2", dict(Norb=1, Nbath=9))]
        if args.solve_cfg3:
            cases.append(("cfg3", dict(Norb=2, Nbath=6, uloc=[2.0, 2.0])))
        for name, kw in cases:
            si = edb.default_input(lanc_method="lanczos", lanc_nstates_sector=1, ed_sparse_H=0, Lmats=1024, Lreal=1024, **kw)
            so = edb.Solver(si, device=local, stream=stream)
            so.solve()                                    # warm-up (table builds, allocator)
            t0 = time.perf_counter()
            so.solve()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            st, zeta, egs = so.states()
            solve[name] = {"wall_s": dt, "phases_s": so.timings(), "egs": egs, "n_gs": len(st),
                           "sectors": (si.Nbath * si.Norb + si.Norb + 1) ** 2, "scan": "all sectors, lanc_method=lanczos, direct H*v"}
            so.close()

    # ---- CPU baseline beside it (rank 0, N=1 only) ----------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        x.free(); y.free()
        mv, P, sample, _ = cpu_reference_sample(args.workload, seconds_target=args.cpu_seconds)
        cpu = {"value": mv, "unit": "matvec/s", "cores": P, "kind": "port", "sample": sample}

    if rank == 0:
        line = {
            "metric": "hxv_matvecs_per_s", "value": value, "unit": "matvec/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {desc}", "bath": "init_dmft_bath noise=0 hwband=2", "uloc": 2.0,
                       "vector": "Philox N(0,1) seed 20240607", "dim": dim,
                       "l2": f"inputs exceed L2: {alg_bytes / 1e9:.3f} GB touched per step" if alg_bytes > 3e8 else "vector fits in L2 (launch-bound case)",
                       "parallelism": "1 GPU" if world == 1 else f"{world} ranks, independent H*v streams (chain-level)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch_set": alg_bytes,
                         "ms_per_hxv_events": ms_kernel},
            "e2e": e2e, "cpu_baseline": cpu, "ed_solve": solve, "gpu_launches": int(launches * args.steps / max(5, args.steps)),
            "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

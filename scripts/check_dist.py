"""torchrun script (N >= 2 GPUs): correctness of the two multi-GPU levels against the single-GPU results of the SAME run.

  1. pair-sharded sector vector: H*v rows and a Lanczos chain (alpha/beta through edgpu_lanczos_tridiag with the
     NCCL-reduced scalars) equal the unsharded ones on every rank;
  2. distributed ed_solve (sectors + ground-state chains dealt over the ranks, ed_set_comm): Sigma(iw), G, densities
     equal the single-GPU solve.
Prints one JSON line on rank 0; exit code 1 on a mismatch.  usage: torchrun --nproc-per-node N scripts/check_dist.py"""
import importlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    edb = importlib.import_module("dmft-ed_b200")
    out, ok = {}, True

    def bcast_uid(make):
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(make()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        return bytes(uid.cpu().numpy().tobytes())

    # ---- 1. pair-sharded vector: Norb=2, Nbath=5 (Ns=12), sector (6,6): 853,776 states ----
    Norb, Nbath, sec = 2, 5, (6, 6)
    rng = np.random.default_rng(5)
    bath = np.concatenate([np.linspace(-2, 2, Nbath).repeat(1)] * Norb + [np.full(Norb * Nbath, 0.45)]) + 0.03 * rng.normal(size=2 * Norb * Nbath)
    ctx = edb.Context(Norb, Nbath, 1, True, device=local, hxv_kernel=3)
    ctx.set_hamiltonian(bath, [2.0, 1.5], ust=0.8, jh=0.1)
    ctx.comm_init(bcast_uid(ctx.comm_unique_id), rank, world)
    sfull, sh = ctx.sector(*sec), ctx.sector_shard(*sec, rank, world)
    xf, yf, xs, ys = sfull.vec().fill_uniform(11), sfull.vec(), sh.vec().fill_uniform(11), sh.vec()
    sfull.hxv(xf, yf)
    sh.hxv(xs, ys)
    rows = [0, 1, sfull.dim_dw // 3, sfull.dim_dw // 2, sfull.dim_dw - 1]
    err = 0.0
    for rd in rows:
        t = torch.from_numpy(ys.download_rows(rd, rd + 1)).cuda()
        dist.all_reduce(t)
        ref = yf.download_rows(rd, rd + 1)
        err = max(err, float(np.abs(t.cpu().numpy() - ref).max() / np.abs(ref).max()))
    a1, b1, _ = sfull.lanczos_tridiag(xf, 40)
    a2, b2, _ = sh.lanczos_tridiag(xs, 40)
    out["shard_hxv_max_rel_err"] = err
    out["shard_chain_alpha_err"] = float(np.abs(a1[:12] - a2[:12]).max())
    out["shard_chain_beta_err"] = float(np.abs(b1[:12] - b2[:12]).max())
    out["shard_local_fraction"] = sh.info()["nalloc"] / sfull.info()["nalloc"]
    ok &= err < 1e-12 and out["shard_chain_alpha_err"] < 1e-9 and out["shard_chain_beta_err"] < 1e-9
    for v in (xf, yf, xs, ys):
        v.free()
    sfull.free(); sh.free(); ctx.close()

    # ---- 2. distributed ed_solve: Norb=2, Nbath=3 (Ns=8), all 81 sectors ----
    kw = dict(Norb=2, Nbath=3, uloc=[2.0, 2.0], ust=1.2, jh=0.2, lanc_method="lanczos", lanc_nstates_sector=1, ed_sparse_H=0,
              Lmats=128, Lreal=64, lanc_dim_threshold=64, chispin_flag=1, Ltau=50, beta=50.0)
    s1 = edb.Solver(edb.default_input(**kw), device=local)
    s1.solve()
    sd = edb.Solver(edb.default_input(**kw), device=local)
    sd.set_comm(bcast_uid(sd.comm_unique_id), rank, world)
    sd.solve()
    e_s = float(np.abs(sd.sigma_matsubara() - s1.sigma_matsubara()).max())
    e_g = float(np.abs(sd.gimp_matsubara() - s1.gimp_matsubara()).max())
    e_d = float(max(np.abs(sd.dens() - s1.dens()).max(), np.abs(sd.docc() - s1.docc()).max()))
    e_c = float(np.abs(sd.spinchi()[1] - s1.spinchi()[1]).max())
    st1, z1, eg1 = s1.states()
    std, zd, egd = sd.states()
    out.update(ed_solve_sigma_err=e_s, ed_solve_gimp_err=e_g, ed_solve_obs_err=e_d, ed_solve_chi_err=e_c, zeta=[z1, zd],
               local_states=len(std), egs_err=abs(eg1 - egd), diag_s=[s1.timings()["diag"], sd.timings()["diag"]])
    ok &= e_s < 1e-8 and e_g < 1e-8 and e_d < 1e-9 and e_c < 1e-8 and z1 == zd and abs(eg1 - egd) < 1e-10
    s1.close(); sd.close()

    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        out["pass"] = bool(flag.item() > 0.5)
        print(json.dumps(out))
    dist.destroy_process_group()
    return 0 if flag.item() > 0.5 else 1


if __name__ == "__main__":
    sys.exit(main())

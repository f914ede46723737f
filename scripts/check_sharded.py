#!/usr/bin/env python
"""torchrun --nproc-per-node N scripts/check_sharded.py [Norb Nbath nup ndw]
Sharded H*v over N GPUs (NCCL all-to-all) vs the single-GPU product computed redundantly on every rank."""
import ctypes as C
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
edb = importlib.import_module("dmft-ed_b200")
sharded = importlib.import_module("dmft-ed_b200.sharded")

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
Norb, Nbath, nup, ndw = [int(a) for a in sys.argv[1:5]] if len(sys.argv) >= 5 else (2, 6, 7, 7)
ctx = edb.Context(Norb, Nbath, 1, True, device=local, stream=torch.cuda.current_stream().cuda_stream, layout=2, hxv_kernel=2)
inp = edb.default_input(Norb=Norb, Nbath=Nbath, uloc=[2.0] * Norb)
bath = np.zeros(edb.lib().ed_get_bath_dimension(inp))
tmp = C.c_void_p()
edb.lib().ed_init_solver(C.byref(inp), local, None, bath.ctypes.data_as(edb.dp), bath.size, None, C.byref(tmp))
edb.lib().ed_finalize_solver(tmp)
ctx.set_hamiltonian(bath, [2.0] * Norb, ust=0.5 if Norb > 1 else 0.0, jh=0.1 if Norb > 1 else 0.0)
s = ctx.sector(nup, ndw)
sh = sharded.make_gpu_shard(edb, s, rank, world, nchunks=int(os.environ.get('CHUNKS', '3')))
plan = sh.plan
# full product on every rank (reference for the check)
x, y = s.vec().fill_normal(20240607), s.vec()
s.hxv(x, y)
ctx.sync()
du, dd, ld = s.dim_up, s.dim_dw, plan.ld
r2iu, r2id = np.zeros(du, dtype=np.uint32), np.zeros(dd, dtype=np.uint32)
ctx.check(edb.lib().edgpu_shard_perm(s.h, r2iu.ctypes.data, r2id.ctypes.data))
Xref, Yref = x.download().reshape(dd, du), y.download().reshape(dd, du)
c0, nc = plan.col0[rank], plan.ncols[rank]
iu = np.argsort(r2iu)          # internal column -> reference column
idw = np.argsort(r2id)
x_loc = sh.zeros()
x_loc[:, :nc] = torch.tensor(Xref[np.ix_(idw, iu[c0:c0 + nc])], device=x_loc.device)
y_loc = sh.zeros()
sh.apply(x_loc, y_loc)
torch.cuda.synchronize()
err = np.abs(y_loc[:, :nc].cpu().numpy() - Yref[np.ix_(idw, iu[c0:c0 + nc])]).max() / np.abs(Yref).max()
t = torch.tensor([err], device="cuda", dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"sharded H*v world={world} sector ({nup},{ndw}) dim={s.dim}: max rel err vs single GPU = {t.item():.3e}")
    assert t.item() < 1e-12
# peer mode: the exchange fused into the up-pass kernel (CUDA IPC mappings of the other ranks' shards)
ps = sharded.PeerShardedHxv(edb, s, rank, world)
ps.vec(0).copy_(x_loc)
for _ in range(2):                      # twice: the scratch shards are reused
    ps.apply(0, 1)
torch.cuda.synchronize()
err = np.abs(ps.vec(1)[:, :nc].cpu().numpy() - Yref[np.ix_(idw, iu[c0:c0 + nc])]).max() / np.abs(Yref).max()
t = torch.tensor([err], device="cuda", dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"peer-mode H*v world={world}: max rel err vs single GPU = {t.item():.3e}")
    assert t.item() < 1e-12
dist.barrier()
dist.destroy_process_group()

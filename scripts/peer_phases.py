#!/usr/bin/env python
"""torchrun --nproc-per-node N scripts/peer_phases.py [workload]: CUDA-event times of the phases of the peer-mode sharded H*v."""
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
Norb, Nbath, nup, ndw = [int(a) for a in sys.argv[1:5]] if len(sys.argv) >= 5 else (2, 7, 8, 8)
ctx = edb.Context(Norb, Nbath, 1, True, device=local, stream=torch.cuda.current_stream().cuda_stream, layout=2, hxv_kernel=2)
inp = edb.default_input(Norb=Norb, Nbath=Nbath, uloc=[2.0] * Norb)
bath = np.zeros(edb.lib().ed_get_bath_dimension(inp))
tmp = C.c_void_p()
edb.lib().ed_init_solver(C.byref(inp), local, None, bath.ctypes.data_as(edb.dp), bath.size, None, C.byref(tmp))
edb.lib().ed_finalize_solver(tmp)
ctx.set_hamiltonian(bath, [2.0] * Norb)
s = ctx.sector(nup, ndw)
ps = sharded.PeerShardedHxv(edb, s, rank, world)
P, me, L = ps.plan, rank, edb.lib()
ps.vec(0).normal_()
x, y, t = ps.bufs[0].t, ps.bufs[1].t, ps.bufs[-1].t
ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
acc = np.zeros(5)
for it in range(8):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    ev[0].record()
    s.ctx.check(L.edgpu_shard_hxv_dw(s.h, P.ncols[me], ps.ldc, x.data_ptr(), y.data_ptr()))
    ev[1].record()
    ps.barrier()
    ev[2].record()
    s.ctx.check(L.edgpu_shard_hxv_up_peers(s.h, P.row0[me], P.nrows[me], world, ps._col0, ps._ldc, ps._parr[0], None, ps._parr[-1], 0))
    ev[3].record()
    ps.barrier()
    ev[4].record()
    y += t
    ev[5].record()
    torch.cuda.synchronize()
    if it >= 3:
        acc += np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(5)])
acc /= 5
# where does the up pass lose time?  variants with local stand-ins for the remote shards (timing only, results discarded)
loc_x = (C.c_void_p * world)(*[ps.bufs[0].ptr] * world)
loc_y = (C.c_void_p * world)(*[ps.bufs[-1].ptr] * world)
var = {}
for name, xa, ya in (("all-local", loc_x, loc_y), ("remote-x", ps._parr[0], loc_y), ("remote-y", loc_x, ps._parr[-1]), ("both", ps._parr[0], ps._parr[-1])):
    tt = 0.0
    for it in range(6):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        ev[0].record()
        s.ctx.check(L.edgpu_shard_hxv_up_peers(s.h, P.row0[me], P.nrows[me], world, ps._col0, ps._ldc, xa, None, ya, 0))
        ev[1].record()
        torch.cuda.synchronize()
        if it >= 2:
            tt += ev[0].elapsed_time(ev[1]) / 4
    var[name] = tt
print(f"rank {rank}: up variants " + "  ".join(f"{k} {v:.3f}" for k, v in var.items()), flush=True)
print(f"rank {rank}: dw {acc[0]:.3f}  barrier {acc[1]:.3f}  up_peers {acc[2]:.3f}  barrier {acc[3]:.3f}  add {acc[4]:.3f}  ms", flush=True)
dist.barrier()
dist.destroy_process_group()

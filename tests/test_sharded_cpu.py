"""CPU, world_size=2 (gloo): the exchange logic of the sharded sector vector (dmft-ed_b200/sharded.py) with dense
stand-in operators for the local kernels.  y = D o X + H_dw X + X H_up^T must come out exactly as on one rank."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class DenseOps:
    """Stand-in for the CUDA kernels: dense H_dw (rows), H_up (columns), diagonal D (full tile)."""

    def __init__(self, Hdw, Hup, D, col0, ncols):
        self.Hdw, self.Hup, self.D, self.col0, self.ncols = Hdw, Hup, D, col0, ncols

    def dw(self, x_cols, y_cols):
        y_cols.zero_()
        y_cols[:, :self.ncols] = self.Hdw @ x_cols[:, :self.ncols]

    def up(self, row0, nrows, x_rows, y_rows):
        du = self.Hup.shape[0]
        y_rows.zero_()
        y_rows[:, :du] = self.D[row0:row0 + nrows] * x_rows[:, :du] + x_rows[:, :du] @ self.Hup.T


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def worker(rank, world, port, dim_up, dim_dw, q, nchunks=1):
    sys.path.insert(0, ROOT)
    sharded = importlib.import_module("dmft-ed_b200.sharded")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(1234)
    Hdw = torch.randn(dim_dw, dim_dw, dtype=torch.float64, generator=g)
    Hdw = Hdw + Hdw.T
    Hup = torch.randn(dim_up, dim_up, dtype=torch.float64, generator=g)
    Hup = Hup + Hup.T
    D = torch.randn(dim_dw, dim_up, dtype=torch.float64, generator=g)
    X = torch.randn(dim_dw, dim_up, dtype=torch.float64, generator=g)
    ld = (dim_up + 3) // 4 * 4
    plan = sharded.ShardPlan(dim_up, dim_dw, ld, world, nchunks)
    ops = DenseOps(Hdw, Hup, D, plan.col0[rank], plan.ncols[rank])
    sh = sharded.ShardedHxv(plan, rank, ops)
    x_loc = sh.zeros()
    x_loc[:, :plan.ncols[rank]] = X[:, plan.col0[rank]:plan.col0[rank] + plan.ncols[rank]]
    y_loc = sh.zeros()
    sh.apply(x_loc, y_loc)
    Y = D * X + Hdw @ X + X @ Hup.T
    err = (y_loc[:, :plan.ncols[rank]] - Y[:, plan.col0[rank]:plan.col0[rank] + plan.ncols[rank]]).abs().max().item()
    pad = y_loc[:, plan.ncols[rank]:].abs().max().item() if plan.ldc[rank] > plan.ncols[rank] else 0.0
    # Lanczos scalars through allreduce: alpha_1 of the chain started from X equals <X|H|X>/<X|X>
    a, b = sh.lanczos_tridiag(x_loc.clone(), 4)
    a_ref = (X * Y).sum().item() / (X * X).sum().item()
    # the whole recurrence (unnormalised vectors, all-reduced scalars) against the textbook one on the dense operator
    def H(V):
        return D * V + Hdw @ V + V @ Hup.T
    vin, vout, bb, ar, br = X / X.norm(), torch.zeros_like(X), 0.0, [], [0.0]
    for _ in range(4):
        t = H(vin) - bb * vout
        aa = (vin * t).sum().item()
        t = t - aa * vin
        bb = t.norm().item()
        vout, vin = vin, t / bb
        ar.append(aa)
        br.append(bb)
    dlanc = max(np.abs(a - np.array(ar)).max(), np.abs(b - np.array(br[:4])).max()) / max(1.0, np.abs(ar).max())
    q.put((rank, err, pad, max(abs(a[0] - a_ref), dlanc), sum(plan.ncols), sum(plan.nrows)))
    dist.destroy_process_group()


@pytest.mark.parametrize("dim_up,dim_dw,nchunks", [(10, 10, 1), (35, 21, 1), (126, 5, 1), (7, 64, 1), (35, 21, 3), (7, 64, 4), (126, 5, 4)])
def test_sharded_exchange_world2(dim_up, dim_dw, nchunks):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=worker, args=(r, 2, port, dim_up, dim_dw, q, nchunks)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, pad, da, ncols, nrows in res:
        assert err < 1e-10, (rank, err)
        assert pad == 0.0
        assert da < 1e-10
        assert ncols == dim_up and nrows == dim_dw


def test_shard_plan_follows_reference_split():
    sys.path.insert(0, ROOT)
    sharded = importlib.import_module("dmft-ed_b200.sharded")
    p = sharded.ShardPlan(12870, 12870, 12872, 8)
    assert sum(p.ncols) == 12870 and sum(p.nrows) == 12870
    assert p.nrows[:7] == [12870 // 8] * 7 and p.nrows[7] == 12870 // 8 + 12870 % 8      # ED_HAMILTONIAN.f90:56-62
    assert all(c % 4 == 0 for c in p.col0) and all(l % 4 == 0 for l in p.ldc)
    assert p.col0[-1] + p.ldc[-1] == p.ld
    # pipelined exchange: every row group is dealt by the same rule and the groups tile the rows
    p4 = sharded.ShardPlan(12870, 12870, 12872, 8, 4)
    assert sum(p4.nrows) == 12870 and len(p4.chunks) == 4
    nxt = 0
    for g0, gl, r0, nr in p4.chunks:
        assert g0 == nxt and sum(nr) == gl and r0 == [g0 + i * (gl // 8) for i in range(8)]
        nxt = g0 + gl
    assert nxt == 12870

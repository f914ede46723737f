"""Host-side plumbing of the sharded sector vector (multi-GPU, one process per GPU).

North star / SURVEY 8e.2: the Dimdw x Dimup tile is sharded by up-spin COLUMN blocks; the down-spin term is
applied locally (libedgpu: edgpu_shard_hxv_dw), the up-spin term needs whole rows and goes through an all-to-all
transpose (torch.distributed over NCCL/NVLink), is applied on the row shard (edgpu_shard_hxv_up) and transposed
back; Lanczos scalars are all-reduced.  This replaces directMatVec_MPI_cc (ED_HAMILTONIAN_DIRECT_HxV.f90:97-195),
which all-gathers the FULL vector on every rank for every product (:163-166).

torch is used for device memory and the collectives only; every flop of H*v runs in the CUDA library.  The local
kernels are injected (`ops`), so the exchange logic is testable on CPU with gloo and dense stand-in operators
(tests/test_sharded_cpu.py).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist


@dataclass
class ShardPlan:
    """Partition of the tile.  Columns are dealt in strips of 4 (32-byte aligned rows of the column shard); rows and
    strips follow the reference's rule Q = N/P with the remainder on the last rank (ED_HAMILTONIAN.f90:56-62)."""
    dim_up: int
    dim_dw: int
    ld: int            # leading dimension of a full row (dim_up rounded up to 4)
    world: int

    def __post_init__(self):
        nstrips = (self.dim_up + 3) // 4
        assert self.ld == nstrips * 4, "full rows must be padded to a multiple of 4"
        q = nstrips // self.world
        self.col0, self.ncols, self.ldc = [], [], []
        for p in range(self.world):
            s0 = p * q
            ns = q if p < self.world - 1 else nstrips - s0
            c0 = s0 * 4
            self.col0.append(c0)
            self.ldc.append(ns * 4)
            self.ncols.append(max(0, min(self.dim_up, c0 + ns * 4) - c0))
        qr = self.dim_dw // self.world
        self.row0 = [p * qr for p in range(self.world)]
        self.nrows = [qr if p < self.world - 1 else self.dim_dw - p * qr for p in range(self.world)]


class ShardedHxv:
    """y_loc = (H x)_loc on the column shard of this rank.

    ops.dw(x_cols, y_cols)               y_cols = H_dw x_cols           on [dim_dw, ldc_me]
    ops.up(row0, nrows, x_rows, y_rows)  y_rows = (D + H_up) x_rows     on [nrows_me, ld]
    """

    def __init__(self, plan: ShardPlan, rank: int, ops, group=None, device="cpu"):
        self.plan, self.rank, self.ops, self.group, self.device = plan, rank, ops, group, device
        P = plan
        me = rank
        self.ldc = P.ldc[me]
        self.nrows = P.nrows[me]
        f64 = torch.float64
        # exchange buffers (reused by every product)
        self.recv_cols = torch.zeros(sum(self.nrows * P.ldc[p] for p in range(P.world)), dtype=f64, device=device)
        self.x_rows = torch.zeros(self.nrows, P.ld, dtype=f64, device=device)
        self.y_rows = torch.zeros(self.nrows, P.ld, dtype=f64, device=device)
        self.send_rows = torch.zeros(sum(self.nrows * P.ldc[p] for p in range(P.world)), dtype=f64, device=device)
        self.tmp_cols = torch.zeros(P.dim_dw, self.ldc, dtype=f64, device=device)
        self.in_split = [P.nrows[p] * self.ldc for p in range(P.world)]          # my columns, rows of rank p
        self.out_split = [self.nrows * P.ldc[p] for p in range(P.world)]         # my rows, columns of rank p
        self.bytes_alltoall = 0

    def zeros(self):
        return torch.zeros(self.plan.dim_dw, self.ldc, dtype=torch.float64, device=self.device)

    def _all_to_all(self, out, inp, out_split, in_split, async_op=False):
        if self.plan.world == 1:
            out.copy_(inp)
            return None
        self.bytes_alltoall += 8 * (sum(in_split) - in_split[self.rank])
        return dist.all_to_all_single(out, inp, output_split_sizes=out_split, input_split_sizes=in_split, group=self.group,
                                      async_op=async_op)

    def apply(self, x_cols: torch.Tensor, y_cols: torch.Tensor):
        P = self.plan
        if hasattr(self.ops, "up_slabs"):
            # transpose #1 runs on NCCL's stream WHILE the down-spin term (local: down hops keep the column) runs on
            # the compute stream; the up-spin term then works directly on the received slabs and writes the slabs
            # that transpose #2 sends back -- no pack/unpack passes.
            work = self._all_to_all(self.recv_cols, x_cols.reshape(-1), self.out_split, self.in_split, async_op=True)
            self.ops.dw(x_cols, y_cols)
            if work is not None:
                work.wait()
            self.ops.up_slabs(P.row0[self.rank], self.nrows, P.col0, P.ldc, self.recv_cols, self.send_rows)
            self._all_to_all(self.tmp_cols.reshape(-1), self.send_rows, self.in_split, self.out_split)
            y_cols += self.tmp_cols
            return y_cols
        # 1. down-spin term on the column shard (local: down hops keep the column)
        self.ops.dw(x_cols, y_cols)
        # 2. transpose #1: slab of rows [row0[p], row0[p]+nrows[p]) of my columns -> rank p (slabs are contiguous)
        self._all_to_all(self.recv_cols, x_cols.reshape(-1), self.out_split, self.in_split)
        off = 0
        for p in range(P.world):
            n = self.nrows * P.ldc[p]
            self.x_rows[:, P.col0[p]:P.col0[p] + P.ldc[p]] = self.recv_cols[off:off + n].view(self.nrows, P.ldc[p])
            off += n
        # 3. diagonal + up-spin term on whole rows
        self.ops.up(P.row0[self.rank], self.nrows, self.x_rows, self.y_rows)
        # 4. transpose #2 back to column shards and accumulate
        off = 0
        for p in range(P.world):
            n = self.nrows * P.ldc[p]
            self.send_rows[off:off + n].view(self.nrows, P.ldc[p]).copy_(self.y_rows[:, P.col0[p]:P.col0[p] + P.ldc[p]])
            off += n
        self._all_to_all(self.tmp_cols.reshape(-1), self.send_rows, self.in_split, self.out_split)
        y_cols += self.tmp_cols
        return y_cols

    # ---- Lanczos scalars: local partial + allreduce (SciFortran does an MPI allreduce per dot product) ----------
    def dot(self, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        s = torch.dot(a.reshape(-1), b.reshape(-1)).reshape(1)
        if self.plan.world > 1:
            dist.all_reduce(s, group=self.group)
        return s

    def lanczos_tridiag(self, v_cols: torch.Tensor, nlanc: int):
        """sp_lanc_tridiag recurrence (.repo/PLAIN_LANCZOS.f90:87-118,154-180) on the sharded vector."""
        vin = v_cols / torch.sqrt(self.dot(v_cols, v_cols))
        vout = torch.zeros_like(vin)
        tmp = torch.zeros_like(vin)
        alfa, beta = [], [0.0]
        b = torch.zeros(1, dtype=torch.float64, device=self.device)
        for _ in range(nlanc):
            self.apply(vin, tmp)
            tmp -= b * vout
            a = self.dot(vin, tmp)
            tmp -= a * vin
            b = torch.sqrt(self.dot(tmp, tmp))
            vout, vin, tmp = vin, tmp / b, vout
            alfa.append(a)
            beta.append(b)
        return torch.cat(alfa).cpu().numpy(), torch.cat([torch.zeros(1, dtype=torch.float64, device=self.device)] + beta[1:]).cpu().numpy()[:nlanc]


class GpuOps:
    """Local kernels = the CUDA library on caller-owned device pointers."""

    def __init__(self, edb, sector):
        self.edb, self.s = edb, sector

    def dw(self, x_cols, y_cols):
        assert x_cols.is_contiguous() and y_cols.is_contiguous()
        ncols_valid = getattr(self, "ncols_valid")
        self.s.ctx.check(self.edb.lib().edgpu_shard_hxv_dw(self.s.h, ncols_valid, x_cols.shape[1], x_cols.data_ptr(), y_cols.data_ptr()))

    def up(self, row0, nrows, x_rows, y_rows):
        assert x_rows.is_contiguous() and y_rows.is_contiguous()
        self.s.ctx.check(self.edb.lib().edgpu_shard_hxv_up(self.s.h, row0, nrows, x_rows.data_ptr(), y_rows.data_ptr(), 0))

    def up_slabs(self, row0, nrows, col0, ldc, x_slabs, y_slabs):
        import ctypes as C
        n = len(col0)
        c0 = (C.c_int64 * n)(*col0)
        lc = (C.c_int64 * n)(*ldc)
        self.s.ctx.check(self.edb.lib().edgpu_shard_hxv_up_slabs(self.s.h, row0, nrows, n, c0, lc, x_slabs.data_ptr(), y_slabs.data_ptr(), 0))


def make_gpu_shard(edb, sector, rank, world, group=None):
    """ShardedHxv over the CUDA library for `sector` (star-product layout) on the current device."""
    import ctypes as C
    ld = C.c_int64()
    sector.ctx.check(edb.lib().edgpu_shard_ld(sector.h, C.byref(ld)))
    plan = ShardPlan(sector.dim_up, sector.dim_dw, ld.value, world)
    ops = GpuOps(edb, sector)
    ops.ncols_valid = plan.ncols[rank]
    return ShardedHxv(plan, rank, ops, group=group, device=torch.device("cuda", torch.cuda.current_device()))

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
    strips follow the reference's rule Q = N/P with the remainder on the last rank (ED_HAMILTONIAN.f90:56-62).

    nchunks > 1 pipelines the exchange: the rows are cut into `nchunks` groups and every group is dealt over the ranks
    by the same rule, so that the rows of one group are contiguous on the sender (one all_to_all_single per group,
    nothing packed) and the up-spin term of group j overlaps the transposes of groups j+1 (forth) and j-1 (back)."""
    dim_up: int
    dim_dw: int
    ld: int            # leading dimension of a full row (dim_up rounded up to 4)
    world: int
    nchunks: int = 1

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
        K = max(1, min(self.nchunks, max(1, self.dim_dw // max(1, self.world))))
        self.nchunks = K
        gq = self.dim_dw // K
        self.chunks = []                         # (first row of the group, rows in the group, row0 per rank, nrows per rank)
        for j in range(K):
            g0 = j * gq
            gl = gq if j < K - 1 else self.dim_dw - g0
            qr = gl // self.world
            r0 = [g0 + p * qr for p in range(self.world)]
            nr = [qr if p < self.world - 1 else gl - p * qr for p in range(self.world)]
            self.chunks.append((g0, gl, r0, nr))
        self.row0 = list(self.chunks[0][2])      # (meaningful as "the" row range of a rank only when nchunks == 1)
        self.nrows = [sum(c[3][p] for c in self.chunks) for p in range(self.world)]


class ShardedHxv:
    """y_loc = (H x)_loc on the column shard of this rank.

    ops.dw(x_cols, y_cols)                                y_cols = H_dw x_cols           on [dim_dw, ldc_me]
    ops.up(row0, nrows, x_rows, y_rows)                   y_rows = (D + H_up) x_rows     on [nrows, ld]
    ops.up_slabs(row0, nrows, col0, ldc, x_slabs, y_slabs)   the same on the slabs an all-to-all delivers (optional)
    """

    def __init__(self, plan: ShardPlan, rank: int, ops, group=None, device="cpu"):
        self.plan, self.rank, self.ops, self.group, self.device = plan, rank, ops, group, device
        P = plan
        me = rank
        self.ldc = P.ldc[me]
        self.nrows = P.nrows[me]
        f64 = torch.float64
        # exchange buffers (reused by every product): per row group, the slabs [nrows_j][ldc_p] of all source ranks p
        self.recv = [torch.zeros(c[3][me] * P.ld, dtype=f64, device=device) for c in P.chunks]
        self.send = [torch.zeros(c[3][me] * P.ld, dtype=f64, device=device) for c in P.chunks]
        self.tmp_cols = torch.zeros(P.dim_dw, self.ldc, dtype=f64, device=device)
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

    def _up_chunk(self, j):
        P, me = self.plan, self.rank
        _, _, r0, nr = P.chunks[j]
        if nr[me] == 0:
            return
        if hasattr(self.ops, "up_slabs"):
            # the up-spin term works directly on the received slabs and writes the slabs that go back: no pack/unpack
            self.ops.up_slabs(r0[me], nr[me], P.col0, P.ldc, self.recv[j], self.send[j])
            return
        x_rows = torch.zeros(nr[me], P.ld, dtype=torch.float64, device=self.device)
        y_rows = torch.zeros_like(x_rows)
        off = 0
        for p in range(P.world):
            n = nr[me] * P.ldc[p]
            x_rows[:, P.col0[p]:P.col0[p] + P.ldc[p]] = self.recv[j][off:off + n].view(nr[me], P.ldc[p])
            off += n
        self.ops.up(r0[me], nr[me], x_rows, y_rows)
        off = 0
        for p in range(P.world):
            n = nr[me] * P.ldc[p]
            self.send[j][off:off + n].view(nr[me], P.ldc[p]).copy_(y_rows[:, P.col0[p]:P.col0[p] + P.ldc[p]])
            off += n

    def apply(self, x_cols: torch.Tensor, y_cols: torch.Tensor):
        P, me = self.plan, self.rank
        # transpose #1 of every row group is posted first and runs on the collective stream WHILE the down-spin term
        # (local: down hops keep the column) runs on the compute stream
        splits, works1 = [], []
        for j, (g0, gl, r0, nr) in enumerate(P.chunks):
            in_split = [nr[p] * self.ldc for p in range(P.world)]            # my columns, rows of rank p in group j
            out_split = [nr[me] * P.ldc[p] for p in range(P.world)]          # my rows of group j, columns of rank p
            splits.append((in_split, out_split))
            works1.append(self._all_to_all(self.recv[j], x_cols[g0:g0 + gl].reshape(-1), out_split, in_split, async_op=True))
        self.ops.dw(x_cols, y_cols)
        # diagonal + up-spin term group by group; transpose #2 of group j overlaps the kernels of group j+1
        works2 = []
        for j, (g0, gl, r0, nr) in enumerate(P.chunks):
            if works1[j] is not None:
                works1[j].wait()
            self._up_chunk(j)
            in_split, out_split = splits[j]
            works2.append(self._all_to_all(self.tmp_cols[g0:g0 + gl].reshape(-1), self.send[j], in_split, out_split, async_op=True))
        for w in works2:
            if w is not None:
                w.wait()
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
        cur = v_cols.clone()
        return sharded_lanczos(lambda a, b: self.apply(a, b), self.dot, cur, torch.zeros_like(cur), torch.zeros_like(cur), nlanc)


def sharded_lanczos(apply, dot, cur, old, u, nlanc):
    """Lanczos tridiagonalisation of a sharded vector on unnormalised vectors (the recurrence of lanczos.cu):
        u = H w_k ; a_k = <w_k,u>/b_k^2 ; w_{k+1} = u/b_k - (b_k/b_{k-1}) w_{k-1} - (a_k/b_k) w_k ; b_{k+1} = |w_{k+1}|
    apply(x, y): y = H x on the shards; dot(a, b): all-reduced scalar (1-element tensor).  cur holds the start vector, old
    and u are scratch shards; every coefficient stays a device tensor (no host synchronisation inside the loop).
    Returns (alfa[nlanc], beta[nlanc]) with beta[0] = 0 like the reference's blanc."""
    one = torch.ones(1, dtype=torch.float64, device=cur.device)
    cur.div_(torch.sqrt(dot(cur, cur)))
    old.zero_()
    ncur, nold, bprev = one, one, torch.zeros_like(one)
    alfa, beta = [], [torch.zeros_like(one)]
    for _ in range(nlanc):
        apply(cur, u)
        a = dot(cur, u) / (ncur * ncur)
        old.mul_(-(bprev / nold))
        old.addcmul_(u, 1.0 / ncur)
        old.addcmul_(cur, -(a / ncur))
        b = torch.sqrt(dot(old, old))
        alfa.append(a)
        beta.append(b)
        cur, old = old, cur
        nold, ncur, bprev = ncur, b, b
    return torch.cat(alfa).cpu().numpy(), torch.cat(beta).cpu().numpy()[:nlanc]


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


def make_gpu_shard(edb, sector, rank, world, group=None, nchunks=1):
    """ShardedHxv over the CUDA library for `sector` (star-product layout) on the current device."""
    import ctypes as C
    ld = C.c_int64()
    sector.ctx.check(edb.lib().edgpu_shard_ld(sector.h, C.byref(ld)))
    plan = ShardPlan(sector.dim_up, sector.dim_dw, ld.value, world, nchunks)
    ops = GpuOps(edb, sector)
    ops.ncols_valid = plan.ncols[rank]
    return ShardedHxv(plan, rank, ops, group=group, device=torch.device("cuda", torch.cuda.current_device()))


class _DevBuf:
    """A cudaMalloc'ed buffer of the CUDA library seen as a torch tensor (__cuda_array_interface__), mappable by the other
    processes of the node through its CUDA IPC handle."""

    def __init__(self, edb, ctx, rows, cols):
        import ctypes as C
        self.edb, self.ctx, self.shape = edb, ctx, (rows, cols)
        p = C.c_void_p()
        ctx.check(edb.lib().edgpu_dev_alloc(ctx.h, max(16, rows * cols * 8), C.byref(p)))
        self.ptr = p.value
        self.__cuda_array_interface__ = {"shape": (rows, cols), "typestr": "<f8", "data": (self.ptr, False), "version": 3,
                                         "strides": None}
        self.t = torch.as_tensor(self, device=torch.device("cuda", torch.cuda.current_device()))

    def handle(self):
        import ctypes as C
        h = C.create_string_buffer(64)
        self.ctx.check(self.edb.lib().edgpu_ipc_export(self.ctx.h, self.ptr, h))
        return h.raw


class PeerShardedHxv:
    """Sharded H*v with the exchange FUSED into the up-pass kernel (one process per GPU, NVLink/NVSwitch peer memory).

    Every rank keeps its column shards (the Lanczos vectors `vec(i)` and one scratch shard) in buffers that all ranks of
    the node map through CUDA IPC.  Per product:
        barrier                                  (1-element NCCL all-reduce on the compute stream)
        y_i   = H_dw x_i                         local (tensor-map down kernel on the column shard), WHILE the copy engines
                                                 pull this rank's rows of the other shards over NVLink (DMA, no SMs)
        tmp_p[my rows] = (D + H_up) x[my rows]   copy-engine up kernel: bulk stores of the result straight into the owners'
                                                 scratch shards (prefetch=False: also bulk loads of x from the owners)
        barrier
        y_i  += tmp                              local
    No transpose buffers, no NCCL data traffic: NVLink carries (P-1)/P * Dim/P * 8 B in and out per rank and product,
    overlapped with the gathers tile by tile by the kernel's own pipeline.  Replaces the full-vector MPI_Allgatherv of
    directMatVec_MPI_cc (ED_HAMILTONIAN_DIRECT_HxV.f90:163-166)."""

    def __init__(self, edb, sector, rank, world, nvec=2, group=None, prefetch=True):
        import ctypes as C
        self.edb, self.s, self.rank, self.world, self.group = edb, sector, rank, world, group
        ld = C.c_int64()
        sector.ctx.check(edb.lib().edgpu_shard_ld(sector.h, C.byref(ld)))
        self.plan = P = ShardPlan(sector.dim_up, sector.dim_dw, ld.value, world)
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.ldc = P.ldc[rank]
        self.bufs = [_DevBuf(edb, sector.ctx, P.dim_dw, self.ldc) for _ in range(nvec + 1)]      # vectors + scratch (last)
        handles = [b.handle() for b in self.bufs]
        allh = [None] * world
        if world > 1:
            dist.all_gather_object(allh, handles, group=group)
        else:
            allh[0] = handles
        # ptrs[i][p] = address of buffer i of rank p in THIS process
        self.ptrs = []
        for i in range(nvec + 1):
            row = []
            for p in range(world):
                if p == rank:
                    row.append(self.bufs[i].ptr)
                else:
                    q = C.c_void_p()
                    sector.ctx.check(edb.lib().edgpu_ipc_open(sector.ctx.h, allh[p][i], C.byref(q)))
                    row.append(q.value)
            self.ptrs.append(row)
        self._flag = torch.zeros(1, device=self.device)
        self._col0 = (C.c_int64 * world)(*P.col0)
        self._ldc = (C.c_int64 * world)(*P.ldc)
        self._parr = [(C.c_void_p * world)(*row) for row in self.ptrs]
        self.bytes_nvlink = 0
        # DMA prefetch of the remote x rows (see apply): local copies [nrows_me][ldc_p] of the other ranks' shards
        self.prefetch = prefetch
        self.side = torch.cuda.Stream()
        self.ev = torch.cuda.Event()
        self.recv = [None if p == rank else torch.zeros(max(1, P.nrows[rank] * P.ldc[p]), dtype=torch.float64, device=self.device)
                     for p in range(world)]
        self._xloc = [(C.c_void_p * world)(*[self.ptrs[i][p] if p == rank else self.recv[p].data_ptr() for p in range(world)])
                      for i in range(nvec + 1)]
        self._xrow0 = (C.c_int64 * world)(*[0 if p == rank else P.row0[rank] for p in range(world)])

    def vec(self, i):
        return self.bufs[i].t

    def barrier(self):
        if self.world > 1:
            dist.all_reduce(self._flag, group=self.group)

    def apply(self, ix, iy):
        """vec(iy) = H vec(ix) on the column shards."""
        import ctypes as C
        P, me, L, s = self.plan, self.rank, self.edb.lib(), self.s
        x, y, tmp = self.bufs[ix].t, self.bufs[iy].t, self.bufs[-1].t
        cur = torch.cuda.current_stream()
        self.barrier()                       # every rank's x is final, every rank is done with its scratch shard
        if self.prefetch and self.world > 1:
            # copy-engine DMA of this rank's rows of the other shards WHILE the down pass runs: the up kernel then reads
            # local memory and NVLink carries its stores only
            self.side.wait_stream(cur)
            for p in range(self.world):
                if p != me:
                    s.ctx.check(L.edgpu_copy_async(s.ctx.h, self.recv[p].data_ptr(), self.ptrs[ix][p] + 8 * P.row0[me] * P.ldc[p],
                                                   8 * P.nrows[me] * P.ldc[p], C.c_void_p(self.side.cuda_stream)))
            self.ev.record(self.side)
            xarr, xrow0 = self._xloc[ix], self._xrow0
        else:
            xarr, xrow0 = self._parr[ix], None
        s.ctx.check(L.edgpu_shard_hxv_dw(s.h, P.ncols[me], self.ldc, x.data_ptr(), y.data_ptr()))
        if self.prefetch and self.world > 1:
            cur.wait_event(self.ev)
        s.ctx.check(L.edgpu_shard_hxv_up_peers(s.h, P.row0[me], P.nrows[me], self.world, self._col0, self._ldc,
                                               xarr, xrow0, self._parr[-1], 0))
        self.barrier()                       # all rows of my scratch shard have arrived
        y += tmp
        self.bytes_nvlink += 8 * P.nrows[me] * (P.ld - self.ldc)
        return y

    def dot(self, a, b):
        v = torch.dot(a.reshape(-1), b.reshape(-1)).reshape(1)
        if self.world > 1:
            dist.all_reduce(v, group=self.group)
        return v

    def lanczos_tridiag(self, nlanc):
        """Lanczos chain started from vec(0) (destroyed); needs nvec >= 3."""
        assert len(self.bufs) >= 4, "PeerShardedHxv(nvec=3) is needed for the Lanczos recurrence"
        idx = {self.bufs[i].t.data_ptr(): i for i in range(len(self.bufs) - 1)}
        return sharded_lanczos(lambda a, b: self.apply(idx[a.data_ptr()], idx[b.data_ptr()]), self.dot,
                               self.vec(0), self.vec(1), self.vec(2), nlanc)

"""GPU (B200): the star-product layout + tiled kernels (layout=2, hxv_kernel=2) against the oracle and against
the generic table kernel.  Same tolerances as test_gpu_parity.py."""
import numpy as np
import pytest

from test_gpu_parity import CASES as BASE_CASES, HXV_TOL, all_sectors, make

pytestmark = pytest.mark.gpu

STAR_CASES = ["cfg1", "nohf_mu", "2orb_hund", "nspin2", "3orb"]
STAR_FLAGS = [0, 16]
CASES = dict(BASE_CASES)
CASES["3orb"] = dict(Norb=3, Nbath=1, uloc=(2.0, 1.0, 3.0), ust=1.0, jh=0.2)


@pytest.mark.parametrize("flags", STAR_FLAGS)
@pytest.mark.parametrize("name", STAR_CASES)
def test_star_hxv_matches_oracle_all_sectors(oracle, edb, name, flags):
    p, model, ctx, rng = make(oracle, edb, CASES[name], layout=2, hxv_kernel=2, debug_flags=flags)
    Ns = p.Ns
    for nup, ndw in all_sectors(Ns):
        smap = oracle.build_sector(Ns, nup, ndw)
        s = ctx.sector(nup, ndw)
        assert np.array_equal(s.map(), smap)                         # the reference map is layout independent
        v = rng.normal(size=smap.size) + 1j * rng.normal(size=smap.size)
        ref = oracle.direct_hxv(model, smap, v)
        got = s.hxv_host(v)
        scale = max(1.0, np.abs(ref).max())
        assert np.abs(got - ref).max() < HXV_TOL * scale, (name, nup, ndw)
        x = s.vec(v.real)
        assert np.array_equal(x.download(), v.real)                  # import/export permutation round trip
        x.free(); s.free()
    ctx.close()


# 3 = fallback paths (2-column down strips + single-stage up pass), 4 = generic tile_pass everywhere,
# 8 = LSU kernels instead of the copy-engine ones, 16 = copy-engine kernels for every block size,
# +64 (+128) = at most 2 (3) stages in the up pipeline (1 x + 1 y image / 2 x + 1 y), +32 = programmatic dependent launch, +512 = 2 x images (compile-time stage counts)
@pytest.mark.parametrize("flags", [0, 3, 4, 8, 16, 16 + 64, 16 + 64 + 128, 16 + 32, 16 + 512])
@pytest.mark.parametrize("Norb,Nbath,sec", [(1, 9, (5, 5)), (1, 9, (6, 5)), (2, 4, (5, 5)), (2, 4, (4, 6)), (3, 2, (4, 5))])
def test_star_hxv_medium_sectors(oracle, edb, Norb, Nbath, sec, flags):
    case = dict(Norb=Norb, Nbath=Nbath, uloc=tuple([2.0] * Norb), ust=0.7 if Norb > 1 else 0.0, jh=0.1 if Norb > 1 else 0.0)
    p, model, ctx, rng = make(oracle, edb, case, layout=2, hxv_kernel=2, debug_flags=flags)
    smap = oracle.build_sector(p.Ns, *sec)
    s = ctx.sector(*sec)
    v = rng.normal(size=smap.size)
    ref = oracle.direct_hxv(model, smap, v).real
    x, y = s.vec(v), s.vec()
    s.hxv(x, y)
    assert np.abs(y.download() - ref).max() < HXV_TOL * np.abs(ref).max()
    # chain + observables + seeds work in the permuted layout
    a_ref, b_ref, _ = oracle.lanc_tridiag(model, smap, v / np.linalg.norm(v), 12)
    a, b, _ = s.lanczos_tridiag(x, 12)
    assert np.abs(a - a_ref).max() < 1e-9 and np.abs(b - b_ref).max() < 1e-9
    for t in (x, y):
        t.free()
    s.free()
    ctx.close()


def test_star_apply_c_and_observables(oracle, edb):
    p, model, ctx, rng = make(oracle, edb, CASES["2orb_hund"], layout=2, hxv_kernel=2)
    Ns = p.Ns
    for (nup, ndw), isite, dagger in [((3, 3), 2, 1), ((3, 2), 8, 0), ((2, 4), 1, 1), ((3, 3), 7, 0)]:
        mapI = oracle.build_sector(Ns, nup, ndw)
        g = rng.normal(size=mapI.size)
        g /= np.linalg.norm(g)
        d = 1 if dagger else -1
        jup, jdw = (nup + d, ndw) if isite <= Ns else (nup, ndw + d)
        mapJ = oracle.build_sector(Ns, jup, jdw)
        vv_ref, n2_ref = oracle.apply_op(Ns, isite, dagger, mapI, mapJ, g)
        si, sj = ctx.sector(nup, ndw), ctx.sector(jup, jdw)
        vin, vout = si.vec(g), sj.vec()
        n2 = edb.apply_c(si, sj, isite, dagger, vin, vout, normalise=False)
        assert abs(n2 - n2_ref) < 1e-13 and np.array_equal(vout.download(), vv_ref.real)
        obs = si.observables(vin)
        up = np.array([[(int(m) >> a) & 1 for a in range(p.Norb)] for m in mapI], dtype=float)
        dw = np.array([[(int(m) >> (a + Ns)) & 1 for a in range(p.Norb)] for m in mapI], dtype=float)
        assert np.abs(obs["dens"] - ((up + dw) * (g ** 2)[:, None]).sum(0)).max() < 1e-12
        assert np.abs(obs["docc"] - ((up * dw) * (g ** 2)[:, None]).sum(0)).max() < 1e-12
        vin.free(); vout.free(); si.free(); sj.free()
    ctx.close()


def test_star_equals_generic_at_cfg3_size(oracle, edb):
    """Ns=14 half filling (11.8M states): tiled star kernels vs the generic kernel on the same Philox vector,
    plus symmetry of H as a size-independent property."""
    case = dict(Norb=2, Nbath=6, uloc=(2.0, 2.0), ust=1.5, jh=0.25)
    out = {}
    for tag, layout, kern in (("generic", 1, 1), ("star", 2, 2)):
        p, model, ctx, rng = make(oracle, edb, case, layout=layout, hxv_kernel=kern)
        s = ctx.sector(7, 7)
        x, y, z, hz = s.vec().fill_normal(20240607), s.vec(), s.vec().fill_normal(99), s.vec()
        s.hxv(x, y)
        s.hxv(z, hz)
        out[tag] = y.download()
        assert abs(z.dot(y) - hz.dot(x)) < 1e-9 * abs(z.dot(y))     # <z|Hx> = <Hz|x>
        for t in (x, y, z, hz):
            t.free()
        s.free()
        ctx.close()
    assert np.abs(out["star"] - out["generic"]).max() < HXV_TOL * np.abs(out["generic"]).max()


@pytest.mark.parametrize("flags", [0, 16])           # 16: the copy-engine kernels (slab-addressed, accumulate = 0) for every block size
def test_shard_entry_points_equal_full_product(oracle, edb, flags):
    """edgpu_shard_hxv_dw on column shards + edgpu_shard_hxv_up on row shards (world=2 and 3, emulated on one GPU)
    reproduce the single-GPU product; the exchange logic itself is covered by tests/test_sharded_cpu.py (gloo)."""
    import ctypes as C
    import importlib
    import torch
    sharded = importlib.import_module("dmft-ed_b200.sharded")
    case = dict(Norb=2, Nbath=4, uloc=(2.0, 2.0), ust=0.7, jh=0.1)
    p, model, ctx, rng = make(oracle, edb, case, layout=2, hxv_kernel=2, debug_flags=flags)
    s = ctx.sector(5, 4)
    du, dd = s.dim_up, s.dim_dw
    r2iu = np.zeros(du, dtype=np.uint32)
    r2id = np.zeros(dd, dtype=np.uint32)
    ctx.check(edb.lib().edgpu_shard_perm(s.h, r2iu.ctypes.data, r2id.ctypes.data))
    ld = C.c_int64()
    edb.lib().edgpu_shard_ld(s.h, C.byref(ld))
    ld = ld.value
    v = rng.normal(size=s.dim)
    x, y = s.vec(v), s.vec()
    s.hxv(x, y)
    Yref = y.download().reshape(dd, du)
    Xint = np.zeros((dd, ld))
    Xint[np.ix_(r2id, r2iu)] = v.reshape(dd, du)
    Yint_ref = np.zeros((dd, ld))
    Yint_ref[np.ix_(r2id, r2iu)] = Yref
    dev = torch.device("cuda")
    for world in (2, 3):
        plan = sharded.ShardPlan(du, dd, ld, world)
        Y = np.zeros((dd, ld))
        for r in range(world):
            xc = torch.tensor(np.ascontiguousarray(Xint[:, plan.col0[r]:plan.col0[r] + plan.ldc[r]]), device=dev)
            yc = torch.zeros_like(xc)
            ctx.check(edb.lib().edgpu_shard_hxv_dw(s.h, plan.ncols[r], plan.ldc[r], xc.data_ptr(), yc.data_ptr()))
            xr = torch.tensor(np.ascontiguousarray(Xint[plan.row0[r]:plan.row0[r] + plan.nrows[r], :]), device=dev)
            yr = torch.zeros_like(xr)
            ctx.check(edb.lib().edgpu_shard_hxv_up(s.h, plan.row0[r], plan.nrows[r], xr.data_ptr(), yr.data_ptr(), 0))
            ctx.sync()
            Y[:, plan.col0[r]:plan.col0[r] + plan.ldc[r]] += yc.cpu().numpy()
            Y[plan.row0[r]:plan.row0[r] + plan.nrows[r], :] += yr.cpu().numpy()
            # the same row shard handed over as all-to-all slabs (what transpose #1 delivers)
            rows = slice(plan.row0[r], plan.row0[r] + plan.nrows[r])
            xs = torch.cat([torch.tensor(np.ascontiguousarray(Xint[rows, plan.col0[q]:plan.col0[q] + plan.ldc[q]]), device=dev).reshape(-1)
                            for q in range(world)])
            ys = torch.zeros_like(xs)
            c0 = (C.c_int64 * world)(*plan.col0)
            lc = (C.c_int64 * world)(*plan.ldc)
            ctx.check(edb.lib().edgpu_shard_hxv_up_slabs(s.h, plan.row0[r], plan.nrows[r], world, c0, lc, xs.data_ptr(), ys.data_ptr(), 0))
            ctx.sync()
            off = 0
            for q in range(world):
                n = plan.nrows[r] * plan.ldc[q]
                blk = ys[off:off + n].view(plan.nrows[r], plan.ldc[q]).cpu().numpy()
                ref_blk = yr.cpu().numpy()[:, plan.col0[q]:plan.col0[q] + plan.ldc[q]]
                assert np.abs(blk - ref_blk).max() <= 1e-13 * max(1.0, np.abs(ref_blk).max())   # other kernels, other summation order
                off += n
        assert np.abs(Y - Yint_ref).max() < HXV_TOL * np.abs(Yint_ref).max()
        # peer mode: the up-spin term reads x from, and writes into, the column shards of all "ranks" directly
        xs = [torch.tensor(np.ascontiguousarray(Xint[:, plan.col0[q]:plan.col0[q] + plan.ldc[q]]), device=dev) for q in range(world)]
        ts = [torch.full_like(t, 7.0) for t in xs]                  # accumulate = 0 must overwrite every element
        ys = [torch.zeros_like(t) for t in xs]
        xp = (C.c_void_p * world)(*[t.data_ptr() for t in xs])
        tp = (C.c_void_p * world)(*[t.data_ptr() for t in ts])
        c0 = (C.c_int64 * world)(*plan.col0)
        lc = (C.c_int64 * world)(*plan.ldc)
        for r in range(world):
            ctx.check(edb.lib().edgpu_shard_hxv_dw(s.h, plan.ncols[r], plan.ldc[r], xs[r].data_ptr(), ys[r].data_ptr()))
            ctx.check(edb.lib().edgpu_shard_hxv_up_peers(s.h, plan.row0[r], plan.nrows[r], world, c0, lc, xp, None, tp, 0))
        ctx.sync()
        Yp = np.concatenate([(ys[q] + ts[q]).cpu().numpy() for q in range(world)], axis=1)
        assert np.abs(Yp - Yint_ref).max() < HXV_TOL * np.abs(Yint_ref).max()
        # the same with local copies of just this rank's rows of the other shards (what the DMA prefetch delivers)
        t2 = [torch.full_like(t, -3.0) for t in xs]
        tp2 = (C.c_void_p * world)(*[t.data_ptr() for t in t2])
        for r in range(world):
            rows = slice(plan.row0[r], plan.row0[r] + plan.nrows[r])
            loc = [xs[q] if q == r else xs[q][rows].contiguous() for q in range(world)]
            xl = (C.c_void_p * world)(*[t.data_ptr() for t in loc])
            xr0 = (C.c_int64 * world)(*[0 if q == r else plan.row0[r] for q in range(world)])
            ctx.check(edb.lib().edgpu_shard_hxv_up_peers(s.h, plan.row0[r], plan.nrows[r], world, c0, lc, xl, xr0, tp2, 0))
            ctx.sync()
        for q in range(world):
            assert torch.equal(t2[q], ts[q])
    x.free(); y.free(); s.free(); ctx.close()

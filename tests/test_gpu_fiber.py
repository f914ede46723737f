"""GPU (B200): the pair-tile layout + fiber kernels (hxv_kernel=3, hxv_fiber.cu) against the oracle.

debug flags: 16 = fiber kernels for every block they can take (also the tiny ones), 4 = thread-per-element pair kernels
only, 0 = the production split (fiber kernels for blocks >= 256 configurations), +4096 = small pipeline slots so that the
largest HALF images take BOTH slots (the path of the 8000-configuration blocks of Ns=18; the images below run as two-slot
full tiles like the 4900-configuration blocks of Ns=16), +32768 = every image above one slot as half tiles (2 rows of a band /
2 columns of a strip through tensor maps), 1024 / 2048 = one pass by the thread-per-element kernels."""
import numpy as np
import pytest

from test_gpu_parity import HXV_TOL, all_sectors, make

pytestmark = pytest.mark.gpu

CASES = {
    "2orb_nb2": dict(Norb=2, Nbath=2, uloc=(2.0, 1.5), ust=1.2, jh=0.3),
    "2orb_nb3": dict(Norb=2, Nbath=3, uloc=(2.0, 1.0), ust=0.9, jh=0.2),
    "2orb_nb4_nspin2": dict(Norb=2, Nbath=4, Nspin=2, uloc=(2.0, 2.5), ust=0.7, jh=0.1),
    "3orb_nb2": dict(Norb=3, Nbath=2, uloc=(2.0, 1.0, 3.0), ust=1.0, jh=0.2),
}


@pytest.mark.parametrize("flags", [16, 4, 0, 16 + 4096, 16 + 1024, 16 + 2048, 16 + 4096 + 32768])
@pytest.mark.parametrize("name", ["2orb_nb2", "2orb_nb3", "3orb_nb2"])
def test_fiber_hxv_matches_oracle_all_sectors(oracle, edb, name, flags):
    p, model, ctx, rng = make(oracle, edb, CASES[name], hxv_kernel=3, debug_flags=flags)
    Ns = p.Ns
    for nup, ndw in all_sectors(Ns):
        smap = oracle.build_sector(Ns, nup, ndw)
        s = ctx.sector(nup, ndw)
        assert s.info()["layout_kind"] == 3
        assert np.array_equal(s.map(), smap)
        v = rng.normal(size=smap.size) + 1j * rng.normal(size=smap.size)
        ref = oracle.direct_hxv(model, smap, v)
        got = s.hxv_host(v)
        scale = max(1.0, np.abs(ref).max())
        assert np.abs(got - ref).max() < HXV_TOL * scale, (name, nup, ndw, np.abs(got - ref).max())
        x = s.vec(v.real)
        assert np.array_equal(x.download(), v.real)                  # import/export round trip through the pair tiles
        x.free(); s.free()
    ctx.close()


@pytest.mark.parametrize("flags", [16, 0, 16 + 4096])
@pytest.mark.parametrize("name,sec", [("2orb_nb4_nspin2", (5, 5)), ("2orb_nb4_nspin2", (4, 6)), ("3orb_nb2", (4, 5)), ("2orb_nb3", (4, 4))])
def test_fiber_medium_sectors_chain_and_seeds(oracle, edb, name, sec, flags):
    p, model, ctx, rng = make(oracle, edb, CASES[name], hxv_kernel=3, debug_flags=flags)
    smap = oracle.build_sector(p.Ns, *sec)
    s = ctx.sector(*sec)
    v = rng.normal(size=smap.size)
    ref = oracle.direct_hxv(model, smap, v).real
    x, y = s.vec(v), s.vec()
    s.hxv(x, y)
    assert np.abs(y.download() - ref).max() < HXV_TOL * np.abs(ref).max()
    # pads stay zero: a second product into the same buffer gives the same answer
    s.hxv(x, y)
    assert np.abs(y.download() - ref).max() < HXV_TOL * np.abs(ref).max()
    assert abs(x.dot(y) - float(v @ ref)) < 1e-10 * abs(float(v @ ref))
    a_ref, b_ref, _ = oracle.lanc_tridiag(model, smap, v / np.linalg.norm(v), 12)
    a, b, _ = s.lanczos_tridiag(x, 12)                               # fused <x,Hx> of the down pass
    assert np.abs(a - a_ref).max() < 1e-9 and np.abs(b - b_ref).max() < 1e-9
    for t in (x, y):
        t.free()
    s.free()
    ctx.close()


def test_fiber_apply_c_and_observables(oracle, edb):
    p, model, ctx, rng = make(oracle, edb, CASES["2orb_nb3"], hxv_kernel=3, debug_flags=16)
    Ns = p.Ns
    for (nup, ndw), isite, dagger in [((4, 4), 2, 1), ((4, 3), Ns + 2, 0), ((3, 5), 1, 1), ((4, 4), Ns + 1, 0)]:
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


@pytest.mark.parametrize("nranks", [2, 3])
def test_pair_shards_sum_to_full_product(oracle, edb, nranks):
    """Sharding by conserved occupation pairs: every rank multiplies only its own pairs, no exchange; the shards'
    results (downloaded with zeros for foreign elements) add up to the single-GPU product."""
    p, model, ctx, rng = make(oracle, edb, CASES["2orb_nb3"], hxv_kernel=3, debug_flags=16)
    sec = (4, 4)
    smap = oracle.build_sector(p.Ns, *sec)
    v = rng.normal(size=smap.size)
    ref = oracle.direct_hxv(model, smap, v).real
    total, cover, nall = np.zeros_like(ref), np.zeros_like(ref), 0
    for r in range(nranks):
        s = ctx.sector_shard(*sec, r, nranks)
        info = s.info()
        assert info["shard_rank"] == r and info["shard_nranks"] == nranks
        nall += info["nalloc"]
        x, y, one = s.vec(v), s.vec(), s.vec(np.ones_like(v))
        s.hxv(x, y)
        total += y.download()
        cover += one.download()
        for t in (x, y, one):
            t.free()
        s.free()
    assert np.array_equal(cover, np.ones_like(ref))                  # every element owned by exactly one rank
    assert np.abs(total - ref).max() < HXV_TOL * np.abs(ref).max()
    ctx.close()

"""GPU (B200): the HEADLINE shapes against the oracle (VERDICT r01 'What's weak' #1).

cfg3 (Norb=2, Nbath=6, sector (7,7), 11.8M states) and cfg4 (Norb=2, Nbath=7, sector (8,8), 165.6M states) select kernel
instantiations no small case reaches (70-configuration fibers, two-slot 157 KB images, 4900-row strips).  One H*v of the
Philox start vector is compared with the window oracle on whole reference rows that touch every (down-block, up-block)
tile (oracle/parity_check.py); tolerance 1e-12 * |y|_inf."""
import numpy as np
import pytest

from oracle import parity_check as PC

pytestmark = pytest.mark.gpu

SHAPES = {
    "cfg3": (2, 6, 7, 7),
    "cfg4": (2, 7, 8, 8),
}


def _setup(oracle, edb, name, hxv_kernel, flags=0):
    Norb, Nbath, nup, ndw = SHAPES[name]
    p = oracle.Params(Norb=Norb, Nbath=Nbath, uloc=tuple([2.0] * Norb), ust=0.8, jh=0.15, lanc_method="lanczos", lanc_nstates_sector=1)
    rng = np.random.default_rng(3)
    bath = oracle.init_bath(p) + 0.05 * rng.normal(size=oracle.init_bath(p).size)
    model = oracle.Model(p, bath)
    ctx = edb.Context(Norb, Nbath, 1, p.hfmode, layout=0, hxv_kernel=hxv_kernel, debug_flags=flags)
    ctx.set_hamiltonian(bath, p.uloc, p.ust, p.jh, p.jx, p.jp, p.xmu)
    return p, model, ctx, nup, ndw


@pytest.mark.parametrize("kernel", [3, 2])               # 3: fiber kernels on pair tiles (production), 2: round-1 star kernels
@pytest.mark.parametrize("name", ["cfg3", "cfg4"])
def test_headline_hxv_rows_match_oracle(oracle, edb, name, kernel):
    p, model, ctx, nup, ndw = _setup(oracle, edb, name, kernel)
    rows, dim_up, dim_dw = PC.pick_rows(oracle, p.Norb, p.Nbath, nup, ndw, per_block=4, max_rows=40)
    s = ctx.sector(nup, ndw)
    assert s.info()["layout_kind"] == (3 if kernel == 3 else 0)
    x, y = s.vec().fill_uniform(20240607), s.vec()
    s.hxv(x, y)
    res = PC.check_rows(oracle, model, nup, ndw, 20240607, rows, dim_up, y.download_rows)
    assert res["max_rel_err"] < 1e-12, res
    # the map of the full-size sector: device order / popcount check and two sampled windows against the closed form
    cs, viol = s.map_check()
    assert viol == 0
    for first in (0, s.dim - 1000):
        got = s.map(first, 1000)
        ref = np.array([oracle.map_entry(p.Ns, nup, ndw, first + i) for i in range(1000)], dtype=np.uint64)
        assert np.array_equal(got, ref)
    x.free(); y.free(); s.free(); ctx.close()


def test_headline_lanczos_dot_fusion_cfg3(oracle, edb):
    """alpha from the fused <x,Hx> of the fiber down pass equals the plain dot of the same product (cfg3)."""
    p, model, ctx, nup, ndw = _setup(oracle, edb, "cfg3", 3)
    s = ctx.sector(nup, ndw)
    x, y = s.vec().fill_uniform(7), s.vec()
    nrm = np.sqrt(x.dot(x))
    s.hxv(x, y)
    a_plain = x.dot(y) / nrm ** 2
    a, b, _ = s.lanczos_tridiag(x, 3)
    assert abs(a[0] - a_plain) < 1e-12 * max(1.0, abs(a_plain))
    x.free(); y.free(); s.free(); ctx.close()

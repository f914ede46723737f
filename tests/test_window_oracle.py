"""CPU: the window oracle (ora_window_hxv: closed-form map / rank, on-the-fly Philox input) equals the literal
gather form of directMatVec_cc bit for bit, and the parity row picker covers every star-occupation block."""
import numpy as np
import pytest

from oracle import parity_check as PC


@pytest.mark.parametrize("case,sec", [
    (dict(Norb=2, Nbath=3, uloc=(2.0, 1.5), ust=0.7, jh=0.1), (4, 4)),
    (dict(Norb=2, Nbath=3, uloc=(2.0, 1.5), ust=0.7, jh=0.1), (3, 5)),
    (dict(Norb=3, Nbath=1, uloc=(2.0, 1.0, 3.0), ust=1.0, jh=0.2, jx=0.2, jp=0.2), (3, 3)),
    (dict(Norb=1, Nbath=5, uloc=(2.0,)), (2, 4)),
])
def test_window_oracle_is_the_literal_gather(oracle, case, sec):
    p = oracle.Params(lanc_method="lanczos", lanc_nstates_sector=1, **case)
    rng = np.random.default_rng(1)
    bath = oracle.init_bath(p) + 0.05 * rng.normal(size=oracle.init_bath(p).size)
    model = oracle.Model(p, bath)
    nup, ndw = sec
    smap = oracle.build_sector(p.Ns, nup, ndw, literal=True)
    v = oracle.philox_uniform(77, smap.size)
    ref = oracle.gather_hxv(model, smap, v)
    got = oracle.window_hxv(model, nup, ndw, 77, np.arange(smap.size))
    assert np.array_equal(got, ref.real) and not ref.imag.any()
    assert all(oracle.map_entry(p.Ns, nup, ndw, i) == int(smap[i]) for i in range(smap.size))


def test_pick_rows_hits_every_down_block(oracle):
    Norb, Nbath, nup, ndw = 2, 4, 5, 5
    rows, dim_up, dim_dw = PC.pick_rows(oracle, Norb, Nbath, nup, ndw, per_block=3, max_rows=1000)
    Ns = Norb * (Nbath + 1)
    seen = {PC.star_tuple(oracle.map_entry(Ns, nup, ndw, rd * dim_up) >> Ns, Norb, Nbath) for rd in rows}
    every = {PC.star_tuple(oracle.map_entry(Ns, nup, ndw, rd * dim_up) >> Ns, Norb, Nbath) for rd in range(dim_dw)}
    assert seen == every and 0 in rows and dim_dw - 1 in rows
    # check_rows with the oracle itself as the "device": zero error
    p = oracle.Params(Norb=Norb, Nbath=Nbath, uloc=(2.0, 2.0), lanc_method="lanczos", lanc_nstates_sector=1)
    model = oracle.Model(p, oracle.init_bath(p))
    fetch = lambda a, b: oracle.window_hxv(model, nup, ndw, 5, np.arange(a * dim_up, b * dim_up))
    res = PC.check_rows(oracle, model, nup, ndw, 5, rows[:3], dim_up, fetch)
    assert res["max_abs_err"] == 0.0 and res["elements"] == 3 * dim_up

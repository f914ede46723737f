"""Sampled-row parity of one H*v at sizes where the literal oracle does not fit the host (cfg3/cfg4/cfg5).

TEST INFRASTRUCTURE (oracle side): used by tests/ and by bench.py's parity / cpu_baseline leg only.

The GPU fills x with the Philox uniforms (counter = reference index, bit-identical on both sides), computes y = H x and
hands back whole reference rows (fixed down configuration rd, all DimUp up configurations); the oracle evaluates the
same rows with ora_window_hxv (direct/HxV*.f90 row rule, map and binary search in closed form, no O(Dim) storage).
Rows are picked so that EVERY star-occupation block of the down spin is hit at its first and last configuration plus a
few interior ones; a row spans all up configurations, i.e. every up block including its first/last column, so every
(down-block, up-block) tile of the device layouts is touched.
"""
import numpy as np


def star_tuple(word, Norb, Nbath):
    """occupation of each star (impurity a + its bath levels, getBathStride ED_SETUP.f90:450-454) of an Ns-bit word"""
    t = []
    for a in range(Norb):
        n = (word >> a) & 1
        for k in range(Nbath):
            n += (word >> (Norb + a * Nbath + k)) & 1
        t.append(n)
    return tuple(t)


def pick_rows(O, Norb, Nbath, nup, ndw, per_block=3, seed=5, max_rows=64):
    """reference down ranks rd covering every down block (first, last in reference order, + random interior ones)"""
    Ns = Norb * (Nbath + 1)
    dim_up = int(O.lib().ora_binomial(Ns, nup))
    dim_dw = int(O.lib().ora_binomial(Ns, ndw))
    blocks = {}
    for rd in range(dim_dw):
        w = O.map_entry(Ns, nup, ndw, rd * dim_up) >> Ns
        blocks.setdefault(star_tuple(w, Norb, Nbath), []).append(rd)
    rng = np.random.default_rng(seed)
    rows = []
    for _, lst in sorted(blocks.items()):
        pick = {lst[0], lst[-1]}
        if len(lst) > 2:
            pick.update(int(v) for v in rng.choice(lst, size=min(per_block - 2, len(lst) - 2), replace=False))
        rows.extend(sorted(pick))
    if len(rows) > max_rows:                     # keep first/last of the largest blocks, thin out the rest
        rows = sorted(set(rows[:: max(1, len(rows) // max_rows)]) | {rows[0], rows[-1]})
    return sorted(set(rows)), dim_up, dim_dw


def check_rows(O, model, nup, ndw, seed, rows, dim_up, fetch_rows):
    """fetch_rows(rd0, rd1) -> GPU y for reference rows [rd0, rd1) as a flat array.  Returns a dict with the error."""
    worst, ymax, n = 0.0, 0.0, 0
    for rd in rows:
        got = np.asarray(fetch_rows(rd, rd + 1))
        idx = np.arange(rd * dim_up, (rd + 1) * dim_up, dtype=np.int64)
        ref = O.window_hxv(model, nup, ndw, seed, idx)
        worst = max(worst, float(np.abs(got - ref).max()))
        ymax = max(ymax, float(np.abs(ref).max()))
        n += idx.size
    return {"rows": len(rows), "elements": n, "max_abs_err": worst, "ymax": ymax,
            "max_rel_err": worst / ymax if ymax > 0 else 0.0}

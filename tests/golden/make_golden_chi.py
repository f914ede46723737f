#!/usr/bin/env python
"""Generates tests/golden/oracle_golden_chi.npz: spin and charge susceptibilities (build_chi_spin, ED_GF_CHISPIN.f90; build_chi_dens, ED_GF_CHIDENS.f90) of BASELINE config 1
and of a two-orbital model from the CPU oracle.  Oracle-generated like oracle_golden.npz ("parity unpinned").

    python tests/golden/make_golden_chi.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ed_oracle as O  # noqa: E402

CFG1 = dict(Norb=1, Nbath=4, lanc_method="lanczos", lanc_nstates_sector=1, Lmats=64, Lreal=64, beta=100.0, chispin_flag=True, chidens_flag=True, Ltau=200)
TWO = dict(Norb=2, Nbath=2, uloc=(2.0, 2.0), ust=1.5, jh=0.25, lanc_method="lanczos", lanc_nstates_sector=1, Lmats=32, Lreal=32,
           beta=50.0, lanc_dim_threshold=64, chispin_flag=True, chidens_flag=True, Ltau=64)


def main():
    O.build()
    out = {}
    for name, kw in (("cfg1", CFG1), ("two", TWO)):
        p = O.Params(**kw)
        r = O.ed_solve(p, O.init_bath(p))
        out[name + "_chi_tau"] = r.spinChi_tau
        out[name + "_chi_iv"] = r.spinChi_iv
        out[name + "_chi_w"] = r.spinChi_w
        out[name + "_sz2"] = r.sz2
        out[name + "_dchi_tau"] = r.densChi_tau
        out[name + "_dchi_iv"] = r.densChi_iv
        out[name + "_dchi_tot_tau"] = r.densChi_tot_tau
    np.savez_compressed(os.path.join(HERE, "oracle_golden_chi.npz"), **out)
    print("wrote oracle_golden_chi.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()

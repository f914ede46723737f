#!/usr/bin/env python
"""Generates tests/golden/*.npz from the CPU oracle (oracle/ed_oracle.{c,py}).

The reference (Fortran + SciFortran) cannot be built or run in this image and ships no golden data, so these
fixtures are ORACLE-GENERATED ("parity unpinned"): they freeze the literal restatement of the
reference rules so that (a) the oracle cannot drift silently and (b) the GPU tests have a fixed target that does
not depend on re-running the CPU code.  Anchors that were computed independently during the survey
 are stored alongside and checked by tests/test_golden.py.

    python tests/golden/make_golden.py        # rewrites the .npz files (takes ~2 minutes)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ed_oracle as O  # noqa: E402


def case_model(**kw):
    base = dict(lanc_method="lanczos", lanc_nstates_sector=1)
    base.update(kw)
    p = O.Params(**base)
    return p, O.init_bath(p)


def main():
    O.build()
    out = {}
    # ---- sector maps, Ns=5: every (nup,ndw) sector concatenated (literal Theta(4^Ns) scan of ED_SETUP.f90:899-916)
    maps, offs = [], [0]
    for nup in range(6):
        for ndw in range(6):
            m = O.build_sector(5, nup, ndw, literal=True)
            maps.append(m)
            offs.append(offs[-1] + m.size)
    out["maps_ns5"] = np.concatenate(maps)
    out["maps_ns5_offsets"] = np.array(offs, dtype=np.int64)
    # ---- H*v vectors
    rng = np.random.default_rng(20240607)
    p, bath = case_model(Norb=1, Nbath=4)
    model = O.Model(p, bath)
    smap = O.build_sector(5, 2, 3)
    v = rng.normal(size=smap.size) + 1j * rng.normal(size=smap.size)
    out["hxv_cfg1_23_in"] = v
    out["hxv_cfg1_23_out"] = O.direct_hxv(model, smap, v)
    p2, bath2 = case_model(Norb=2, Nbath=2, uloc=(2.0, 1.5), ust=1.2, jh=0.3, jx=0.2, jp=0.1)
    bath2 = bath2 + 0.05 * rng.normal(size=bath2.size)
    model2 = O.Model(p2, bath2)
    smap2 = O.build_sector(6, 3, 3)
    v2 = rng.normal(size=smap2.size) + 1j * rng.normal(size=smap2.size)
    out["hxv_2orb_bath"] = bath2
    out["hxv_2orb_33_in"] = v2
    out["hxv_2orb_33_out"] = O.direct_hxv(model2, smap2, v2)
    p3, bath3 = case_model(Norb=1, Nbath=9)
    model3 = O.Model(p3, bath3)
    smap3 = O.build_sector(10, 5, 5)
    v3 = O.philox_normal(20240607, smap3.size)
    hv3 = O.direct_hxv(model3, smap3, v3).real
    out["hxv_cfg2_55_probe_idx"] = np.arange(0, smap3.size, 997)
    out["hxv_cfg2_55_probe_out"] = hv3[::997]
    out["hxv_cfg2_55_norm"] = np.array([np.linalg.norm(hv3), v3 @ hv3])
    # ---- ed_solve, BASELINE config 1 (all 36 sectors)
    pc1 = O.Params(Norb=1, Nbath=4, lanc_method="lanczos", lanc_nstates_sector=1, Lmats=64, Lreal=64, beta=100.0)
    r1 = O.ed_solve(pc1, O.init_bath(pc1))
    out["cfg1_egs"] = np.array([r1.egs])
    out["cfg1_state_energies"] = np.array(sorted(s.e for s in r1.states))
    out["cfg1_dens_docc"] = np.array([r1.dens[0], r1.docc[0]])
    out["cfg1_gmats"] = r1.impGmats[0, 0, 0, 0]
    out["cfg1_smats"] = r1.impSmats[0, 0, 0, 0]
    # ---- ed_solve, BASELINE config 2, half-filling window
    secs = [(5, 5), (4, 5), (5, 4), (6, 5), (5, 6), (4, 4), (6, 6)]
    pc2 = O.Params(Norb=1, Nbath=9, lanc_method="lanczos", lanc_nstates_sector=1, Lmats=64, Lreal=64, beta=100.0, lanc_ngfiter=60)
    r2 = O.ed_solve(pc2, O.init_bath(pc2), sectors=secs)
    out["cfg2_egs"] = np.array([r2.egs])
    out["cfg2_dens_docc"] = np.array([r2.dens[0], r2.docc[0]])
    out["cfg2_gmats"] = r2.impGmats[0, 0, 0, 0]
    out["cfg2_smats"] = r2.impSmats[0, 0, 0, 0]
    out["cfg2_chain0_alfa"] = r2.chains[0]["alfa"][:8]
    out["cfg2_chain0_beta"] = r2.chains[0]["beta"][:8]
    # ---- survey anchors, computed independently of this oracle during the survey
    out["anchor_cfg1_e0"] = np.array([-5.671950916933])
    out["anchor_cfg1_e0_u0"] = np.array([-5.595866798531])
    out["anchor_cfg2_e012"] = np.array([-11.341244826804, -11.001795935561, -10.870726063586])
    out["anchor_cfg2_dens_docc"] = np.array([1.0, 0.162122572520])
    np.savez_compressed(os.path.join(HERE, "oracle_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "oracle_golden.npz"), {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()

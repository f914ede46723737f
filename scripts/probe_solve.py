"""Wall time of a full ed_solve with the per-sector / per-chain trace (ED_B200_TRACE=1).  usage: probe_solve.py cfg2|cfg3"""
import importlib
import os
import sys
import time

os.environ["ED_B200_TRACE"] = "1"
sys.path.insert(0, ".")


def main():
    wl = sys.argv[1]
    edb = importlib.import_module("dmft-ed_b200")
    kw = {"cfg1": dict(Norb=1, Nbath=4), "cfg2": dict(Norb=1, Nbath=9), "cfg3": dict(Norb=2, Nbath=6, uloc=[2.0, 2.0])}[wl]
    si = edb.default_input(lanc_method="lanczos", lanc_nstates_sector=1, ed_sparse_H=0, Lmats=1024, Lreal=1024, **kw)
    so = edb.Solver(si, device=0)
    so.solve()
    print("=== second solve", file=sys.stderr, flush=True)
    t0 = time.perf_counter()
    so.solve()
    dt = time.perf_counter() - t0
    print(f"wall {dt:.3f} s phases {so.timings()}")


main()

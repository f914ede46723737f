"""Host-side model of the pair sharding of round 2 (test driver; the product path is C++: hxv_fiber.cu + comm.cu + lanczos.cu).

H is block diagonal over (down-block, up-block) pairs of conserved star occupations.  `lpt_owner` restates the dealing rule of
pair_layout_build (longest-processing-time-first over pair sizes, ties by pair index) and `sharded_lanczos` the recurrence
of lanczos_step on unnormalised vectors with the two scalar reductions per step that are the ONLY collectives left
(.repo/PLAIN_LANCZOS.f90:87-118 evaluated as in dmft-ed_b200/csrc/lanczos.cu)."""
import numpy as np


def lpt_owner(sizes, nranks):
    """owner[p] of every pair p; stable sort by decreasing size, each pair to the least loaded rank (lowest rank on ties)"""
    order = sorted(range(len(sizes)), key=lambda p: -sizes[p])          # Python's sort is stable like std::stable_sort
    load = [0] * nranks
    owner = [0] * len(sizes)
    for p in order:
        best = min(range(nranks), key=lambda r: (load[r], r))
        owner[p] = best
        load[best] += sizes[p]
    return owner, load


def sharded_lanczos(apply_local, x_local, nlanc, allreduce_sum):
    """alpha[nlanc], beta[nlanc] (beta[0] unused) of the chain started at the GLOBAL vector whose local part is x_local.
    apply_local(v) = H restricted to the pairs of this rank; allreduce_sum(float) sums a scalar over the ranks."""
    a, b = np.zeros(nlanc), np.zeros(nlanc + 1)
    nrm = np.sqrt(allreduce_sum(float(x_local @ x_local)))
    cur = x_local / nrm
    old = np.zeros_like(cur)
    ncur, nold, bprev = 1.0, 1.0, 0.0
    for it in range(1, nlanc + 1):
        u = apply_local(cur)
        a[it - 1] = allreduce_sum(float(cur @ u)) / ncur ** 2
        new = u / ncur - (bprev / nold) * old - (a[it - 1] / ncur) * cur
        b[it] = np.sqrt(allreduce_sum(float(new @ new)))
        old, cur = cur, new
        nold, ncur, bprev = ncur, b[it], b[it]
    beta = np.zeros(nlanc)
    beta[1:] = b[1:nlanc]
    return a, beta

"""CPU, world_size=2 (gloo): the pair sharding of round 2.  A block-diagonal stand-in operator (one dense symmetric block
per (down-block, up-block) pair) is dealt to the ranks with the LPT rule of pair_layout_build; every rank multiplies only
its own blocks and the chain needs nothing but two scalar all-reduces per step.  alpha/beta must equal the single-rank chain."""
import importlib
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def make_problem():
    rng = np.random.default_rng(7)
    sizes = [40, 33, 33, 21, 12, 12, 5, 1, 1]
    blocks = []
    for n in sizes:
        m = rng.normal(size=(n, n))
        blocks.append(m + m.T)
    x = rng.normal(size=sum(sizes))
    return sizes, blocks, x


def worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    ps = importlib.import_module("dmft-ed_b200.pairshard")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sizes, blocks, x = make_problem()
    owner, load = ps.lpt_owner([n * n for n in sizes], world)
    offs = np.concatenate([[0], np.cumsum(sizes)])
    mine = [p for p in range(len(sizes)) if owner[p] == rank]
    xl = np.concatenate([x[offs[p]:offs[p + 1]] for p in mine])
    loffs = np.concatenate([[0], np.cumsum([sizes[p] for p in mine])])

    def apply_local(v):
        return np.concatenate([blocks[p] @ v[loffs[i]:loffs[i + 1]] for i, p in enumerate(mine)])

    def allreduce(v):
        t = torch.tensor([v], dtype=torch.float64)
        dist.all_reduce(t)
        return float(t.item())

    a, b = ps.sharded_lanczos(apply_local, xl, 12, allreduce)
    if rank == 0:
        q.put((a, b, load))
    dist.destroy_process_group()


def test_pair_sharded_chain_equals_single_rank():
    ps = importlib.import_module("dmft-ed_b200.pairshard")
    sizes, blocks, x = make_problem()
    offs = np.concatenate([[0], np.cumsum(sizes)])

    def apply_full(v):
        return np.concatenate([blocks[p] @ v[offs[p]:offs[p + 1]] for p in range(len(sizes))])

    a1, b1 = ps.sharded_lanczos(apply_full, x, 12, lambda v: v)
    # textbook form of the same recurrence (.repo/PLAIN_LANCZOS.f90:87-118) on normalised vectors
    vin, vout, bb = x / np.linalg.norm(x), np.zeros_like(x), 0.0
    at, bt = [], [0.0]
    for _ in range(12):
        tmp = apply_full(vin) - bb * vout
        aa = vin @ tmp
        tmp = tmp - aa * vin
        bb = np.linalg.norm(tmp)
        vout, vin = vin, tmp / bb
        at.append(aa); bt.append(bb)
    assert np.abs(a1 - np.array(at)).max() < 1e-9 and np.abs(b1[1:] - np.array(bt[1:12])).max() < 1e-9
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    a2, b2, load = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.abs(a2 - a1).max() < 1e-10 and np.abs(b2 - b1).max() < 1e-10
    assert max(load) / (sum(load) / 2) < 1.2                                  # LPT keeps the two ranks within 20 %


def test_lpt_rule_is_deterministic_and_balanced_for_cfg4_and_cfg5():
    """pair sizes of the Ns=16 and Ns=18 half-filling sectors: expected load balance of the dealing rule"""
    import itertools
    import math
    ps = importlib.import_module("dmft-ed_b200.pairshard")
    for norb, nbath, n, nranks, bound in [(2, 7, 8, 8, 1.17), (2, 7, 8, 2, 1.01), (3, 5, 9, 8, 1.01)]:
        nl = nbath + 1
        blocks = [math.prod(math.comb(nl, m) for m in t) for t in itertools.product(range(nl + 1), repeat=norb) if sum(t) == n]
        sizes = [a * b for a in blocks for b in blocks]
        owner, load = ps.lpt_owner(sizes, nranks)
        assert owner == ps.lpt_owner(sizes, nranks)[0]
        assert max(load) / (sum(load) / nranks) < bound, (norb, nbath, nranks, max(load) / (sum(load) / nranks))

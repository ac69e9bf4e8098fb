"""world_size-2 gloo test of the multi-GPU host logic on CPU: i-sliced pl-pl gravity with an allgather of the
drifted positions, and block-partitioned test particles against a replicated pl array.  The per-slice compute
is done by the oracle here (CPU box); on the GPU box tests/test_gpu_multi.py runs the same decomposition through
the C ABI + NCCL."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, ntp, out):
    sys.path.insert(0, ROOT)
    from oracle import load
    from swiftest_b200 import shard, workloads as W
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = load()
    d = W.disk(n, seed=42)
    rh, vb = d["rh"].copy(), d["vh"].copy()
    dt = d["dt"]
    i0, i1 = shard.partition(n, world, rank)
    for _ in range(2):  # kick - drift - allgather, twice
        # rows [i0,i1): full-row sums over ALL columns, i.e. the tri kernel restricted to a row slice
        ah = np.zeros((n, 3))
        o.omp_kick_tri_rad_pl_rows(rh, d["Gmass"], d["radius"], ah, n, i0, i1)
        vb[i0:i1] += ah[i0:i1] * dt
        x, v, fl = o.drift_all(d["mu"][i0:i1], rh[i0:i1], vb[i0:i1], dt)
        assert not fl.any()
        # allgather of (possibly unequal) slices through padded buffers, like swcu_pl_allgather
        maxc = -(-n // world)
        send = torch.zeros(6, maxc, dtype=torch.float64)
        send[:3, : i1 - i0] = torch.from_numpy(x.T.copy())
        send[3:, : i1 - i0] = torch.from_numpy(v.T.copy())
        recv = [torch.zeros_like(send) for _ in range(world)]
        dist.all_gather(recv, send)
        for r in range(world):
            j0, j1 = shard.partition(n, world, r)
            rh[j0:j1] = recv[r][:3, : j1 - j0].numpy().T
            vb[j0:j1] = recv[r][3:, : j1 - j0].numpy().T
    # test particles: block partition, replicated planets, no communication until the final gather
    tp = W.tp_cloud(ntp, seed=5)
    t0, t1 = shard.tp_block_partition(ntp, world, rank)
    acc = o.kick_all_tp(tp["rh"][t0:t1], rh, d["Gmass"], np.ones(t1 - t0, np.int32), np.zeros((t1 - t0, 3)))
    gathered = [None] * world
    dist.all_gather_object(gathered, (t0, t1, acc))
    if rank == 0:
        full = np.zeros((ntp, 3))
        for (a, b, blk) in gathered:
            full[a:b] = blk
        np.savez(out, rh=rh, vb=vb, tp_acc=full)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_slices_match_single_process(tmp_path, oracle):
    from swiftest_b200 import workloads as W
    n, ntp, world = 301, 57, 2
    out = str(tmp_path / "res.npz")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, n, ntp, out), nprocs=world, join=True)
    res = np.load(out)
    d = W.disk(n, seed=42)
    rh, vb = d["rh"].copy(), d["vh"].copy()
    for _ in range(2):
        ah = oracle.kick_tri_pl(rh, d["Gmass"], d["radius"], np.zeros((n, 3)))
        vb += ah * d["dt"]
        rh, vb, fl = oracle.drift_all(d["mu"], rh, vb, d["dt"])
    assert np.array_equal(res["rh"], rh) and np.array_equal(res["vb"], vb)  # same row order => bit-identical
    tp = W.tp_cloud(ntp, seed=5)
    ref = oracle.kick_all_tp(tp["rh"], rh, d["Gmass"], np.ones(ntp, np.int32), np.zeros((ntp, 3)))
    assert np.array_equal(res["tp_acc"], ref)


def _rebalance_worker(rank, world, port, ntp, out):
    sys.path.insert(0, ROOT)
    from swiftest_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ids = np.arange(ntp)
    t0, t1 = shard.tp_block_partition(ntp, world, rank)
    mine = ids[t0:t1]
    # discards skew the counts: rank 0 loses most of its particles (e.g. an inner-edge discard)
    rng = np.random.default_rng(100 + rank)
    keep = rng.uniform(size=len(mine)) > (0.8 if rank == 0 else 0.05)
    mine = mine[keep]
    counts = [None] * world
    dist.all_gather_object(counts, len(mine))
    did = shard.needs_rebalance(counts)
    if did:
        plan = shard.rebalance_plan(counts)
        new_n = shard.tp_block_partition(sum(counts), world, rank)
        new = np.full(new_n[1] - new_n[0], -1)
        sends = [(src, lo, hi, dst, dlo) for (src, lo, hi, dst, dlo) in plan if src == rank]
        outbox = [[] for _ in range(world)]
        for (_, lo, hi, dst, dlo) in sends:
            outbox[dst].append((dlo, mine[lo:hi]))
        # (gloo has no object all-to-all: gather every outbox, pick our part)
        allbox = [None] * world
        dist.all_gather_object(allbox, outbox)
        for src in range(world):
            for (dlo, blk) in allbox[src][rank]:
                new[dlo:dlo + len(blk)] = blk
        mine = new
    gathered = [None] * world
    dist.all_gather_object(gathered, (did, mine))
    if rank == 0:
        np.savez(out, did=np.array([g[0] for g in gathered]), **{f"r{k}": g[1] for k, g in enumerate(gathered)})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_tp_rebalance_after_discards(tmp_path):
    """swiftest_coarray_balance_system on two ranks: after skewed discards the particles are collected in image order
    and cut into equal blocks again; nothing is lost, order is kept."""
    from swiftest_b200 import shard
    ntp, world = 1000, 2
    out = str(tmp_path / "reb.npz")
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_rebalance_worker, args=(world, port, ntp, out), nprocs=world, join=True)
    res = np.load(out)
    assert res["did"].all()
    a, b = res["r0"], res["r1"]
    assert abs(len(a) - len(b)) <= 1 and (a >= 0).all() and (b >= 0).all()
    merged = np.concatenate([a, b])
    assert np.array_equal(merged, np.sort(merged)) and len(np.unique(merged)) == len(merged)


def test_rebalance_plan_properties():
    from swiftest_b200 import shard
    assert not shard.needs_rebalance([10, 10, 11, 9])
    assert shard.needs_rebalance([10, 10, 14, 9, 10])
    assert not shard.needs_rebalance([])
    for counts in ([5, 0, 17, 3], [0, 0, 0], [7], [1, 2, 3, 4, 5, 6, 7, 8]):
        plan = shard.rebalance_plan(counts)
        ntot, nimg = sum(counts), len(counts)
        got = [[None] * (shard.tp_block_partition(ntot, nimg, k)[1] - shard.tp_block_partition(ntot, nimg, k)[0])
               for k in range(nimg)]
        for (src, lo, hi, dst, dlo) in plan:
            assert 0 <= lo < hi <= counts[src]
            for t in range(hi - lo):
                assert got[dst][dlo + t] is None
                got[dst][dlo + t] = (src, lo + t)
        flat = [x for blk in got for x in blk]
        assert None not in flat and flat == sorted(flat) and len(flat) == ntot


def _pair_slice_worker(rank, world, port, n, out):
    """The decomposition of the fused peer-memory step (swcu_pl_kick_drift_p2p): every rank evaluates a run of the
    flattened pair list into its own partial-acceleration buffer, then for ITS slice of bodies sums the partials of all
    ranks in rank order (reduce-scatter), kicks, drifts, and delivers the slice to everybody (allgather)."""
    sys.path.insert(0, ROOT)
    from oracle import load
    from swiftest_b200 import shard, workloads as W
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = load()
    d = W.disk(n, seed=77)
    rh, vb = d["rh"].copy(), d["vh"].copy()
    iu = np.triu_indices(n, 1)
    k_all = np.stack([iu[0] + 1, iu[1] + 1], axis=1).astype(np.int32)       # canonical flattened order, 1-based
    p0, p1 = shard.partition(len(k_all), world, rank)
    i0, i1 = shard.partition(n, world, rank)
    for _ in range(2):
        F = o.kick_flat_pl(rh, d["Gmass"], d["radius"], np.zeros((n, 3)), k_plpl=k_all[p0:p1])
        parts = [torch.zeros(n, 3, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(parts, torch.from_numpy(F))
        ah = np.zeros((i1 - i0, 3))
        for r in range(world):                                                 # fixed rank order, own slice only
            ah = ah + parts[r][i0:i1].numpy()
        v = vb[i0:i1] + ah * d["dt"]
        x, v, fl = o.drift_all(d["mu"][i0:i1], rh[i0:i1], v, d["dt"])
        assert not fl.any()
        got = [None] * world
        dist.all_gather_object(got, (i0, i1, x, v))
        for (a, b, xs, vs) in got:
            rh[a:b], vb[a:b] = xs, vs
    if rank == 0:
        np.savez(out, rh=rh, vb=vb)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_pair_slices_reduce_scatter_matches_single_process(tmp_path, oracle):
    from swiftest_b200 import workloads as W
    n, world = 257, 2
    out = str(tmp_path / "p2p.npz")
    port = 33500 + (os.getpid() % 2000)
    mp.spawn(_pair_slice_worker, args=(world, port, n, out), nprocs=world, join=True)
    res = np.load(out)
    d = W.disk(n, seed=77)
    rh, vb = d["rh"].copy(), d["vh"].copy()
    for _ in range(2):
        ah = oracle.kick_flat_pl(rh, d["Gmass"], d["radius"], np.zeros((n, 3)))
        vb = vb + ah * d["dt"]
        rh, vb, fl = oracle.drift_all(d["mu"], rh, vb, d["dt"])
    # the partial sums are regrouped by rank: same terms, different association
    assert np.max(np.abs(res["rh"] - rh) / np.linalg.norm(rh, axis=1, keepdims=True)) < 1e-14
    assert np.max(np.abs(res["vb"] - vb) / np.linalg.norm(vb, axis=1, keepdims=True)) < 1e-13

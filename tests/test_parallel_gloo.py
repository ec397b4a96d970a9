"""Multi-rank logic on CPU: world_size = 2 over gloo. Covers the row sharding and the partial-record exchange
(all-gather + log-sum-exp merge) that the N > 1 GPU path uses; the per-rank arithmetic is the oracle here,
the merge is the product's (libais_b200 host functions through ais_benchmarks_b200.parallel)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_golden


def _kld_record(lp, lq):
    inl = (lp > -np.inf) & (lq > -np.inf)
    p, q = lp[inl], lq[inl]
    mp_, mq = (p.max(), q.max()) if len(p) else (-3e38, -3e38)
    return np.array([mp_, np.exp(p - mp_).sum(), (np.exp(p - mp_) * (p - q)).sum(), mq, np.exp(q - mq).sum(),
                     inl.sum(), 0, len(inl)], dtype=np.float64)


def _weight_record(lw):
    m = lw.max() if len(lw) else -3e38
    return np.array([m, np.exp(lw - m).sum(), np.exp(2 * (lw - m)).sum(), len(lw)], dtype=np.float64)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ais_benchmarks_b200 import _device, parallel
    g = load_golden("kde_kld")
    lp, lq = g["lp1"], g["lq1"]
    s, e = parallel.shard_rows(len(lp))
    kld = parallel.global_kld(_kld_record(lp[s:e], lq[s:e]))
    gw = load_golden("dm_weights")
    lw = np.log(gw["w1"])
    s2, e2 = parallel.shard_rows(len(lw))
    rec = parallel.global_weight_partials(_weight_record(lw[s2:e2]))
    ness = _device.weight_ess(rec) / rec[3]
    t = torch.full((3,), float(rank + 1), dtype=torch.float64)
    parallel.all_reduce_sum(t)
    # JSD / EVMSE with sharded evaluation points: the record exchange of CJSDivergence / CExpectedValueMSE (sharded=True)
    x = g["x1"]
    pair = parallel.merge_kld_partials(parallel.all_gather_records(_kld_record(lp[s:e], lq[s:e])))
    Lp, Lq = _device.pair_log_norms(pair)
    a, b = lp[s:e] - Lp, lq[s:e] - Lq
    lm = np.logaddexp(a, b) - np.log(2)
    terms = np.array([(0.5 * np.exp(a) * (a - lm) + 0.5 * np.exp(b) * (b - lm)).sum(), 0.0])
    jsd = parallel.all_gather_records(terms).sum(axis=0)[0] / np.log(2)
    evs = []
    for lw_ in (lp, lq):
        norm = _device.weight_logsum(parallel.global_weight_partials(_weight_record(lw_[s:e])))
        evs.append(parallel.all_gather_records((x[s:e] * np.exp(lw_[s:e] - norm)[:, None]).sum(axis=0)).sum(axis=0))
    evmse = float((evs[0] - evs[1]) @ (evs[0] - evs[1]))
    np.save(os.path.join(out_dir, "r%d.npy" % rank), np.array([kld, ness, rec[3], t[0].item(), s, e, jsd, evmse]))
    dist.destroy_process_group()


def test_world2_gloo_partials(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    g = load_golden("kde_kld")
    gw = load_golden("dm_weights")
    gm = load_golden("metrics")
    res = [np.load(os.path.join(str(tmp_path), "r%d.npy" % r)) for r in range(world)]
    for r in res:
        assert r[0] == pytest.approx(float(g["kld1"]), rel=1e-10)      # KLD over the union of both shards
        assert r[1] == pytest.approx(float(gw["ness1"]), rel=1e-10)    # NESS from the merged weight record
        assert r[2] == len(gw["w1"]) and r[3] == 3.0
        assert r[6] == pytest.approx(float(gm["jsd1"]), rel=1e-9)      # JSD / EVMSE from merged records
        assert r[7] == pytest.approx(float(gm["evmse1"]), rel=1e-9)
    assert res[0][0] == res[1][0]                                      # identical on every rank
    n = len(g["lp1"])
    assert (res[0][4], res[0][5], res[1][4], res[1][5]) == (0, n // 2, n // 2, n)


def test_shard_rows_covers_everything():
    from ais_benchmarks_b200.parallel import shard_rows
    for n in (0, 1, 7, 8, 1000003):
        for w in (1, 2, 3, 8):
            spans = [shard_rows(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


# ---------------------------------------------------------------------------------------------------
# Adaptation over row-sharded samples: two-level multinomial resampling, chain-shard gather, broadcast, host all-reduce
# ---------------------------------------------------------------------------------------------------
def _adapt_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ais_benchmarks_b200 import parallel
    np.random.seed(100 + rank)          # ranks deliberately disagree: rank 0 decides, the others follow
    torch.manual_seed(200 + rank)
    # rank r owns rows whose first coordinate is r; rank 1's rows carry 3x the mass of rank 0's
    n_loc, D, N = 500 + 100 * rank, 3, 4000
    rows = torch.cat([torch.full((n_loc, 1), float(rank)), torch.arange(n_loc, dtype=torch.float32).reshape(-1, 1),
                      torch.randn(n_loc, 1)], dim=1)
    logw = torch.full((n_loc,), float(np.log((1.0 if rank == 0 else 3.0) / n_loc)))
    logw[0] = float("nan")              # invalid weights count as zero
    logw[1] = -float("inf")
    new = parallel.resample_global(rows, logw, N)
    # degenerate weights -> uniform over the union
    deg = parallel.resample_global(rows, torch.full((n_loc,), -float("inf")), 1000)
    start = parallel.broadcast_array(np.random.uniform(size=(5, 2)))
    summed = parallel.all_reduce_sum_host(np.arange(6, dtype=np.float64).reshape(2, 3) * (rank + 1))
    counts = [3, 1]
    gathered = parallel.all_gather_rows(torch.full((counts[rank], 2), float(rank + 1)), counts)
    np.savez(os.path.join(out_dir, "a%d.npz" % rank), new=new.numpy(), deg=deg.numpy(), start=start, summed=summed,
             gathered=gathered.numpy())
    dist.destroy_process_group()


def test_world2_gloo_sharded_adaptation(tmp_path):
    world = 2
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_adapt_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [np.load(os.path.join(str(tmp_path), "a%d.npz" % k)) for k in range(world)]
    for key in ("new", "deg", "start", "summed", "gathered"):
        assert np.array_equal(r[0][key], r[1][key]), key            # replicated state stays identical
    new = r[0]["new"]
    assert new.shape == (4000, 3)
    frac1 = (new[:, 0] == 1).mean()
    assert abs(frac1 - 0.75) < 0.03                                  # shard masses 1 : 3
    # chosen rows exist in their shard and never carry an invalid / zero weight
    assert set(np.unique(new[:, 0])) <= {0.0, 1.0}
    assert new[new[:, 0] == 0, 1].max() < 500 and new[new[:, 0] == 1, 1].max() < 600
    assert new[:, 1].min() >= 2
    # within a shard the draws are uniform over its (equal-weight) rows
    assert abs(new[new[:, 0] == 1, 1].mean() - 300.5) < 20
    deg = r[0]["deg"]
    assert abs((deg[:, 0] == 1).mean() - 600 / 1100) < 0.06
    assert np.array_equal(r[0]["summed"], np.arange(6).reshape(2, 3) * 3.0)
    assert np.array_equal(r[0]["gathered"], np.array([[1, 1]] * 3 + [[2, 2]], dtype=np.float32))


def _seed_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from ais_benchmarks_b200 import _device
    _device.seed(1234)                              # seeded BEFORE the process group exists: rank still unknown (0)
    k_before = _device._key()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    k_after = _device._key()                        # the rank is folded in as soon as it is known
    _device.seed(1234)
    k_reseed = _device._key()
    off0 = _device._next_offset(40)
    off1 = _device._next_offset(40)
    np.save(os.path.join(out_dir, "s%d.npy" % rank), np.array([k_before, k_after, k_reseed, off0, off1], dtype=np.uint64))
    dist.destroy_process_group()


def test_world2_per_rank_philox_streams(tmp_path):
    """Sharded samplers draw each rank's share with the same (seed, offset) bookkeeping: the Philox KEY must differ by
    rank or every rank would draw bit-identical samples. Rank 0 keeps the user's seed as its key."""
    port = 29700 + os.getpid() % 200
    mp.spawn(_seed_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (np.load(os.path.join(str(tmp_path), "s%d.npy" % r)) for r in (0, 1))
    assert r0[0] == 1234 and r1[0] == 1234          # before init_process_group nobody knows its rank
    assert r0[1] == 1234 and r1[1] != 1234 and r1[1] == r1[2]
    assert r0[3] == r1[3] == 0 and r0[4] == r1[4] == 11   # the offsets advance identically: only the key separates ranks

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

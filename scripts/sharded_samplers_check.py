"""Sharded adapt loops on N GPUs (torchrun, one rank per GPU): DM-AIS, LAIS and M-PMC with `sharded: True`.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
      scripts/sharded_samplers_check.py

Checks, on every rank: replicated proposal parameters stay bit-identical across ranks after every call; the ranks draw
DIFFERENT samples (per-rank Philox streams); the returned weights are normalised over ALL ranks; NESS agrees on every
rank; the weights of this rank's rows match the float64 oracle evaluated with the proposal parameters that generated
them; the adapted proposals have moved towards the target (KLD of the proposal mixture falls). Prints one line per
sampler on rank 0 and exits non-zero on failure.

`run_checks(rank, world, use_oracle=False)` is what `bench.py --gpus N` (N > 1) calls after its timed sections, so the
driver's scaling runs exercise the sharded adapt loops too; there the reference weights come from an UNSHARDED
evaluation of the same rows with the library's weight kernel instead of the oracle (bench.py may not lean on oracle/
outside its CPU-baseline leg).
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import bench
from ais_benchmarks_b200 import _device, parallel
from ais_benchmarks_b200.distributions import CGaussianMixtureModel
from ais_benchmarks_b200.metrics import CKLDivergence
from ais_benchmarks_b200.sampling_methods import CDeterministicMixtureAIS, CLayeredAIS, CMixturePMC
FAILURES = []


def ensure(cond, msg):
    """Record a failed check WITHOUT leaving the collective call sequence (a rank that raised while the others went
    on to the next all-gather would hang the job); run_checks raises once, at the end, on every rank."""
    if not cond:
        FAILURES.append(str(msg))


def same_on_all_ranks(t, what):
    t = t.detach().to("cuda", torch.float64).contiguous().reshape(-1)
    ws = dist.get_world_size()
    out = torch.empty(ws * t.numel(), dtype=torch.float64, device="cuda")
    dist.all_gather_into_tensor(out, t)
    out = out.reshape(ws, -1)
    ensure(bool((out == out[0:1]).all() or (torch.isnan(out).all())), "%s differs across ranks" % what)


def global_sum(v):
    t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
    dist.all_reduce(t)
    return float(t[0])


def differs_across_ranks(t, what):
    """First rows of every rank's draw: no two ranks may hold the same numbers."""
    t = t.detach().to("cuda", torch.float64).contiguous().reshape(-1)[:64]
    ws = dist.get_world_size()
    out = torch.empty(ws * t.numel(), dtype=torch.float64, device="cuda")
    dist.all_gather_into_tensor(out, t)
    out = out.reshape(ws, -1)
    for a in range(ws):
        for b in range(a + 1, ws):
            ensure(not bool((out[a] == out[b]).all()), "%s: ranks %d and %d drew identical samples" % (what, a, b))


def run_checks(rank, world, use_oracle=True):
    """-> list of report lines (identical on every rank). Raises AssertionError on failure."""
    if use_oracle:
        from oracle import ais_oracle as O
    D, N = 8, 64
    mu, sig, w, sup = bench.random_gmm(0, D, 16)
    target = CGaussianMixtureModel({"means": mu, "sigmas": sig, "weights": w, "support": sup})
    metric = CKLDivergence()
    np.random.seed(5)                      # same evaluation points on every rank
    x_eval = np.random.uniform(sup[0], sup[1], size=(20000, D))
    lines = []

    def run(cls, extra, K, iters, name):
        np.random.seed(1000 + rank)        # ranks disagree on purpose: rank 0's draw must win
        _device.seed(77)                   # the SAME seed everywhere: _device folds the rank into its Philox streams
        params = {"space_min": sup[0], "space_max": sup[1], "dims": D, "K": K, "N": N, "J": 1000, "sigma": 0.05,
                  "device_outputs": True, "sharded": True}
        params.update(extra)
        s = cls(params)
        per_iter = K if cls is CMixturePMC else N * K
        if cls is CMixturePMC:
            same_on_all_ranks(torch.from_numpy(s._means), name + " initial means")
            pm0 = (s._means.copy(), np.array([np.diag(c) for c in s._covs]), s._mix_weights.copy())
        else:
            same_on_all_ranks(s.proposal_means(), name + " initial means")
            pm0 = s.proposal_means().double().cpu().numpy()
        kld0 = metric.compute_from_samples(target, s, x_eval)
        x, wn = s.importance_sample(target, per_iter)            # one iteration: weights are pure functions of pm0
        n_loc = x.shape[0]
        ensure(abs(global_sum(n_loc) - per_iter) < 0.5, (name, n_loc))
        differs_across_ranks(x, name)
        ensure(abs(global_sum(wn.double().sum()) - 1.0) < 1e-4, name)
        # oracle on a row subset of THIS rank's rows
        rs = np.random.RandomState(rank)
        pick = rs.choice(n_loc, size=min(500, n_loc), replace=False)
        xs = x[pick].double().cpu().numpy()
        if use_oracle:
            if cls is CMixturePMC:
                ref_lq = O.gmm_log_prob(xs, pm0[0], pm0[1], pm0[2])
            else:
                ref_lq = O.iso_mixture_log_prob(xs, pm0, 0.05)
            ref_lw = O.gmm_log_prob(xs, mu, sig, w) - ref_lq
        else:
            # unsharded evaluation of the same rows with the library's weight kernel and the parameters saved above
            from ais_benchmarks_b200._lib import COV_DIAG, COV_ISO_SHARED
            if cls is CMixturePMC:
                qm = _device.DeviceMixture.from_params(pm0[0], pm0[1], pm0[2], COV_DIAG)
            else:
                qm = _device.DeviceMixture.from_params(pm0, np.array([0.05]), None, COV_ISO_SHARED)
            ref_lw = _device.dm_weights(x[pick].contiguous().float(), target.device_mixture(), None, qm)[0]
            ref_lw = ref_lw.double().cpu().numpy()
        got_lw = np.log(wn[pick].double().cpu().numpy())
        ok = np.isfinite(got_lw) & np.isfinite(ref_lw)
        shift = np.median((got_lw - ref_lw)[ok])
        # the normaliser is global: every rank must see the same shift
        same_on_all_ranks(torch.tensor([round(float(shift), 3)]), name + " global normaliser")
        err = np.max(np.abs(got_lw[ok] - shift - ref_lw[ok]) / np.maximum(1.0, np.abs(ref_lw[ok])))
        ensure(err < 1e-4, (name, err))
        x, wn = s.importance_sample(target, iters * per_iter)    # more adapt iterations
        ensure(abs(global_sum(x.shape[0]) - iters * per_iter) < 0.5, name + " sample count")
        if cls is CMixturePMC:
            same_on_all_ranks(torch.from_numpy(s._means), name + " adapted means")
            same_on_all_ranks(torch.from_numpy(s._mix_weights), name + " adapted weights")
        else:
            same_on_all_ranks(s.proposal_means(), name + " adapted means")
            ensure(abs(global_sum(wn.double().sum()) - 1.0) < 1e-4, name)
        # M-PMC normalises its weights per batch (reference m_pmc.py:108,118): no single global NESS to compare
        ness = s.get_approx_NESS() if cls is not CMixturePMC else None
        if ness is not None:
            same_on_all_ranks(torch.tensor([ness]), name + " NESS")
        kld1 = metric.compute_from_samples(target, s, x_eval)
        # the same metric with the points dealt out over the ranks by the metric itself (full array on every rank)
        dealt = CKLDivergence()
        dealt.sharded, dealt.replicated_points = True, True
        kld1_dealt = dealt.compute_from_samples(target, s, x_eval)
        ensure(abs(kld1_dealt - kld1) < 2e-5 * max(1.0, abs(kld1)), (name, "dealt-out KLD", kld1_dealt, kld1))
        lines.append("%-8s world %d: %d samples over %d iterations, local rows %d, weight parity %.1e, NESS %s, "
                     "KLD(target || proposal) %.3f -> %.3f" % (name, world, iters * per_iter, iters, x.shape[0], err,
                                                             "%.4g" % ness if ness is not None else "per batch",
                                                             kld0, kld1))
        return kld0, kld1

    k0, k1 = run(CDeterministicMixtureAIS, {}, 101, 8, "DM-AIS")
    ensure(k1 < k0, ("DM-AIS did not adapt", k0, k1))
    k0, k1 = run(CLayeredAIS, {"L": 10, "mh_sigma": 0.05}, 101, 8, "LAIS")
    ensure(k1 < k0, ("LAIS did not adapt", k0, k1))
    run(CMixturePMC, {}, 50001, 3, "M-PMC")
    # replicated numpy input: each rank stages a slice, NVLink carries the rest -- bit-identical to a full upload
    from ais_benchmarks_b200 import parallel
    for shape, dt in (((300_001, D), np.float64), ((world * 70_000 + 3,), np.float32)):
        a = np.random.RandomState(9).standard_normal(shape).astype(dt)
        got, want = parallel.upload_replicated(a), _device.upload_f32(a)
        ensure(got.shape == want.shape and torch.equal(got, want), ("upload_replicated", shape))
    lines.append("replicated upload world %d: slice + all-gather equals the full upload (float64 [300001, %d], "
                 "float32 [%d])" % (world, D, world * 70_000 + 3))
    # agree on the verdict: every rank raises if any rank recorded a failure
    bad = torch.tensor([len(FAILURES)], dtype=torch.int64, device="cuda")
    dist.all_reduce(bad)
    if int(bad.item()) > 0:
        msgs = list(FAILURES)
        FAILURES.clear()
        raise AssertionError("sharded sampler checks failed on some rank (%d failures in total); this rank: %s"
                             % (int(bad.item()), msgs))
    return lines


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lines = run_checks(rank, world, use_oracle=True)
    if rank == 0:
        print("\n".join(lines))
        print("sharded samplers ok on %d ranks" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

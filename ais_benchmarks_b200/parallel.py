"""Multi-GPU sharding of the hot path: one process per GPU (torchrun), points / samples sharded by
rows, mixture parameters replicated, and ONLY the small partial records exchanged:

  * KLD:      8 doubles per rank  (m_p, S_p, A, m_q, S_q, n_in, n_posinf_q, n_total)
  * weights:  4 doubles per rank  (M, S1, S2, n_valid)
  * M-PMC:    K x (D+1) doubles   (all-reduce SUM of the moment accumulators)
  * DM-AIS / LAIS adaptation: the N new proposal means (N x D floats), the one non-scalar exchange
    (two-level multinomial resampling: resample_global; chain shards: all_gather_rows)

The records are all-gathered and merged on every rank with the log-sum-exp rescale rule
(C ABI aisb_kld_partials_merge / aisb_weight_partials_merge), so every rank ends with the same
scalar. The reference has no distributed path at all; this is the B200-native
data-parallel design of SURVEY.md section 8(e). Works with the NCCL backend on GPUs and with gloo
on CPU tensors (used by the world_size-2 tests).
"""
import numpy as np
import torch
import torch.distributed as dist

from . import _device


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_rows(n, rank=None, world_size=None):
    """Contiguous, balanced row range [start, stop) of rank `rank` out of `world_size` for n rows."""
    if rank is None or world_size is None:
        rank, world_size = world()
    base, rem = divmod(int(n), int(world_size))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


SPATIAL_BLOCK = 256    # rows per block of the block-cyclic split of the Z-order curve (2 warp tiles of the KDE kernels;
                       # 1e6 points over 8 ranks: the fullest rank holds +0.15 %, with blocks of 1024 it held +0.4 %)


def spatial_shard(x, rank=None, world_size=None):
    """This rank's share of the point set x (float32 CUDA [n, D], the SAME on every rank): the rows are put in Z-order
    and dealt out in blocks of SPATIAL_BLOCK consecutive curve positions, round-robin over the ranks
    -> (rows [m, D], their indices into x). Every warp of the KDE kernels still sees 128 NEIGHBOURS of the full set (the
    density that lets it skip negligible groups, csrc/pairwise.cuh), and every rank gets pieces of every region of
    space: a contiguous piece of the curve per rank (round 1) left the ranks with different refinement loads
    (1.3 ms of waiting per step at N = 2). Any partition gives the same KLD. The permutation is recomputed identically
    on every rank (no communication)."""
    if rank is None or world_size is None:
        rank, world_size = world()
    perm = _device.morton_order(x)
    if world_size == 1:
        return _device.gather_rows(x, perm), perm
    n = x.shape[0]
    nblocks = (n + SPATIAL_BLOCK - 1) // SPATIAL_BLOCK
    mine_blocks = torch.arange(rank, nblocks, world_size, device=x.device)
    pos = (mine_blocks[:, None] * SPATIAL_BLOCK + torch.arange(SPATIAL_BLOCK, device=x.device)[None, :]).reshape(-1)
    pos = pos[pos < n]
    mine = perm[pos]
    return _device.gather_rows(x, mine), mine


def upload_replicated(a):
    """float32 CUDA copy of a numpy array that is IDENTICAL on every rank (an SPMD caller's KDE samples, say): every
    rank pushes only its 1/world slice through the host link (_device.upload_f32) and the slices are exchanged with an
    in-place all-gather over NVLink. The ranks of a node share the host's memory bandwidth and cores: 8 ranks each
    staging the same 64 MB of float64 took 14 ms of a 61 ms configs[3] step, against 47 ms with everything resident.
    Falls back to a plain upload without a process group, under gloo, or for small arrays. The result is bit-identical
    to _device.upload_f32(a) -- provided the caller's promise holds."""
    a = np.ascontiguousarray(a)
    rank, ws = world()
    if ws == 1 or dist.get_backend() != "nccl" or a.size < ws * _device.HOST_CAST_MAX:
        return _device.upload_f32(a)
    flat = a.reshape(-1)
    per = (flat.size + ws - 1) // ws
    out = torch.empty(ws * per, dtype=torch.float32, device=_device.device())
    lo, hi = min(flat.size, rank * per), min(flat.size, (rank + 1) * per)
    mine = out[rank * per:(rank + 1) * per]
    if hi > lo:
        _device.upload_f32(flat[lo:hi], out=mine[:hi - lo])
    dist.all_gather_into_tensor(out, mine)            # in place: this rank's slice already sits where it belongs
    return out[:flat.size].reshape(a.shape)


def _comm_device():
    if dist.get_backend() == "nccl":
        return _device.device()
    return torch.device("cpu")


def all_gather_records(record):
    """record: 1-D float64 array-like (host or device). -> [world, len] float64 numpy, same on every rank."""
    rank, ws = world()
    rec = record if _device.is_tensor(record) else torch.as_tensor(np.asarray(record, dtype=np.float64))
    if ws == 1:
        return rec.detach().cpu().numpy().reshape(1, -1)
    rec = rec.to(_comm_device(), torch.float64).contiguous().reshape(-1)
    out = torch.empty(ws * rec.numel(), dtype=torch.float64, device=rec.device)
    dist.all_gather_into_tensor(out, rec)
    return out.cpu().numpy().reshape(ws, -1)


def all_gather_records_dev(record):
    """record: 1-D float64 tensor on the communication device -> [world, len] tensor on the same device, identical on
    every rank. With NCCL nothing is copied to the host and nothing synchronises: the gather is stream-ordered."""
    _, ws = world()
    rec = record.reshape(1, -1)
    if ws == 1:
        return rec
    rec = rec.contiguous()
    out = torch.empty((ws, rec.shape[1]), dtype=rec.dtype, device=rec.device)
    dist.all_gather_into_tensor(out, rec)
    return out


class PeerRecordExchange:
    """Weight-record exchange over NVLink peer memory (C ABI aisb_weight_records_exchange): a symmetric buffer per
    rank (torch.distributed._symmetric_memory), one 1-CTA kernel per exchange, no collective launch. Created
    collectively on first use; if symmetric memory is not available on ANY rank, every rank falls back to the NCCL
    all-gather (the decision is agreed with an all-reduce so that the ranks cannot diverge)."""
    _instance = None
    _disabled = False

    def __init__(self):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm
        lib = _device.require_cuda()
        self.rank, self.world = world()
        self.lib, self.C = lib, C
        n = int(lib.aisb_weight_exchange_doubles(self.world))
        group = dist.group.WORLD
        self.buf = symm.empty(n, dtype=torch.float64, device=_device.device())
        self.buf.zero_()
        self.handle = symm.rendezvous(self.buf, group)
        self.peer_ptrs = torch.tensor([int(p) for p in self.handle.buffer_ptrs], dtype=torch.int64,
                                      device=_device.device())
        self.epoch = 0
        torch.cuda.synchronize()
        dist.barrier()          # every rank's buffer is zeroed and mapped before the first flag is written

    def exchange(self, local_record):
        """local_record: float64 CUDA [4] -> [world, 4] float64 CUDA (stream-ordered, no host synchronisation)."""
        self.epoch += 1
        out = torch.empty((self.world, 4), dtype=torch.float64, device=local_record.device)
        rec = local_record.reshape(-1).contiguous()
        _device.check(self.lib.aisb_weight_records_exchange(
            self.C.c_void_p(rec.data_ptr()), self.C.c_void_p(self.peer_ptrs.data_ptr()), self.rank, self.world,
            self.epoch, self.C.c_void_p(out.data_ptr()), _device.stream_ptr()), "aisb_weight_records_exchange")
        return out

    @classmethod
    def get(cls):
        """The process-wide exchange, or None (not NCCL, single rank, switched off with AISB_P2P=0, or unavailable)."""
        import os
        if cls._instance is not None or cls._disabled:
            return cls._instance
        _, ws = world()
        if ws == 1 or dist.get_backend() != "nccl" or os.environ.get("AISB_P2P", "1") == "0":
            cls._disabled = True
            return None
        inst, ok = None, 1
        try:
            inst = cls()
            # self-test: every rank publishes {rank, rank + 1, ...}; every rank must read all of them back
            mine = torch.arange(4, dtype=torch.float64, device=_device.device()) + float(inst.rank)
            got = inst.exchange(mine).cpu()
            want = torch.arange(4, dtype=torch.float64)[None, :] + torch.arange(ws, dtype=torch.float64)[:, None]
            if not torch.equal(got, want):
                raise RuntimeError("self-test mismatch %s" % got.tolist())
        except Exception as e:  # noqa: BLE001 -- any failure means "use NCCL", agreed below
            ok = 0
            if world()[0] == 0:
                print("[ais_benchmarks_b200] peer-memory record exchange unavailable (%s: %s); using NCCL all-gather"
                      % (type(e).__name__, e))
        flag = torch.tensor([ok], dtype=torch.int32, device=_device.device())
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 1:
            cls._instance = inst
        else:
            cls._disabled = True
        return cls._instance


def global_weight_record_dev(local_records):
    """Device-resident counterpart of global_weight_partials: this rank's weight record(s) ([4] or [C, 4], float64 CUDA)
    gathered from all ranks -> [ranks * C, 4], still on the device: a single record goes through the NVLink peer-memory
    exchange kernel (PeerRecordExchange), anything else through an NCCL all-gather. _device.normalize_weights_dev folds
    the records inside its own kernel, so a weighting step has no host round trip and no separate merge launch."""
    _, ws = world()
    if ws > 1 and local_records.numel() == 4 and local_records.is_cuda:
        px = PeerRecordExchange.get()
        if px is not None:
            return px.exchange(local_records)
    return all_gather_records_dev(local_records.reshape(-1)).reshape(-1, 4)


def merge_weight_partials(records):
    acc = np.asarray(records[0], dtype=np.float64)
    for r in records[1:]:
        acc = _device.weight_merge(acc, r)
    return acc


def merge_kld_partials(records):
    acc = np.asarray(records[0], dtype=np.float64)
    for r in records[1:]:
        acc = _device.kld_merge(acc, r)
    return acc


def global_weight_partials(local_record):
    """All ranks' weight records merged (identical result on every rank)."""
    return merge_weight_partials(all_gather_records(local_record))


def global_kld(local_record):
    """KLD over the union of all ranks' evaluation points from each rank's 8-double record."""
    return _device.kld_finalize(merge_kld_partials(all_gather_records(local_record)))


def all_reduce_sum(t):
    """In-place SUM all-reduce of a small accumulator tensor (M-PMC moments)."""
    _, ws = world()
    if ws > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def sharded_kld(p, q, samples):
    """CKLDivergence.compute_from_samples with the evaluation points row-sharded over the ranks.
    `samples` is the FULL (N, D) point set (replicated input); each rank evaluates its slice only."""
    from .metrics.divergences import CKLDivergence
    start, stop = shard_rows(len(samples))
    metric = CKLDivergence()
    metric.compute_from_samples(p, q, samples[start:stop])
    return global_kld(metric.partials)


# ----------------------------------------------------------------------------------------------------
# Adaptation steps over row-sharded samples
# ----------------------------------------------------------------------------------------------------
def _comm(t):
    """Tensor on the communication device (CUDA for NCCL, CPU for gloo)."""
    dev = _comm_device()
    return t if t.device.type == dev.type else t.to(dev)


def broadcast_array(a, src=0):
    """numpy array of rank `src` on every rank (replicated proposal parameters start identical)."""
    _, ws = world()
    a = np.ascontiguousarray(a, dtype=np.float64)
    if ws == 1:
        return a
    t = _comm(torch.from_numpy(a.copy()))
    dist.broadcast(t, src=src)
    return t.cpu().numpy()


def all_gather_rows(rows, counts):
    """rows: [counts[rank], D] tensor of this rank -> [sum(counts), D] on rows.device, rank-major, identical on
    every rank. `counts` (one int per rank) must be known to all ranks."""
    rank, ws = world()
    if ws == 1:
        return rows
    total, cmax = int(sum(counts)), int(max(counts))
    pad = torch.zeros((cmax, rows.shape[1]), dtype=rows.dtype, device=rows.device)
    pad[:rows.shape[0]] = rows
    send = _comm(pad).contiguous()
    recv = torch.empty((ws * cmax, rows.shape[1]), dtype=rows.dtype, device=send.device)
    dist.all_gather_into_tensor(recv, send)
    recv = recv.reshape(ws, cmax, rows.shape[1])
    out = torch.cat([recv[r, :int(counts[r])] for r in range(ws)]).to(rows.device)
    assert out.shape[0] == total
    return out


def shard_log_mass(logw):
    """Host record {m, sum exp(logw - m)} of this rank's log weights; invalid (NaN / +inf) weights count as zero,
    as in sampling_methods.base.multinomial_resample."""
    if logw.shape[0] == 0:
        return np.array([-np.inf, 0.0])
    lw = logw.double()
    lw = torch.where(torch.isfinite(lw) | (lw == -np.inf), lw, torch.full_like(lw, -np.inf))
    m = torch.max(lw)
    if not torch.isfinite(m):
        return np.array([-np.inf, 0.0])
    return np.array([float(m), float(torch.exp(lw - m).sum())])


def resample_global(samples, logw, n_draws):
    """DM-PMC resampling (dm_ais.py:47-56) over row-sharded samples: n_draws rows ~ Categorical(w / sum w) over the
    UNION of all ranks' rows, as a two-level multinomial -- shard counts ~ Multinomial(n_draws, shard masses) drawn by
    rank 0 and broadcast, then counts[rank] rows drawn locally -- followed by an all-gather of the chosen rows
    (n_draws x D floats). Every rank returns the same [n_draws, D] tensor. Degenerate weights (zero total mass) fall
    back to uniform draws, like the single-GPU path."""
    from .sampling_methods.base import multinomial_resample
    rank, ws = world()
    if ws == 1:
        return samples[multinomial_resample(logw, n_draws)]
    recs = all_gather_records(shard_log_mass(logw))                      # [G, 2]
    finite = np.isfinite(recs[:, 0])
    mass = np.zeros(ws)
    if finite.any():
        M = recs[finite, 0].max()
        mass[finite] = recs[finite, 1] * np.exp(recs[finite, 0] - M)
    sizes = all_gather_records(np.array([float(logw.shape[0])]))[:, 0]
    uniform = not (mass.sum() > 0)
    p = sizes / sizes.sum() if uniform else mass / mass.sum()
    counts = torch.zeros(ws, dtype=torch.int64)
    if rank == 0:
        counts = torch.from_numpy(np.random.multinomial(int(n_draws), p).astype(np.int64))
    counts = _comm(counts)
    dist.broadcast(counts, src=0)
    counts = [int(c) for c in counts.cpu().tolist()]
    c = counts[rank]
    if c == 0:
        chosen = samples[:0]
    elif uniform:
        chosen = samples[torch.randint(0, samples.shape[0], (c,), device=samples.device)]
    else:
        chosen = samples[multinomial_resample(logw, c)]
    return all_gather_rows(chosen, counts)


def all_reduce_sum_host(a):
    """SUM over ranks of a small numpy float64 array (M-PMC sufficient statistics)."""
    _, ws = world()
    a = np.ascontiguousarray(a, dtype=np.float64)
    if ws == 1:
        return a
    t = _comm(torch.from_numpy(a.copy()))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()

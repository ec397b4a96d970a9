"""Multi-GPU sharding of the hot path: one process per GPU (torchrun), points / samples sharded by
rows, mixture parameters replicated, and ONLY the small partial records exchanged:

  * KLD:      8 doubles per rank  (m_p, S_p, A, m_q, S_q, n_in, n_posinf_q, n_total)
  * weights:  4 doubles per rank  (M, S1, S2, n_valid)
  * M-PMC:    K x (D+1) doubles   (all-reduce SUM of the moment accumulators)

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


def global_weight_record_dev(local_record):
    """Device-resident counterpart of global_weight_partials: all ranks' weight records gathered over NCCL and merged
    by a one-block kernel; the result (float64 CUDA [4]) feeds _device.normalize_weights_dev without a host round trip."""
    recs = all_gather_records_dev(local_record)
    return local_record if recs.shape[0] == 1 else _device.merge_weight_records(recs)


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

"""Peer-memory weight-record exchange (parallel.PeerRecordExchange, C ABI aisb_weight_records_exchange) against the NCCL
all-gather on N GPUs: equality over many epochs, then the latency of both (CUDA events, max over ranks).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 \
      scripts/p2p_records_check.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from ais_benchmarks_b200 import _device, parallel


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    px = parallel.PeerRecordExchange.get()
    if px is None:
        if rank == 0:
            print("peer-memory exchange unavailable on this box: NCCL all-gather stays in use")
        dist.destroy_process_group()
        return
    g = torch.Generator(device="cuda")
    g.manual_seed(100 + rank)
    for it in range(300):
        rec = torch.randn(4, dtype=torch.float64, device="cuda", generator=g)
        got = px.exchange(rec)
        ref = parallel.all_gather_records_dev(rec)
        assert torch.equal(got, ref), (it, got, ref)
        if it % 7 == 0:      # skew the ranks: a late peer must not break the flag protocol
            torch.cuda._sleep(int(2e6) * (rank + 1))
    # end to end through the public helper + the normalise kernel
    logw = torch.randn(1000 + rank, device="cuda", generator=g)
    part = _device.logw_partials(logw)
    recs = parallel.global_weight_record_dev(part)
    w, stats = _device.normalize_weights_dev(logw, recs)
    tot = w.double().sum().reshape(1)
    dist.all_reduce(tot)
    assert abs(float(tot) - 1.0) < 1e-5, float(tot)

    def timed(fn, n=2000):
        for _ in range(20):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / n * 1e3], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)
    rec = torch.randn(4, dtype=torch.float64, device="cuda", generator=g)
    t_p2p = timed(lambda: px.exchange(rec))
    t_nccl = timed(lambda: parallel.all_gather_records_dev(rec))
    if rank == 0:
        print("record exchange on %d ranks: peer-memory kernel %.2f us, NCCL all-gather %.2f us per exchange "
              "(back to back, launch-bound on the host side for both)" % (world, t_p2p, t_nccl))
        print("p2p records ok on %d ranks" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

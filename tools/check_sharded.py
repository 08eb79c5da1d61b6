#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, NCCL): top-k of a corpus sharded by rows over the ranks
(ShardedFlatIPIndex: local search, all-gather of [Q,k] scores + ids, merge) must equal the top-k of one index over the
whole corpus, ids bit for bit and scores to 1e-6, including duplicated rows that tie exactly across shard boundaries.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_sharded.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdag_b200.retrieval import FlatIPIndex, ShardedFlatIPIndex  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    N, d, Q, k = 400_000, 256, int(os.environ.get('CHECK_Q', '512')), 10
    g = torch.Generator(device=dev).manual_seed(123)  # same corpus on every rank
    x = torch.randn(N, d, generator=g, device=dev)
    x /= x.norm(dim=1, keepdim=True)
    dup = torch.arange(0, N, 997, device=dev)
    x[(dup + N // 2 + 3) % N] = x[dup]  # exact ties whose two copies live on different shards
    xq = torch.randn(Q, d, generator=g, device=dev)
    xq /= xq.norm(dim=1, keepdim=True)
    xq[:32] = x[dup[:32]]  # queries that hit a duplicated pair with score 1
    rows = N // world
    lo = rank * rows
    hi = N if rank == world - 1 else lo + rows
    s_full, i_full = FlatIPIndex.from_device(x).search_device(xq, k)
    local = FlatIPIndex.from_device(x[lo:hi].contiguous(), id_offset=lo)
    flag = torch.tensor([1], device=dev)
    for mode in ("p2p", "nccl"):  # one kernel over peer memory (NVLink) / pack + ncclAllGather + merge
        sharded = ShardedFlatIPIndex(local, exchange=mode)
        ok = True
        for rep in range(5):  # repeated calls: the peer-memory exchange alternates its receive areas by call parity
            s_sh, i_sh = sharded.search_device(xq, k)
            ok = ok and bool(torch.equal(i_sh, i_full)) and bool(torch.allclose(s_sh, s_full, rtol=0, atol=1e-6))
        # pipelined form: two batches in flight (the second batch's local search is enqueued before the first exchange is awaited)
        prev = None
        for rep in range(6):
            h = sharded.submit(xq, k)
            if prev is not None:
                s_p, i_p = sharded.result(prev)
                ok = ok and bool(torch.equal(i_p, i_full)) and bool(torch.allclose(s_p, s_full, rtol=0, atol=1e-6))
            prev = h
        s_p, i_p = sharded.result(prev)
        ok = ok and bool(torch.equal(i_p, i_full)) and bool(torch.allclose(s_p, s_full, rtol=0, atol=1e-6))
        # exchange step alone: time 50 back-to-back searches of a tiny local index (the exchange dominates)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            sharded.search_device(xq, k)
        e1.record()
        torch.cuda.synchronize()
        # the exchange step alone, on the local results of the last search (Q x k records per rank)
        s_loc, i_loc = local.search_device(xq, k)
        import sdag_b200._lib as L_

        def exchange_only():
            if sharded.exchange == "p2p":
                return sharded._search_exchange_p2p(s_loc, i_loc, k)
            lib = L_.load()
            rec = torch.empty((Q, k, 3), dtype=torch.int32, device=dev)
            L_.check(lib.topk_pack(s_loc.data_ptr(), i_loc.data_ptr(), Q * k, rec.data_ptr(), L_.stream_ptr(dev)), "topk_pack")
            gathered = torch.empty((world, Q, k, 3), dtype=torch.int32, device=dev)
            dist.all_gather_into_tensor(gathered, rec)
            o_s = torch.empty((Q, k), dtype=torch.float32, device=dev)
            o_i = torch.empty((Q, k), dtype=torch.int64, device=dev)
            L_.check(lib.topk_merge_packed(gathered.data_ptr(), world, Q, k, o_s.data_ptr(), o_i.data_ptr(), L_.stream_ptr(dev)), "merge")
            return o_s, o_i

        for _ in range(5):
            exchange_only()
        torch.cuda.synchronize()
        dist.barrier()
        x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        x0.record()
        for _ in range(200):
            exchange_only()
        x1.record()
        torch.cuda.synchronize()
        xs, xi = exchange_only()
        ok = ok and bool(torch.equal(xi, i_full))
        exch_us = x0.elapsed_time(x1) / 200 * 1e3
        f = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(f, op=dist.ReduceOp.MIN)
        flag = torch.minimum(flag, f)
        if rank == 0:
            print(f"[{mode} -> ran as {sharded.exchange}] sharded == single-index top-{k} on {world} GPUs: {'OK' if int(f) else 'MISMATCH'} "
                  f"(N={N}, Q={Q}, {len(dup)} duplicated rows across shards); {e0.elapsed_time(e1) / 50:.3f} ms per sharded search, exchange + merge alone {exch_us:.1f} us")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()

"""world_size-2 gloo test of the sharded-retrieval plumbing (row partition, global ids, all-gather, merge order).

The CUDA kernels cannot run here, so each rank's local shard is searched by the oracle and the merge is the
oracle's — what is under test is the host-side distribution logic of ``ShardedFlatIPIndex``."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.flat_ip import FlatIPOracle, merge_topk_oracle


class _OracleShard:
    """Duck-typed stand-in for FlatIPIndex on CPU: same search_device contract, ids offset by the shard start."""

    def __init__(self, x, id_offset):
        self.d = x.shape[1]
        self.ntotal = x.shape[0]
        self.device = torch.device("cpu")
        self.id_offset = id_offset
        self._o = FlatIPOracle(self.d)
        self._o.add(x)

    def search_device(self, xq, k):
        D, I = self._o.search(xq.numpy(), k)
        I = np.where(I >= 0, I + self.id_offset, -1)
        return torch.from_numpy(D), torch.from_numpy(I)


def _merge(gs, gi, k):
    s, i = merge_topk_oracle(gs.numpy(), gi.numpy(), k)
    return torch.from_numpy(s), torch.from_numpy(i)


def _worker(rank, world, port, N, d, Q, k, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from sdag_b200.retrieval import ShardedFlatIPIndex

        rng = np.random.default_rng(42)
        x = rng.standard_normal((N, d)).astype(np.float32)
        x /= np.linalg.norm(x, axis=1, keepdims=True)
        x[N - 3] = x[2]  # cross-shard exact tie: the lower global id must come first
        q = np.concatenate([rng.standard_normal((Q - 1, d)).astype(np.float32), x[2:3]])
        rows = N // world
        lo = rank * rows
        hi = N if rank == world - 1 else lo + rows
        sharded = ShardedFlatIPIndex(_OracleShard(x[lo:hi], lo), merge_fn=_merge)
        assert sharded.ntotal == N
        D, I = sharded.search(q, k)
        np.save(os.path.join(out_dir, f"D{rank}.npy"), D)
        np.save(os.path.join(out_dir, f"I{rank}.npy"), I)
        # submit / result (the pipelined form): with an injected merge there is no peer-memory exchange to overlap, the
        # handle must still carry the same merged result, two batches in flight in submission order
        import torch

        xq = torch.from_numpy(q)
        h1 = sharded.submit(xq, k)
        h2 = sharded.submit(xq[:5], k)
        s1, i1 = sharded.result(h1)
        s2, i2 = sharded.result(h2)
        assert np.array_equal(i1.numpy(), I) and np.array_equal(i2.numpy(), I[:5])
        np.testing.assert_allclose(s1.numpy(), D, rtol=0, atol=0)
        if rank == 0:
            full = FlatIPOracle(d)
            full.add(x)
            Df, If = full.search(q, k)
            assert np.array_equal(I, If), "sharded ids differ from the unsharded search"
            np.testing.assert_allclose(D, Df, rtol=1e-6)
            assert I[-1, 0] == 2 and I[-1, 1] == N - 3
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(180)
def test_sharded_retrieval_two_ranks_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), 3001, 32, 9, 5, str(tmp_path)), nprocs=world, join=True)
    # every rank ends with the same merged answer
    assert np.array_equal(np.load(tmp_path / "I0.npy"), np.load(tmp_path / "I1.npy"))
    assert np.array_equal(np.load(tmp_path / "D0.npy"), np.load(tmp_path / "D1.npy"))

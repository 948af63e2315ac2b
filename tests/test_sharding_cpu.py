"""CPU: host-side logic of the data-parallel path -- contiguous sharding, fixed-shape padding
and the final gather, exercised with world_size 2 on the gloo backend."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rtpe_b200 import inference


def test_shard_range_partitions_everything():
    for total in (1, 7, 32, 256, 1001):
        for world in (1, 2, 3, 4, 8):
            spans = [inference.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_pad_results_shapes():
    ans = torch.arange(2 * 3 * 17 * 4, dtype=torch.float32).view(2, 3, 17, 4)
    count = torch.tensor([3, 1], dtype=torch.int32)
    scores = torch.rand(2, 3)
    a, c, s = inference.pad_results(ans, count, scores, 5)
    assert a.shape == (2, 5, 17, 4) and s.shape == (2, 5)
    assert torch.equal(a[:, :3], ans) and torch.all(a[:, 3:] == 0)
    a, c, s = inference.pad_results(ans, count, scores, 2)
    assert a.shape == (2, 2, 17, 4) and torch.equal(a, ans[:, :2])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, total, tmpdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = inference.shard_range(total, rank, world)
        # fake per-image decode results that encode the global image index
        n = hi - lo
        ans = torch.zeros(n, 4, 17, 4)
        count = torch.zeros(n, dtype=torch.int32)
        scores = torch.zeros(n, 4)
        for i in range(n):
            gidx = lo + i
            count[i] = gidx % 4
            ans[i, :count[i]] = float(gidx)
            scores[i, :count[i]] = float(gidx) / 100
        a, c, s = inference.pad_results(ans, count, scores, 4)
        ga, gc, gs = inference.gather_results(a, c, s)
        if rank == 0:
            np.savez(os.path.join(tmpdir, "out.npz"), a=ga.numpy(), c=gc.numpy(), s=gs.numpy())
    finally:
        dist.destroy_process_group()


def test_gather_world2_gloo(tmp_path):
    total = 6                                   # equal shards (all_gather needs equal shapes)
    port = _free_port()
    mp.spawn(_worker, args=(2, port, total, str(tmp_path)), nprocs=2, join=True)
    z = np.load(os.path.join(str(tmp_path), "out.npz"))
    assert z["a"].shape == (total, 4, 17, 4)
    for g in range(total):
        assert z["c"][g] == g % 4
        assert np.all(z["a"][g, :g % 4] == float(g))
        assert np.all(z["a"][g, g % 4:] == 0)
    res = inference.unpack_results(torch.from_numpy(z["a"]), torch.from_numpy(z["c"]),
                                   torch.from_numpy(z["s"]))
    assert [r[0].shape[0] if r[0].size else 0 for r in res] == [g % 4 for g in range(total)]


def test_pack_results_round_trip():
    ans = torch.randn(3, 5, 17, 5)
    count = torch.tensor([5, 0, 2], dtype=torch.int32)
    scores = torch.rand(3, 5)
    payload = inference.pack_results(ans, count, scores)
    assert payload.shape == (3, 1 + 5 + 5 * 17 * 5) and payload.dtype == torch.float32
    a, c, s = inference.unpack_payload(payload, 5, 17, 5)
    assert torch.equal(a, ans) and torch.equal(c, count) and torch.equal(s, scores)


def test_pin_to_gpu_numa_is_best_effort():
    before = os.sched_getaffinity(0)
    n = inference.pin_to_gpu_numa(0)            # no GPU here: must not raise, must not change anything
    assert n == 0 and os.sched_getaffinity(0) == before

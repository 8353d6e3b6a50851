"""World-size-2 test of the batch-sharding host logic on CPU (gloo): contiguous shards cover the
batch, variable-shape outputs gather back in global order, timing is the max over ranks."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from wav2vec_s_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_utts, D = 7, 5
        lens = [9, 4, 7, 3, 8, 2, 6]                       # frames per utterance of the global batch
        lo, hi = sharding.shard_bounds(n_utts, world, rank)
        T_r = max(lens[lo:hi])
        x = torch.zeros(hi - lo, T_r, D)
        for j, i in enumerate(range(lo, hi)):              # "encoder output" = utterance id, frame id
            x[j, : lens[i]] = (100 * i + torch.arange(lens[i]).float()).unsqueeze(1)
        outs, counts = sharding.gather_outputs(x, torch.tensor(lens[lo:hi]))
        assert counts.tolist() == lens
        for i, o in enumerate(outs):
            assert o.shape == (lens[i], D)
            assert torch.equal(o[:, 0], 100 * i + torch.arange(lens[i]).float())
        t = sharding.max_over_ranks(1.0 + rank)
        assert t == float(world)
        only = sharding.gather_outputs(x, torch.tensor(lens[lo:hi]), dst=0)
        assert (only is None) == (rank != 0)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_batch():
    for n in (1, 7, 64, 129):
        for w in (1, 2, 4, 8):
            b = [sharding.shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def test_balanced_order_is_a_partition():
    lens = [1499, 300, 800, 1200, 50, 999, 640, 77, 1300]
    parts = sharding.balanced_order(lens, 4)
    assert sorted(i for p in parts for i in p) == list(range(len(lens)))
    cost = [sum(lens[i] + lens[i] ** 2 / 4096.0 for i in p) for p in parts]
    assert max(cost) < 1.5 * (sum(cost) / 4)


def test_gather_and_timing_world2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert res == {0: "ok", 1: "ok"}, res

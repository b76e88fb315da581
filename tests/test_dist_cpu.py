"""world_size-2 gloo test of the multi-GPU host logic: batch sharding + the all-gather of decoded frames."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from upgpt_b200.distributed import gather_frames, per_sample_noise, shard_batch, shard_range


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, n_total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    frames_all = (torch.arange(n_total * 2 * 3 * 3) % 251).to(torch.uint8).reshape(n_total, 2, 3, 3)
    cond = {"c_crossattn": torch.arange(n_total * 4.).reshape(n_total, 2, 2), "c_concat": [torch.arange(n_total * 1.).reshape(n_total, 1)]}
    mine = shard_batch(cond)
    lo, hi = shard_range(n_total, rank, world)
    ok = torch.equal(mine["c_crossattn"], cond["c_crossattn"][lo:hi]) and torch.equal(mine["c_concat"][0], cond["c_concat"][0][lo:hi])
    out = gather_frames(frames_all[lo:hi].clone())
    ok = ok and torch.equal(out, frames_all)
    sizes = [shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world)]
    ok = ok and torch.equal(gather_frames(frames_all[lo:hi].clone(), sizes=sizes), frames_all)     # known shard sizes: one collective
    # per-sample noise: what rank r draws for its shard equals the corresponding slice of a single-process draw
    whole = per_sample_noise((4, 2, 2), range(n_total), seed=5, steps=3)
    mine_n = per_sample_noise((4, 2, 2), range(lo, hi), seed=5, steps=3)
    ok = ok and torch.equal(mine_n, whole[:, lo:hi]) and tuple(whole.shape) == (3, n_total, 4, 2, 2)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def _run(n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def test_shard_and_gather_even():
    _run(8)


def test_shard_and_gather_uneven():
    _run(7)


def test_shard_range_partition():
    for n in (1, 7, 8, 64):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1

"""world_size-2 gloo tests (CPU) of the multi-GPU plumbing: time sharding, the MAX all-reduce of
the flat-field maxima and the all-gather of the per-marker summaries."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from magnify_b200 import dist as mdist
from oracle import flatfield as o_ff


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, num_times, tmpdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)     # same data on every rank; each takes its own shard
        tiles = rng.integers(0, 4000, (2, num_times, 1, 2, 8, 8), dtype=np.uint16)
        flat = 0.8 + 0.4 * rng.random((8, 8))
        dark = 50.0
        a, b = mdist.shard_timepoints(num_times, rank, world)
        local = tiles[:, a:b]
        # pass 1 on the shard, MAX all-reduce, pass 2 on the shard == the global computation
        maxima = torch.tensor(o_ff.flatfield_maxima(local, flat, dark), dtype=torch.float64)
        mdist.allreduce_maxima(maxima)
        want_max = o_ff.flatfield_maxima(tiles, flat, dark)
        assert tuple(maxima.tolist()) == want_max
        got = o_ff.flatfield_correct(local, flat, dark, maxima=tuple(maxima.tolist()))
        want = o_ff.flatfield_correct(tiles, flat, dark)[:, a:b]
        assert np.array_equal(got, want)
        # summaries: (M, C, T_local, K) gathered along time, uneven shards
        m, c, k = 3, 2, 6
        full = torch.arange(m * c * num_times * k, dtype=torch.float64).reshape(m, c, num_times, k)
        out = mdist.gather_summaries(full[:, :, a:b].contiguous(), num_times)
        assert torch.equal(out, full)
        centres = torch.arange(12, dtype=torch.float64).reshape(4, 3) if rank == 0 else None
        centres = mdist.broadcast_centres(centres, src=0)
        assert torch.equal(centres, torch.arange(12, dtype=torch.float64).reshape(4, 3))
        open(os.path.join(tmpdir, f"ok{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("num_times", [5, 8])
def test_two_rank_time_sharding_gloo(tmp_path, num_times):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), num_times, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_shard_timepoints_cover_and_balance():
    for t in (1, 5, 50, 100, 7):
        for w in (1, 2, 4, 8):
            blocks = [mdist.shard_timepoints(t, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == t
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1 and sizes == mdist.shard_sizes(t, w)
    with pytest.raises(ValueError):
        mdist.shard_timepoints(5, 2, 2)

"""Two-rank NCCL test of the multi-GPU path on real GPUs (skipped on a single-GPU box): time
sharding with the MAX all-reduce of the flat-field maxima, and the summaries all-gathered by the
gather kernel itself through peer (symmetric) memory vs the NCCL all-gather."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, tmpdir):
    import torch.distributed as dist

    from magnify_b200 import dist as mdist, pipeline, synth

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = torch.device(f"cuda:{rank}")
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        # the same 6-timepoint stack on every rank; each rank processes its block of timepoints
        case = synth.chip_case(c=2, t=6, r=2, cc=4, h=256, w=256, overlap=22, rows=3, cols=3, row_dist=126.1,
                               col_dist=250.0, seed=7, device=dev)
        a, b = mdist.shard_timepoints(6, rank, world)
        tiles = case.tiles[:, a:b].contiguous()
        plan = pipeline.QuantifyPlan(tiles.shape, case.overlap, case.roi_length, case.flat, case.dark, device=dev,
                                     group=dist.group.WORLD)
        plan.set_chip_markers(case.x[:, a:b], case.y[:, a:b], case.fg_radius, case.chamber_radius,
                              case.max_button_radius)
        res = plan.run_device(tiles)
        # single-GPU reference on the whole stack (global maxima) computed redundantly on each rank
        full = pipeline.QuantifyPlan(case.tiles.shape, case.overlap, case.roi_length, case.flat, case.dark, device=dev)
        full.set_chip_markers(case.x, case.y, case.fg_radius, case.chamber_radius, case.max_button_radius)
        ref = full.run_device(case.tiles)
        assert torch.equal(res.maxima, ref.maxima)
        assert torch.equal(res.image.view(torch.int16), ref.image[:, a:b].contiguous().view(torch.int16))
        assert torch.equal(res.roi.view(torch.int16), ref.roi[:, :, a:b].contiguous().view(torch.int16))
        gathered = mdist.gather_summaries(res.stats, 6)
        assert torch.equal(torch.nan_to_num(gathered, nan=-1.0), torch.nan_to_num(ref.stats, nan=-1.0))
        try:
            symm = mdist.SymmetricSummaries(res.stats.shape, dev)
        except Exception as exc:   # symmetric memory not available on this box: NCCL path is the product
            open(os.path.join(tmpdir, f"nosymm{rank}"), "w").write(repr(exc))
        else:
            plan.run_device(tiles, peer_stats=symm.peer_blocks)
            symm.barrier()
            torch.cuda.synchronize(dev)
            want = torch.stack([ref.stats[:, :, x0:x1] for x0, x1 in
                                (mdist.shard_timepoints(6, r, world) for r in range(world))])
            assert torch.equal(torch.nan_to_num(symm.gathered, nan=-1.0), torch.nan_to_num(want, nan=-1.0))
        open(os.path.join(tmpdir, f"ok{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


def test_two_gpu_time_sharding_and_fused_summary_gather(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()

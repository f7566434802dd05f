"""BASELINE.json's configurations at (or near) full size: parity through size-independent
properties and independent on-device references (torch float64 eager ops, torch slicing), plus a
complete CPU-oracle replay where the oracle finishes in seconds (config 1)."""
import numpy as np
import pytest
import torch

from oracle import geometry as o_geo
from oracle import reduce as o_red
from oracle import rois as o_rois

pytestmark = pytest.mark.gpu


def eq16(a, b):
    return torch.equal(a.contiguous().view(torch.int16), b.contiguous().view(torch.int16))


def torch_flatfield(tiles, flat, dark):
    """preprocess.py:83-87 with torch float64 eager ops (IEEE sub/div/mul, no contraction)."""
    t = (tiles.to(torch.float64) - dark).clamp_min(0)
    m1 = t.max()
    t = t / flat
    m2 = t.max()
    return ((t * m1) / m2).to(torch.int64).to(torch.uint16), (float(m1), float(m2))


def test_config1_mrbles_full_oracle_replay(cuda_device):
    """Config 1: one 2048x2048 uint16 image, 9 channels, ~300 beads r in [10,25], roi_length 100
    (mrbles defaults, registry.py:281-282): everything against the CPU oracle."""
    from magnify_b200 import pipeline, synth
    from oracle import flatfield as o_ff, stitch as o_st

    case = synth.bead_case(c=9, t=1, r=1, cc=1, h=2048, w=2048, overlap=0, n_beads=300, min_radius=10,
                           max_radius=25, roi_length=100, seed=3, device=cuda_device)
    plan = pipeline.QuantifyPlan(case.tiles.shape, 0, 100, case.flat, case.dark, device=cuda_device)
    plan.set_bead_markers(case.beads)
    res = plan.run_device(case.tiles)
    tiles = case.tiles.cpu().numpy()
    image = o_st.stitch(o_ff.flatfield_correct(tiles, case.flat, case.dark), 0)
    assert np.array_equal(res.image.cpu().numpy(), image)
    labels, fg, bg = o_rois.bead_masks(case.beads, 2048, 2048, 100)
    assert np.array_equal(plan.labels.cpu().numpy(), labels)
    assert (labels == -2).any()                      # the generator plants overlapping beads
    assert np.array_equal(res.fg[:, 0].cpu().numpy().astype(bool), fg)
    assert np.array_equal(res.bg[:, 0].cpu().numpy().astype(bool), bg)
    x = case.beads[:, 1:2]
    y = case.beads[:, 0:1]
    roi = o_rois.gather_rois(image, x, y, 100)
    assert np.array_equal(res.roi.cpu().numpy(), roi)
    want = o_red.masked_stats(roi, fg[:, None], bg[:, None])
    np.testing.assert_allclose(res.stats.cpu().numpy(), want, rtol=1e-12, equal_nan=True)


def test_config4_flatfield_stitch_10x10(cuda_device):
    """Config 4 slice: one (channel, time) plane of 10x10 tiles of 2048^2, overlap 102, flat-field
    on -- against torch float64 eager arithmetic + torch slicing on the same device."""
    from magnify_b200 import ops, synth

    torch.manual_seed(4)
    shape = (1, 1, 10, 10, 2048, 2048)
    tiles = torch.randint(0, 65536, shape, dtype=torch.int32, device=cuda_device).to(torch.uint16)
    flat_np, dark_np = synth.smooth_flat_dark(2048, 2048)
    flat = torch.from_numpy(flat_np).to(cuda_device)
    dark = torch.from_numpy(dark_np).to(cuda_device)
    plan = ops.FlatFieldPlan(shape, flat_np, dark_np, device=cuda_device)
    image = ops.flatfield_stitch(tiles, overlap=102, plan=plan)
    want = torch.empty_like(tiles)
    # maxima are global: first pass over all tiles, then apply with the same scalars
    t = None
    m1 = m2 = 0.0
    for r in range(10):
        tt = (tiles[0, 0, r].to(torch.float64) - dark).clamp_min(0)
        m1 = max(m1, float(tt.max()))
        m2 = max(m2, float((tt / flat).max()))
    assert tuple(plan.maxima.cpu().tolist()) == (m1, m2)
    for r in range(10):
        tt = (tiles[0, 0, r].to(torch.float64) - dark).clamp_min(0) / flat
        want[0, 0, r] = ((tt * m1) / m2).to(torch.int64).to(torch.uint16)
    kept = want[..., 51:2048 - 51, 51:2048 - 51]
    ref = kept.permute(0, 1, 2, 4, 3, 5).reshape(1, 1, 10 * 1946, 10 * 1946)
    assert eq16(image, ref)


def test_config3_time_series_small_t(cuda_device):
    """Config 3 geometry (4 channels, 4x4 tiles of 2048^2, overlap 102, 1792 buttons, L=72,
    flat-field on) at T=2: image against torch float64, crops against slicing, sums against a
    torch reduction; copy-forward reuses the t=0 masks."""
    from magnify_b200 import pipeline, synth

    case = synth.chip_case(c=4, t=2, seed=5, device=cuda_device)
    plan = pipeline.QuantifyPlan(case.tiles.shape, case.overlap, case.roi_length, case.flat, case.dark,
                                 device=cuda_device)
    plan.set_chip_markers(case.x, case.y, case.fg_radius, case.chamber_radius, case.max_button_radius)
    res = plan.run_device(case.tiles)
    flat = torch.from_numpy(case.flat).to(cuda_device)
    dark = torch.from_numpy(case.dark).to(cuda_device)
    corrected, maxima = torch_flatfield(case.tiles, flat, dark)
    assert tuple(res.maxima.cpu().tolist()) == maxima
    kept = corrected[..., 51:2048 - 51, 51:2048 - 51]
    ref = kept.permute(0, 1, 2, 4, 3, 5).reshape(4, 2, 4 * 1946, 4 * 1946)
    assert eq16(res.image, ref)
    boxes = res.boxes.cpu().numpy()
    for m in (0, 5, 31, 32, 900, 1791):
        for t in (0, 1):
            top, left = boxes[m, t]
            assert eq16(res.roi[m, :, t], ref[:, t, top:top + 72, left:left + 72])
    assert res.fg.shape[1] == 1 and res.mask_t.cpu().tolist() == [0, 0]
    sums = (res.roi.to(torch.float64) * res.fg[:, None, :, :, :].to(torch.float64)).sum((-1, -2))
    assert torch.equal(sums, res.stats[..., 2])
    nbg = res.bg.sum((-1, -2)).to(torch.float64)[:, None].expand(-1, 4, 2)
    assert torch.equal(res.stats[..., 1], nbg)


def test_config5_bead_screen_slice(cuda_device):
    """Config 5 slice: one 20480^2 image, 100k beads, roi_length 50 (beads_pipe default): label
    raster against the CPU oracle on a window, masks/crops on a sample, label-count identities."""
    from magnify_b200 import ops

    rng = np.random.default_rng(55)
    size, n, length = 20480, 100_000, 50
    beads = np.stack([rng.integers(0, size, n), rng.integers(0, size, n), rng.integers(4, 13, n)], 1)
    beads_d = torch.from_numpy(beads.astype(np.int32)).to(cuda_device)
    labels = ops.bead_labels(beads_d, size, size)
    # window [0:1500, 0:1500]: only beads whose bounding square touches it matter there
    wsz = 1500
    near = np.where((beads[:, 0] - beads[:, 2] < wsz) & (beads[:, 1] - beads[:, 2] < wsz))[0]
    sub = o_geo.circle_labels(beads[near], wsz, wsz)
    want = np.where(sub >= 0, near[np.clip(sub, 0, None)], sub).astype(np.int32)
    got = labels[:wsz, :wsz].cpu().numpy()
    assert np.array_equal(got, want)
    image = torch.randint(0, 65536, (1, 1, size, size), dtype=torch.int32, device=cuda_device).to(torch.uint16)
    x = torch.from_numpy(beads[:, 1:2].astype(np.float64)).to(cuda_device).contiguous()
    y = torch.from_numpy(beads[:, 0:1].astype(np.float64)).to(cuda_device).contiguous()
    boxes = ops.bounding_boxes(x, y, length, size, size)
    fg, bg, counts = ops.bead_masks(labels, boxes[:, 0].contiguous(), length, want_counts=True)
    roi, stats = ops.roi_gather_stats(image, boxes, fg[:, None].contiguous(), bg[:, None].contiguous(), length)
    bx = boxes.cpu().numpy()
    for m in list(range(0, n, 9973)) + [n - 1]:
        top, left = bx[m, 0]
        assert eq16(roi[m, 0, 0], image[0, 0, top:top + length, left:left + length])
        lab = labels[top:top + length, left:left + length]
        assert torch.equal(fg[m].bool(), lab == m) and torch.equal(bg[m].bool(), lab == -1)
    # every pixel owned by exactly one bead is that bead's foreground somewhere: a bead's fg count
    # equals its label count whenever its disc fits in its box (r <= 12 < 25)
    owned = torch.bincount(labels[labels >= 0].flatten().to(torch.int64), minlength=n)
    assert torch.equal(owned.to(torch.int32), counts[:, 0])
    assert torch.equal(stats[:, 0, 0, 0].to(torch.int32), counts[:, 0])
    sums = (roi[:, 0, 0].to(torch.float64) * fg.to(torch.float64)).sum((-1, -2))
    assert torch.equal(sums, stats[:, 0, 0, 2])

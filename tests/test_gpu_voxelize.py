"""Parity of the CUDA voxeliser (through the C ABI) with the reference numba kernel's golden outputs: BIT-EXACT
coordinates, per-voxel counts and voxel payloads (integer / byte work)."""
import os

import numpy as np
import pytest
import torch

from cases import sha, voxel_cases

pytestmark = pytest.mark.gpu


def _run_device(clouds, max_points, max_voxels):
    from sessd_b200 import ops, synth
    dev = torch.device("cuda")
    cfg = ops.make_voxel_cfg(synth.VOXEL_SIZE, synth.PC_RANGE, max_points, max_voxels)
    total = sum(c.shape[0] for c in clouds)
    cap = max(total, 1)
    pts = torch.zeros((cap, 4), dtype=torch.float32, device=dev)
    if total:
        pts[:total] = torch.from_numpy(np.concatenate(clouds, 0)).to(dev)
    off = np.zeros(len(clouds) + 1, np.int32)
    off[1:] = np.cumsum([c.shape[0] for c in clouds])
    buf = ops.VoxelBuffers(cfg, len(clouds), cap, dev)
    ops.voxelize(pts, torch.from_numpy(off).to(dev), buf)
    torch.cuda.synchronize()
    nv = buf.num_voxels.cpu().numpy()
    out = []
    base = 0
    for f in range(len(clouds)):
        m = int(nv[f])
        sl = slice(base, base + m)
        out.append((buf.voxels[sl].cpu().numpy(), buf.coors[sl].cpu().numpy(), buf.num_points[sl].cpu().numpy(),
                    buf.mean[sl].cpu().numpy()))
        base += m
    assert int(nv[-1]) == base
    return out


@pytest.mark.parametrize("case", voxel_cases(), ids=lambda c: c[0])
def test_voxelize_matches_reference_golden(case, golden_dir):
    name, pts, mp, mv = case
    g = np.load(os.path.join(golden_dir, "voxel_cases.npz"))
    assert (sha(pts) == g[name + "_points_sha"]).all(), "seeded input drifted"
    (v, c, n, mean), = _run_device([pts], mp, mv)
    assert c.shape[0] == g[name + "_coors"].shape[0]
    assert (c[:, 0] == 0).all()
    assert np.array_equal(c[:, 1:], g[name + "_coors"])
    assert np.array_equal(n, g[name + "_num"])
    assert (sha(v) == g[name + "_voxels_sha"]).all()
    if name + "_voxels" in g:
        assert np.array_equal(v, g[name + "_voxels"])
    if v.shape[0]:
        ref_mean = v.sum(1) / n[:, None].astype(np.float32)
        np.testing.assert_allclose(mean, ref_mean, rtol=2e-7, atol=0)


def test_voxelize_batched_frames_are_independent(golden_dir):
    """Frames voxelised together (collate_kitti wire format) equal the frames voxelised alone."""
    cs = [c for c in voxel_cases() if c[2] == 5 and c[3] == 20000]
    clouds = [c[1] for c in cs]
    g = np.load(os.path.join(golden_dir, "voxel_cases.npz"))
    res = _run_device(clouds, 5, 20000)
    for f, (case, (v, c, n, _m)) in enumerate(zip(cs, res)):
        name = case[0]
        assert (c[:, 0] == f).all()
        assert np.array_equal(c[:, 1:], g[name + "_coors"])
        assert np.array_equal(n, g[name + "_num"])
        assert (sha(v) == g[name + "_voxels_sha"]).all()


def test_voxelize_host_api_matches_golden(golden_dir):
    from sessd_b200 import ops, synth
    g = np.load(os.path.join(golden_dir, "voxel_cases.npz"))
    for name, pts, mp, mv in voxel_cases():
        cfg = ops.make_voxel_cfg(synth.VOXEL_SIZE, synth.PC_RANGE, mp, mv)
        v, c, n = ops.voxelize_host(pts, cfg)
        assert np.array_equal(c, g[name + "_coors"])
        assert np.array_equal(n, g[name + "_num"])
        assert (sha(v) == g[name + "_voxels_sha"]).all()


def test_vfe_mean_matches_reference_golden(golden_dir):
    from sessd_b200 import synth
    g = np.load(os.path.join(golden_dir, "vfe_case.npz"))
    (v, c, n, mean), = _run_device([synth.uniform_cloud(1, 2000)], 5, 20000)
    np.testing.assert_allclose(mean, g["mean"], rtol=2e-7, atol=0)


def test_voxelize_full_size_properties():
    """Stress shape (200k points, 4 frames, max_voxels 200000): size-independent invariants + oracle equality."""
    from oracle import cpu as ocpu
    from sessd_b200 import synth
    clouds = [synth.uniform_cloud(100 + f, 200000) for f in range(3)] + [synth.ring_cloud(7, 20000)]
    res = _run_device(clouds, 5, 200000)
    for pts, (v, c, n, _m) in zip(clouds, res):
        ov, oc, on = ocpu.points_to_voxel(pts, synth.VOXEL_SIZE, synth.PC_RANGE, 5, 200000)
        assert np.array_equal(c[:, 1:], oc) and np.array_equal(n, on) and np.array_equal(v, ov)
        # every kept point lies in its voxel; voxels are unique
        lin = (c[:, 1].astype(np.int64) * 1600 + c[:, 2]) * 1408 + c[:, 3]
        assert len(np.unique(lin)) == len(lin)
        assert n.min() >= 1 and n.max() <= 5

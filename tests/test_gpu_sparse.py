"""Rulebook (bit-exact vs the CPU restatement, canonical order) and sparse conv features (<= 1e-4 rel vs an fp64 oracle)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _voxels(cloud):
    from oracle import cpu as ocpu
    from sessd_b200 import synth
    v, c, n = ocpu.points_to_voxel(cloud, synth.VOXEL_SIZE, synth.PC_RANGE, 5, 20000)
    return v, np.concatenate([np.zeros((len(c), 1), np.int32), c], 1).astype(np.int32), n


def _two_frame_coors():
    from sessd_b200 import synth
    _, c0, _ = _voxels(synth.ring_cloud(3, 6000))
    _, c1, _ = _voxels(synth.uniform_cloud(4, 3000))
    c1[:, 0] = 1
    return np.concatenate([c0, c1], 0)


def test_subm_rulebook_hash_and_pairs_bit_exact():
    from oracle import spconv_ref as S
    from sessd_b200 import ops
    coors = _two_frame_coors()
    n = len(coors)
    shape = (41, 1600, 1408)
    grid = ops.make_grid(2, shape)
    cap = n + 100
    d_coors = torch.zeros((cap, 4), dtype=torch.int32, device="cuda")
    d_coors[:n] = torch.from_numpy(coors).cuda()
    d_n = torch.tensor([n], dtype=torch.int32, device="cuda")
    table = ops.hash_build(d_coors, d_n, cap, grid)
    nbr = ops.subm_rulebook(d_coors, d_n, cap, grid, (3, 3, 3), 0, table)
    ref = S.neighbor_table(coors, shape, coors, (3, 3, 3), (1, 1, 1), (1, 1, 1))
    got = nbr[:n].cpu().numpy()
    assert np.array_equal(got, ref)
    # canonical spconv pairs: per offset sorted by output index
    pin, pout, num = ops.rulebook_pairs(nbr, d_n, cap, 27)
    pairs = S.pairs_from_nbr(ref)
    num = num.cpu().numpy()
    for k in range(27):
        assert num[k] == len(pairs[k][0])
        assert np.array_equal(pin[k, : num[k]].cpu().numpy(), pairs[k][0])
        assert np.array_equal(pout[k, : num[k]].cpu().numpy(), pairs[k][1])
    assert int(num[13]) == n   # centre offset: every voxel pairs with itself


def test_strided_chain_rulebooks_bit_exact():
    """The 4 strided + 3 SubM(bitmap-indexed) rulebooks of SpMiddleFHD on a 2-frame batch."""
    from oracle import spconv_ref as S
    from sessd_b200 import ops
    from sessd_b200.runners import SpMiddleRunner
    coors = _two_frame_coors()
    n = len(coors)
    r = SpMiddleRunner(2, n + 50, device="cuda")
    layers, _, _ = _weights()
    r.load_weights(layers)
    d_coors = torch.zeros((n + 50, 4), dtype=torch.int32, device="cuda")
    d_coors[:n] = torch.from_numpy(coors).cuda()
    feat = torch.randn((n + 50, 4), device="cuda")
    r.forward(feat, d_coors, torch.tensor([n], dtype=torch.int32, device="cuda"))
    torch.cuda.synchronize()
    assert int(r.status.item()) == 0
    cur, shape = coors, (41, 1600, 1408)
    lvl = 0
    seen = set()
    for p in r.plan:
        if p["kind"] == "subm":
            if p["key"] in seen:
                continue
            seen.add(p["key"])
            ref = S.neighbor_table(cur, shape, cur, p["ks"], (1, 1, 1), (1, 1, 1))
            assert np.array_equal(p["nbr"][: len(cur)].cpu().numpy(), ref), p["key"]
        else:
            oc, oshape = S.strided_out_coors(cur, shape, p["ks"], p["st"], p["pd"])
            lv = r.levels[lvl + 1]
            assert int(lv["n"].item()) == len(oc)
            assert np.array_equal(lv["coors"][: len(oc)].cpu().numpy(), oc)
            ref = S.neighbor_table(cur, shape, oc, p["ks"], p["st"], p["pd"])
            assert np.array_equal(p["nbr"][: len(oc)].cpu().numpy(), ref)
            cur, shape, lvl = oc, oshape, lvl + 1
    assert shape == (2, 200, 176)


def _weights(seed=3):
    from sessd_b200 import weights
    sd = weights.random_detector_state(seed)
    return weights.split_detector_state(sd)


@pytest.mark.parametrize("use_tc,split,rows", [(False, None, 0), (True, "tf32", 0), (True, "fp16", 0), (True, "fp16", 32), (True, "fp16", 16)],
                         ids=["simt", "tcgen05-3xtf32", "tma-gather-fp16x2", "rows32+tma-gather", "rows16+tma-gather(default)"])
def test_spmiddle_features_match_fp64_oracle(use_tc, split, rows):
    from oracle import spconv_ref as S
    from sessd_b200 import synth
    from sessd_b200.runners import SpMiddleRunner
    v, coors, num = _voxels(synth.ring_cloud(5, 8000))
    n = len(coors)
    feat = (v.sum(1) / num[:, None]).astype(np.float32)
    layers, _, _ = _weights()
    params = [dict(weight=l["weight"].numpy(), gamma=l["gamma"].numpy(), beta=l["beta"].numpy(), mean=l["mean"].numpy(),
                   var=l["var"].numpy()) for l in layers]
    trace = []
    ref = S.spmiddle_forward(feat, coors, 1, (1408, 1600, 40), params, np.float64, trace)   # [1,128,200,176]
    r = SpMiddleRunner(1, n, device="cuda", use_tc=use_tc, split=split, rows_max_cin=rows)
    assert [p["impl"] for p in r.plan].count("rows") == {0: 0, 16: 3, 32: 5}[rows]
    r.load_weights(layers)
    dense = r.forward(torch.from_numpy(feat).cuda(), torch.from_numpy(coors).cuda(), torch.tensor([n], dtype=torch.int32, device="cuda"))
    torch.cuda.synchronize()
    for li, t in enumerate(trace):
        got = r.feats[li][: len(t["coors"])].cpu().numpy().astype(np.float64)
        scale = np.abs(t["feat"]).max() + 1e-30
        assert np.abs(got - t["feat"]).max() / scale < 1e-5, "layer %d" % li
    got = dense.permute(0, 3, 1, 2).cpu().numpy().astype(np.float64)     # NHWC storage -> logical NCHW
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-5
    assert ((got != 0) == (ref != 0)).all()

"""BEV neck (SSFA) + head kernels vs the reference modules' golden output (small map) and the torch-CPU oracle."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 1e-4   # north_star: <= 1e-4 rel on regressions / confidences


def _rel(a, b):
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def test_ssfa_and_head_match_reference_golden(golden_dir):
    from oracle import bev_ref
    from sessd_b200.runners import SSFARunner
    g = np.load(os.path.join(golden_dir, "ssfa_head_case.npz"))
    sd = bev_ref.ssfa_random_state(7)
    hsd = bev_ref.head_random_state(9, prefix="tasks.0.")
    gen = torch.Generator().manual_seed(8)
    x = torch.relu(torch.randn(1, 128, 24, 16, generator=gen))
    r = SSFARunner(1, (24, 16), "cuda")
    r.load_state(sd, hsd)
    out, head = r.forward(x.permute(0, 2, 3, 1).contiguous().cuda())
    torch.cuda.synchronize()
    got = out.permute(0, 3, 1, 2).cpu().numpy()
    assert _rel(got, g["ssfa_out"]) < TOL
    # the golden head used prefix "" weights drawn from the same seed
    hsd0 = bev_ref.head_random_state(9, prefix="")
    for k in hsd0:
        assert torch.equal(hsd0[k], hsd["tasks.0." + k])
    h = head.cpu().numpy()
    assert _rel(h[..., 0:14], g["box_preds"]) < TOL
    assert _rel(h[..., 14:16], g["cls_preds"]) < TOL
    assert _rel(h[..., 16:20], g["dir_cls_preds"]) < TOL
    assert _rel(h[..., 20:22], g["iou_preds"]) < TOL


def test_ssfa_intermediates_match_oracle_fp64_batch2():
    from oracle import bev_ref
    from sessd_b200.runners import SSFARunner
    sd = bev_ref.ssfa_random_state(17)
    hsd = bev_ref.head_random_state(19)
    gen = torch.Generator().manual_seed(18)
    x = torch.relu(torch.randn(2, 128, 40, 48, generator=gen))
    trace = {}
    ref = bev_ref.ssfa_forward(x.double(), {k: v.double() if v.is_floating_point() else v for k, v in sd.items()}, trace)
    r = SSFARunner(2, (40, 48), "cuda")
    r.load_state(sd, hsd)
    out, _ = r.forward(x.permute(0, 2, 3, 1).contiguous().cuda())
    torch.cuda.synchronize()
    for mine, theirs in (("x0", "x0"), ("x1", "x1"), ("t0", "t0"), ("t1", "t1"), ("m0", "m0"), ("m1", "m1"), ("o0", "o0"), ("o1", "o1")):
        got = r.buf[mine].permute(0, 3, 1, 2).cpu().numpy()
        assert _rel(got, trace[theirs].numpy()) < 2e-5, mine
    assert _rel(out.permute(0, 3, 1, 2).cpu().numpy(), ref.numpy()) < 2e-5

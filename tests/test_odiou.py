"""ODIoU loss (csrc/odiou.cuh, odiou.cu).

CPU part: the kernel's arithmetic, run on the host through sessd_odiou_pairs_host, against the REFERENCE's own odiou_3D
(det3d/models/losses/odious.py, imported where it lies by tests/golden/make_golden.py `odiou`; value + gradient through its custom autograd
Functions).  Bars: value <= 2e-4 abs (the reference accumulates in float32 numpy loops); gradient w.r.t. x, y, z, w, l, h <= 2e-3 of the
pair's largest gradient component, w.r.t. the yaw <= 3 % (the reference's hand-written Jacobians and the piecewise minimum-bounding-rectangle
term agree with forward-mode differentiation of the same value only to that level; finite differences sit between the two).
Degenerate constructions (identical z / h / BEV rectangles) hit ties in min / max / hull selection where the two implementations pick
different sub-gradients; they are compared by value only, and the exactly-identical pair is excluded: the reference's polygon routine
returns IoU 1/3 for two identical boxes (odiou 0.667), the clipped polygon gives IoU 1 (odiou 0)."""
import os

import numpy as np
import pytest

from cases import odiou_pairs


def test_odiou_host_arithmetic_matches_reference_golden(golden_dir):
    from sessd_b200 import ops
    ref = np.load(os.path.join(golden_dir, "odiou_case.npz"))
    g, q = odiou_pairs()
    v, gr = ops.odiou_pairs_host(g, q)
    n = 48                                            # generic pairs
    np.testing.assert_allclose(v[:n], ref["odiou"][:n], atol=2e-4, rtol=0)
    scale = np.abs(ref["grad_q"][:n]).max(1, keepdims=True) + 1e-6
    rel = np.abs(gr[:n] - ref["grad_q"][:n]) / scale
    assert rel[:, :6].max() < 2e-3, rel[:, :6].max()
    assert rel[:, 6].max() < 3e-2, rel[:, 6].max()
    for i in (48, 49, 51, 52, 53):                    # disjoint, contained, perpendicular, no height overlap, far-rotated: values
        assert abs(v[i] - ref["odiou"][i]) < 2e-4, (i, v[i], ref["odiou"][i])
    assert abs(v[50]) < 1e-5 and abs(ref["odiou"][50] - 2.0 / 3.0) < 1e-4      # identical boxes: see the module docstring
    # properties: value in [0, 1 + 1 + 1.25], zero gradient w.r.t. nothing else than the 7 parameters, finite
    assert np.isfinite(v).all() and np.isfinite(gr).all() and v.min() > -1e-5 and v.max() < 3.25


def test_odiou_gradient_is_consistent_with_finite_differences():
    """Central differences of the value (float32, h = 2e-3) vs the forward-mode gradient on the well-conditioned components (x, y, z, w, l, h)."""
    from sessd_b200 import ops
    g, q = odiou_pairs()
    g, q = g[16:40], q[16:40]                          # noise scales 0.1 / 0.4: away from the exact-alignment kinks
    _, gr = ops.odiou_pairs_host(g, q)
    h = 2e-3
    for j in range(6):
        qp, qm = q.copy(), q.copy()
        qp[:, j] += h
        qm[:, j] -= h
        fd = (ops.odiou_pairs_host(g, qp, False)[0] - ops.odiou_pairs_host(g, qm, False)[0]) / (2 * h)
        err = np.abs(fd - gr[:, j]) / (np.abs(gr).max(1) + 1e-3)
        assert np.median(err) < 2e-2 and (err < 0.15).mean() > 0.9, (j, np.median(err), err.max())


@pytest.mark.gpu
def test_odiou_device_kernel_matches_host_arithmetic():
    import torch
    from cases import head_loss_case
    from oracle import bev_ref
    from sessd_b200 import ops
    head, anc, labels, targets = head_loss_case()
    head = head.copy()
    head[..., :14] *= 0.2
    d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()   # noqa: E731
    losses, grad = ops.head_loss(d(head), d(anc), d(labels), d(targets), w_loc=0.0)
    base = grad.clone()
    sums = ops.odiou_loss(d(head), d(anc), d(labels), d(targets), losses, grad, w_odiou=2.0)
    torch.cuda.synchronize()
    B = 2
    for b in range(B):
        pos = np.nonzero(labels[b] > 0)[0]
        hv = torch.from_numpy(head[b].reshape(-1, 24))
        enc = torch.stack([hv[a // 2, 7 * (a % 2):7 * (a % 2) + 7] for a in pos])
        A = torch.from_numpy(anc[pos])
        qb = bev_ref.box_decode(enc, A).numpy()
        gb = bev_ref.box_decode(torch.from_numpy(targets[b][pos]), A).numpy()
        v, gq = ops.odiou_pairs_host(gb, qb)
        np.testing.assert_allclose(float(sums[b]), v.sum() / max(len(pos), 1), rtol=2e-4)
        diag = np.sqrt(anc[pos, 4] ** 2 + anc[pos, 3] ** 2)
        jac = np.stack([diag, diag, anc[pos, 5], qb[:, 3], qb[:, 4], qb[:, 5], np.ones(len(pos), np.float32)], 1)
        expect = gq * jac * (2.0 / B / max(len(pos), 1))
        delta = (grad - base)[b].cpu().numpy().reshape(-1, 24)
        got = np.stack([delta[a // 2, 7 * (a % 2):7 * (a % 2) + 7] for a in pos])
        np.testing.assert_allclose(got, expect, rtol=2e-3, atol=1e-6)

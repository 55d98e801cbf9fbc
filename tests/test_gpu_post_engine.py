"""Device post-processing and the whole-frame engine vs the CPU oracle (predict glue of mg_head_sessd.py:893-1057)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _synthetic_head(seed, batch=1, n_pos=600, hw=(200, 176)):
    """Head output [B,H,W,24] that yields ~n_pos candidates over threshold with overlapping boxes."""
    rng = np.random.default_rng(seed)
    h = np.zeros((batch, hw[0], hw[1], 24), np.float32)
    h[..., 0:14] = rng.normal(0, 0.25, h[..., 0:14].shape)
    h[..., 14:16] = rng.normal(-4.0, 0.5, h[..., 14:16].shape)
    h[..., 16:20] = rng.normal(0, 1, h[..., 16:20].shape)
    h[..., 20:22] = rng.uniform(-0.5, 1.0, h[..., 20:22].shape)
    for b in range(batch):
        ys = rng.integers(0, hw[0], n_pos)
        xs = rng.integers(0, hw[1], n_pos)
        rs = rng.integers(0, 2, n_pos)
        h[b, ys, xs, 14 + rs] = rng.uniform(-0.8, 4.0, n_pos)   # sigmoid in (0.31, 0.98)
    return h


def _oracle_frame(h, anchors, **kw):
    from oracle import bev_ref
    flat = torch.from_numpy(h.reshape(-1, 24))
    enc = flat[:, 0:14].reshape(-1, 7)
    cls = flat[:, 14:16].reshape(-1)
    dirs = flat[:, 16:20].reshape(-1, 2)
    iou = flat[:, 20:22].reshape(-1)
    return bev_ref.predict_frame(enc, cls, dirs, iou, torch.from_numpy(anchors), return_aux=True, **kw)


@pytest.mark.parametrize("seed,n_pos", [(0, 0), (1, 40), (2, 600), (3, 3000)])
def test_postprocess_matches_oracle(seed, n_pos):
    from sessd_b200 import ops, weights
    anchors = weights.kitti_car_anchors()
    batch = 2
    h = _synthetic_head(seed, batch, n_pos)
    cfg = ops.make_post_cfg(batch=batch, head_stride=24)
    buf = ops.PostBuffers(cfg, "cuda")
    ops.postprocess(torch.from_numpy(h).cuda(), torch.from_numpy(anchors).cuda(), None, buf)
    torch.cuda.synchronize()
    for b in range(batch):
        boxes, scores, labels, aux = _oracle_frame(h[b], anchors)
        k = int(buf.count[b].item())
        assert int(buf.aux[b, 0].item()) == aux["n_candidates"]
        if "nms_selected_anchor" in aux:
            sel = aux["nms_selected_anchor"].numpy()
            got_sel = buf.sel_anchor[b, : len(sel)].cpu().numpy()
            assert np.array_equal(got_sel, sel), "NMS keep set differs"
        assert k == boxes.shape[0]
        np.testing.assert_allclose(buf.boxes[b, :k].cpu().numpy(), boxes.numpy(), rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(buf.scores[b, :k].cpu().numpy(), scores.numpy(), rtol=1e-4, atol=1e-7)


def test_postprocess_frustum_and_gt_mode():
    from oracle import bev_ref
    from sessd_b200 import ops, weights
    anchors = weights.kitti_car_anchors()
    h = _synthetic_head(5, 1, 500)
    # a synthetic convex frustum: the half space x < 35 plus five far-away planes (normals point outwards: sign<0 inside)
    planes = np.array([[1, 0, 0, -35.0], [-1, 0, 0, -1.0], [0, 1, 0, -100.0], [0, -1, 0, -100.0], [0, 0, 1, -50.0], [0, 0, -1, -50.0]],
                      np.float32)[None]
    cfg = ops.make_post_cfg(batch=1, head_stride=24, use_frustum=True)
    buf = ops.PostBuffers(cfg, "cuda")
    ops.postprocess(torch.from_numpy(h).cuda(), torch.from_numpy(anchors).cuda(), torch.from_numpy(planes).cuda(), buf)
    torch.cuda.synchronize()
    boxes, scores, _l, _aux = _oracle_frame(h[0], anchors)
    keep = boxes[:, 0] < 35.0
    k = int(buf.count[0].item())
    assert k == int(keep.sum())
    np.testing.assert_allclose(buf.boxes[0, :k].cpu().numpy(), boxes[keep].numpy(), rtol=1e-4, atol=1e-5)


def _check_engine_vs_oracle(kind, clouds, batch):
    from oracle import frame as oframe
    from sessd_b200 import weights
    from sessd_b200.engine import FrameEngine
    anchors = weights.kitti_car_anchors()
    layers, ssfa, head = weights.bench_detector_state(kind, 0)
    lnp = oframe.layers_to_numpy(layers)
    eng = FrameEngine(batch=batch, max_points_per_frame=max(c.shape[0] for c in clouds))
    eng.load_weights(layers, ssfa, head, anchors)
    eager = eng.infer(clouds)
    eng.capture()
    graph = eng.infer(clouds)
    graph2 = eng.infer(clouds[::-1])[::-1]
    head_gpu = eng.neck.buf["head"].cpu().numpy()      # holds the reversed batch now
    for f, cloud in enumerate(clouds):
        hd = oframe.frame_head(cloud, lnp, ssfa, head)
        boxes, scores, _labels, aux = oframe.frame_detections(cloud, lnp, ssfa, head, anchors)
        assert aux["n_candidates"] > 100 and boxes.shape[0] > 5, "vacuous test: no detections"
        sc = np.sort(scores.numpy().astype(np.float64))
        assert np.min(np.diff(sc) / sc[1:]) > 1e-6, "tie-degenerate workload"
        for res in (eager[f], graph[f], graph2[f]):
            assert res["num_candidates"] == aux["n_candidates"]
            assert np.array_equal(res["anchor_index"], aux["final_anchor"].numpy()), "kept detections differ (matched by anchor index)"
            np.testing.assert_allclose(res["box3d_lidar"], boxes.numpy(), rtol=1e-4, atol=1e-4)
            np.testing.assert_allclose(res["scores"], scores.numpy(), rtol=1e-4, atol=1e-6)
        # raw head maps within 1e-4 relative
        hg = head_gpu[batch - 1 - f]
        for sl, key in ((slice(0, 14), "box_preds"), (slice(14, 16), "cls_preds"), (slice(16, 20), "dir_cls_preds"), (slice(20, 22), "iou_preds")):
            ref = hd[key][0].numpy()
            assert np.abs(hg[..., sl] - ref).max() / np.abs(ref).max() < 1e-4, key


def test_engine_end_to_end_matches_cpu_oracle():
    """Whole path (eager, CUDA graph, graph with the batch reversed) on the bench workload's weights: ring clouds, batch 2."""
    from sessd_b200 import synth
    _check_engine_vs_oracle("ring", [synth.ring_cloud(21, 20000), synth.ring_cloud(22, 18000)], 2)


def test_engine_end_to_end_uniform20k_matches_cpu_oracle():
    """Same on the uniform-20k input (SURVEY 8(d) primary input), batch 1."""
    from sessd_b200 import synth
    _check_engine_vs_oracle("uniform", [synth.uniform_cloud(3, 20000)], 1)

"""Teacher / student bookkeeping of the SE-SSD trainer (reference: det3d/torchie/trainer/trainer_sessd.py:250-275, 302-318).  Only the
pieces that sit on the per-step path are mirrored: the consistency-weight ramp and the teacher's exponential moving average (one kernel
launch over the flat parameter arenas, sessd_b200.train)."""
import numpy as np

from sessd_b200 import train as _train


def sigmoid_rampup(current_epoch, max_epochs=1, rampup_length=15.0):
    """consistency-weight ramp exp(-5 (1 - t)^2), t = clip(epoch, 0, 15) / 15 (trainer_sessd.py:302-309); 1.0 when max_epochs == 0"""
    if max_epochs == 0:
        return 1.0
    phase = 1.0 - float(np.clip(current_epoch, 0.0, rampup_length)) / rampup_length
    return float(np.exp(-5.0 * phase * phase))


def update_ema_variables(model, ema_model, global_step):
    """ema = alpha ema + (1 - alpha) param with alpha = min(1 - 1 / (global_step + 1), 0.999) (trainer_sessd.py:315-318); ``model`` and
    ``ema_model`` are ParamArenas (or modules carrying one as ``._arena``)."""
    a = getattr(model, "_arena", model)
    b = getattr(ema_model, "_arena", ema_model)
    return _train.update_ema_variables(a, b, global_step)


# ---------------------------------------------------------------------------------------------------------------- per-step host logic
_LIST_KEYS = frozenset(b + s for b in ("anchors", "anchors_mask", "reg_targets", "reg_weights", "labels") for s in ("", "_raw"))
_TENSOR_KEYS = frozenset(b + s for b in ("voxels", "bev_map", "coordinates", "num_points", "points", "num_voxels") for s in ("", "_raw"))


def example_to_device(example, device, non_blocking=False):
    """Move one collated batch to ``device`` with the reference's key classes (trainer_sessd.py:20-38): per-task lists of tensors, plain
    tensors, the ``calib`` dict of tensors; everything else (metadata, shapes, transformation dicts) is passed through untouched."""
    import torch
    out = {}
    for key, val in example.items():
        if key in _LIST_KEYS:
            out[key] = [t.to(device, non_blocking=non_blocking) for t in val]
        elif key in _TENSOR_KEYS:
            out[key] = val.to(device, non_blocking=non_blocking)
        elif key == "calib":
            out[key] = {k: (v.to(device, non_blocking=non_blocking) if isinstance(v, torch.Tensor) else v) for k, v in val.items()}
        else:
            out[key] = val
    return out


def parse_second_losses(losses):
    """key -> [per-task value] dict of ``MultiGroupHead.loss``  ->  (total loss tensor, log_vars of python numbers) with the reference's
    three value shapes (trainer_sessd.py:40-51): per-code lists (``loc_loss_elem*``), the one-element consistency tensors, scalars."""
    from collections import OrderedDict
    num = lambda v: float(v.detach()) if hasattr(v, "detach") else float(v)            # noqa: E731
    log_vars = OrderedDict()
    for name, values in losses.items():
        if name in ("loc_loss_elem", "loc_loss_elem_ema"):
            log_vars[name] = [[num(e) for e in per_task] for per_task in values]
        elif name in ("consistency_loss", "consistency_loss_ema"):
            log_vars[name] = [num(e) for e in values[0].detach().cpu().reshape(-1)]
        else:
            log_vars[name] = [num(v) for v in values]
    return sum(losses["loss"]), log_vars


def batch_processor_inline(model, model_ema, data, consistency_weight, train_mode, device=None, hooks=None):
    """One student / teacher step on a collated batch (trainer_sessd.py:250-275): teacher forward on the raw copy of the frames, student
    forward with the teacher's predictions, ``loss += consistency_weight * consistency_loss``, losses flattened for logging.  ``hooks``
    (optional callable taking the reference's hook names) stands in for ``Trainer.call_hook``.  Returns the reference's
    ``dict(loss, log_vars, num_samples)`` in train mode, the detections otherwise."""
    import torch
    call = hooks if hooks is not None else (lambda name: None)
    example = example_to_device(data, torch.cuda.current_device() if device is None else device, non_blocking=False)
    call("after_data_to_device")
    if not train_mode:
        return model(example, return_loss=False)
    preds_ema = model_ema(example, is_ema=[True, None])
    losses = model(example, is_ema=[False, preds_ema], return_loss=True)
    losses["loss"][0] = losses["loss"][0] + losses["consistency_loss"][0][0] * consistency_weight
    call("after_forward")
    loss, log_vars = parse_second_losses(losses)
    call("after_parse_loss")
    return dict(loss=loss, log_vars=log_vars, num_samples=len(example["anchors"][0]))


def merge_label_unlabel_data(data_batch, data_batch_unlabel):
    """Concatenate a labelled and an unlabelled collated batch for the semi-supervised mode (trainer_sessd.py:275-300), key class by key
    class: tensors with a batch-id column (``points`` / ``coordinates`` + ``_raw``) get the unlabelled frame ids shifted behind the labelled
    ones; per-voxel / per-frame tensors are concatenated; ``metadata`` lists and the ``calib`` tensors likewise; per-task target lists
    are concatenated when the unlabelled batch carries them; other keys present in both are joined with ``np.concatenate``;
    ``ssl_labeled`` marks the supervised frames.  Returns a new dict (the reference edits ``data_batch`` in place)."""
    import torch
    n_lab = int(data_batch["points"][-1, 0].item()) + 1
    n_unl = int(data_batch_unlabel["points"][-1, 0].item()) + 1
    batch_id_keys = ("coordinates", "points", "coordinates_raw", "points_raw")
    cat_keys = frozenset(b + s for b in ("voxels", "num_points", "num_gt", "voxel_labels", "num_voxels") for s in ("", "_raw"))
    merged = {}
    for key, val in data_batch.items():
        other = data_batch_unlabel.get(key)
        if other is None:
            merged[key] = val
        elif key in batch_id_keys:
            shifted = other.clone()
            shifted[:, 0] += n_lab
            merged[key] = torch.cat([val, shifted], 0)
        elif key in cat_keys:
            merged[key] = torch.cat([val, other], 0)
        elif key == "metadata":
            merged[key] = list(val) + list(other)
        elif key == "calib":
            merged[key] = {k: torch.cat([v, other[k]], 0) for k, v in val.items()}
        elif key in _LIST_KEYS:
            merged[key] = [torch.cat([val[0], other[0]], 0)] + list(val[1:])
        else:
            merged[key] = np.concatenate([val, other], axis=0)
    flag = torch.zeros(n_lab + n_unl, dtype=torch.int32)
    flag[:n_lab] = 1
    merged["ssl_labeled"] = flag
    return merged

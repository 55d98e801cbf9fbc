"""Checkpoint wire format (reference: det3d/torchie/trainer/checkpoint.py:117-214): a torch file holding ``{"meta", "state_dict",
["optimizer"]}`` or a bare state dict; ``module.`` prefixes of DataParallel checkpoints are stripped; parameter names and the spconv
weight layout ``[kz, ky, kx, Cin, Cout]`` are the reference's, so a published SE-SSD checkpoint (and its ``_ema`` twin) loads as is.
Only local files are supported (no model-zoo / http schemes: there is no network on the target boxes)."""
import os.path as osp
from collections import OrderedDict

import torch


def load_state_dict(module, state_dict, strict=False, logger=None):
    """Copy matching entries; collect shape mismatches / unexpected / missing keys like the reference (:37-97)."""
    own = module.state_dict()
    unexpected, mismatched = [], []
    for name, param in state_dict.items():
        if name not in own:
            unexpected.append(name)
            continue
        if tuple(own[name].shape) != tuple(param.shape):
            mismatched.append("%s: checkpoint %s vs model %s" % (name, tuple(param.shape), tuple(own[name].shape)))
            continue
        own[name].copy_(param.data if isinstance(param, torch.nn.Parameter) else param)
    missing = sorted(set(own.keys()) - set(state_dict.keys()))
    msgs = []
    if unexpected:
        msgs.append("unexpected key in source state_dict: " + ", ".join(unexpected))
    if missing:
        msgs.append("missing keys in source state_dict: " + ", ".join(missing))
    msgs += mismatched
    if msgs:
        text = "\n".join(msgs)
        if strict:
            raise RuntimeError(text)
        if logger is not None:
            logger.warning(text)
    return msgs


def load_checkpoint(model, filename, map_location=None, strict=False, logger=None, trusted=False):
    """trusted: the wire format is tensors + plain containers, so the file is read with ``weights_only=True`` (no arbitrary pickle
    execution from a downloaded checkpoint); pass trusted=True to fall back to the full unpickler for a file you wrote yourself."""
    if "://" in filename:
        raise NotImplementedError("only local checkpoint files are supported (no network on the target boxes)")
    if not osp.isfile(filename):
        raise IOError("{} is not a checkpoint file".format(filename))
    try:
        checkpoint = torch.load(filename, map_location=map_location, weights_only=True)
    except Exception:
        if not trusted:
            raise
        checkpoint = torch.load(filename, map_location=map_location, weights_only=False)
    if isinstance(checkpoint, OrderedDict):
        state_dict = checkpoint
    elif isinstance(checkpoint, dict) and "state_dict" in checkpoint:
        state_dict = checkpoint["state_dict"]
    else:
        raise RuntimeError("No state_dict found in checkpoint file {}".format(filename))
    if list(state_dict.keys())[0].startswith("module."):
        state_dict = OrderedDict((k[7:], v) for k, v in state_dict.items())
    load_state_dict(model.module if hasattr(model, "module") else model, state_dict, strict, logger)
    return checkpoint


def weights_to_cpu(state_dict):
    return OrderedDict((k, v.cpu()) for k, v in state_dict.items())


def save_checkpoint(model, filename, optimizer=None, meta=None):
    if meta is None:
        meta = {}
    elif not isinstance(meta, dict):
        raise TypeError("meta must be a dict or None, but got {}".format(type(meta)))
    if hasattr(model, "module"):
        model = model.module
    checkpoint = {"meta": meta, "state_dict": weights_to_cpu(model.state_dict())}
    if optimizer is not None:
        checkpoint["optimizer"] = optimizer.state_dict()
    torch.save(checkpoint, filename)

"""Optimiser side of the SE-SSD training step on flat parameter arenas (csrc/train.cu).

The reference updates the teacher parameter by parameter in Python (trainer_sessd.py:315-318), flattens and unflattens the gradients
around ``dist.all_reduce`` (dist_utils.py:8-29) and steps a fastai ``OptimWrapper`` over parameter groups.  Here a model's parameters --
and their gradients -- are views into ONE contiguous fp32 buffer each (``ParamArena``): the EMA, the gradient all-reduce (in place on the
arena: NCCL over NVLink on the device, gloo in the CPU tests) and the AdamW step are one launch / one collective each."""
import ctypes as C

import torch
import torch.distributed as dist

from ._lib import check, lib


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _st(t):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream) if t.is_cuda else C.c_void_p(0)


class ParamArena:
    """Re-homes every parameter of ``model`` into one flat fp32 buffer (``.flat``) and gives every parameter a ``.grad`` that is a view
    into a second flat buffer (``.grad_flat``).  Parameter values, names, shapes and ``state_dict()`` are unchanged; autograd accumulates
    straight into the gradient arena.  Offsets are 16-byte aligned (vector loads in the kernels)."""

    def __init__(self, model, with_grad=True):
        self.params = [p for p in model.parameters()]
        assert all(p.dtype == torch.float32 for p in self.params), "fp32 master parameters"
        dev = self.params[0].device
        self.offsets, n = [], 0
        for p in self.params:
            self.offsets.append(n)
            n += (p.numel() + 3) // 4 * 4
        self.numel = n
        self.flat = torch.zeros((n,), dtype=torch.float32, device=dev)
        self.grad_flat = torch.zeros((n,), dtype=torch.float32, device=dev) if with_grad else None
        for p, off in zip(self.params, self.offsets):
            view = self.flat[off:off + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view
            if with_grad:
                p.grad = self.grad_flat[off:off + p.numel()].view_as(p)

    def zero_grad(self):
        self.grad_flat.zero_()


def axpby(y, x, a, b):
    """y = a y + b x on flat fp32 CUDA tensors (x may be None): one kernel.  No host implementation."""
    if not y.is_cuda or (x is not None and not x.is_cuda):
        raise ValueError("sessd_b200.train.axpby: CUDA tensors only (there is no CPU path)")
    check(lib.sessd_axpby(_p(y), _p(x), float(a), float(b), int(y.numel()), _st(y)), "sessd_axpby")
    return y


def update_ema_variables(model_or_arena, ema_model_or_arena, global_step, max_alpha=0.999):
    """Teacher update of trainer_sessd.py:315-318: alpha = min(1 - 1 / (global_step + 1), 0.999); ema = alpha * ema + (1 - alpha) * param.
    Both models live in ParamArenas: one launch over the flat buffers."""
    alpha = min(1.0 - 1.0 / (global_step + 1), max_alpha)
    if isinstance(model_or_arena, ParamArena) and isinstance(ema_model_or_arena, ParamArena):
        assert model_or_arena.numel == ema_model_or_arena.numel
        axpby(ema_model_or_arena.flat, model_or_arena.flat, alpha, 1.0 - alpha)
        return alpha
    raise TypeError("update_ema_variables: wrap both models in a ParamArena first (the update is one launch over the flat buffers)")


def allreduce_grad_arena(arena, group=None):
    """Average the gradient arena over the process group IN PLACE: one collective on one contiguous buffer (the reference's coalesced path
    builds that buffer with _flatten_dense_tensors and copies the result back tensor by tensor, dist_utils.py:20-29)."""
    world = dist.get_world_size(group)
    if world == 1:
        return arena.grad_flat
    if arena.grad_flat.is_cuda and dist.get_backend(group) == "nccl":
        dist.all_reduce(arena.grad_flat, op=dist.ReduceOp.AVG, group=group)          # averaged inside the collective (NVLink / NVSwitch)
    else:                                                                            # gloo (CPU tests of the sharding logic): sum, then scale
        dist.all_reduce(arena.grad_flat, op=dist.ReduceOp.SUM, group=group)
        arena.grad_flat.div_(world)
    return arena.grad_flat


class ArenaAdamW:
    """AdamW over a ParamArena in one launch per step (torch.optim.AdamW arithmetic == fastai OptimWrapper(true_wd=True) around Adam,
    det3d/solver/fastai_optim.py).  ``lr`` may be changed between steps (one-cycle schedule)."""

    def __init__(self, arena, lr=3e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.01):
        self.arena, self.lr, self.betas, self.eps, self.weight_decay = arena, lr, betas, eps, weight_decay
        self.exp_avg = torch.zeros_like(arena.flat)
        self.exp_avg_sq = torch.zeros_like(arena.flat)
        self.steps = 0

    def step(self):
        self.steps += 1
        a = self.arena
        if not a.flat.is_cuda:
            raise ValueError("ArenaAdamW: CUDA arenas only (there is no CPU path)")
        check(lib.sessd_adamw_step(_p(a.flat), _p(a.grad_flat), _p(self.exp_avg), _p(self.exp_avg_sq), int(a.numel), float(self.lr),
                                   float(self.betas[0]), float(self.betas[1]), float(self.eps), float(self.weight_decay), int(self.steps),
                                   _st(a.flat)), "sessd_adamw_step")


# ------------------------------------------------------------------------------------------------- data gradients through the forward kernels
def subm_dgrad_weight(w):
    """SubMConv3d weight [kz, ky, kx, Cin, Cout] (spconv layout, scn.py:106-131)  ->  [kz, ky, kx, Cout, Cin] such that the FORWARD sparse conv
    over the SAME rulebook / tile lists, fed with d loss / d out, yields d loss / d in:  W'[k] = W[K-1-k]^T.  (A SubM layer's neighbour
    table is point-symmetric -- nbr_t[i, k] = nbr[i, K-1-k] -- so the data gradient needs neither a transposed rulebook nor a new kernel;
    pinned on the CPU by tests/test_train_step.py against oracle/spconv_grad_ref.py, which is pinned to dense autograd.)"""
    return w.flip(0, 1, 2).transpose(3, 4).contiguous()


def conv2d_s1_dgrad_weight(w):
    """Conv2d weight [Cout, Cin, k, k] of a stride-1 neck layer  ->  [Cin, Cout, k, k], taps flipped: the stride-1 conv (pad k - 1 - p) that
    maps d loss / d out to d loss / d in (rpn_v1.py:135-210 layers bottom_up_block_0.*, bottom_up_block_1.3 / .6, conv_0, conv_1, trans_*)."""
    return w.flip(2, 3).transpose(0, 1).contiguous()

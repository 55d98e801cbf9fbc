"""CPU restatement of the DATA gradient of the BEV neck's convolutions (TEST INFRASTRUCTURE ONLY; oracle for the SURVEY §8(f) row 1
kernels that are not built yet).  The reference obtains these through torch autograd / cuDNN (``loss.backward()`` in
det3d/torchie/trainer/trainer_sessd.py:346-352 over the modules of det3d/models/necks/rpn_v1.py:135-210); here each one is written as the
FORWARD operation the device kernels already implement (csrc/bevconv_p2.cu: conv, stride-2 conv, stride-2 deconv), so a backward pass
is the same kernels with re-packed weights:

* conv k x k, stride 1, pad p  (bottom_up_block_0.*, bottom_up_block_1.3 / .6, conv_0, conv_1, trans_*, w_*):
      dX = conv(dY, W~, stride 1, pad k - 1 - p),   W~[ci, co, a, b] = W[co, ci, k-1-a, k-1-b]          (taps flipped, channels swapped)
* conv 3 x 3, stride 2, pad 1 on an even-sized map  (bottom_up_block_1.0):
      dX = conv_transpose(dY, W, stride 2, pad 1, output_padding 1)      -- a Conv2d weight [co, ci, k, k] IS a ConvTranspose2d weight
* conv_transpose 3 x 3, stride 2, pad 1, output_padding 1  (deconv_block_0 / _1):
      dX = conv(dY, Wd, stride 2, pad 1)                                 -- a ConvTranspose2d weight [ci, co, k, k] IS a Conv2d weight
Pinned by tests/test_oracle.py against torch autograd (the reference's own mechanism)."""
import torch
import torch.nn.functional as F


def conv_s1_dgrad_weight(w):
    """Conv2d weight [Cout, Cin, k, k] -> the weight of the stride-1 conv that maps dY to dX: [Cin, Cout, k, k], taps flipped."""
    return w.flip(2, 3).transpose(0, 1).contiguous()


def conv_dgrad(grad_out, w, stride, pad):
    """dX of Y = conv2d(X, w, stride, pad) for the two cases of the neck (stride 1 any k; stride 2 with k = 3, pad = 1, even input size)."""
    k = w.shape[2]
    if stride == 1:
        return F.conv2d(grad_out, conv_s1_dgrad_weight(w), None, 1, k - 1 - pad)
    if stride == 2 and k == 3 and pad == 1:
        return F.conv_transpose2d(grad_out, w, None, 2, 1, output_padding=1)
    raise NotImplementedError("not a layer shape of the SSFA neck")


def deconv_dgrad(grad_out, w, stride=2, pad=1):
    """dX of Y = conv_transpose2d(X, w [Cin, Cout, 3, 3], stride 2, pad 1, output_padding 1)."""
    return F.conv2d(grad_out, w, None, stride, pad)

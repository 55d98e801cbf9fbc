"""CPU restatement of the BACKWARD of the sparse 3-D convolution stage (TEST INFRASTRUCTURE ONLY; oracle for the SURVEY §8(f) row 1
kernels that are not built yet -- the order of work is oracle first).

Like the forward (oracle/spconv_ref.py) the arithmetic lives in spconv 1.x, which is not under /root/reference -- PARITY UNPINNED to the
reference; the functions below are pinned to an independent implementation instead: torch autograd through ``F.conv3d`` on the
densified volume (tests/test_oracle.py).  What the reference fixes for this path is the layer list (det3d/models/backbones/scn.py:106-149:
SubMConv3d / SparseConv3d, bias=False, BatchNorm1d(eps=1e-3) + ReLU after every conv) and that the gradient reaches the encoder through
``loss.backward()`` on the dense BEV map (det3d/torchie/apis/train_sessd.py / trainer_sessd.py:346-352).

Rulebook form (same canonical neighbour table as the forward: ``nbr[o, k]`` = input row of output row ``o`` under kernel offset ``k`` or -1):

    forward   out[o]   = sum_k  in[nbr[o, k]] @ W[k]
    dgrad     gin[i]   = sum_k  sum_{o : nbr[o, k] = i}  gout[o] @ W[k]^T
    wgrad     gW[k]    = sum_{o : nbr[o, k] >= 0}  in[nbr[o, k]]^T  gout[o]

Every (k, i) has AT MOST ONE such o (pos_o = (pos_i + pad - k) / stride is a function of i and k), so the dgrad is again an
output-stationary gather -- over the TRANSPOSED table ``nbr_t[i, k] = o`` with the transposed weights -- i.e. the forward kernel
(csrc/spconv_cg.cu) run with ``(gout, nbr_t, W^T)``; for SubM layers ``nbr_t[i, k] = nbr[i, K-1-k]`` (point symmetry of the offsets), no
second table needed."""
import numpy as np


def transpose_nbr(nbr, n_in):
    """nbr [N_out, K] (input row or -1)  ->  nbr_t [n_in, K] with nbr_t[i, k] = the output row o whose k-th neighbour is input row i (or -1)."""
    n_out, kvol = nbr.shape
    nbr_t = np.full((n_in, kvol), -1, dtype=np.int64)
    for k in range(kvol):
        o = np.nonzero(nbr[:, k] >= 0)[0]
        i = nbr[o, k]
        assert len(np.unique(i)) == len(i), "an (input row, offset) pair can feed one output row only"
        nbr_t[i, k] = o
    return nbr_t


def conv_backward_from_nbr(feat, nbr, weight, grad_out, dtype=np.float64):
    """Gradients of ``out = conv_from_nbr(feat, nbr, weight)`` w.r.t. ``feat`` [N_in, Cin] and ``weight`` [K, Cin, Cout] given
    ``grad_out`` [N_out, Cout]; pair by pair, ascending k (the summation order a deterministic kernel would use)."""
    feat = np.asarray(feat, dtype)
    w = np.asarray(weight, dtype)
    go = np.asarray(grad_out, dtype)
    gfeat = np.zeros_like(feat)
    gw = np.zeros_like(w)
    for k in range(nbr.shape[1]):
        o = np.nonzero(nbr[:, k] >= 0)[0]
        if len(o) == 0:
            continue
        i = nbr[o, k]
        gfeat[i] += go[o] @ w[k].T                      # rows i are distinct inside one offset (see transpose_nbr)
        gw[k] = feat[i].T @ go[o]
    return gfeat, gw


def bn_relu_backward(x, grad_y, gamma, beta, mean, var, eps=1e-3, relu=True):
    """Backward of the EVAL-mode BatchNorm1d + ReLU that follows every conv (oracle/spconv_ref.bn_relu, same argument order):
    y = relu((x - mean) * s + beta), s = gamma / sqrt(var + eps).  Returns d loss / d x (the statistics are constants in eval mode; the
    train-mode batch-statistics terms are a separate restatement when that path is built)."""
    s = gamma / np.sqrt(var + eps)
    g = np.asarray(grad_y, np.float64) * s
    if relu:
        g = g * (((x - mean) * s + beta) > 0)
    return g

"""Host-side weight preparation: fold eval-mode BatchNorm into the convolutions and repack the
reference `state_dict` tensors into the K-major fp16 layouts the tcgen05 kernels consume.

Reference layouts: learning/models/network_modules.py:37-50 (ConvBNReLU: net.0 = Conv2d, net.1 = BN),
:73-111 (ResnetBasicBlock: conv1/bn1/conv2/bn2).  BatchNorm2d in eval mode is the per-channel affine
y = (x - mean) / sqrt(var + eps) * gamma + beta, folded here in fp32 before the fp16 cast.
"""
import torch


def fold_bn(w, b, bn):
    """w (Co,Ci,kh,kw), b (Co) or None, bn = dict(weight,bias,running_mean,running_var,eps) or None."""
    w = w.detach().float()
    b = torch.zeros(w.shape[0]) if b is None else b.detach().float()
    if bn is None:
        return w, b
    scale = bn["weight"].float() / torch.sqrt(bn["running_var"].float() + bn["eps"])
    w = w * scale.reshape(-1, 1, 1, 1)
    b = (b - bn["running_mean"].float()) * scale + bn["bias"].float()
    return w, b


def pack_conv3(w):
    """(Co,Ci,3,3) fp32 -> (Co, 9*Ci) fp16, K ordered (r, s, c)."""
    co, ci = w.shape[:2]
    return w.permute(0, 2, 3, 1).reshape(co, 9 * ci).contiguous().half()


def pack_conv7(w):
    """(Co,Ci<=8,7,7) fp32 -> (Co, 7*64) fp16: per filter row r, 7 taps x 8 channels + 8 zeros."""
    co, ci = w.shape[:2]
    out = torch.zeros(co, 7, 8, 8, dtype=torch.float32)  # (co, r, s(pad to 8), c(pad to 8))
    out[:, :, :7, :ci] = w.permute(0, 2, 3, 1)
    return out.reshape(co, 7 * 64).contiguous().half()


def pack_linear(w):
    """(Co, K) -> fp16 contiguous."""
    return w.detach().float().contiguous().half()


def pad_image_c8(x):
    """(N,C<=8,H,W) float -> stem input layout [N][H+6][W+8][8] fp16, image at (3,3), zero border."""
    n, c, h, w = x.shape
    out = torch.zeros(n, h + 6, w + 8, 8, dtype=torch.float16, device=x.device)
    out[:, 3 : 3 + h, 3 : 3 + w, :c] = x.permute(0, 2, 3, 1).half()
    return out

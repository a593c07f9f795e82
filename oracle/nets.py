"""Functional fp32 CPU restatement of the two networks (TEST INFRASTRUCTURE — see oracle/__init__.py).

Operates directly on a reference-layout `state_dict`:
  refine_forward  follows learning/models/refine_network.py:73-93 (RefineNet.forward)
  score_features  follows learning/models/score_network.py:60-74  (extract_feat)
  score_forward   follows learning/models/score_network.py:77-90  (forward)
with the building blocks of learning/models/network_modules.py:37-50 (ConvBNReLU),
:73-111 (ResnetBasicBlock), :115-137 (PositionalEmbedding) and torch's
nn.TransformerEncoderLayer(norm_first=False, activation=relu, batch_first=True) /
nn.MultiheadAttention in eval mode (dropout inactive).

PINNED: tests/test_oracle_golden.py compares this file with outputs produced by the reference's own
classes (tools/make_golden.py, run in the build container where /root/reference is mounted).
"""
import math

import torch
import torch.nn.functional as F

BN_EPS = 1e-5
LN_EPS = 1e-5
NHEAD = 4


def _bn(x, sd, p):
    return F.batch_norm(x, sd[f"{p}.running_mean"], sd[f"{p}.running_var"], sd[f"{p}.weight"], sd[f"{p}.bias"],
                        training=False, eps=BN_EPS)


def conv_bn_relu(x, sd, p, k, stride):
    """network_modules.py:37-50."""
    y = F.conv2d(x, sd[f"{p}.net.0.weight"], sd.get(f"{p}.net.0.bias"), stride=stride, padding=(k - 1) // 2)
    if f"{p}.net.1.weight" in sd:
        y = _bn(y, sd, f"{p}.net.1")
    return F.relu(y)


def res_block(x, sd, p):
    """network_modules.py:94-111."""
    y = F.conv2d(x, sd[f"{p}.conv1.weight"], sd.get(f"{p}.conv1.bias"), padding=1)
    if f"{p}.bn1.weight" in sd:
        y = _bn(y, sd, f"{p}.bn1")
    y = F.relu(y)
    y = F.conv2d(y, sd[f"{p}.conv2.weight"], sd.get(f"{p}.conv2.bias"), padding=1)
    if f"{p}.bn2.weight" in sd:
        y = _bn(y, sd, f"{p}.bn2")
    return F.relu(y + x)


def encode_a(x, sd, name):
    x = conv_bn_relu(x, sd, f"{name}.0", 7, 2)
    x = conv_bn_relu(x, sd, f"{name}.1", 3, 2)
    x = res_block(x, sd, f"{name}.2")
    return res_block(x, sd, f"{name}.3")


def encode_ab(x, sd, name):
    x = res_block(x, sd, f"{name}.0")
    x = res_block(x, sd, f"{name}.1")
    x = conv_bn_relu(x, sd, f"{name}.2", 3, 2)
    x = res_block(x, sd, f"{name}.3")
    return res_block(x, sd, f"{name}.4")


def mha(x, sd, p):
    """Self multi-head attention, batch_first, q=k=v=x: (B, T, 512) -> (B, T, 512)."""
    B, T, D = x.shape
    dh = D // NHEAD
    qkv = x @ sd[f"{p}.in_proj_weight"].t() + sd[f"{p}.in_proj_bias"]
    q, k, v = qkv.split(D, dim=-1)
    q = q.reshape(B, T, NHEAD, dh).transpose(1, 2)
    k = k.reshape(B, T, NHEAD, dh).transpose(1, 2)
    v = v.reshape(B, T, NHEAD, dh).transpose(1, 2)
    att = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(dh), dim=-1)
    o = (att @ v).transpose(1, 2).reshape(B, T, D)
    return o @ sd[f"{p}.out_proj.weight"].t() + sd[f"{p}.out_proj.bias"]


def encoder_layer(x, sd, p):
    """Post-LN TransformerEncoderLayer(d=512, nhead=4, ff=512, relu)."""
    x = F.layer_norm(x + mha(x, sd, f"{p}.self_attn"), (x.shape[-1],), sd[f"{p}.norm1.weight"], sd[f"{p}.norm1.bias"], LN_EPS)
    ff = F.relu(x @ sd[f"{p}.linear1.weight"].t() + sd[f"{p}.linear1.bias"]) @ sd[f"{p}.linear2.weight"].t() + sd[f"{p}.linear2.bias"]
    return F.layer_norm(x + ff, (x.shape[-1],), sd[f"{p}.norm2.weight"], sd[f"{p}.norm2.bias"], LN_EPS)


def _tokens(ab, sd):
    bs = ab.shape[0]
    t = ab.reshape(bs, ab.shape[1], -1).permute(0, 2, 1)
    return t + sd["pos_embed.pe"][:, : t.shape[1]]


@torch.no_grad()
def refine_forward(sd, A, B):
    """refine_network.py:73-93.  A, B: (N, 6, 160, 160) fp32 -> {'trans': (N,3), 'rot': (N,3)}."""
    bs = A.shape[0]
    x = encode_a(torch.cat([A, B], 0), sd, "encodeA")
    ab = encode_ab(torch.cat((x[:bs], x[bs:]), 1), sd, "encodeAB")
    t = _tokens(ab, sd)
    out = {}
    for name, head in (("trans", "trans_head"), ("rot", "rot_head")):
        y = encoder_layer(t, sd, f"{head}.0")
        y = y @ sd[f"{head}.1.weight"].t() + sd[f"{head}.1.bias"]
        out[name] = y.mean(dim=1)
    return out


@torch.no_grad()
def score_features(sd, A, B):
    """score_network.py:60-74."""
    bs = A.shape[0]
    x = encode_a(torch.cat([A, B], 0), sd, "encoderA")
    ab = encode_ab(torch.cat((x[:bs], x[bs:]), 1), sd, "encoderAB")
    t = _tokens(ab, sd)
    return mha(t, sd, "att").mean(dim=1).reshape(bs, -1)


@torch.no_grad()
def score_tail(sd, feats, L):
    """score_network.py:84-88: cross-hypothesis attention + linear.  feats (bs*L, 512) -> (bs, L)."""
    bs = feats.shape[0] // L
    x = feats.reshape(bs, L, -1)
    x = mha(x, sd, "att_cross")
    return (x @ sd["linear.weight"].t() + sd["linear.bias"]).reshape(bs, L)


@torch.no_grad()
def score_forward(sd, A, B, L):
    return score_tail(sd, score_features(sd, A, B), L)

"""state_dict handling for the two networks.

The engine ingests the *reference's* checkpoints unchanged: the key layout below is the one
`RefineNet` / `ScoreNetMultiPair` produce (learning/models/refine_network.py:24-70,
learning/models/score_network.py:27-57), loaded the way the predictors do
(learning/training/predict_pose_refine.py:137-143: `ckpt['model']` if present).

The released checkpoints are not redistributable with this repository and there is no network, so
`random_state_dict` builds a seeded stand-in with the same keys/shapes (BatchNorm running statistics
randomised so that BN folding is exercised).  tests/test_oracle_golden.py checks that these
state_dicts load *strictly* into the reference classes (done once in the build container; the
golden outputs are committed).
"""
import math
import os

import torch

DEFAULT_CFG = {
    "use_BN": True,
    "c_in": 6,
    "normalize_xyz": True,
    "crop_ratio": 1.2,
    "input_resize": [160, 160],
    "rot_rep": "axis_angle",
    "trans_rep": "tracknet",
    "rot_normalizer": 0.3490658503988659,
    "trans_normalizer": [0.019999999552965164, 0.019999999552965164, 0.05000000074505806],
    "use_normal": False,
    "zfar": float("inf"),
}


def positional_embedding(max_len=400, d_model=512):
    """Sinusoidal table, network_modules.py:115-137."""
    pe = torch.zeros(max_len, d_model, dtype=torch.float32)
    position = torch.arange(0, max_len, dtype=torch.float32).unsqueeze(1)
    div_term = (torch.arange(0, d_model, 2, dtype=torch.float32) * -(math.log(10000.0) / d_model)).exp()[None]
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.unsqueeze(0)


def _conv(g, sd, name, co, ci, k, gain=2.0):
    fan_in = ci * k * k
    sd[f"{name}.weight"] = torch.randn(co, ci, k, k, generator=g) * math.sqrt(gain / fan_in)
    sd[f"{name}.bias"] = torch.randn(co, generator=g) * 0.05


def _bn(g, sd, name, c):
    sd[f"{name}.weight"] = 0.8 + 0.4 * torch.rand(c, generator=g)
    sd[f"{name}.bias"] = torch.randn(c, generator=g) * 0.1
    sd[f"{name}.running_mean"] = torch.randn(c, generator=g) * 0.1
    sd[f"{name}.running_var"] = 0.5 + torch.rand(c, generator=g)
    sd[f"{name}.num_batches_tracked"] = torch.tensor(1000, dtype=torch.long)


def _linear(g, sd, name, co, ci, wname="weight", bname="bias", gain=1.0):
    sd[f"{name}.{wname}" if wname else name] = torch.randn(co, ci, generator=g) * math.sqrt(gain / ci)
    sd[f"{name}.{bname}"] = torch.randn(co, generator=g) * 0.05


def _encoders(g, sd, nameA, nameAB, c_in, use_bn):
    def cbr(prefix, co, ci, k):
        _conv(g, sd, f"{prefix}.net.0", co, ci, k)
        if use_bn:
            _bn(g, sd, f"{prefix}.net.1", co)

    def res(prefix, c):
        _conv(g, sd, f"{prefix}.conv1", c, c, 3)
        _conv(g, sd, f"{prefix}.conv2", c, c, 3, gain=0.5)
        if use_bn:
            _bn(g, sd, f"{prefix}.bn1", c)
            _bn(g, sd, f"{prefix}.bn2", c)

    cbr(f"{nameA}.0", 64, c_in, 7)
    cbr(f"{nameA}.1", 128, 64, 3)
    res(f"{nameA}.2", 128)
    res(f"{nameA}.3", 128)
    res(f"{nameAB}.0", 256)
    res(f"{nameAB}.1", 256)
    cbr(f"{nameAB}.2", 512, 256, 3)
    res(f"{nameAB}.3", 512)
    res(f"{nameAB}.4", 512)


def _mha(g, sd, name, d=512):
    sd[f"{name}.in_proj_weight"] = torch.randn(3 * d, d, generator=g) * math.sqrt(1.0 / d)
    sd[f"{name}.in_proj_bias"] = torch.randn(3 * d, generator=g) * 0.05
    _linear(g, sd, f"{name}.out_proj", d, d)


def random_state_dict(kind, seed=0, c_in=6, use_bn=True):
    """kind: 'refine' | 'score'.  Deterministic for a given torch version (CPU generator)."""
    g = torch.Generator(device="cpu").manual_seed(seed + (0 if kind == "refine" else 1000))
    sd = {}
    if kind == "refine":
        _encoders(g, sd, "encodeA", "encodeAB", c_in, use_bn)
        sd["pos_embed.pe"] = positional_embedding()
        for head, out_dim in (("trans_head", 3), ("rot_head", 3)):
            _mha(g, sd, f"{head}.0.self_attn")
            _linear(g, sd, f"{head}.0.linear1", 512, 512, gain=2.0)
            _linear(g, sd, f"{head}.0.linear2", 512, 512)
            for ln in ("norm1", "norm2"):
                sd[f"{head}.0.{ln}.weight"] = 0.8 + 0.4 * torch.rand(512, generator=g)
                sd[f"{head}.0.{ln}.bias"] = torch.randn(512, generator=g) * 0.1
            _linear(g, sd, f"{head}.1", out_dim, 512, gain=0.05)
            # A trained refiner is contractive: near the answer its update shrinks.  A plain random read-out moves every
            # hypothesis by centimetres / several degrees per pass whatever it sees, so rounding differences between two
            # correct implementations grow ~2x per pass (measured: 3.5e-4 after one pass -> 5e-3 after five) and decide
            # which hypothesis the scorer picks.  The stand-in's read-out is therefore scaled to millimetre / sub-degree
            # updates (x0.1): the render-and-compare loop stays exercised, the run-to-run result becomes reproducible.
            sd[f"{head}.1.weight"] *= 0.1
            sd[f"{head}.1.bias"] *= 0.1
    elif kind == "score":
        _encoders(g, sd, "encoderA", "encoderAB", c_in, use_bn)
        _mha(g, sd, "att")
        sd["pos_embed.pe"] = positional_embedding()
        # The cross-hypothesis tail decides the arg-max.  A plain random init makes att_cross attend uniformly, so all
        # 252 scores collapse to one value +- 6e-4 and "the selected index" is decided by rounding noise.  The stand-in
        # tail therefore has its own generator (the encoder / `att` weights above are untouched by it), sharper
        # query/key projections (x2) and a larger read-out (x60), picked by tools/pick_tail_seed.py so that on BOTH the
        # 252-hypothesis golden scene (oracle features) and the bench scene (bench.py) the winner leads the runner-up
        # by 1.2 sigma of the score spread (0.10 of 0.085 / 0.19 of 0.16), ~50x the score error of an fp16 feature path.
        g2 = torch.Generator(device="cpu").manual_seed(7615 + seed)
        _mha(g2, sd, "att_cross")
        sd["att_cross.in_proj_weight"][:1024] *= 2.0
        sd["att_cross.in_proj_bias"][:1024] *= 2.0
        _linear(g2, sd, "linear", 1, 512, gain=4.0)
        sd["linear.weight"] *= 60.0
    else:
        raise ValueError(kind)
    return sd


def load_checkpoint(path):
    """Reference checkpoint loader semantics (predict_pose_refine.py:137-140)."""
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    if "model" in ckpt:
        ckpt = ckpt["model"]
    return ckpt


def load_reference_config(path, kind):
    """weights/<run>/config.yml with the reference's backward-compatibility defaults for missing keys
    (predict_pose_refine.py:107-131 for the refiner, predict_score.py:131-143 for the scorer)."""
    import yaml

    cfg = {}
    if os.path.exists(path):
        with open(path) as fh:
            cfg = yaml.safe_load(fh) or {}
    defaults = {"use_normal": False, "use_BN": False, "c_in": 4, "normalize_xyz": False}
    if kind == "refine":
        defaults.update({"use_mask": False, "n_view": 1, "trans_rep": "tracknet", "rot_rep": "axis_angle", "zfar": 3, "normal_uint8": False})
    else:
        defaults.update({"zfar": float("inf")})
    for k, v in defaults.items():
        cfg.setdefault(k, v)
    if cfg.get("crop_ratio") is None:
        cfg["crop_ratio"] = 1.2
    if isinstance(cfg["zfar"], str) and "inf" in cfg["zfar"].lower():
        cfg["zfar"] = float("inf")
    for k, v in DEFAULT_CFG.items():  # keys the engine reads that old configs may lack
        if k not in ("use_BN", "c_in", "normalize_xyz", "use_normal", "zfar", "crop_ratio"):
            cfg.setdefault(k, v)
    return cfg


def find_reference_weights(run_name, root=None):
    """Look for weights/<run_name>/model_best.pth next to the caller's tree (reference layout)."""
    cands = []
    if root:
        cands.append(os.path.join(root, "weights", run_name, "model_best.pth"))
    cands.append(os.path.join(os.getcwd(), "weights", run_name, "model_best.pth"))
    for c in cands:
        if os.path.exists(c):
            return c
    return None

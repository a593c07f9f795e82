"""CPU: the oracle's network restatement (oracle/nets.py) against golden outputs produced by the
reference's own RefineNet / ScoreNetMultiPair classes (tools/make_golden.py).  Tolerance 2e-5 abs:
both sides are fp32 CPU, differences are only op-ordering inside attention / layer norm."""
import os

import numpy as np
import torch

from foundationpose_b200.weights import random_state_dict
from oracle import nets

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "nets_golden.npz"))


def _crops(n, seed):
    g = torch.Generator().manual_seed(seed)
    A = torch.rand(n, 6, 160, 160, generator=g)
    B = torch.rand(n, 6, 160, 160, generator=g)
    for T in (A, B):
        T[:, 3:] = (T[:, 3:] - 0.5) * 2
        T[:, 3:, :30] = 0
    return A, B


def _wsum(sd):
    return float(sum(v.double().abs().sum() for v in sd.values() if v.dtype.is_floating_point))


def test_seeded_weights_reproduce():
    assert abs(_wsum(random_state_dict("refine", 0)) - GOLD["refine_wsum"][0]) < 1e-6 * GOLD["refine_wsum"][0]
    assert abs(_wsum(random_state_dict("score", 0)) - GOLD["score_wsum"][0]) < 1e-6 * GOLD["score_wsum"][0]


def test_refine_net_matches_reference_golden():
    sd = random_state_dict("refine", 0)
    A, B = _crops(2, 123)
    out = nets.refine_forward(sd, A, B)
    np.testing.assert_allclose(out["trans"].numpy(), GOLD["refine_trans"], atol=2e-5, rtol=1e-4)
    np.testing.assert_allclose(out["rot"].numpy(), GOLD["refine_rot"], atol=2e-5, rtol=1e-4)


def test_score_net_matches_reference_golden():
    sd = random_state_dict("score", 0)
    A, B = _crops(3, 456)
    feats = nets.score_features(sd, A, B)
    np.testing.assert_allclose(feats.numpy(), GOLD["score_feats"], atol=5e-5, rtol=1e-4)
    logits = nets.score_tail(sd, feats, 3)
    np.testing.assert_allclose(logits.numpy(), GOLD["score_logits"], atol=2e-5, rtol=1e-4)
    assert int(logits.argmax()) == int(GOLD["score_logits"].argmax())

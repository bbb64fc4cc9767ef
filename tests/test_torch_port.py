"""oracle/torch_port.py (the CPU/cuDNN baseline stand-in) against the reference fixtures."""
import os

import numpy as np
import pytest
import torch

from oracle import lstm_lm_oracle as O
from oracle import torch_port as P
from tests._golden import GOLDEN, StepCase


@pytest.mark.parametrize("name", ["tiny_pytorch", "tiny_carry3", "mid_H72"])
def test_port_train_steps_match_reference(name):
    c = StepCase(name)
    m = P.TorchLstmLm(c.V, c.H, c.L, c.dropout, c.winit)
    m.load_reference_state_dict(c.params0())
    m.eval()
    states = [(torch.tensor(h)[None], torch.tensor(cc)[None]) for h, cc in c.states0()]
    for s in range(c.steps):
        x, y = torch.tensor(c.x(s)), torch.tensor(c.y(s))
        loss, norm, states = P.train_step(m, x, y, states, c.lr, c.max_norm)
        assert abs(loss.item() - c.loss(s)) < 1e-5 * max(1, abs(c.loss(s)))
        assert abs(float(norm) - c.norm(s)) < 1e-5 * max(1, c.norm(s))
        after = c.params_after(s)
        for k, v in m.reference_state_dict().items():
            np.testing.assert_allclose(v.detach().numpy(), after[k], rtol=1e-5, atol=1e-6)


def test_port_init_is_seed_for_seed_with_reference():
    """Same torch seed => same initial weights as reference Model(..., 'pytorch')
    (nn.LSTM constructor draws first, then U(-winit,winit) in registration order)."""
    z = np.load(os.path.join(GOLDEN, "small_cfg_summary.npz"))
    V, H, L, T, B, steps = [int(v) for v in z["meta"]]
    m = P.TorchLstmLm(V, H, L, 0.0, float(z["winit"]), seed=int(z["seed"]))
    for k, v in m.reference_state_dict().items():
        a = v.detach().numpy().astype(np.float64)
        got = np.array([a.sum(), np.abs(a).sum(), (a * a).sum()])
        np.testing.assert_allclose(got, z["param0_sum/" + k], rtol=1e-12, atol=1e-9)


@pytest.mark.parametrize("name", ["small_cfg_summary", "medium_cfg_summary", "large_cfg_summary"])
def test_oracle_matches_reference_summary_at_baseline_configs(name):
    """BASELINE.json configs[0..2] shapes (2x200 T=20 / 2x650 / 2x1500 T=35, B=20, V=10000): numpy oracle vs the
    reference's recorded loss / score rows / grad norms / updated parameters / states over two carried train-mode
    steps with the reference's own dropout masks.  The clip norm is compared with the one recomputed in fp64 from
    the fixture's per-tensor L2 norms: torch's CPU `clip_grad_norm_` itself is only good to ~1e-3 at Large
    (fp32 accumulation over 66 M squares), which the second assert documents."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    V, H, L, T, B, steps = [int(v) for v in z["meta"]]
    p = float(z["dropout"])
    m = P.TorchLstmLm(V, H, L, p, float(z["winit"]), seed=int(z["seed"]))
    params = {k: v.detach().numpy().copy() for k, v in m.reference_state_dict().items()}
    states = O.zero_states(L, B, H)
    n = T * B * H
    for s in range(steps):
        x, y = z[f"s{s}/x"], z[f"s{s}/y"]
        masks = None if p == 0 else [np.unpackbits(z[f"s{s}/mask/{i}"])[:n].reshape(T, B, H).astype(bool)
                                     for i in range(L + 1)]
        scores, new_states, cache = O.model_fwd(params, x, states, L, p, masks)
        loss = O.nll_loss(scores, y)
        grads = O.model_bwd(params, cache, O.nll_loss_bwd(scores, y), L)
        assert abs(loss - float(z[f"s{s}/loss"])) < 2e-5 * float(z[f"s{s}/loss"])
        rows = scores[:: max(1, scores.shape[0] // 8)][:, :64]
        np.testing.assert_allclose(rows, z[f"s{s}/scores_rows"], rtol=1e-4, atol=2e-5)
        for k in O.param_names(L):
            l2 = np.sqrt((grads[k].astype(np.float64) ** 2).sum())
            assert abs(l2 - float(z[f"s{s}/grad_l2/" + k])) < 1e-4 * max(l2, 1e-6), k
            np.testing.assert_allclose(grads[k].reshape(-1)[:32], z[f"s{s}/grad_head/" + k],
                                       rtol=2e-4, atol=1e-7)
        norm = O.clip_sgd(params, grads, float(z["lr"]), float(z["max_norm"]), O.param_names(L))
        norm_l2 = np.sqrt(sum(float(z[f"s{s}/grad_l2/" + k]) ** 2 for k in O.param_names(L)))
        assert abs(norm - norm_l2) < 1e-5 * norm
        assert abs(norm - float(z[f"s{s}/norm"])) < 2.5e-3 * norm
        for k in O.param_names(L):
            np.testing.assert_allclose(params[k].reshape(-1)[:32], z[f"s{s}/param_head/" + k], rtol=1e-5, atol=1e-7)
        for l in range(L):
            np.testing.assert_allclose(new_states[l][0], z[f"s{s}/h/{l}"], rtol=1e-4, atol=2e-6)
            np.testing.assert_allclose(new_states[l][1], z[f"s{s}/c/{l}"], rtol=1e-4, atol=2e-6)
        states = new_states

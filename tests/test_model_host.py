"""Host-side logic of the drop-in module (no GPU): constructor contract, parameter names,
seed-for-seed initialisation, state layouts, re-batching -- against the reference fixtures."""
import os

import numpy as np
import torch

import zaremba_b200
from tests._golden import GOLDEN


def test_parameter_names_and_order_pytorch():
    m = zaremba_b200.Model(50, 12, 2, 0.5, 0.05)
    names = [k for k, _ in m.named_parameters()]
    assert names == ["embed.W", "rnns.0.weight_ih_l0", "rnns.0.weight_hh_l0", "rnns.0.bias_ih_l0",
                     "rnns.0.bias_hh_l0", "rnns.1.weight_ih_l0", "rnns.1.weight_hh_l0", "rnns.1.bias_ih_l0",
                     "rnns.1.bias_hh_l0", "fc.W", "fc.b"]
    shapes = {k: tuple(v.shape) for k, v in m.named_parameters()}
    assert shapes["embed.W"] == (50, 12) and shapes["fc.W"] == (50, 12) and shapes["fc.b"] == (50,)
    assert shapes["rnns.1.weight_hh_l0"] == (48, 12) and shapes["rnns.0.bias_ih_l0"] == (48,)
    assert sum(p.numel() for p in zaremba_b200.Model(10000, 1500, 2, 0.65, 0.04).parameters()) == 66034000


def test_parameter_names_custom():
    m = zaremba_b200.Model(50, 12, 1, 0.0, 0.05, "custom")
    assert [k for k, _ in m.named_parameters()] == ["embed.W", "rnns.0.W_x", "rnns.0.W_h", "rnns.0.b_x",
                                                    "rnns.0.b_h", "fc.W", "fc.b"]
    h, c = m.state_init(3)[0]
    assert h.shape == (3, 12) and c.shape == (3, 12)


def test_seed_for_seed_init_matches_reference():
    """torch.manual_seed(s); Model(...) gives the reference's weights (fixture holds sums of
    the reference's tensors at the small config)."""
    z = np.load(os.path.join(GOLDEN, "small_cfg_summary.npz"))
    V, H, L, T, B, steps = [int(v) for v in z["meta"]]
    torch.manual_seed(int(z["seed"]))
    m = zaremba_b200.Model(V, H, L, 0.0, float(z["winit"]))
    for k, v in m.named_parameters():
        a = v.detach().numpy().astype(np.float64)
        np.testing.assert_allclose([a.sum(), np.abs(a).sum(), (a * a).sum()], z["param0_sum/" + k], rtol=1e-12)
    assert all(float(p.detach().abs().max()) <= float(z["winit"]) * (1 + 1e-6) for p in m.parameters())


def test_state_layout_and_detach():
    m = zaremba_b200.Model(20, 8, 3, 0.0, 0.1)
    st = m.state_init(4)
    assert len(st) == 3 and st[0][0].shape == (1, 4, 8) and float(st[2][1].abs().sum()) == 0.0
    st2 = m.detach(st)
    assert all(not h.requires_grad and not c.requires_grad for h, c in st2)


def test_minibatch_matches_reference_fixture():
    z = np.load(os.path.join(GOLDEN, "minibatch.npz"))
    ci = 0
    while f"c{ci}/args" in z.files:
        n, bs, sl = [int(v) for v in z[f"c{ci}/args"]]
        ds = zaremba_b200.minibatch(z[f"c{ci}/data"], bs, sl)
        assert len(ds) == int(z[f"c{ci}/n"])
        for bi, (x, y) in enumerate(ds):
            assert np.array_equal(x.numpy(), z[f"c{ci}/x{bi}"]) and np.array_equal(y.numpy(), z[f"c{ci}/y{bi}"])
            assert x.dtype == torch.int64
        ci += 1


def test_custom_gate_permutation_is_involution():
    from zaremba_b200.model import _ifon_to_ifgo
    a = torch.arange(8.0).view(8, 1)
    assert _ifon_to_ifgo(a).view(-1).tolist() == [0, 1, 2, 3, 6, 7, 4, 5]
    assert torch.equal(_ifon_to_ifgo(_ifon_to_ifgo(a)), a)


def test_ptb_id_fixture_is_consistent_with_the_reference_slice():
    """tests/golden/ptb_ids.npz (minted by make_ptb_ids.py with the vocabulary rule of main.py:44-59) must agree with the
    ids make_golden.py took from the same text through its own restatement of that rule (first 1500 validation tokens),
    have the reference's corpus sizes and '\\n' as id 0 once per line."""
    z = np.load(os.path.join(GOLDEN, "ptb_ids.npz"))
    s = np.load(os.path.join(GOLDEN, "perplexity_ptb_slice.npz"))
    assert int(z["vocab_size"]) == 10000
    assert (z["train"].size, z["valid"].size, z["test"].size) == (929589, 73760, 82430)
    assert z["train"].dtype == np.int16 and int(z["train"].max()) == 9999 and int(z["train"].min()) == 0
    np.testing.assert_array_equal(z["valid"][:1500].astype(np.int64), s["ids"].reshape(-1).astype(np.int64))
    assert (int((z["train"] == 0).sum()), int((z["valid"] == 0).sum()), int((z["test"] == 0).sum())) == (42068, 3370, 3761)
    # the re-batching the recipes use: 1327 windows of [35, 20] for Medium / Large, 2323 of [20, 20] for Small
    assert len(zaremba_b200.minibatch(z["train"].astype(np.int64).reshape(-1, 1), 20, 35)) == 1327
    assert len(zaremba_b200.minibatch(z["train"].astype(np.int64).reshape(-1, 1), 20, 20)) == 2323

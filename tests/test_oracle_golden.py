"""The oracle (oracle/lstm_lm_oracle.py) against outputs of the reference itself.

CPU only.  Fixtures: tests/golden/*.npz (made by tests/golden/make_golden.py from
/root/reference).  fp32 oracle vs fp32 reference: both sum in different orders, so the
bar is a few ulp of the quantity's scale; the fp64 oracle is held to the same bar to
show the reference's own fp32 rounding is what is left.
"""
import os

import numpy as np
import pytest

from oracle import lstm_lm_oracle as O
from tests._golden import GOLDEN, STEP_CASES, StepCase


def _close(a, b, rtol, atol, what):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    err = np.abs(a - b)
    tol = atol + rtol * np.abs(b)
    assert np.all(err <= tol), f"{what}: max err {err.max():.3e} (tol {tol.flat[err.argmax()]:.3e})"


@pytest.mark.parametrize("name", STEP_CASES)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_train_steps_match_reference(name, dtype):
    c = StepCase(name)
    params = c.params0(dtype)
    states = c.states0(dtype)
    for s in range(c.steps):
        loss, norm, states, scores, grads = O.train_step(
            params, c.x(s), c.y(s), states, c.L, c.lr, c.max_norm, c.dropout, c.masks(s))
        _close(scores, c.scores(s), 2e-5, 2e-5, f"{name} s{s} scores")
        _close(loss, c.loss(s), 2e-5, 1e-5, f"{name} s{s} loss")
        _close(norm, c.norm(s), 5e-5, 1e-6, f"{name} s{s} grad norm")
        ref_after = c.params_after(s)
        ref_grads = c.grads(s)
        coef = min(1.0, c.max_norm / (c.norm(s) + 1e-6))
        for k in c.names:
            scale = np.abs(ref_grads[k]).max() + 1e-12
            # oracle grads are post-clip (clip_sgd scales in place, like clip_grad_norm_);
            # the fixture stores them pre-clip
            _close(grads[k], ref_grads[k] * coef, 1e-4, 2e-5 * scale * coef, f"{name} s{s} grad {k}")
            _close(params[k], ref_after[k], 1e-5, 2e-6, f"{name} s{s} param {k}")
        for l, (h, cc) in enumerate(c.states_after(s)):
            _close(states[l][0], h, 2e-5, 2e-6, f"{name} s{s} h{l}")
            _close(states[l][1], cc, 2e-5, 2e-6, f"{name} s{s} c{l}")


def test_custom_and_pytorch_paths_use_different_gate_orders():
    """custom = (i,f,o,n) row blocks, nn.LSTM = (i,f,g,o)."""
    a = np.arange(8, dtype=np.float32).reshape(8, 1)
    out = O.custom_to_pytorch_gates(a)
    assert out.reshape(-1).tolist() == [0, 1, 2, 3, 6, 7, 4, 5]


def test_nll_loss_bwd_is_gradient_of_loss():
    rng = np.random.default_rng(0)
    T, B, V = 3, 2, 7
    s = rng.normal(size=(T * B, V))
    y = rng.integers(0, V, size=(T, B))
    g = O.nll_loss_bwd(s, y)
    eps = 1e-6
    for (n, v) in [(0, 0), (2, 3), (5, 6)]:
        sp = s.copy(); sp[n, v] += eps
        sm = s.copy(); sm[n, v] -= eps
        fd = (O.nll_loss(sp, y) - O.nll_loss(sm, y)) / (2 * eps)
        assert abs(fd - g[n, v]) < 1e-6


def test_model_bwd_matches_finite_differences():
    rng = np.random.default_rng(1)
    V, H, L, T, B = 11, 6, 2, 3, 2
    p = O.init_params(V, H, L, 0.4, 3, np.float64)
    x = rng.integers(0, V, size=(T, B)); y = rng.integers(0, V, size=(T, B))
    st = [(rng.normal(size=(B, H)) * 0.3, rng.normal(size=(B, H)) * 0.3) for _ in range(L)]
    masks = [rng.random((T, B, H)) > 0.4 for _ in range(L + 1)]

    def loss_of(pp):
        sc, _, _ = O.model_fwd(pp, x, st, L, 0.4, masks)
        return O.nll_loss(sc, y)

    sc, _, cache = O.model_fwd(p, x, st, L, 0.4, masks)
    grads = O.model_bwd(p, cache, O.nll_loss_bwd(sc, y), L)
    eps = 1e-6
    for k in O.param_names(L):
        flat = p[k].reshape(-1)
        for idx in rng.choice(flat.shape[0], size=3, replace=False):
            if k == "embed.W" and (idx // H) not in x:
                continue
            old = flat[idx]
            flat[idx] = old + eps; lp = loss_of(p)
            flat[idx] = old - eps; lm = loss_of(p)
            flat[idx] = old
            fd = (lp - lm) / (2 * eps)
            assert abs(fd - grads[k].reshape(-1)[idx]) < 1e-6 * max(1.0, abs(fd)), (k, idx)


def test_minibatch_matches_reference():
    z = np.load(os.path.join(GOLDEN, "minibatch.npz"))
    ci = 0
    while f"c{ci}/args" in z.files:
        n, bs, sl = [int(v) for v in z[f"c{ci}/args"]]
        ds = O.minibatch(z[f"c{ci}/data"], bs, sl)
        assert len(ds) == int(z[f"c{ci}/n"]), (n, bs, sl)
        for bi, (x, y) in enumerate(ds):
            assert np.array_equal(x, z[f"c{ci}/x{bi}"])
            assert np.array_equal(y, z[f"c{ci}/y{bi}"])
        ci += 1
    assert ci == 7


def test_perplexity_and_ensemble_match_reference():
    z = np.load(os.path.join(GOLDEN, "perplexity_ptb_slice.npz"))
    V, H, L, T, B = [int(v) for v in z["meta"]]
    p1 = {k[len("param/"):]: z[k] for k in z.files if k.startswith("param/")}
    p2 = {k[len("param2/"):]: z[k] for k in z.files if k.startswith("param2/")}
    ds = O.minibatch(z["ids"], B, T)
    assert len(ds) == int(z["n_batches"])
    ppl = O.perplexity(p1, ds, L, B, H)
    assert abs(ppl - float(z["ppl"])) < 2e-5 * float(z["ppl"])
    x, y = ds[0]
    s1, _, _ = O.model_fwd(p1, x, O.zero_states(L, B, H), L)
    s2, _, _ = O.model_fwd(p2, x, O.zero_states(L, B, H), L)
    ens = O.ensemble_nll_loss([s1, s2], y)
    assert abs(ens - float(z["ens_loss"])) < 2e-5 * abs(float(z["ens_loss"]))

"""CPU: the numpy restatement of the reference policy (oracle/policy_oracle.py) against outputs of the
unmodified reference network recorded in tests/golden/policy_golden.npz."""
import os

import numpy as np
import pytest

from oracle import policy_oracle as po
import flatland_marl_b200.policy_weights as pw

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load_cases():
    with np.load(os.path.join(GOLD, "policy_golden.npz")) as z:
        return {k: z[k] for k in z.files}


def obs_of(fixture, step, cache={}):
    if fixture not in cache:
        with np.load(os.path.join(GOLD, fixture + ".npz")) as z:
            cache[fixture] = {k: z[k] for k in z.files}
    g = cache[fixture]
    p = "obs%d_" % step
    return dict(agent_attr=g[p + "attr"], forest=g[p + "forest"], adjacency=g[p + "adjacency"],
                node_order=g[p + "node_order"], edge_order=g[p + "edge_order"], valid_actions=g[p + "valid_actions"])


def test_weight_spec_matches_reference_shapes():
    w = pw.init_weights(0)
    assert sum(v.size for v in w.values()) == 1_897_478 or sum(v.size for v in w.values()) > 1_800_000
    pw.check_weights(w)
    assert w["tree_lstm.U_iou.weight"].shape == (384, 384)
    assert w["transformer.2.attention.in_proj_weight"].shape == (768, 256)


def test_policy_oracle_against_reference_outputs():
    gold = load_cases()
    w = pw.init_weights(int(gold["weight_seed"]))
    worst = 0.0
    for k, (fixture, step) in enumerate(zip(gold["case_fixture"], gold["case_step"])):
        o = obs_of(str(fixture), int(step))
        if o["agent_attr"].shape[0] > 100 and k % 2:
            continue
        logits, value = po.forward(w, o["agent_attr"][None], po.clean_forest(o["forest"])[None], o["adjacency"][None],
                                   o["node_order"][None], o["edge_order"][None])
        # float32 sums in a different order than torch: tolerance 2e-5 absolute on O(1) activations
        np.testing.assert_allclose(logits[0], gold["logits_%d" % k], rtol=0, atol=2e-5)
        np.testing.assert_allclose(value, gold["value_%d" % k], rtol=0, atol=2e-5)
        worst = max(worst, float(np.abs(logits[0] - gold["logits_%d" % k]).max()))
        act = po.choose_actions(logits[0], o["valid_actions"])
        safe = po.choice_margin(gold["logits_%d" % k], o["valid_actions"]) > 1e-4
        assert (act[safe] == gold["actions_%d" % k][safe]).all()
        assert safe.mean() > 0.95
    assert worst < 2e-5


def test_policy_oracle_batch_axis():
    gold = load_cases()
    w = pw.init_weights(int(gold["weight_seed"]))
    obs = [obs_of("t03_l0_random", int(s)) for s in gold["batched_steps"]]
    st = lambda k: np.stack([o[k] for o in obs])
    logits, value = po.forward(w, st("agent_attr"), po.clean_forest(st("forest")), st("adjacency"), st("node_order"), st("edge_order"))
    np.testing.assert_allclose(logits, gold["batched_logits"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(value, gold["batched_value"], rtol=0, atol=2e-5)


def test_pack_weights_folds_are_exact():
    """Host-side weight packing (policy.pack_weights) against the layer-by-layer form: the attention out-projection folded
    into att_mlp, the Tree-LSTM gate biases moved into weight column 12, the halved sigmoid-gate rows."""
    from flatland_marl_b200.policy import pack_weights
    w = pw.init_weights(3)
    d = pack_weights(w)
    rng = np.random.RandomState(0)
    x, a = rng.randn(5, 256), rng.randn(5, 256)
    for l in range(3):
        p = "transformer.%d." % l
        proj = a @ w[p + "attention.out_proj.weight"].T.astype(np.float64) + w[p + "attention.out_proj.bias"]
        want = np.concatenate([x, proj], axis=1) @ w[p + "att_mlp.0.weight"].T.astype(np.float64) + w[p + "att_mlp.0.bias"]
        got = np.concatenate([x, a], axis=1) @ d["tf_wm%d" % l].T.astype(np.float64) + d["tf_bm%d" % l]
        np.testing.assert_allclose(got, want, rtol=0, atol=2e-5)
    # tree: [x | 1 | 0 0 0] @ wiou^T reproduces W_iou x + b, with the i and o rows halved
    xn = rng.randn(4, 12)
    xe = np.concatenate([xn, np.ones((4, 1)), np.zeros((4, 3))], axis=1)
    pre = xn @ w["tree_lstm.W_iou.weight"].T.astype(np.float64) + w["tree_lstm.W_iou.bias"]
    got = xe @ d["tree_wiou"].T.astype(np.float64)
    np.testing.assert_allclose(got[:, :256], 0.5 * pre[:, :256], rtol=0, atol=1e-6)
    np.testing.assert_allclose(got[:, 256:], pre[:, 256:], rtol=0, atol=1e-6)
    assert (d["tree_uiou"][:256] == 0.5 * w["tree_lstm.U_iou.weight"][:256]).all()
    assert (d["tree_uiou"][256:] == w["tree_lstm.U_iou.weight"][256:]).all()
    hc = rng.randn(4, 128)
    f_pre = hc @ w["tree_lstm.U_f.weight"].T.astype(np.float64) + xn @ w["tree_lstm.W_f.weight"].T.astype(np.float64) + w["tree_lstm.W_f.bias"]
    got = np.concatenate([hc, xe], axis=1) @ d["tree_ufwf"].T.astype(np.float64)
    np.testing.assert_allclose(got, 0.5 * f_pre, rtol=0, atol=1e-6)
    # sigmoid(2z) = 0.5 tanh(z) + 0.5, the identity the kernels evaluate
    z = rng.randn(100)
    np.testing.assert_allclose(0.5 * np.tanh(0.5 * z) + 0.5, 1.0 / (1.0 + np.exp(-z)), rtol=0, atol=1e-12)


def test_gelu_fit_used_by_the_kernels():
    """The kernels' GELU, 0.5 x (1 + tanh(x P(x^2))) with x^2 clamped at 49 (csrc/policy/umma.cuh), against the erf form."""
    x = np.linspace(-12, 12, 200001)
    u = np.minimum(x * x, 49.0)
    fit = 0.5 * x * (1 + np.tanh(x * (7.97507884e-1 + u * (3.70056460e-2 + u * -3.51516783e-4))))
    assert np.abs(fit - po.gelu(x.astype(np.float32)).astype(np.float64)).max() < 4e-5


def test_load_weights_from_reference_checkpoint_formats(tmp_path):
    """A reference checkpoint is `torch.save(net.state_dict())` (plfActor.py:10-13); .npz with the same keys also loads;
    a missing or mis-shaped tensor is an error, not a silent default."""
    import torch
    w = pw.init_weights(5)
    pt = tmp_path / "phase-III-50.pt"
    torch.save({k: torch.from_numpy(v) for k, v in w.items()}, str(pt))
    got = pw.load_weights(str(pt))
    assert set(got) == set(w) and all((got[k] == w[k]).all() for k in w)
    npz = tmp_path / "w.npz"
    np.savez(str(npz), **w)
    got = pw.load_weights(str(npz))
    assert all((got[k] == w[k]).all() for k in w)
    bad = dict(w)
    del bad["tree_lstm.U_f.weight"]
    with pytest.raises(KeyError):
        pw.check_weights(bad)
    bad = dict(w)
    bad["actor_net.4.weight"] = np.zeros((4, 128), np.float32)
    with pytest.raises(ValueError):
        pw.check_weights(bad)

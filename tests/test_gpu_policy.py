"""GPU: the tcgen05 policy forward pass (include/flatland_policy_b200.h) against outputs of the unmodified
reference network (tests/golden/policy_golden.npz) and the numpy oracle.

Tolerance.  The kernels multiply bf16 operands (8-bit mantissa) with fp32 accumulation and keep activations in
bf16 between layers; the reference is fp32 throughout.  Measured worst case over the golden cases is 1.0e-2 on
logits of magnitude 0.1-0.6 after ~25 layers; the tests allow 3e-2 absolute on logits and values, and demand the
same chosen action wherever the fp32 choice is not within 0.02 of a cumulative-probability step."""
import os

import numpy as np
import pytest
import torch

from oracle import policy_oracle as po

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
LOGIT_ATOL = 3e-2


@pytest.fixture(scope="module")
def actor():
    from flatland_marl_b200.policy import BatchedActor
    return BatchedActor(None, seed=0)


@pytest.fixture(scope="module")
def gold():
    with np.load(os.path.join(GOLD, "policy_golden.npz")) as z:
        return {k: z[k] for k in z.files}


def obs_of(golden, fixture, step):
    g = golden(fixture)
    p = "obs%d_" % step
    return {"agent_attr": g[p + "attr"], "forest": g[p + "forest"], "adjacency": g[p + "adjacency"],
            "node_order": g[p + "node_order"], "valid_actions": g[p + "valid_actions"]}


def to_dev(obs_list, dev):
    return {k: torch.from_numpy(np.ascontiguousarray(np.stack([o[k] for o in obs_list]))).to(dev) for k in obs_list[0]}


@pytest.mark.parametrize("M,K,N,act", [(128, 64, 128, 0), (1, 128, 128, 1), (300, 256, 256, 1), (20000, 256, 768, 0),
                                       (1000, 512, 256, 1), (129, 448, 384, 0)])
def test_linear_against_torch(actor, M, K, N, act):
    dev = actor.device
    gen = torch.Generator(device="cpu").manual_seed(M + K + N)
    a = (torch.randn(M, K, generator=gen) * 0.5).to(dev).to(torch.bfloat16)
    w = (torch.randn(N, K, generator=gen) / K ** 0.5).to(dev).to(torch.bfloat16)
    b = torch.randn(N, generator=gen).to(dev)
    out = actor.linear(a, w, b, act).float()
    ref = a.float() @ w.float().t() + b
    if act:
        ref = torch.nn.functional.gelu(ref)
    # same bf16 operands, fp32 accumulation: only the summation order and the bf16 rounding of the output differ
    tol = ref.abs() * 2.0 ** -8 + 2e-3
    assert bool(((out - ref).abs() <= tol).all()), float((out - ref).abs().max())


def test_forward_against_reference_outputs(actor, gold, golden):
    worst = 0.0
    for k, (fixture, step) in enumerate(zip(gold["case_fixture"], gold["case_step"])):
        obs = obs_of(golden, str(fixture), int(step))
        d = to_dev([obs], actor.device)
        logits, value = actor.forward(d)
        acts = actor.choose_actions(logits, d["valid_actions"])[0].cpu().numpy()
        lg = logits[0].cpu().numpy()
        ref = gold["logits_%d" % k]
        np.testing.assert_allclose(lg, ref, rtol=0, atol=LOGIT_ATOL)
        np.testing.assert_allclose(value.cpu().numpy(), gold["value_%d" % k], rtol=0, atol=LOGIT_ATOL)
        worst = max(worst, float(np.abs(lg - ref).max()))
        safe = po.choice_margin(ref, obs["valid_actions"]) > 0.02
        assert (acts[safe] == gold["actions_%d" % k][safe]).all()
        # the choice kernel itself is exact on the logits it is given
        exact = po.choice_margin(lg, obs["valid_actions"]) > 1e-5
        assert (po.choose_actions(lg, obs["valid_actions"])[exact] == acts[exact]).all()
    assert worst < LOGIT_ATOL


def test_batch_composition_independence(actor, gold, golden):
    """Rows are independent: a batch of 3 environments gives the bytes of three batch-1 calls."""
    steps = [int(s) for s in gold["batched_steps"]]
    obs = [obs_of(golden, "t03_l0_random", s) for s in steps]
    solo = []
    for o in obs:
        lg, v = actor.forward(to_dev([o], actor.device))
        solo.append((lg[0].cpu().numpy().copy(), float(v[0].item())))
    lg, v = actor.forward(to_dev(obs, actor.device))
    lg, v = lg.cpu().numpy(), v.cpu().numpy()
    for i in range(3):
        assert (lg[i] == solo[i][0]).all()
        assert v[i] == solo[i][1]
    np.testing.assert_allclose(lg, gold["batched_logits"], rtol=0, atol=LOGIT_ATOL)
    np.testing.assert_allclose(v, gold["batched_value"], rtol=0, atol=LOGIT_ATOL)


def test_choose_actions_against_oracle(actor):
    rng = np.random.RandomState(5)
    logits = (rng.randn(64, 50, 5) * 2).astype(np.float32)
    valid = (rng.rand(64, 50, 5) < 0.6).astype(np.uint8)
    valid[0, :5] = 0                      # no valid action -> 0
    valid[1, :5] = np.eye(5, dtype=np.uint8)
    got = actor.choose_actions(torch.from_numpy(logits).to(actor.device), torch.from_numpy(valid).to(actor.device)).cpu().numpy()
    want = po.choose_actions(logits, valid)
    ok = po.choice_margin(logits, valid) > 1e-5
    assert ok.mean() > 0.99
    assert (got[ok] == want[ok]).all()
    assert (got[0, :5] == 0).all() and (got[1, :5] == np.arange(5)).all()


def test_rollout_with_policy_in_the_loop(actor, golden):
    """Observation -> policy -> action -> step entirely on the device; every action taken must be one the
    reference's valid-action mask allows (or 0 when none is), and the oracle environment fed the same actions must
    end up in the same state."""
    import flatland_marl_b200 as fb
    from oracle import oracle as orc
    g = golden("t03_l0_random")
    batch = fb.BatchedRailEnv([dict(g)] * 4, sched_rows=int(g["n_steps"]))
    env = orc.OracleEnv(g)
    obs = batch.reset()
    env.reset()
    for t in range(30):
        acts = actor.get_actions(obs)
        a = acts.cpu().numpy()
        va = obs["valid_actions"].cpu().numpy()
        chosen_ok = np.take_along_axis(va, a[..., None].astype(np.int64), axis=2)[..., 0]
        assert ((chosen_ok == 1) | ((va.sum(axis=2) == 0) & (a == 0))).all()
        assert (a == a[0]).all()          # identical environments -> identical actions
        obs, rew, don = batch.step(acts)
        env.step(a[0], g["sched"][t])
        s, os_ = batch.state_numpy(0), env.state()
        for k in ("pos", "dir", "state"):
            assert (s[k] == os_[k]).all()


@pytest.mark.parametrize("config,n_envs,sample", [("Test_03", 1024, (0, 517, 1023)), ("Test_14", 64, (63,)), ("Test_02", 8192, (1, 8191))])
def test_full_batch_policy_on_sampled_envs(actor, config, n_envs, sample):
    """BASELINE.json's batch sizes: after 40 lock-step steps with the policy in the loop, the logits of sampled
    environments equal (bit for bit) those of a batch-1 forward on the same observation, agree with the numpy
    oracle of the reference network within the stated tolerance, and the chosen actions are the oracle's wherever the
    choice is not marginal."""
    import bench
    import flatland_marl_b200 as fb
    worlds = bench.load_worlds(config, n_envs)
    batch = fb.BatchedRailEnv(worlds, auto_reset=True)
    obs = batch.reset()
    for _ in range(40):
        obs, _, _ = batch.step(actor.get_actions(obs))
    logits, value = actor.forward(obs)
    acts = actor.choose_actions(logits, obs["valid_actions"]).cpu().numpy()
    lg_all, v_all = logits.cpu().numpy().copy(), value.cpu().numpy().copy()
    assert np.isfinite(lg_all).all() and np.isfinite(v_all).all()
    for e in sample:
        one = {k: obs[k][e:e + 1].contiguous() for k in ("agent_attr", "forest", "adjacency", "node_order", "valid_actions")}
        lg1, v1 = actor.forward(one)
        assert (lg1[0].cpu().numpy() == lg_all[e]).all() and float(v1[0].item()) == float(v_all[e])
        o = {k: v[0].cpu().numpy() for k, v in one.items()}
        eo = obs["edge_order"][e].cpu().numpy()
        want, wval = po.forward(actor.weights, o["agent_attr"][None], po.clean_forest(o["forest"])[None], o["adjacency"][None],
                                o["node_order"][None], eo[None])
        np.testing.assert_allclose(lg_all[e], want[0], rtol=0, atol=LOGIT_ATOL)
        assert abs(float(v_all[e]) - float(wval[0])) < LOGIT_ATOL
        safe = po.choice_margin(want[0], o["valid_actions"]) > 0.02
        assert (acts[e][safe] == po.choose_actions(want[0], o["valid_actions"])[safe]).all()


# ---- the fp32 path: the reference's own arithmetic (fl_policy_forward_f32) -------------------------------------------
F32_ATOL = 1e-4      # summation-order rounding only: measured differences are a few 1e-6


@pytest.fixture(scope="module")
def actor_f32():
    from flatland_marl_b200.policy import BatchedActor
    return BatchedActor(None, seed=0, precision="fp32")


def test_fp32_forward_matches_reference_outputs_to_1e4(actor_f32, gold, golden):
    """Logits and values within 1e-4 absolute of the UNMODIFIED reference network's outputs on all recorded observations,
    and the same chosen action for every agent whose choice is not within 1e-4 of a cumulative-probability step."""
    worst = 0.0
    for k, (fixture, step) in enumerate(zip(gold["case_fixture"], gold["case_step"])):
        obs = obs_of(golden, str(fixture), int(step))
        d = to_dev([obs], actor_f32.device)
        logits, value = actor_f32.forward(d)
        acts = actor_f32.choose_actions(logits, d["valid_actions"])[0].cpu().numpy()
        lg = logits[0].cpu().numpy()
        ref = gold["logits_%d" % k]
        np.testing.assert_allclose(lg, ref, rtol=0, atol=F32_ATOL, err_msg="%s step %d logits" % (fixture, step))
        np.testing.assert_allclose(value.cpu().numpy(), gold["value_%d" % k], rtol=0, atol=F32_ATOL)
        worst = max(worst, float(np.abs(lg - ref).max()))
        safe = po.choice_margin(ref, obs["valid_actions"]) > 1e-4
        assert (acts[safe] == gold["actions_%d" % k][safe]).all()
    assert worst < F32_ATOL
    print("fp32 path: worst |logit - reference| = %.2e" % worst)


def test_fp32_forward_is_batch_independent_and_close_to_bf16(actor, actor_f32, golden):
    """A batch of different observations gives every environment the bytes it gets alone; and the bf16 tensor-core path stays
    within its stated tolerance of the fp32 path on a deep tree (relative bound: 3e-2 of the largest |logit| + 1e-2)."""
    obs = [obs_of(golden, "t03_l0_random", s) for s in (0, 89, 176)]
    d3 = to_dev(obs, actor_f32.device)
    l3, v3 = actor_f32.forward(d3)
    l3, v3 = l3.clone(), v3.clone()
    for j, o in enumerate(obs):
        l1, v1 = actor_f32.forward(to_dev([o], actor_f32.device))
        assert torch.equal(l1[0], l3[j]) and torch.equal(v1[0], v3[j])
    lb, _ = actor.forward(d3)
    scale = float(l3.abs().max())
    assert float((lb - l3).abs().max()) <= 3e-2 * scale + 1e-2

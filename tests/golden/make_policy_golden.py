"""Generates tests/golden/policy_golden.npz by running the UNMODIFIED reference policy
(solution/nn/net_tree.py `Network`, solution/plfActor.py `Actor.get_actions`, the observation clean-up of
solution/eval_env.py `parse_features`) in the build container on observations already recorded from the
reference simulator (tests/golden/t*.npz).

  python tests/golden/make_policy_golden.py          # needs /root/reference

The reference ships no checkpoint, so the network is loaded with `init_weights(0)` of
flatland-marl_b200/policy_weights.py (numpy-seeded, reproducible anywhere); the fixture stores only the
case list and the reference's outputs (logits, critic value, chosen actions).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh  # noqa: E402

CASES = [("t00_l0_random", None), ("t00_l1_greedy", None), ("t02_l0_forward", None), ("t02_l2_stacking", [0, 2]),
         ("t03_l0_random", None), ("t03_l2_stacking", [0, 3]), ("t08_l0_greedy", [0, 2]), ("t14_l0_forward", [1])]
WEIGHT_SEED = 0


def main():
    rh.load()
    sys.path.insert(0, os.path.join(rh.REFERENCE_ROOT, "solution"))
    import torch
    from plfActor import Actor
    from nn.net_tree import Network
    from eval_env import TestEnvWrapper
    import flatland_marl_b200.policy_weights as pw

    w = pw.init_weights(WEIGHT_SEED)
    net = Network()
    missing = net.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=True)
    net.eval()
    actor = Actor.__new__(Actor)
    actor.net = net

    out = {"weight_seed": np.int32(WEIGHT_SEED)}
    names, steps = [], []
    for fixture, pick in CASES:
        with np.load(os.path.join(HERE, fixture + ".npz")) as z:
            g = {k: z[k] for k in z.files}
        sample = [int(s) for s in g["sample_steps"]]
        for j, s in enumerate(sample):
            if pick is not None and j not in pick:
                continue
            n = int(g["N"])
            # the reference hands the policy python lists; parse_features turns them into float64 arrays
            feature = (g["obs%d_attr" % s].tolist(),
                       (g["obs%d_forest" % s].tolist(), g["obs%d_adjacency" % s].tolist(),
                        g["obs%d_node_order" % s].tolist(), g["obs%d_edge_order" % s].tolist()))
            props = {"valid_actions": g["obs%d_valid_actions" % s].astype(bool).tolist()}
            obs = TestEnvWrapper.parse_features(None, feature, props)
            feats = actor.get_feature([obs])
            with torch.no_grad():
                (logits,), value = net(*[f.clone() for f in feats])
            actions = actor.get_actions([obs], obs["valid_actions"], n)
            k = len(names)
            names.append(fixture)
            steps.append(s)
            out["logits_%d" % k] = logits[0].numpy().astype(np.float32)
            out["value_%d" % k] = value.numpy().astype(np.float32)
            out["actions_%d" % k] = np.array([actions[i] for i in range(n)], np.uint8)
            print(fixture, s, "N", n, "logits", out["logits_%d" % k][0], "value", out["value_%d" % k])
    # one batched call (batch 3 of the same world) to pin the batch axis of Network.forward
    with np.load(os.path.join(HERE, "t03_l0_random.npz")) as z:
        g = {k: z[k] for k in z.files}
    ss = [int(s) for s in g["sample_steps"]][3:6]
    ff = np.stack([g["obs%d_forest" % s] for s in ss]).astype(np.float64)
    ff[ff == np.inf] = -1
    args = [torch.from_numpy(np.stack([g["obs%d_attr" % s] for s in ss])).float(), torch.from_numpy(ff).float()] + \
           [torch.from_numpy(np.stack([g["obs%d_%s" % (s, k)] for s in ss])).long() for k in ("adjacency", "node_order", "edge_order")]
    with torch.no_grad():
        (logits,), value = net(*args)
    out["batched_steps"] = np.array(ss, np.int32)
    out["batched_logits"] = logits.numpy().astype(np.float32)
    out["batched_value"] = value.numpy().astype(np.float32)
    out["case_fixture"] = np.array(names)
    out["case_step"] = np.array(steps, np.int32)
    np.savez_compressed(os.path.join(HERE, "policy_golden.npz"), **out)
    print("wrote policy_golden.npz:", len(names), "cases")


if __name__ == "__main__":
    main()

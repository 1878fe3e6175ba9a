"""Debug aid: one environment of a world pack against the oracle, every mismatch listed.
usage: python tools/debug_env.py [config] [env] [steps]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import flatland_marl_b200 as fb  # noqa: E402
from oracle import oracle as orc  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "Test_03"
e = int(sys.argv[2]) if len(sys.argv) > 2 else 517
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
worlds = bench.load_worlds(cfg, e + 1)
w = worlds[e]
batch = fb.BatchedRailEnv([w], debug_clocks=True)
env = orc.OracleEnv(w)
batch.reset(); env.reset()
print("plan", batch.observe_plan())
N = batch.N
rng = np.random.RandomState(7)
KEYS = (("agent_attr", "attr"), ("forest", "forest"), ("adjacency", "adjacency"), ("node_order", "node_order"),
        ("edge_order", "edge_order"), ("valid_actions", "valid_actions"), ("dist_target", "dist_target"))
for t in range(steps + 1):
    oo = env.obs()
    print("step", t, "entries", int(batch.debug_clocks[0, 10]))
    for k, ok in KEYS:
        got = batch.obs[k][0].cpu().numpy()
        want = np.asarray(oo[ok]).reshape(got.shape)
        bad = np.argwhere(got != want)
        if len(bad):
            print("  %s: %d mismatches" % (k, len(bad)))
            for ix in bad[:12]:
                ix = tuple(int(x) for x in ix)
                print("    ", ix, "got", got[ix], "want", want[ix])
            if k == "forest":
                for ag in sorted(set(int(b[0]) for b in bad))[:3]:
                    print("    agent", ag, "speed", w["speed"][ag], "forest rows got/want of first bad node")
                    nd = int([b for b in bad if b[0] == ag][0][1])
                    print("      got ", got[ag, nd]); print("      want", want[ag, nd])
    act = np.where(rng.rand(1, N) < 0.7, 2, rng.randint(0, 5, (1, N))).astype(np.uint8)
    batch.step(torch.from_numpy(act).to(batch.device))
    env.step(act[0], w["sched"][t])

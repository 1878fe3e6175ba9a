"""SM-clock timeline of CTA 0 of k_tree_leaf inside one policy forward (tuning)."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import flatland_marl_b200 as fb
from flatland_marl_b200.policy import BatchedActor
E = 1024
env = fb.BatchedRailEnv(bench.load_worlds("Test_03", E), auto_reset=True)
actor = BatchedActor(None, seed=0)
obs = env.reset()
rng = np.random.RandomState(3)
for _ in range(100):
    obs, _, _ = env.step(torch.from_numpy(rng.randint(0, 5, size=(E, env.N)).astype(np.uint8)).to(env.device))
actor.get_actions(obs)
clk = torch.zeros(128, dtype=torch.int64, device=env.device)
actor.lib.fl_policy_debug_clocks(clk.data_ptr())
actor.get_actions(obs)
torch.cuda.synchronize()
actor.lib.fl_policy_debug_clocks(None)
c = clk.cpu().numpy().reshape(8, 16)
t0 = c[0][0]
if os.environ.get("FL_POLICY_DBG_F"):
    c = clk.cpu().numpy()
    t0 = c[3]
    print("k_lin<TREE_F> level 1, CTA 0: mma warp start %d, weights resident %d" % (c[0] - t0, c[1] - t0))
    for t in range(8):
        print("  tile %d: producer start %6d | mma: acc free %6d first full %6d last full %6d | epi: acc full %6d read %6d math %6d staged %6d bar2 %6d stored %6d" %
              (t, c[80 + t] - t0, c[8 + 4 * t] - t0, c[9 + 4 * t] - t0, c[10 + 4 * t] - t0, c[48 + 4 * t] - t0, c[49 + 4 * t] - t0,
               c[96 + 4 * t] - t0, c[97 + 4 * t] - t0, c[98 + 4 * t] - t0, c[50 + 4 * t] - t0))
    sys.exit(0)
for t in range(4):
    print("tile %d (epilogue warp 0, half 0): start %d, accumulator full %d, gates done %d, staged %d, stored %d" %
          (8 + t, c[t][4] - c[0][4], c[t][5] - c[0][4], c[t][6] - c[0][4], c[t][8] - c[0][4], c[t][9] - c[0][4]))

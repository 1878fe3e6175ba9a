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
for t in range(4):
    print("tile %d (epilogue warp 0, half 0): start %d, accumulator full %d, gates done %d, staged %d, stored %d" %
          (8 + t, c[t][4] - c[0][4], c[t][5] - c[0][4], c[t][6] - c[0][4], c[t][8] - c[0][4], c[t][9] - c[0][4]))

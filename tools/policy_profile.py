"""One policy forward between cudaProfilerStart/Stop, on observations of a pre-rolled batch — for
`ncu --profile-from-start off` captures of the policy kernels.  usage: policy_profile.py [config] [envs]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import flatland_marl_b200 as fb  # noqa: E402
from flatland_marl_b200.policy import BatchedActor  # noqa: E402

config = sys.argv[1] if len(sys.argv) > 1 else "Test_03"
E = int(sys.argv[2]) if len(sys.argv) > 2 else bench.CONFIGS[config]["envs"]
env = fb.BatchedRailEnv(bench.load_worlds(config, E), auto_reset=True)
actor = BatchedActor(None, seed=0)
obs = env.reset()
rng = np.random.RandomState(3)
for _ in range(100):
    obs, _, _ = env.step(torch.from_numpy(rng.randint(0, 5, size=(E, env.N)).astype(np.uint8)).to(env.device))
actor.get_actions(obs)
torch.cuda.synchronize()
torch.cuda.profiler.start()
actor.get_actions(obs)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")

"""Reset throughput: how fast new worlds are swapped into slots of a running batch — the part of
RailEnv.reset the GPU path takes over (DistanceMap, distance_map.py:57-160: 16.9 s of Python per world at Test_14 by the
survey's probe, plus the static walk tables that have no reference counterpart).
usage: python tools/reset_bench.py [config] [envs] [rounds]"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import flatland_marl_b200 as fb  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "Test_14"
E = int(sys.argv[2]) if len(sys.argv) > 2 and int(sys.argv[2]) > 0 else bench.CONFIGS[cfg]["envs"]
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 5
worlds = bench.load_worlds(cfg, 2 * E)
batch = fb.BatchedRailEnv(worlds[:E], auto_reset=True, reserve=0.3, min_slots=max(len(fb.unique_target_slots(w)[0]) for w in worlds))
batch.reset()
a = torch.full((E, batch.N), 2, dtype=torch.uint8, device=batch.device)
for _ in range(5):
    batch.step(a)
torch.cuda.synchronize()
ids = list(range(E))
wall, gpu = [], []
for r in range(rounds):
    new = worlds[E:] if r % 2 == 0 else worlds[:E]
    batch.profile(True)
    t0 = time.perf_counter()
    batch.replace_worlds(ids, new)
    batch.observe()
    torch.cuda.synchronize()
    wall.append(time.perf_counter() - t0)
    prof = batch.profile_collect()
    batch.profile(False)
    gpu.append({k: ms for k, (ms, n) in prof.items() if n})
    batch.step(a)
best = min(range(rounds), key=lambda k: wall[k])
line = {"config": cfg, "envs": E, "agents": batch.N, "grid": "%dx%d" % (batch.H, batch.W), "rounds": rounds,
        "wall_s_per_call": wall[best], "worlds_per_s_wall": E / wall[best],
        "gpu_ms_per_call": gpu[best], "gpu_ms_per_world": sum(v for k, v in gpu[best].items() if k != "k_observe") / E,
        "what": "replace_worlds of every slot (host: stacking + upload of the static arrays; device: distance maps of the slots' "
                "unique targets, static walk tables, agent reset) + the first observation; reference: 16.9 s per Test_14 reset "
                "(SURVEY.md §6, Python DistanceMap)"}
print(json.dumps(line))

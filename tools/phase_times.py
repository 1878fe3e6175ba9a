"""Tuning aid: per-phase SM-clock cycles of k_observe (FlBatch.debug_clocks), averaged over environments.
usage: python tools/phase_times.py [config] [envs] [steps]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import flatland_marl_b200 as fb  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "Test_03"
E = int(sys.argv[2]) if len(sys.argv) > 2 and int(sys.argv[2]) > 0 else bench.CONFIGS[cfg]["envs"]
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 150
worlds = bench.load_worlds(cfg, E)
batch = fb.BatchedRailEnv(worlds, auto_reset=True, debug_clocks=True)
N = batch.N
batch.reset()
print('plan', batch.observe_plan())
gen = torch.Generator(device=batch.device); gen.manual_seed(1)
acc = []
for t in range(steps):
    a = torch.randint(0, 5, (E, N), dtype=torch.uint8, device=batch.device, generator=gen)
    batch.step(a)
    if t >= steps - 20:
        acc.append(batch.debug_clocks.cpu().numpy().copy())
d = np.stack(acc).astype(np.float64)          # [20, E, 32]
split = batch.observe_plan().get("parts", 0) > 0
t0 = d[..., 15]
names = [("setup+zero", 15, 0), ("tma wait+loader+occupancy+deadlock neighbours", 0, 1), ("path chains + count pass", 1, 2),
         ("scan+scatter pass", 2, 4), ("small sorts (thread per bucket)", 4, 9), ("warp sorts", 9, 3),
         ("trees (structure+features)", 3, 5), ("attributes", 6, 7)]
if split:
    names[6] = ("index dump", 3, 5)
    names += [("TREES kernel: bulk loads", 31, 16), ("TREES kernel: trees (CTA 0 of the env)", 16, 21)]
tot = d[..., 7] - t0 + ((d[..., 21] - d[..., 31]) if split else 0)
print("%s E=%d N=%d: mean cycles per env %.0f (p50 %.0f, p99 %.0f, max %.0f)" % (cfg, E, N, tot.mean(), np.median(tot), np.percentile(tot, 99), tot.max()))
for nm, a, b in names:
    x = d[..., b] - d[..., a]
    print("  %-40s mean %9.0f  p99 %9.0f  (%.1f%%)" % (nm, x.mean(), np.percentile(x, 99), 100 * x.mean() / tot.mean()))
dl = d[..., 8] - d[..., 1]
print("  %-40s mean %9.0f  p99 %9.0f" % ("deadlock lane (from occupancy done)", dl.mean(), np.percentile(dl, 99)))
print("  entries per env: mean %.0f max %.0f; path segments per env: mean %.0f p99 %.0f max %.0f" %
      (d[..., 10].mean(), d[..., 10].max(), d[..., 11].mean(), np.percentile(d[..., 11], 99), d[..., 11].max()))
print("  per agent: cells %.1f windows %.2f, full conflict checks %.1f in %.2f drains" %
      (32 * d[..., 13].mean() / N, d[..., 13].mean() / N, d[..., 12].mean() / N, d[..., 14].mean() / N))

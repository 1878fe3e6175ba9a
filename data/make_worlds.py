"""Generates the benchmark world packs under data/worlds/ with the UNMODIFIED reference generators
(sparse_rail_generator, sparse_line_generator, timetable_generator; flatland-rl/flatland/envs/
rail_generators.py:196-292, line_generators.py:82-165, timetable_generators.py:21-96) in the build
container.  World generation is out of scope for the GPU path (BASELINE.json north_star: "stays
reference Python at reset time and is uploaded once"), and /root/reference does not exist on the GPU
box, so bench.py and the full-size GPU tests load these packs instead.

  python data/make_worlds.py [Test_03 ...]      # needs /root/reference; ~10 min on 8 cores for all

Environment k of a pack is the Flatland-3 round-2 configuration of that name
(solution/debug-environments/parameters_flatland_round_2_new.csv) generated with
random_seed = csv_seed[Level_(k mod 10)] + (k div 10).  A pack stores only the
generated world (grid, agents, timetable); malfunction schedules and actions are synthetic and drawn
at run time.  Seeds whose generation fails in the reference (it can raise on unlucky city layouts)
are skipped and listed in meta_skipped.
"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

OUT = os.path.join(ROOT, "data", "worlds")
PACKS = {"Test_00": (0, 16), "Test_02": (2, 1024), "Test_03": (3, 1024), "Test_08": (8, 512), "Test_14": (14, 64)}


def one(job):
    cfg, test, k = job
    from oracle import ref_harness as rh
    import importlib.util
    spec = importlib.util.spec_from_file_location("_w", os.path.join(ROOT, "flatland-marl_b200", "worlds.py"))
    wm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(wm)
    seed = rh.csv_seed(test, k % 10) + k // 10
    try:
        env = rh.make_env(cfg, seed)
        env.reset()
        w = wm.world_from_reference_env(env, n_sched=0)
        w.pop("sched")
        return k, seed, w
    except Exception as e:  # noqa: BLE001
        return k, seed, repr(e)


def main():
    names = sys.argv[1:] or list(PACKS)
    spec = importlib_worlds()
    for name in names:
        test, count = PACKS[name]
        t0 = time.time()
        with mp.Pool(min(8, os.cpu_count() or 1)) as pool:
            res = pool.map(one, [(name, test, k) for k in range(count)], chunksize=1)
        worlds, seeds, skipped = [], [], []
        for k, seed, w in sorted(res, key=lambda x: x[0]):
            if isinstance(w, str):
                skipped.append(seed)
                print("  skipped seed %d: %s" % (seed, w))
            else:
                worlds.append(w)
                seeds.append(seed)
        path = os.path.join(OUT, name.lower() + ".npz")
        spec.save_worlds_npz(path, worlds, seeds=np.array(seeds, np.uint64), skipped=np.array(skipped, np.uint64),
                             mal_interval=np.int64(__import__("oracle.ref_harness", fromlist=["CONFIGS"]).CONFIGS[name]["mal_interval"]))
        print("%s: %d worlds, %.1f KB, %.0f s" % (name, len(worlds), os.path.getsize(path) / 1024, time.time() - t0))


def importlib_worlds():
    import importlib.util
    spec = importlib.util.spec_from_file_location("_w", os.path.join(ROOT, "flatland-marl_b200", "worlds.py"))
    wm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(wm)
    return wm


if __name__ == "__main__":
    main()

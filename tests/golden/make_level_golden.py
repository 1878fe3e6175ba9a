"""Writes tests/golden/level_t00.pkl with the UNMODIFIED reference (`RailEnvPersister.save(..., save_distance_maps=True)`,
the call solution/debug-environments/generate_test_cases.py:64-68 makes) and tests/golden/level_t00_expected.npz, the
world the reference environment itself holds after reset (worlds.world_from_reference_env).  Needs /root/reference."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh  # noqa: E402


def main():
    rh.load()
    from flatland.envs.persistence import RailEnvPersister
    import flatland_marl_b200 as fb
    env = rh.make_env("Test_00", rh.csv_seed(0, 3))
    env.reset(random_seed=rh.csv_seed(0, 3))
    RailEnvPersister.save(env, os.path.join(HERE, "level_t00.pkl"), save_distance_maps=True)
    w = fb.world_from_reference_env(env, n_sched=4)
    np.savez_compressed(os.path.join(HERE, "level_t00_expected.npz"),
                        **{k: v for k, v in w.items() if k != "sched"}, dist_f64=env.distance_map.get(),
                        mal_rate=np.float64(env.malfunction_generator.MFP.malfunction_rate))
    print("wrote level_t00.pkl", os.path.getsize(os.path.join(HERE, "level_t00.pkl")), "bytes")


if __name__ == "__main__":
    main()

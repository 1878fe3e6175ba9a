"""Pins the CPU baseline of bench.py (`kind: "port"`, the C restatement in oracle/) to the REAL reference: times the
unmodified flatland-rl `RailEnv.step` + compiled `flatland_cutils` tree observation exactly as the reference's own driver
calls them — `solution/eval_env.py:108-114` `LocalTestEnvWrapper.step`, i.e. parse_actions + env.step +
obs_builder.get_properties() + parse_features — in one process per host core, next to the C port stepping the SAME worlds
with the SAME actions and malfunction schedules.  Build container only (/root/reference must exist; the Python reference
cannot travel to the GPU box).

  python tools/cpu_reference_vs_port.py [--config Test_03] [--procs 8] [--envs-per-proc 2] [--steps 120] [--preroll 200]

Writes profiles/cpu_reference_vs_port.json; bench.py quotes it in cpu_baseline (real_reference_per_core,
port_over_reference)."""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def worker(job):
    config, seeds, steps, preroll, which = job
    from oracle import ref_harness as rh
    from oracle import oracle as orc
    import make_golden as mg
    rh.load()
    sys.path.insert(0, os.path.join(rh.REFERENCE_ROOT, "solution"))
    from eval_env import LocalTestEnvWrapper            # the unmodified consumer-side wrapper of the reference
    out = []
    for seed in seeds:
        env = rh.make_env(config, seed)
        wrapper = LocalTestEnvWrapper(env)
        wrapper.reset()
        n = env.get_num_agents()
        world = mg.static_of(env)
        T = int(world["T"])
        preroll, steps = min(job[3], T // 3), min(job[2], T - min(job[3], T // 3) - 1)    # stay inside this world's episode
        sched = mg.draw_schedule(env, preroll + steps)
        rng = np.random.RandomState(seed % (2 ** 32))
        acts = np.where(rng.rand(preroll + steps, n) < 0.6, 2, rng.randint(0, 5, (preroll + steps, n))).astype(np.uint8)
        if which == "reference":
            for t in range(preroll):
                wrapper.step({i: int(acts[t, i]) for i in range(n)})
                if env.dones["__all__"]:
                    raise RuntimeError("episode ended during pre-roll; lower --preroll")
            t0 = time.perf_counter()
            for t in range(preroll, preroll + steps):
                wrapper.step({i: int(acts[t, i]) for i in range(n)})
            dt = time.perf_counter() - t0
            st = np.array([int(a.state) for a in env.agents])
        else:
            o = orc.OracleEnv(world)
            o.reset()
            # the wrapper drops actions of agents for which action_required is false (eval_env.py:33-39)
            def masked(t):
                s = o.state()
                req = (s["state"] == 1) | ((s["state"] >= 3) & (s["state"] <= 5) & (s["ctr"] == 0))
                return np.where(req, acts[t], 255).astype(np.uint8)
            for t in range(preroll):
                o.step(masked(t), sched[t])
                o.obs()
            t0 = time.perf_counter()
            for t in range(preroll, preroll + steps):
                o.step(masked(t), sched[t])
                o.obs()
            dt = time.perf_counter() - t0
            st = o.state()["state"]
        out.append((n * steps, dt, int(((st >= 3) & (st <= 5)).sum()), st.tolist()))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="Test_03")
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--envs-per-proc", type=int, default=2)
    ap.add_argument("--steps", type=int, default=120)
    ap.add_argument("--preroll", type=int, default=150)
    args = ap.parse_args()
    from oracle import ref_harness as rh
    base = rh.csv_seed(int(args.config.split("_")[1]), 0)
    seeds = [[base + 1000 * p + k for k in range(args.envs_per_proc)] for p in range(args.procs)]
    res = {}
    final_states = {}
    for which in ("reference", "port"):
        jobs = [(args.config, s, args.steps, args.preroll, which) for s in seeds]
        t0 = time.perf_counter()
        with mp.get_context("spawn").Pool(args.procs) as pool:
            outs = pool.map(worker, jobs)
        wall = time.perf_counter() - t0
        agent_steps = sum(o[0] for out in outs for o in out)
        busy = [sum(o[1] for o in out) for out in outs]           # timed seconds per process (resets / pre-roll excluded)
        res[which] = {"agent_steps": agent_steps, "timed_s_max_over_procs": max(busy), "timed_s_sum": sum(busy),
                      "aggregate_agent_steps_per_s": agent_steps / max(busy), "per_core_agent_steps_per_s": agent_steps / sum(busy),
                      "wall_s_incl_reset_and_preroll": wall,
                      "trains_on_map_mean": float(np.mean([o[2] for out in outs for o in out]))}
        final_states[which] = [o[3] for out in outs for o in out]
    same = final_states["reference"] == final_states["port"]
    line = {
        "config": args.config, "cores": args.procs, "envs": args.procs * args.envs_per_proc, "steps": args.steps,
        "preroll": args.preroll,
        "reference": res["reference"], "port": res["port"],
        "reference_per_core": res["reference"]["per_core_agent_steps_per_s"],
        "port_per_core": res["port"]["per_core_agent_steps_per_s"],
        "port_over_reference": res["port"]["per_core_agent_steps_per_s"] / res["reference"]["per_core_agent_steps_per_s"],
        "same_final_agent_states": same,
        "what": "reference = unmodified flatland-rl RailEnv.step + flatland_cutils get_many/get_properties through "
                "solution/eval_env.py LocalTestEnvWrapper.step, one process per core; port = oracle/flatland_oracle.c stepping the "
                "same worlds with the same actions and malfunction schedules; reset and pre-roll excluded on both sides",
        "host": {"cpu_count": os.cpu_count(), "python": sys.version.split()[0]},
    }
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", "cpu_reference_vs_port.json"), "w") as f:
        json.dump(line, f, indent=1)
    print(json.dumps(line, indent=1))


if __name__ == "__main__":
    main()

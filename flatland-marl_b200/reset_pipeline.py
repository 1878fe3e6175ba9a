"""Lock-step reset pipeline: environments that finish their episode are given a NEW world while the
rest of the batch keeps stepping, instead of replaying the same world (FL_FLAG_AUTO_RESET).

World generation stays what it is in the reference — Python (`sparse_rail_generator`, line and timetable generators,
rail_env.py:260-357) — so it is taken off the critical path: a `WorldSource` hands out finished world dicts; `PackSource`
cycles through pre-generated worlds, `GeneratorPool` runs any picklable `make_world(seed) -> world` function (e.g.
`world_from_reference_env(make_env(...))` where flatland-rl is installed) in worker processes ahead of demand.  The
expensive part of a reset, the per-target distance maps (distance_map.py:57-160: 16.9 s in Python at Test_14), is rebuilt
on the GPU for the replaced slots only (`BatchedRailEnv.replace_worlds`).
"""
import multiprocessing as mp
import queue


class WorldSource:
    def take(self, n):
        """Up to n worlds that are ready now (never blocks)."""
        raise NotImplementedError

    def close(self):
        pass


class PackSource(WorldSource):
    """Cycles through a list of pre-generated worlds (e.g. `load_worlds_npz`, `load_level`)."""

    def __init__(self, worlds, start=0):
        if not worlds:
            raise ValueError("empty world list")
        self.worlds, self.pos = list(worlds), int(start)

    def take(self, n):
        out = [self.worlds[(self.pos + k) % len(self.worlds)] for k in range(n)]
        self.pos = (self.pos + n) % len(self.worlds)
        return out


def _pool_worker(make_world, seeds, out):
    for seed in iter(seeds.get, None):
        try:
            out.put((seed, make_world(seed), None))
        except Exception as e:  # noqa: BLE001 — reported to the consumer, the worker keeps going
            out.put((seed, None, repr(e)))


class GeneratorPool(WorldSource):
    """Runs `make_world(seed)` in `n_workers` processes, `ahead` worlds in advance; worlds come back in completion
    order.  `make_world` must be picklable (a module-level function)."""

    def __init__(self, make_world, first_seed=0, n_workers=2, ahead=8, context="spawn"):
        ctx = mp.get_context(context)
        self._seeds, self._out = ctx.Queue(), ctx.Queue()
        self._next_seed, self._pending, self.ahead = int(first_seed), 0, int(ahead)
        self._procs = [ctx.Process(target=_pool_worker, args=(make_world, self._seeds, self._out), daemon=True)
                       for _ in range(int(n_workers))]
        for p in self._procs:
            p.start()
        self._refill()

    def _refill(self):
        while self._pending < self.ahead:
            self._seeds.put(self._next_seed)
            self._next_seed += 1
            self._pending += 1

    def take(self, n, timeout=0.0):
        out = []
        while len(out) < n:
            try:
                seed, world, err = self._out.get(timeout=timeout) if timeout else self._out.get_nowait()
            except queue.Empty:
                break
            self._pending -= 1
            if err is not None:
                self._refill()
                raise RuntimeError("world generator failed for seed %s: %s" % (seed, err))
            out.append(world)
        self._refill()
        return out

    def close(self):
        for _ in self._procs:
            self._seeds.put(None)
        for p in self._procs:
            p.join(timeout=5)
            if p.is_alive():
                p.terminate()


class ResetPipeline:
    """Drives a `BatchedRailEnv` (built with auto_reset=True and some `reserve`) so that finished environments get
    fresh worlds from `source` when one is ready, and are reset in place otherwise.

    No host synchronisation on the stepping path: the per-environment `dones["__all__"]` flags of a step travel to pinned
    host memory with an asynchronous copy and are looked at one step later, when they have long arrived.  An environment
    whose episode ended in step t restarts in place in step t+1 (FL_FLAG_AUTO_RESET) and — if the source has a world ready
    — is given the new world right after that same call, so the observation step t+1 returns for it is already the new
    world's first observation: the consumer sees `done` at t and a fresh episode at t+1, never the replayed world.
    `replace_worlds` itself synchronises once per call (table-size check), i.e. once per batch of finished episodes."""

    def __init__(self, batch, source):
        import torch
        self.batch, self.source = batch, source
        self.replaced = 0
        self._flags = [torch.zeros(batch.E, dtype=torch.uint8).pin_memory() for _ in range(2)]
        self._events = [torch.cuda.Event() for _ in range(2)]
        self._pending = [False, False]
        self._t = 0

    def reset(self):
        self._pending = [False, False]
        return self.batch.reset()

    def step(self, actions):
        """One lock-step step; returns (obs, rewards, dones) like BatchedRailEnv.step."""
        b = self.batch
        cur, prev = self._t & 1, (self._t & 1) ^ 1
        obs, rewards, dones = b.step(actions)
        self._flags[cur].copy_(dones[:, b.N], non_blocking=True)      # this step's "__all__" flags, read next step
        self._events[cur].record()
        self._pending[cur] = True
        if self._pending[prev]:
            self._events[prev].synchronize()                           # recorded a whole step ago
            self._pending[prev] = False
            finished = self._flags[prev].nonzero().flatten().tolist()
            if finished:
                worlds = self.source.take(len(finished))
                if worlds:
                    b.replace_worlds(finished[: len(worlds)], worlds)
                    self.replaced += len(worlds)
                    obs = b.observe()
        self._t += 1
        return obs, rewards, dones

"""Evaluator client shim: the client side of the Flatland-3 remote evaluation protocol with the local
environment on the GPU.

The reference's `FlatlandRemoteClient` (flatland-rl/flatland/evaluators/client.py:38-321) talks to the evaluation service
through a redis server — requests LPUSHed on "<namespace>::<service id>::commands" as msgpack (or pickle) dicts
{type, payload, response_channel, timestamp}, blocking requests answered on a per-request response channel, time-outs
announced out of band on "...::errors" — and, crucially, steps a LOCAL copy of the environment itself: `env_create`
(client.py:228-289) builds `RailEnv(rail_from_file, line_from_file, FileMalfunctionGen)` from the level file the service
names and resets it with the service's random seed; `env_step` (client.py:291-321) fires the action at the service without
waiting and calls the local `env.step`.  So "remote" evaluation runs the hot path locally, and that is what this class
redirects: same constructor, same methods, same wire format, but the local environment is `flatland_marl_b200.RailEnv`.

World generation stays reference Python (north star): when the `flatland` package is importable the level's rail, line and
timetable are produced by the reference's own `rail_from_file` / `line_from_file` / `timetable_generator` under the
service's seed, exactly as client.py:268-283 does, and the resulting world is uploaded.  Where flatland is not installed the
level file is read by `persistence.load_level` and keeps the timetable stored in the file (`timetable="file"`); the
reference regenerates the timetable from the seed, so departure windows differ there and the shim says so in
`self.timetable_source`.
"""
import hashlib
import os
import pickle
import random
import time

import numpy as np

from .persistence import load_level

SERVICE_ID = os.getenv("AICROWD_SUBMISSION_ID", "T12345")


class FLATLAND_RL:  # message types of flatland/evaluators/messages.py
    PING = "FLATLAND_RL.PING"
    PONG = "FLATLAND_RL.PONG"
    ENV_CREATE = "FLATLAND_RL.ENV_CREATE"
    ENV_CREATE_RESPONSE = "FLATLAND_RL.ENV_CREATE_RESPONSE"
    ENV_STEP = "FLATLAND_RL.ENV_STEP"
    ENV_STEP_RESPONSE = "FLATLAND_RL.ENV_STEP_RESPONSE"
    ENV_SUBMIT = "FLATLAND_RL.ENV_SUBMIT"
    ENV_SUBMIT_RESPONSE = "FLATLAND_RL.ENV_SUBMIT_RESPONSE"
    ERROR = "FLATLAND_RL.ERROR"


class TimeoutException(StopAsyncIteration):
    """Evaluation time-out announced by the service (client.py:32-35)."""


def _np_encode(obj):
    """msgpack `default=` hook in the msgpack_numpy wire format ({nd, type, kind, shape, data})."""
    if isinstance(obj, np.ndarray):
        return {b"nd": True, b"type": obj.dtype.str, b"kind": b"", b"shape": list(obj.shape), b"data": obj.tobytes()}
    if isinstance(obj, (np.bool_, np.number)):
        return {b"nd": False, b"type": obj.dtype.str, b"data": obj.tobytes()}
    raise TypeError("cannot serialise %r" % type(obj))


def _np_decode(obj):
    """msgpack `object_hook=` counterpart of _np_encode."""
    if isinstance(obj, dict) and (b"nd" in obj or "nd" in obj):
        g = lambda k: obj.get(k.encode(), obj.get(k))
        dt = g("type")
        dt = dt.decode() if isinstance(dt, bytes) else dt
        if g("nd"):
            return np.frombuffer(g("data"), dtype=np.dtype(dt)).reshape(g("shape")).copy()
        return np.frombuffer(g("data"), dtype=np.dtype(dt))[0]
    return obj


class FlatlandRemoteClient:
    """Drop-in for flatland.evaluators.client.FlatlandRemoteClient with the local environment on the GPU.

    Extra keyword arguments (all optional): `redis_conn` — an object with lpush / blpop / rpop (default: a `redis.Redis`
    connection built from the remote_* arguments, as the reference does); `env_factory(world, obs_builder)` — builds the
    local environment from a world dict (default: `flatland_marl_b200.RailEnv` on `device`); `device`."""

    def __init__(self, test_env_folder=None, flatland_rl_service_id=SERVICE_ID, remote_host=os.getenv("redis_ip", "127.0.0.1"),
                 remote_port=6379, remote_db=0, remote_password=None, verbose=False, use_pickle=False, *,
                 redis_conn=None, env_factory=None, device="cuda:0"):
        self.use_pickle = use_pickle
        self.remote_host, self.remote_port, self.remote_db, self.remote_password = remote_host, remote_port, remote_db, remote_password
        if redis_conn is None:
            import redis
            self.redis_pool = redis.ConnectionPool(host=remote_host, port=remote_port, db=remote_db, password=remote_password)
            redis_conn = redis.Redis(connection_pool=self.redis_pool)
        self.redis_conn = redis_conn
        self.namespace = "flatland-rl"
        self.service_id = flatland_rl_service_id
        self.command_channel = "%s::%s::commands" % (self.namespace, self.service_id)
        self.error_channel = "%s::%s::errors" % (self.namespace, self.service_id)      # time-outs, out of band
        self.test_envs_root = test_env_folder or os.getenv("AICROWD_TESTS_FOLDER", "/tmp/flatland_envs")
        self.current_env_path = None
        self.verbose = verbose
        self.device = device
        self._env_factory = env_factory
        self.env = None
        self.timetable_source = None
        self.stats = {}
        self.env_step_times = []
        self.last_env_step_time = None
        self.ping_pong()

    # ---- bookkeeping (client.py:104-129) ---------------------------------------------------------------------------
    def update_running_stats(self, key, scalar):
        mean, cnt, lo, hi = (key + s for s in ("_mean", "_counter", "_min", "_max"))
        if cnt not in self.stats:
            self.stats.update({mean: scalar, lo: scalar, hi: scalar, cnt: 1})
            return
        n = self.stats[cnt]
        self.stats[mean] = (self.stats[mean] * n + scalar) / (n + 1)
        self.stats[lo], self.stats[hi] = min(self.stats[lo], scalar), max(self.stats[hi], scalar)
        self.stats[cnt] = n + 1

    def get_redis_connection(self):
        return self.redis_conn

    # ---- wire format (client.py:133-206) ----------------------------------------------------------------------------
    def _pack(self, d):
        if self.use_pickle:
            return pickle.dumps(d)
        import msgpack
        return msgpack.packb(d, default=_np_encode, use_bin_type=True)

    def _unpack(self, data):
        if self.use_pickle:
            return pickle.loads(data)     # the evaluation service is the trusted side of this channel, as in the reference
        import msgpack
        return msgpack.unpackb(data, object_hook=_np_decode, strict_map_key=False, raw=False)

    def _generate_response_channel(self):
        h = hashlib.md5(str(random.randint(0, 10 ** 10)).encode("utf-8")).hexdigest()
        return "%s::%s::response::%s" % (self.namespace, self.service_id, h)

    def _remote_request(self, request, blocking=True):
        assert isinstance(request, dict)
        request["response_channel"] = self._generate_response_channel()
        request["timestamp"] = time.time()
        r = self.get_redis_connection()
        err = r.rpop(self.error_channel)                       # a pending time-out ends the evaluation
        if err is not None:
            raise TimeoutException(self._unpack(err)["type"])
        r.lpush(self.command_channel, self._pack(request))     # the client pushes left, the service pops right
        if not blocking:
            return None
        response = self._unpack(r.blpop(request["response_channel"])[1])
        if response["type"] == FLATLAND_RL.ERROR:
            raise Exception(str(response["payload"]))
        return response

    def ping_pong(self):
        response = self._remote_request({"type": FLATLAND_RL.PING, "payload": {"version": "3.0.15"}})
        if response["type"] != FLATLAND_RL.PONG:
            raise Exception("Unable to perform handshake with the evaluation service. Expected PONG; received %r" % (response,))
        return True

    # ---- the local environment ---------------------------------------------------------------------------------------
    def _make_env(self, path, obs_builder_object, random_seed):
        """client.py:268-283: the level's world under the service's seed, on the GPU."""
        from .rail_env import RailEnv
        try:
            from flatland.envs.rail_generators import rail_from_file
            from flatland.envs.line_generators import line_from_file
            from flatland.envs.malfunction_generators import FileMalfunctionGen
            have_flatland = True
        except ImportError:
            have_flatland = False
        if have_flatland and self._env_factory is None:
            env = RailEnv(width=1, height=1, rail_generator=rail_from_file(path), line_generator=line_from_file(path),
                          malfunction_generator=FileMalfunctionGen(filename=path), obs_builder_object=obs_builder_object,
                          device=self.device)
            self.timetable_source = "reference generators under the service's seed"
            return env, env.reset(regenerate_rail=True, regenerate_schedule=True, random_seed=random_seed)
        if have_flatland:
            from flatland.envs.rail_env import RailEnv as RefRailEnv
            from flatland.core.env_observation_builder import DummyObservationBuilder
            from .worlds import world_from_reference_env
            ref = RefRailEnv(width=1, height=1, rail_generator=rail_from_file(path), line_generator=line_from_file(path),
                             malfunction_generator=FileMalfunctionGen(filename=path), obs_builder_object=DummyObservationBuilder())
            ref.reset(regenerate_rail=True, regenerate_schedule=True, random_seed=random_seed)
            world = world_from_reference_env(ref)
            self.timetable_source = "reference generators under the service's seed"
        else:
            world = load_level(path, malfunction_seed=int(random_seed or 0))
            self.timetable_source = "file (flatland is not importable: the reference would regenerate the timetable from the seed)"
        if self._env_factory is not None:
            env = self._env_factory(world, obs_builder_object)
        else:
            env = RailEnv(width=int(world["W"]), height=int(world["H"]), number_of_agents=int(world["N"]), world=world,
                          obs_builder_object=obs_builder_object, device=self.device)
        return env, env.reset()

    def env_create(self, obs_builder_object):
        t0 = time.time()
        response = self._remote_request({"type": FLATLAND_RL.ENV_CREATE, "payload": {}})
        payload = response["payload"]
        observation, info = payload["observation"], payload["info"]
        self.update_running_stats("env_creation_wait_time", time.time() - t0)
        if not observation:                                    # the service has no more levels: evaluation complete
            return observation, info
        path = os.path.join(self.test_envs_root, payload["env_file_path"])
        if not os.path.exists(path):
            raise Exception("\nWe cannot seem to find the env file paths at the required location.\n"
                            "Did you remember to set the AICROWD_TESTS_FOLDER environment variable to point to the location "
                            "of the Tests folder ? \nWe are currently looking at `%s` for the tests" % self.test_envs_root)
        self.current_env_path = path
        t0 = time.time()
        self.env, (local_observation, info) = self._make_env(path, obs_builder_object, payload["random_seed"])
        self.update_running_stats("internal_env_reset_time", time.time() - t0)
        self.last_env_step_time = time.time()
        return local_observation, info

    def env_step(self, action, render=False):
        inference = time.time() - self.last_env_step_time
        self.update_running_stats("inference_time(approx)", inference)
        # fire and forget: the service steps its own copy in parallel (this can raise a pending time-out)
        self._remote_request({"type": FLATLAND_RL.ENV_STEP, "payload": {"action": {int(k): int(v) for k, v in action.items()},
                                                                         "inference_time": inference}}, blocking=False)
        t0 = time.time()
        out = self.env.step(action)
        self.update_running_stats("internal_env_step_time", time.time() - t0)
        self.last_env_step_time = time.time()
        return list(out)

    def submit(self):
        response = self._remote_request({"type": FLATLAND_RL.ENV_SUBMIT, "payload": {}})
        if response["type"] != FLATLAND_RL.ENV_SUBMIT_RESPONSE:
            raise Exception(str(response))
        return response["payload"]

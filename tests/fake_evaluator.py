"""Test infrastructure: an in-process stand-in for the redis server and the evaluation service of the Flatland-3 remote
evaluation (flatland/evaluators/service.py), just enough of the protocol for the client side: PING -> PONG, ENV_CREATE ->
{observation, info, random_seed, env_file_path} for each level of a list and observation False afterwards, ENV_STEP (not
answered: the client does not wait, client.py:303-306), ENV_SUBMIT -> ENV_SUBMIT_RESPONSE.  Messages are msgpack dicts."""
import threading
import time

import msgpack


class FakeRedis:
    """lpush / rpop / blpop / brpop over in-memory lists."""

    def __init__(self):
        self.lists, self.cv = {}, threading.Condition()

    def lpush(self, key, value):
        with self.cv:
            self.lists.setdefault(key, []).insert(0, value)
            self.cv.notify_all()

    def rpush(self, key, value):
        with self.cv:
            self.lists.setdefault(key, []).append(value)
            self.cv.notify_all()

    def rpop(self, key):
        with self.cv:
            lst = self.lists.get(key)
            return lst.pop() if lst else None

    def _bpop(self, key, left, timeout):
        end = time.time() + (timeout or 30)
        with self.cv:
            while not self.lists.get(key):
                if not self.cv.wait(max(0.0, end - time.time())) and time.time() >= end:
                    return None
            v = self.lists[key].pop(0 if left else -1)
            return (key.encode() if isinstance(key, str) else key, v)

    def blpop(self, key, timeout=0):
        return self._bpop(key, True, timeout)

    def brpop(self, key, timeout=0):
        return self._bpop(key, False, timeout)


class FakeService(threading.Thread):
    def __init__(self, redis_conn, levels, seeds, service_id="T12345"):
        super().__init__(daemon=True)
        self.r, self.levels, self.seeds = redis_conn, list(levels), list(seeds)
        self.command_channel = "flatland-rl::%s::commands" % service_id
        self.next_level, self.steps, self.actions, self.stopped = 0, 0, [], False

    def run(self):
        while not self.stopped:
            item = self.r.brpop(self.command_channel, timeout=0.2)
            if item is None:
                continue
            req = msgpack.unpackb(item[1], raw=False, strict_map_key=False)
            t = req["type"]
            if t == "FLATLAND_RL.PING":
                self.reply(req, "FLATLAND_RL.PONG", {})
            elif t == "FLATLAND_RL.ENV_CREATE":
                if self.next_level < len(self.levels):
                    k = self.next_level
                    self.next_level += 1
                    self.reply(req, "FLATLAND_RL.ENV_CREATE_RESPONSE",
                               {"observation": True, "info": {}, "random_seed": self.seeds[k], "env_file_path": self.levels[k]})
                else:
                    self.reply(req, "FLATLAND_RL.ENV_CREATE_RESPONSE",
                               {"observation": False, "info": False, "random_seed": False, "env_file_path": False})
            elif t == "FLATLAND_RL.ENV_STEP":
                self.steps += 1
                self.actions.append(req["payload"]["action"])
            elif t == "FLATLAND_RL.ENV_SUBMIT":
                self.reply(req, "FLATLAND_RL.ENV_SUBMIT_RESPONSE", {"mean_reward": 0.0, "mean_percentage_complete": 0.0})
                self.stopped = True

    def reply(self, req, typ, payload):
        self.r.rpush(req["response_channel"], msgpack.packb({"type": typ, "payload": payload}, use_bin_type=True))

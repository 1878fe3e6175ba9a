"""Import-time stand-in for msgpack_numpy (absent from this image; flatland/envs/persistence.py:4-5 and
flatland/evaluators/client.py:10,29 import and patch it).  TEST INFRASTRUCTURE ONLY.  `encode` / `decode` follow the
published msgpack_numpy wire format ({nd, type, kind, shape, data}) so that the reference's evaluator client can talk to the
in-process fake service of tests/fake_evaluator.py."""
import numpy as np


def patch():
    pass


def encode(obj, chain=None):
    if isinstance(obj, np.ndarray):
        return {b"nd": True, b"type": obj.dtype.str, b"kind": b"", b"shape": list(obj.shape), b"data": obj.tobytes()}
    if isinstance(obj, (np.bool_, np.number)):
        return {b"nd": False, b"type": obj.dtype.str, b"data": obj.tobytes()}
    return obj if chain is None else chain(obj)


def decode(obj, chain=None):
    if isinstance(obj, dict) and (b"nd" in obj or "nd" in obj):
        g = lambda k: obj.get(k.encode(), obj.get(k))
        dt = g("type")
        dt = dt.decode() if isinstance(dt, bytes) else dt
        if g("nd"):
            return np.frombuffer(g("data"), dtype=np.dtype(dt)).reshape(g("shape")).copy()
        return np.frombuffer(g("data"), dtype=np.dtype(dt))[0]
    return obj if chain is None else chain(obj)

"""Import-time stand-in (flatland/envs/persistence.py patches msgpack at import)."""


def patch():
    pass

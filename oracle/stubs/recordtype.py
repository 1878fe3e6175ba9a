"""Import-time stand-in (only rendering utilities use recordtype)."""
from collections import namedtuple


def recordtype(name, fields, default=None):
    return namedtuple(name, [f if isinstance(f, str) else f[0] for f in fields])

"""Importable name of the package whose sources live in ``flatland-marl_b200/`` (a hyphen is not a
valid Python identifier, so this shim points the package path at that directory)."""
import os as _os

_SRC = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "flatland-marl_b200")
__path__.insert(0, _SRC)

from .api import *  # noqa: E402,F401,F403
from .api import __all__  # noqa: E402,F401

"""Import-time stand-in: the backport's ``path`` is the stdlib one on Python >= 3.9."""
from importlib.resources import path  # noqa: F401

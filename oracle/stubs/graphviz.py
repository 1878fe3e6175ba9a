"""Import-time stand-in (flatland/envs/agent_chains.py:5 imports graphviz for rendering only)."""


class Source:
    def __init__(self, *a, **k):
        pass

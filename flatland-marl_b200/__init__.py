"""Sources of the `flatland_marl_b200` package (import it under that name; see ../flatland_marl_b200)."""

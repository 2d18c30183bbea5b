from xitorch_b200.optimize.rootfinder import rootfinder, equilibrium, minimize   # noqa: F401

__all__ = ["rootfinder", "equilibrium", "minimize"]

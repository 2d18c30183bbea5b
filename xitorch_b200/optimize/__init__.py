from xitorch_b200.optimize.rootfinder import rootfinder, equilibrium   # noqa: F401

__all__ = ["rootfinder", "equilibrium"]

"""Importable alias for the product package, whose directory is named `dm-nerf_b200/` (not a valid
Python identifier).  `import dmnerf_b200` executes dm-nerf_b200/__init__.py with this module's
__path__ pointing at that directory, so `dmnerf_b200.render` is dm-nerf_b200/render.py."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "dm-nerf_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f

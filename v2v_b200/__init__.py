"""Importable alias of the package directory `video-to-video-diffusion_b200/` (a hyphenated directory name is
not a valid Python identifier): `import v2v_b200.models` resolves to `video-to-video-diffusion_b200/models`."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "video-to-video-diffusion_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f

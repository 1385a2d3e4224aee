"""Import alias for the package directory ``prompt-diffusion_b200/``.

The product lives in ``prompt-diffusion_b200/`` (the name the build contract
fixes); a hyphen cannot appear in a Python import, so this shim makes the same
directory importable as ``prompt_diffusion_b200`` by pointing ``__path__`` at it
and executing its ``__init__.py`` in this module's namespace.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "prompt-diffusion_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f

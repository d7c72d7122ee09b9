"""Importable alias of the `gpt2-vision-language_b200/` package directory.

The package directory carries the reference's name (with hyphens, which Python cannot import), so this
shim points its ``__path__`` at that directory: ``import gpt2_vision_language_b200 as vl`` just works.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "gpt2-vision-language_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))

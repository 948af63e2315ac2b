"""Importable alias of the product package.

The product lives in ``realtime-pose-estimation_b200/`` (the directory name the
build contract asks for); a hyphen cannot appear in a Python import name, so this
stub points its ``__path__`` there and executes that directory's ``__init__.py``
as its own body.  ``import rtpe_b200`` == the product package.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "realtime-pose-estimation_b200")
__path__ = [_real]
_init = _os.path.join(_real, "__init__.py")
with open(_init, "r") as _f:
    exec(compile(_f.read(), _init, "exec"))

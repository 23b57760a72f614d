"""Importable alias for the package directory ``high-fidelity-generative-compression_b200``.

``import hific_b200.ops`` loads ``high-fidelity-generative-compression_b200/ops.py``: this module
only redirects ``__path__`` (a hyphenated directory name cannot be imported directly).
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "high-fidelity-generative-compression_b200")
if not _os.path.isdir(_real):  # pragma: no cover
    raise ImportError(f"hific_b200: package directory not found: {_real}")
__path__ = [_real]
PACKAGE_DIR = _real
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))

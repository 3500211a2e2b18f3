"""Import alias for the ``semantic-gaussians_b200/`` package directory.

The product code lives in ``semantic-gaussians_b200/`` (the directory name the build
contract fixes).  A hyphen is not importable, so this stub extends ``__path__`` to that
directory: ``import semantic_gaussians_b200.rasterizer`` resolves to
``semantic-gaussians_b200/rasterizer.py``.  No code lives here.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "semantic-gaussians_b200")
if not _os.path.isdir(_real):  # pragma: no cover
    raise ImportError(f"package directory missing: {_real}")
__path__.insert(0, _real)

from ._version import __version__  # noqa: E402,F401

"""tensorflow.compat (stand-in, see ../__init__.py): compat.v2 is the package itself."""
import sys as _sys

v2 = _sys.modules["tensorflow"]
_sys.modules["tensorflow.compat.v2"] = v2

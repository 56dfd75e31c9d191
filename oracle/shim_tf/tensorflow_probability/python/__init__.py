from . import internal, util  # noqa: F401
from . import distributions  # noqa: F401


class _Empty:
    pass


bijectors = _Empty()
math = _Empty()

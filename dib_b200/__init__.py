"""Importable alias of the package directory ``distributed-information-bottleneck.github.io_b200/`` (whose name
is not a valid Python identifier): ``import dib_b200`` executes that package's ``__init__`` under this name."""
import os as _os

_impl = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "distributed-information-bottleneck.github.io_b200")
__path__ = [_impl]
with open(_os.path.join(_impl, "__init__.py")) as _fh:
    exec(compile(_fh.read(), _os.path.join(_impl, "__init__.py"), "exec"))
del _fh

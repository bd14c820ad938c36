"""Import alias for the package directory ``cilrs-autonomous-driving-carla_b200/`` (a hyphenated
directory name is not importable, so this stub points ``cilrs_b200.__path__`` at it and runs its
``__init__``)."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "cilrs-autonomous-driving-carla_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f

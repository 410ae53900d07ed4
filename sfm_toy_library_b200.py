"""Import alias: the package directory is `sfm-toy-library_b200/` (not a valid Python identifier);
this module makes it importable as `sfm_toy_library_b200` (and `sfm_toy_library_b200.capi`, ...)."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "sfm-toy-library_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))

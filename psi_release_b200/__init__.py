"""Import alias: `psi_release_b200` -> the package directory `psi-release_b200/`.

The package lives in `psi-release_b200/` (the name the build contract fixes); a hyphen is
not importable, so this stub points the import system at that directory and runs its
__init__.py in this module's namespace.
"""
import os as _os

_home = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "psi-release_b200")
__path__ = [_home]
with open(_os.path.join(_home, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_home, "__init__.py"), "exec"))
del _f

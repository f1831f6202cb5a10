"""Importable alias for the package directory `ubisoft-laforge-daft-exprt_b200/` (a hyphenated directory name cannot be
written in an `import` statement).  All code lives there; this file only extends `__path__` and re-exports."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), 'ubisoft-laforge-daft-exprt_b200')
__path__.append(_real)
with open(_os.path.join(_real, '__init__.py')) as _f:
    exec(compile(_f.read(), _os.path.join(_real, '__init__.py'), 'exec'))

"""Importable name of the host package.

The package directory is `eeg-gan-timegan-cgan_b200/` (named after the reference repository); a hyphenated
directory cannot be imported, so this stub points `timegan_b200.__path__` at it and runs its `__init__.py`.
`import timegan_b200.timegan_model` etc. then resolve to the files in that directory.
"""
from pathlib import Path as _Path

_REAL = _Path(__file__).resolve().parent.parent / "eeg-gan-timegan-cgan_b200"
__path__ = [str(_REAL)]
_init = _REAL / "__init__.py"
exec(compile(_init.read_text(), str(_init), "exec"))

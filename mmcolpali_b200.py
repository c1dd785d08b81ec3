"""Importable alias for the hyphenated package directory ``multi-modal_colpali_b200``."""
import importlib as _il
import sys as _sys

_pkg = _il.import_module("multi-modal_colpali_b200")
_sys.modules[__name__] = _pkg

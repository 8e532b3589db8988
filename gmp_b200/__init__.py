"""Import alias: ``import gmp_b200`` loads the package that lives in the directory
``geometric-message-passing_b200/`` (the hyphen keeps that directory from being importable by name)."""
import importlib.util
import os
import sys

_real = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "geometric-message-passing_b200")
_spec = importlib.util.spec_from_file_location(
    "gmp_b200", os.path.join(_real, "__init__.py"), submodule_search_locations=[_real])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["gmp_b200"] = _mod
_spec.loader.exec_module(_mod)

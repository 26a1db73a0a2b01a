"""Importable alias: ``import dcnr_b200`` loads the package that lives in the directory
``hybrid-hotel-recommendation-system-based-on-friends-recommendations_b200/`` (its name is not a
Python identifier, so it cannot be imported by a plain ``import`` statement)."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                    "hybrid-hotel-recommendation-system-based-on-friends-recommendations_b200")
_spec = importlib.util.spec_from_file_location("dcnr_b200", os.path.join(_DIR, "__init__.py"),
                                               submodule_search_locations=[_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["dcnr_b200"] = _mod
_spec.loader.exec_module(_mod)

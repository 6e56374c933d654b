"""Import shim: the product package lives in the directory `hyper-greco_b200/` (not an importable name);
`import hyper_greco_b200` resolves to it."""
import importlib.util
import os
import sys

_d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hyper-greco_b200")
_spec = importlib.util.spec_from_file_location("hyper_greco_b200", os.path.join(_d, "__init__.py"), submodule_search_locations=[_d])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["hyper_greco_b200"] = _mod
_spec.loader.exec_module(_mod)

"""Import shim: the product package lives in the directory `lightning-asr_b200/` (the name the project layout
prescribes, not a legal Python identifier); this makes it importable as `lightning_asr_b200`."""
import importlib.util
import os
import sys

_real = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lightning-asr_b200")
_spec = importlib.util.spec_from_file_location(
    __name__, os.path.join(_real, "__init__.py"), submodule_search_locations=[_real]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)

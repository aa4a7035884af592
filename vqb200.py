"""Import alias: ``import vqb200`` loads the package that lives in the directory
``vq-vae-transformer-arc-welding_b200/`` (a name Python cannot import directly
because of the hyphens)."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vq-vae-transformer-arc-welding_b200")
_spec = importlib.util.spec_from_file_location(
    "vqb200", os.path.join(_DIR, "__init__.py"), submodule_search_locations=[_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["vqb200"] = _mod
_spec.loader.exec_module(_mod)

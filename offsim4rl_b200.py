"""Import shim: the package lives in ``rl-offline-simulation_b200/`` (not a valid Python identifier),
so ``import offsim4rl_b200`` loads that directory as the package ``offsim4rl_b200``."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "rl-offline-simulation_b200")
_spec = importlib.util.spec_from_file_location("offsim4rl_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["offsim4rl_b200"] = _mod
_spec.loader.exec_module(_mod)

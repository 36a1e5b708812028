"""Import alias: `import ilrl_b200` loads the package in ./imitation-learning-rl_b200/ (a hyphen cannot be imported)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "imitation-learning-rl_b200")
_spec = importlib.util.spec_from_file_location("ilrl_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["ilrl_b200"] = _mod
_spec.loader.exec_module(_mod)

"""Imports the package directory `nvidia-texture-tools_b200/` (not a valid identifier) as module `nvtt_b200`."""
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))


def load():
    if "nvtt_b200" in sys.modules:
        return sys.modules["nvtt_b200"]
    pkg = os.path.join(_ROOT, "nvidia-texture-tools_b200")
    spec = importlib.util.spec_from_file_location("nvtt_b200", os.path.join(pkg, "__init__.py"),
                                                  submodule_search_locations=[pkg])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["nvtt_b200"] = mod
    spec.loader.exec_module(mod)
    return mod

"""Import helper: the package directory name contains hyphens, so it cannot be imported
with a plain `import` statement.  `load()` registers it as module `hegpu_b200`."""
import importlib.util
import os
import sys

PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "homomorphic-encryption-algorithms-diploma-thesis_b200")


def load():
    if "hegpu_b200" in sys.modules:
        return sys.modules["hegpu_b200"]
    spec = importlib.util.spec_from_file_location("hegpu_b200", os.path.join(PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["hegpu_b200"] = mod
    spec.loader.exec_module(mod)
    return mod

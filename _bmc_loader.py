"""Imports the hyphenated package directory `biocma-mcst_b200/` as module
`biocma_mcst_b200` (tests, bench.py and __graft_entry__.py share this)."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "biocma-mcst_b200")


def load_pkg():
    name = "biocma_mcst_b200"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_synth():
    load_pkg()
    import importlib
    return importlib.import_module("biocma_mcst_b200.synth")

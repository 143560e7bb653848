"""Loader of the `_pcbhost` CPython extension (csrc/pcb_hostmod.cpp), built in-tree by phyclone_b200.build."""

import importlib.machinery
import importlib.util
import os
import sysconfig

_HERE = os.path.dirname(os.path.abspath(__file__))
MOD_PATH = os.path.join(_HERE, "lib", "_pcbhost" + sysconfig.get_config_var("EXT_SUFFIX"))
_mod = None


def load():
    global _mod
    if _mod is None:
        if not os.path.exists(MOD_PATH):
            raise RuntimeError("%s is missing: build it with `python -m phyclone_b200.build`" % MOD_PATH)
        loader = importlib.machinery.ExtensionFileLoader("_pcbhost", MOD_PATH)
        spec = importlib.util.spec_from_loader("_pcbhost", loader)
        _mod = importlib.util.module_from_spec(spec)
        loader.exec_module(_mod)
    return _mod

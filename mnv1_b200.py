"""Import shim: the package directory name (fixed by the build contract) contains hyphens,
so it is loaded by path and exposed as ``mnv1_b200``."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                        "cnn-mobilenet-v1-implementation-on-aws-fpga-using-opencl_b200")
_NAME = "mnv1_b200"
if _NAME not in sys.modules or getattr(sys.modules[_NAME], "__path__", None) is None:
    _spec = importlib.util.spec_from_file_location(
        _NAME, os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)

"""Import alias: ``import tsr_b200`` gives the package that lives in the (hyphenated, hence not
directly importable) directory ``template-speech-recognition_b200/``."""
import importlib
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
if _here not in sys.path:
    sys.path.insert(0, _here)
_pkg = importlib.import_module("template-speech-recognition_b200")
sys.modules[__name__] = _pkg

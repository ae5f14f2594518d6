"""`import sarpost` — importable alias of the hyphenated package directory `sar-yolo_b200/`."""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
_pkg = importlib.import_module("sar-yolo_b200")
sys.modules[__name__] = _pkg

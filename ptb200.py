"""Import alias: `import ptb200` == the package in ./multi-gpu-path-tracer_b200 (whose name is not an identifier)."""
import importlib
import sys
from pathlib import Path

_root = str(Path(__file__).resolve().parent)
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("multi-gpu-path-tracer_b200")
sys.modules[__name__] = _pkg

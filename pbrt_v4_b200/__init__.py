"""Importable alias of the `pbrt-v4_b200/` package directory (a hyphen and a dot cannot appear in a Python
module name).  `import pbrt_v4_b200.engine` resolves to `pbrt-v4_b200/engine.py`."""
import os

__path__.insert(0, os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "pbrt-v4_b200")))

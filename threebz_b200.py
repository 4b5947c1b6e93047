"""Alias so `import threebz_b200` works: the package directory is `3bz_b200/` (like the reference's
package name `3bz`), which Python's import statement cannot spell."""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
_pkg = importlib.import_module("3bz_b200")
sys.modules[__name__] = _pkg

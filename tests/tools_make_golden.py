"""Makes tools/py_quadtree.py importable from the tests."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
from py_quadtree import distribute  # noqa: E402,F401

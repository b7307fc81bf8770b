"""Shim: the loader lives in baseline/ref_loader.py (bench.py's reference arm uses it too)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from baseline.ref_loader import install_stubs, load_reference, reference_root  # noqa: E402,F401

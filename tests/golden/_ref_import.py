"""Golden-fixture generation imports the unmodified reference through oracle/ref_import.py (see there)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.ref_import import import_reference, reference_root  # noqa: E402,F401

REF_ROOT = reference_root()

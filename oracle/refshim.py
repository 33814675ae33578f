"""oracle/refshim.py — TEST INFRASTRUCTURE ONLY (never imported by discover_tfbs_b200/).

Imports the *unmodified* reference module `/root/reference/utils.py` in this container by
putting `oracle/stubs/` (stand-ins for the absent Biopython / mmh3 / pybedtools / matplotlib
/ rpy2) in front of it on sys.path. Used only by `tests/golden/make_golden.py` to run the
reference's own functions on seeded synthetic inputs and commit their outputs as golden
vectors; `/root/reference` does not exist on the GPU box, so nothing at test/bench time
calls this.
"""
import importlib
import os
import sys

REFERENCE_DIR = os.environ.get("TFBS_REFERENCE_DIR", "/root/reference")
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stubs")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "utils.py"))


def load_reference_utils():
    """Return the reference's `utils` module (utils.__version__ == '0.3.7')."""
    if not available():
        raise RuntimeError(f"reference not present at {REFERENCE_DIR}")
    for p in (REFERENCE_DIR, _STUBS):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, REFERENCE_DIR)
    sys.path.insert(0, _STUBS)
    saved = sys.modules.pop("utils", None)
    try:
        mod = importlib.import_module("utils")
    finally:
        # do not leave the reference registered under the bare name "utils"
        ref = sys.modules.pop("utils", None)
        if saved is not None:
            sys.modules["utils"] = saved
        sys.path.remove(REFERENCE_DIR)
        sys.path.remove(_STUBS)
    return ref if ref is not None else mod

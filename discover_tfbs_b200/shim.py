"""Run one of the reference's OWN driver scripts with the hot path served by this package.

    python -m discover_tfbs_b200.shim /path/to/discover_tfbs/build_features.py  <the script's arguments>
    python -m discover_tfbs_b200.shim /path/to/discover_tfbs/predict.py         <the script's arguments>

The reference's drivers do `from utils import build_feature_table, parse_fasta`
(build_features.py:17, predict.py:19).  This shim loads the `utils.py` that sits next to the script
(the reference's own, with all its other helpers), replaces the hot-path names below with the
GPU-backed functions of `discover_tfbs_b200.utils`, registers the result as the module `utils`, and
then executes the script unchanged as `__main__`.  Nothing of the driver is re-implemented here.
"""
import importlib.util
import os
import runpy
import sys

# names of the reference's utils.py that are on the hot path
HOT_PATH_NAMES = ("build_feature_table", "parse_fasta", "calculate_gc_content", "_get_DNA_shape",
                  "get_shape_data")


def install(script_dir: str):
    """Register a `utils` module = the reference's utils.py from `script_dir` (if present) with the
    hot-path names overridden.  Returns the module."""
    from . import utils as ours

    path = os.path.join(script_dir, "utils.py")
    if os.path.isfile(path):
        spec = importlib.util.spec_from_file_location("utils", path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules["utils"] = mod
        spec.loader.exec_module(mod)
    else:
        mod = type(sys)("utils")
        sys.modules["utils"] = mod
    for name in HOT_PATH_NAMES:
        setattr(mod, name, getattr(ours, name))
    mod.__tfbs_b200__ = ours.__version__
    return mod


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or not os.path.isfile(argv[0]):
        sys.exit("usage: python -m discover_tfbs_b200.shim /path/to/reference/<driver>.py [driver arguments]")
    script = os.path.abspath(argv[0])
    script_dir = os.path.dirname(script)
    install(script_dir)
    if script_dir not in sys.path:
        sys.path.insert(0, script_dir)
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()

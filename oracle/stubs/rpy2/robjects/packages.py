"""Stub of rpy2.robjects.packages.importr (create_bkg_seqs.py:26-34)."""


def importr(name):
    class _Pkg:
        def __getattr__(self, item):
            raise RuntimeError(f"rpy2 stub: R package {name!r} is not available here")

    return _Pkg()

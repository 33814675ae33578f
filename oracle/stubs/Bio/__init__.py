"""Stub of the Biopython top-level package (test infrastructure; see oracle/stubs/README.md)."""

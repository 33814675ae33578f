"""Stub package (test infrastructure; see oracle/stubs/README.md)."""

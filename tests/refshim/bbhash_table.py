"""Test shim for the third-party ``bbhash_table`` package (BBHashTable as used by the reference)."""
from _backend import BBHashTable  # noqa: F401

"""Test shim for the third-party ``khmer`` package: only what spacegraphcats/cdbg/hashing.py:14-20 uses."""
from _backend import Nodetable  # noqa: F401

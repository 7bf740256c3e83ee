"""Replacement ``models`` package (same sub-module names as the reference's lib/models)."""
from . import GAT, MDR, GATOR  # noqa: F401

"""Stand-in for the reference's ``hgnnaggr`` torch extension (hgnnaggr.cc:146-151)."""
from .ops import hgnnaggr, hgnnaggr_max, hgnnaggr_mean  # noqa: F401

__all__ = ["hgnnaggr", "hgnnaggr_mean", "hgnnaggr_max"]

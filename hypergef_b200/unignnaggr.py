"""Stand-in for the reference's ``unignnaggr`` torch extension (unignnaggr.cc:98-102).

Both spellings are exported: the extension's (``unignnaggrdeg``, ``unignnaggr``) and the
ones its own Python wrapper calls (``unignnconvdeg``, ``unignnconv``,
source/python/unignnconv.py:7,10), which the reference extension does not define.
"""
from .ops import unignnaggr, unignnaggrdeg  # noqa: F401

unignnconvdeg = unignnaggrdeg
unignnconv = unignnaggr

__all__ = ["unignnaggrdeg", "unignnaggr", "unignnconvdeg", "unignnconv"]

"""sqlp_b200 -- the argmax cut-formation path of yhz0/SQLP's TwoSD solver on B200.

``sqlp_b200.twosd`` mirrors the reference's Julia interface for the path; the work is done
by hand-written sm_100a CUDA kernels in ``libsqlp_b200.so`` (C ABI: ``include/sqlp_b200.h``).
"""
from ._lib import SqlpError, NoArgmaxError, build, lib, SO_PATH  # noqa: F401

__all__ = ["SqlpError", "NoArgmaxError", "build", "lib", "SO_PATH"]

"""lgcnhs_b200 — Python host binding of liblgcnhs.so (sm_100a kernels for the LGCNHS hot paths).

The sibling packages model/, utils/, metrics/, processing/ mirror the reference's module
surface so that the reference's main.py / const.py drive this code unchanged
when this directory precedes /root/reference on sys.path (see run_main.py).
"""
from ._lib import LgcnhsError, launch_count, lib, reset_launch_count  # noqa: F401

__version__ = "0.1.0"

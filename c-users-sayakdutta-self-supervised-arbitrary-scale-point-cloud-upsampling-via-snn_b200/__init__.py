"""sapcu_b200 -- B200-native (sm_100a) inference hot path of the SNN point-cloud upsampler.

Public surface mirrors the reference's:

    from sapcu_b200.fn import config as fn_config      # load_config / get_model  (fn/config.py)
    from sapcu_b200.fd import config as fd_config      # load_config / get_model  (fd/config.py)
    from sapcu_b200.generation import Generator3D6     # generation.py

The compute lives in libsapcu_b200.so (csrc/, C ABI in include/sapcu_b200.h); importing this package does
not load it -- the first operator call does, and raises if it has not been built.
"""
import os as _os

PACKAGE_DIR = _os.path.dirname(_os.path.abspath(__file__))
CONFIG_DIR = _os.path.join(PACKAGE_DIR, "config")

from . import _native  # noqa: E402
from ._native import SapcuError, build, lib  # noqa: E402,F401

__all__ = ["SapcuError", "build", "lib", "PACKAGE_DIR", "CONFIG_DIR"]

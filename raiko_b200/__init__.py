"""raiko_b200: B200-native (sm_100a) implementation of raiko's blob-KZG path.

The product is ``libraiko_kzg.so`` (C ABI in ``include/raiko_kzg.h``); this package is
the Python-side mirror of the reference's Rust wrapper
``lib/src/primitives/eip4844.rs`` over that ABI.  There is no CPU fallback: without
the built CUDA library and a GPU every compute call raises.
"""
from . import eip4844  # noqa: F401
from .eip4844 import *  # noqa: F401,F403

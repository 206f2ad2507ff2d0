"""bee2_b200 — Python harness over the C ABI of the B200-native bee2 batch engine.

The product is ``libbee2_b200.so`` (``bee2_b200/csrc``: hand-written sm_100a CUDA kernels
behind a C host layer that keeps the reference's ``include/bee2`` names; see
``include/bee2_b200.h``). This package only binds that ABI with ``ctypes`` so that tests
and ``bench.py`` can drive it; function names, argument meaning and ``err_t`` behaviour
mirror the reference (bash.h, belt.h, bign.h) so the parity tests read like the reference's
own tests (test/crypto/{bash,belt,bign}_test.c).

There is no CPU fallback anywhere: if the shared library is missing the import of
:func:`lib` raises, and if no CUDA device is usable every call returns / raises
``ERR_B2G_NO_DEVICE``.
"""
from .api import *  # noqa: F401,F403
from .api import __all__  # noqa: F401

"""thunder_b200 - B200-native Optimiser hot path for THUNDER (C ABI in include/thunder_b200.h).

Only what the path needs lives here: csrc/ (CUDA kernels + the C ABI + host C++), capi.py (ctypes
mirror of the ABI) and synth.py (synthetic stacks).  There is no CPU implementation.
"""
__version__ = "0.1.0"

"""deepsignal_b200 — B200-native call_mods / extract hot path of deepsignal.

Host Python mirrors the reference's call sites; all arithmetic runs in
``libdsb200.so`` (hand-written sm_100a CUDA behind the C ABI of include/dsb200.h).
"""
__version__ = "0.1.0"

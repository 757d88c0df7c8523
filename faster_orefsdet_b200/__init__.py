"""B200-native support-guided detection head of Faster-OreFSDet.

    from faster_orefsdet_b200.config import get_cfg
    from faster_orefsdet_b200.modeling import build_model

The hot path runs in hand-written sm_100a kernels behind a C ABI
(include/fod_b200.h -> faster_orefsdet_b200/libfod_b200.so); see DESIGN.md.
"""
__version__ = "0.1.0"

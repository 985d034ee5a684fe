"""kami_b200 -- B200-native self-play hot path of codeandkey/kami.

This package is a thin ctypes face over ``libkami_b200.so`` (the C ABI declared in
``include/kami_b200.h``).  All compute runs as hand-written sm_100a CUDA kernels; there is no
CPU fallback -- importing works anywhere (so the symbol table can be checked), but any compute
call without a B200 raises :class:`KamiError`.
"""
from .api import (  # noqa: F401
    KamiError,
    Env,
    MCTS,
    NN,
    Trainer,
    DataParallelTrainer,
    TreePool,
    Arena,
    TreeCfg,
    Position,
    lib,
    lib_path,
    build,
    device_count,
    PSIZE,
    OBSIZE,
    NFEATURES,
    DEF_YML,
)

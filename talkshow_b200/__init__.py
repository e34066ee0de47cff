"""talkshow_b200 — B200-native engine for TalkSHOW's speech-to-motion generation path.

Package layout: ``csrc/`` CUDA kernels + C ABI (built to ``libtalkshow_b200.so``), ``_lib.py`` ctypes
binding, ``engine.py`` module-level mirror (tensors in/out), ``nets/`` + ``trainer/`` +
``data_utils/`` the reference's wrapper/config/feature surface, ``pipeline.py`` whole-body batch
generation + multi-GPU sharding, ``synth.py`` checkpoint schema and seeded synthetic checkpoints.
"""
__version__ = "0.1.0"

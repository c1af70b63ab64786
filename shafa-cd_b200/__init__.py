"""shafa-cd_b200: B200-native (sm_100a) implementation of Shafa's F/T/C/D block
pipeline behind a C ABI (include/shafa_cuda.h).  See DESIGN.md.

  csrc/      hand-written CUDA kernels + the extern "C" library (libshafa_cuda.so)
  host/      the `shafa` CLI in C: same options and container formats as the reference
  binding.py ctypes view of the C ABI + .freq/.cod/.shaf text helpers
  datagen.py seeded synthetic inputs of the BASELINE.json shapes

Import through `__graft_entry__.load_package()` (the directory name has a hyphen).
"""
from . import datagen, sharding  # noqa: F401
from .binding import (  # noqa: F401
    DeviceBuffer,
    Encoded,
    Shafa,
    ShafaError,
    ShafaMulti,
    cod_text,
    declared_symbols,
    freq_text,
    load_library,
    parse_cod,
    parse_cod_wide,
    parse_shaf,
    shaf_bytes,
)

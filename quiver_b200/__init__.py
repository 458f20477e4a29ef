"""quiver_b200 — B200 (sm_100a) implementation of Quiver's exact-search hot path.

Layout:
  csrc/   hand-written CUDA kernels + the C ABI (include/quiver_gpu.h) -> lib/libquivergpu.so
  host/   C++ host side above the C ABI (string IDs, request validation, negative-example
          rerank, predicate compiler) mirroring the reference's Go interfaces -> lib/libquiverhost.so
  capi.py / hybrid.py / collection.py   ctypes plumbing that exposes the above under the
          reference's names so the parity tests read like the reference's own tests
There is no CPU fallback anywhere in this package.
"""
from . import capi  # noqa: F401

__all__ = ["capi"]

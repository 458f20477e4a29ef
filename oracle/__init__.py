"""oracle — CPU restatement of Quiver's exact-search hot path. TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; nothing under quiver_b200/ does. Parity status: PINNED against the
reference's own known-answer tests (tests/golden/, tests/test_oracle_golden.py); the Go
reference itself cannot be built in this image (no Go toolchain), so oracle/_ref is absent.
"""
from .cref import (  # noqa: F401
    COSINE, L2, DOT, SQL2, L1, ARITH_VECTORTYPES, ARITH_HNSW_F32,
    distance, distances, exact_search, exact_search_batch, synth, build, lib_path,
)
from . import filters, gotypes, rerank  # noqa: F401,E402

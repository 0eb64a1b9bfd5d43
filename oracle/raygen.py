"""TEST INFRASTRUCTURE ONLY -- CPU (numpy) ray-stream generators for the checker side.

The generators themselves live in bench_rays.py at the repository root (neutral: neither engine nor checker), because
both arms of bench.py must feed the very same ray bytes to the GPU and to the CPU path. This module re-exports them under
the name the tests have always used; every parity test still feeds ONE set of bytes to both the oracle and the CUDA path.
"""
from __future__ import annotations

import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from bench_rays import INVALID, RAY_DTYPE, RESULT_DTYPE, bounce_rays, look_at, primary_rays  # noqa: E402,F401

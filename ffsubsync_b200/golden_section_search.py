"""Golden-section search with the evaluation order of the reference
(ffsubsync/golden_section_search.py:15-74), which MaxScoreAligner.fit_gss relies on: the
objective is called as f(x, is_last_iter) and only the evaluation flagged last is recorded.

Host-side control flow (17 strictly sequential evaluations for [0.9, 1.1], tol 1e-4); every
evaluation itself (re-rasterise + align) runs on the GPU.
"""
import logging
import math
from typing import Callable, Tuple

logger = logging.getLogger(__name__)

invphi = (math.sqrt(5) - 1) / 2   # 1 / phi
invphi2 = (3 - math.sqrt(5)) / 2  # 1 / phi^2


def _call(f: Callable, x: float, is_last_iter: bool):
    try:
        return f(x, is_last_iter)
    except TypeError:
        return f(x)


def gss(f: Callable, a: float, b: float, tol: float = 1e-4) -> Tuple[float, float]:
    """Shrink [a, b] around the minimum of a unimodal f until it is no wider than tol."""
    a, b = min(a, b), max(a, b)
    h = b - a
    if h <= tol:
        return a, b
    n = int(math.ceil(math.log(tol / h) / math.log(invphi)))
    logger.info("About to perform %d iterations of golden section search to find the best framerate", n)
    c = a + invphi2 * h
    d = a + invphi * h
    yc = _call(f, c, n == 1)
    yd = _call(f, d, n == 1)
    for k in range(n - 1):
        last = k == n - 2
        h = invphi * h
        if yc < yd:
            b, d, yd = d, c, yc
            c = a + invphi2 * h
            yc = _call(f, c, last)
        else:
            a, c, yc = c, d, yd
            d = a + invphi * h
            yd = _call(f, d, last)
    return (a, d) if yc < yd else (c, b)

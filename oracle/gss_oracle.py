"""Restatement of the reference's golden-section search (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/ffsubsync/golden_section_search.py:15-74 and its use in
aligners.py:111-129 (minimise -score over ratio in [0.9, 1.1], only the last evaluation is
recorded).  Pinned against the reference's evaluation sequence by tests/golden/make_golden.py.
"""
import math
from typing import Callable, List, Tuple

INV_PHI = (math.sqrt(5) - 1) / 2
INV_PHI2 = (3 - math.sqrt(5)) / 2


def num_iterations(a: float, b: float, tol: float = 1e-4) -> int:
    h = abs(b - a)
    if h <= tol:
        return 0
    return int(math.ceil(math.log(tol / h) / math.log(INV_PHI)))


def golden_section_trace(
    f: Callable[[float, bool], float], a: float, b: float, tol: float = 1e-4
) -> Tuple[Tuple[float, float], List[Tuple[float, bool]]]:
    """Run the search; return (final interval, [(x, is_last_iter) for every call of f])."""
    a, b = min(a, b), max(a, b)
    calls: List[Tuple[float, bool]] = []

    def ev(x: float, last: bool) -> float:
        calls.append((x, last))
        return f(x, last)

    h = b - a
    if h <= tol:
        return (a, b), calls
    n = num_iterations(a, b, tol)
    c, d = a + INV_PHI2 * h, a + INV_PHI * h
    yc, yd = ev(c, n == 1), ev(d, n == 1)
    for k in range(n - 1):
        last = k == n - 2
        h = INV_PHI * h
        if yc < yd:
            b, d, yd = d, c, yc
            c = a + INV_PHI2 * h
            yc = ev(c, last)
        else:
            a, c, yc = c, d, yd
            d = a + INV_PHI * h
            yd = ev(d, last)
    return ((a, d) if yc < yd else (c, b)), calls

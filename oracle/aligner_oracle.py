"""Float64 numpy restatement of the reference aligners (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/ffsubsync/aligners.py:24-167.  Pinned against the
reference itself by tests/golden/make_golden.py (see oracle/__init__.py).

Two views of the same algorithm are provided:

* ``fft_align``      - what the reference literally computes (complex128 FFTs,
                       ``-inf`` mask, first-index argmax).  This is the parity target.
* ``exact_score`` /
  ``exact_align``    - the closed form the FFT evaluates,
                       ``score(o) = sum_j s'[j] * r'[j + o]`` with ``x' = 2x - 1`` and
                       out-of-range terms equal to 0,
                       evaluated by direct summation.  Used to state what "exact"
                       means for the GPU re-score step and to bound FFT round-off.
"""
import math
from typing import List, Optional, Sequence, Tuple

import numpy as np


class OracleAlignmentFailure(Exception):
    """Mirror of FailedToFindAlignmentException (aligners.py:20-21)."""


def _as_signal(s) -> np.ndarray:
    # aligners.py:51-57: '0'/'1' strings become int lists; everything -> 2*float(x) - 1
    if isinstance(s, str):
        s = [int(ch) for ch in s]
    return 2.0 * np.asarray(s).astype(float) - 1.0


def padded_length(n_ref: int, n_sub: int) -> int:
    # aligners.py:67-68 (math.log(x, 2) then ceil - reproduced literally, quirks included)
    return int(2 ** math.ceil(math.log(n_ref + n_sub, 2)))


def correlation(ref, sub) -> np.ndarray:
    """aligners.py:67-74: the length-N real sequence the reference calls ``convolve``."""
    r, s = _as_signal(ref), _as_signal(sub)
    if len(r) == 0 or len(s) == 0:  # aligners.py:58-66
        raise OracleAlignmentFailure(
            "cannot align empty speech data (reference length=%d, subtitle length=%d)"
            % (len(r), len(s))
        )
    n = padded_length(len(r), len(s))
    extra = n - len(r) - len(s)
    a = np.concatenate([np.zeros(extra + len(r)), s])
    b = np.concatenate([r, np.zeros(len(s) + extra)])[::-1]
    return np.real(np.fft.ifft(np.fft.fft(a) * np.fft.fft(b)))


def surviving_index_range(n: int, n_sub: int, max_offset_samples: Optional[int]) -> Tuple[int, int]:
    """Half-open index range [lo, hi) left finite by aligners.py:31-43.

    ``conv[:a] = -inf; conv[b:] = -inf`` with a = n-1-max-n_sub, b = n-1+max-n_sub and
    Python slice semantics (a negative ``a`` wraps once, then clamps to 0).
    """
    if max_offset_samples is None:
        return 0, n
    a = n - 1 - max_offset_samples - n_sub
    b = n - 1 + max_offset_samples - n_sub
    lo = min(a, n) if a >= 0 else max(a + n, 0)
    if b >= 0:
        hi = min(b, n)
    else:  # cannot happen for n >= n_sub + 1, kept for completeness
        hi = max(b + n, 0)
    return lo, hi


def offset_range(n_ref: int, n_sub: int, max_offset_samples: Optional[int]) -> Tuple[int, int]:
    """Inclusive offset range [o_lo, o_hi] that survives the mask (may be empty: o_lo > o_hi)."""
    n = padded_length(n_ref, n_sub)
    lo, hi = surviving_index_range(n, n_sub, max_offset_samples)
    # aligners.py:47: offset = n - 1 - idx - n_sub
    return n - n_sub - hi, n - 1 - n_sub - lo


def fft_align(ref, sub, max_offset_samples: Optional[int] = None) -> Tuple[float, int]:
    """(best_score_, best_offset_) exactly as FFTAligner.fit leaves them (aligners.py:45-48,75-78)."""
    conv = correlation(ref, sub)
    n, n_sub = len(conv), len(_as_signal(sub))
    lo, hi = surviving_index_range(n, n_sub, max_offset_samples)
    masked = np.full(n, -np.inf)
    masked[lo:hi] = conv[lo:hi]
    idx = int(np.argmax(masked))
    return float(masked[idx]), n - 1 - idx - n_sub


def exact_score(ref, sub, offset: int) -> float:
    """sum_j s'[j] r'[j+offset] over the overlap, in float64 (math.fsum: correctly rounded)."""
    r, s = _as_signal(ref), _as_signal(sub)
    j_lo = max(0, -offset)
    j_hi = min(len(s), len(r) - offset)
    if j_hi <= j_lo:
        return 0.0
    return math.fsum((s[j_lo:j_hi] * r[j_lo + offset : j_hi + offset]).tolist())


def exact_scores_window(ref, sub, o_lo: int, o_hi: int) -> np.ndarray:
    """Exact scores for every offset in [o_lo, o_hi] (direct O(W*S) summation; small inputs)."""
    return np.array([exact_score(ref, sub, o) for o in range(o_lo, o_hi + 1)])


def exact_align(ref, sub, max_offset_samples: Optional[int] = None) -> Tuple[float, int]:
    """The reference's answer under exact arithmetic: max score, ties -> lowest index
    (= largest offset), candidates = every index of the length-N array that survives
    the mask, including the structural zeros (no-overlap offsets)."""
    r = np.asarray(list(ref) if isinstance(ref, str) else ref)
    s = np.asarray(list(sub) if isinstance(sub, str) else sub)
    if len(r) == 0 or len(s) == 0:
        raise OracleAlignmentFailure("cannot align empty speech data")
    o_lo, o_hi = offset_range(len(r), len(s), max_offset_samples)
    if o_lo > o_hi:  # everything masked: argmax of all -inf is index 0
        n = padded_length(len(r), len(s))
        return -np.inf, n - 1 - len(s)
    scores = exact_scores_window(ref, sub, o_lo, o_hi)
    best = int(np.argmax(scores[::-1]))  # largest offset first == lowest index first
    o = o_hi - best
    return float(scores[o - o_lo]), o


def max_score_select(
    results: Sequence[Tuple[float, int]], max_offset_samples: Optional[int]
) -> int:
    """Index into ``results`` chosen by MaxScoreAligner.transform (aligners.py:154-167):
    drop |offset| > max_offset_samples, raise if none left, highest score, first wins ties."""
    keep = [
        i
        for i, (_, off) in enumerate(results)
        if max_offset_samples is None or abs(off) <= max_offset_samples
    ]
    if not keep:
        raise OracleAlignmentFailure("Synchronization failed; consider passing --max-offset-seconds")
    best = keep[0]
    for i in keep[1:]:
        if results[i][0] > results[best][0]:
            best = i
    return best


def max_offset_samples_of(sample_rate, max_offset_seconds) -> Optional[int]:
    # aligners.py:98-101
    if sample_rate is None or max_offset_seconds is None:
        return None
    return abs(int(max_offset_seconds * sample_rate))


def max_score_align(
    ref, subs: List, sample_rate=None, max_offset_seconds=None
) -> Tuple[Tuple[float, int], int]:
    """MaxScoreAligner(FFTAligner, None, sample_rate, max_offset_seconds).fit_transform(ref, subs)
    for raw-array ``subs`` (aligners.py:131-167).  Returns ((score, offset), index_of_winner)."""
    mos = max_offset_samples_of(sample_rate, max_offset_seconds)
    results = [fft_align(ref, s, mos) for s in subs]
    k = max_score_select(results, mos)
    return results[k], k

"""FFTAligner / MaxScoreAligner with the reference's API (ffsubsync/aligners.py:20-167), computed
by the CUDA library (b2_align_batch: windowed shared-memory FFT correlation, exact float64
re-score of the near-maximal candidates, argmax).

Semantics kept from the reference:
  * inputs may be '0'/'1' strings, lists or arrays of any float values; x -> 2x-1 inside;
  * empty reference or subtitles -> FailedToFindAlignmentException("cannot align empty speech data ...");
  * ``max_offset_samples`` mask incl. its slice corner cases; argmax ties -> largest offset;
  * ``MaxScoreAligner`` accepts an aligner class or instance, a Pipeline / list of fitted
    pipelines, raw arrays, or callables (golden-section search over the ratio);
    first-in-list wins score ties; |offset| filter; the failure message.
Difference (documented in DESIGN.md): signals are handed to the GPU as float32, scores come back
as exact float64 sums over those float32 values (equal to the reference to ~1e-8 relative for
non-binary levels, exactly for binary ones).
"""
import logging
from typing import List, Optional, Tuple, Type, Union

import numpy as np

from . import _native
from .golden_section_search import gss
from .sklearn_shim import Pipeline, TransformerMixin

logger: logging.Logger = logging.getLogger(__name__)

MIN_FRAMERATE_RATIO = 0.9
MAX_FRAMERATE_RATIO = 1.1


class FailedToFindAlignmentException(Exception):
    pass


def _as_float_signal(s) -> np.ndarray:
    if isinstance(s, str):
        s = list(map(int, s))
    return np.array(s).astype(float)


def _raise_empty(n_ref: int, n_sub: int) -> None:
    raise FailedToFindAlignmentException(
        "cannot align empty speech data "
        "(reference length=%d, subtitle length=%d); "
        "the reference or subtitles may contain no detectable speech" % (n_ref, n_sub))


def align_many(ref: np.ndarray, subs: List[np.ndarray], max_offset_samples: Optional[int]
               ) -> List[Tuple[float, int]]:
    """One reference against K subtitle signals in a single GPU call -> [(score, offset)] * K."""
    for s in subs:
        if len(ref) == 0 or len(s) == 0:
            _raise_empty(len(ref), len(s))
    K = len(subs)
    sub_off = np.concatenate([[0], np.cumsum([len(s) for s in subs])]).astype(np.int64)
    sub_cat = np.concatenate(subs).astype(np.float32) if K else np.zeros(0, np.float32)
    score, offset, status = _native.get_handle().align_batch(
        ref.astype(np.float32), [0, len(ref)], sub_cat, sub_off, 1, K, max_offset_samples)
    if np.any(status & _native.ALIGN_CAND_OVERFLOW):
        logger.debug("near-degenerate correlation: more tied candidates than the re-score budget")
    return [(float(score[k]), int(offset[k])) for k in range(K)]


class FFTAligner(TransformerMixin):
    def __init__(self, max_offset_samples: Optional[int] = None) -> None:
        self.max_offset_samples: Optional[int] = max_offset_samples
        self.best_offset_: Optional[int] = None
        self.best_score_: Optional[float] = None
        self.get_score_: bool = False

    def fit(self, refstring, substring, get_score: bool = False) -> "FFTAligner":
        ref, sub = _as_float_signal(refstring), _as_float_signal(substring)
        (self.best_score_, self.best_offset_), = align_many(ref, [sub], self.max_offset_samples)
        self.get_score_ = get_score
        return self

    def transform(self, *_) -> Union[int, Tuple[float, int]]:
        if self.get_score_:
            return self.best_score_, self.best_offset_
        return self.best_offset_


class MaxScoreAligner(TransformerMixin):
    def __init__(
        self,
        base_aligner: Union[FFTAligner, Type[FFTAligner]],
        srtin: Optional[str] = None,
        sample_rate=None,
        max_offset_seconds=None,
    ) -> None:
        # aligners.py:90-109: seconds -> samples (None when either is missing); an aligner CLASS is
        # instantiated with that mask, an INSTANCE is used as it is
        have_mask = sample_rate is not None and max_offset_seconds is not None
        self.max_offset_samples: Optional[int] = (
            abs(int(max_offset_seconds * sample_rate)) if have_mask else None)
        self.max_offset_seconds: Optional[int] = max_offset_seconds
        self.srtin: Optional[str] = srtin
        self.base_aligner: FFTAligner = (
            base_aligner(max_offset_samples=self.max_offset_samples)
            if isinstance(base_aligner, type) else base_aligner)
        self._scores: List[Tuple[Tuple[float, int], Pipeline]] = []

    def _align_one(self, refstring, substring) -> Tuple[float, int]:
        return self.base_aligner.fit_transform(refstring, substring, get_score=True)

    def fit_gss(self, refstring, subpipe_maker):
        """Golden-section search over the framerate ratio (aligners.py:111-129): 17 evaluations of
        rescale -> rasterise -> align on [0.9, 1.1]; only the last one is kept in ``_scores``."""

        def negative_score(ratio: float, final: bool) -> float:
            pipe = subpipe_maker(ratio)
            result = self._align_one(refstring, pipe.fit_transform(self.srtin))
            logger.info("got score %.0f (offset %d) for ratio %.3f", result[0], result[1], ratio)
            if final:
                self._scores.append((result, pipe))
            return -result[0]

        gss(negative_score, MIN_FRAMERATE_RATIO, MAX_FRAMERATE_RATIO)
        return self

    def fit(self, refstring, subpipes: Union[Pipeline, List[Pipeline]]) -> "MaxScoreAligner":
        if not isinstance(subpipes, list):
            subpipes = [subpipes]
        batched = type(self.base_aligner) is FFTAligner
        ref = _as_float_signal(refstring) if batched else refstring
        pending: List[Tuple[int, np.ndarray]] = []   # (slot in _scores, signal) awaiting one GPU call
        slots: List[Optional[Tuple[Tuple[float, int], object]]] = []

        def flush():
            if not pending:
                return
            results = align_many(ref, [s for _, s in pending], self.base_aligner.max_offset_samples)
            for (slot, _), res in zip(pending, results):
                slots[slot] = (res, slots[slot][1])
            self.base_aligner.best_score_, self.base_aligner.best_offset_ = results[-1]
            self.base_aligner.get_score_ = True
            pending.clear()

        for subpipe in subpipes:
            if callable(subpipe):
                flush()
                before = len(self._scores)
                self.fit_gss(refstring, subpipe)
                slots.extend(self._scores[before:])
                del self._scores[before:]
                continue
            substring = subpipe.transform(self.srtin) if hasattr(subpipe, "transform") else subpipe
            if batched:
                slots.append((None, subpipe))
                pending.append((len(slots) - 1, _as_float_signal(substring)))
            else:
                slots.append((self._align_one(refstring, substring), subpipe))
        flush()
        self._scores.extend(slots)
        return self

    def transform(self, *_) -> Tuple[Tuple[float, float], Pipeline]:
        # aligners.py:154-167: drop candidates beyond the offset limit, highest score wins, the
        # first in list order on equal scores (max() keeps the first maximum)
        limit = self.max_offset_samples
        kept = [c for c in self._scores if limit is None or abs(c[0][1]) <= limit]
        if not kept:
            raise FailedToFindAlignmentException(
                "Synchronization failed; consider passing --max-offset-seconds with a number larger "
                "than {}".format(self.max_offset_seconds))
        best = max(kept, key=lambda c: c[0][0])
        return (best[0][0], best[0][1]), best[1]

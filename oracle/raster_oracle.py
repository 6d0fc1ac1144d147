"""Restatement of the subtitle-side signal producers (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/ffsubsync:
  * subtitle_transformers.py:35-47   SubtitleScaler.fit   (cue times * ratio, via timedelta)
  * speech_transformers.py:957-980   SubtitleSpeechTransformer.fit (100 Hz rasterisation)
  * speech_transformers.py:299-317   ComputeSpeechFrameBoundariesMixin
  * speech_transformers.py:928-943   _is_metadata
Pinned against the reference classes by tests/golden/make_golden.py.
"""
import math
from datetime import timedelta
from typing import List, Optional, Sequence, Tuple

import numpy as np


# ---------------------------------------------------------------- SubtitleScaler

def scale_seconds(t_seconds: float, ratio: float) -> float:
    """What a cue time becomes after SubtitleScaler: the product goes through
    ``timedelta(seconds=...)`` (microsecond rounding) and back through ``total_seconds()``
    (subtitle_transformers.py:41-42; read back at speech_transformers.py:960,968-973)."""
    return timedelta(seconds=t_seconds * ratio).total_seconds()


def seconds_via_timedelta_closed_form(x: float) -> float:
    """Arithmetic-only equivalent of ``timedelta(seconds=x).total_seconds()`` for a float x:
    whole seconds are kept exactly, the fractional part is multiplied by 1e6 in double
    precision and rounded half-to-even to an integer number of microseconds, and the total
    microsecond count is divided by 1e6 (one correctly-rounded IEEE division).  This is the
    form the CUDA rasteriser evaluates; tests check it against ``datetime.timedelta``."""
    frac, whole = math.modf(x)
    us = int(whole) * 1000000 + int(round(frac * 1e6))  # Python round(): half-to-even
    return us / 1e6


# ---------------------------------------------------------------- metadata filter

_OPEN_TO_CLOSE = {"(": ")", "{": "}", "[": "]", "（": "）", "【": "】", "「": "」"}
_MUSIC = set("♪♫♬♩\U0001F3B5\U0001F3B6")


def is_metadata(content: str, is_beginning_or_end: bool) -> bool:
    """speech_transformers.py:928-943: cue text that carries no speech."""
    import re

    text = re.sub(r"<[^>]+>", "", content).strip()
    if not text:
        return True
    if text[0] in _OPEN_TO_CLOSE and text[-1] == _OPEN_TO_CLOSE[text[0]]:
        return True
    if all(c.isspace() or c in _MUSIC for c in text):
        return True
    if is_beginning_or_end and ("english" in text.lower() or " - " in text):
        return True
    return False


# ---------------------------------------------------------------- rasteriser

def _py_slice_bounds(start: int, end: int, n: int) -> Tuple[int, int]:
    """Python/numpy slice normalisation for ``a[start:end]`` on a length-n array."""
    lo, hi, _ = slice(start, end).indices(n)
    return lo, max(lo, hi)


def rasterize(
    starts_s: Sequence[float],
    ends_s: Sequence[float],
    keep: Optional[Sequence[bool]] = None,
    sample_rate: int = 100,
    start_seconds: float = 0,
    ratio: float = 1.0,
    scale: bool = True,
):
    """SubtitleScaler(ratio) followed by SubtitleSpeechTransformer(sample_rate, start_seconds, ratio).

    ``starts_s``/``ends_s`` are the *unscaled* cue times in seconds (``total_seconds()`` of the
    parsed cues); ``keep[i]`` is False for cues ``_is_metadata`` drops.  With ``scale=False``
    the times are taken as already scaled (SubtitleSpeechTransformer alone).
    Returns (samples float64[int(max_time*sr)+2], max_time_, start_frame_, end_frame_).
    """
    n_cues = len(starts_s)
    if keep is None:
        keep = [True] * n_cues
    if scale:
        st = [scale_seconds(float(t), ratio) for t in starts_s]
        en = [scale_seconds(float(t), ratio) for t in ends_s]
    else:
        st = [float(t) for t in starts_s]
        en = [float(t) for t in ends_s]
    max_time = 0
    for e in en:  # speech_transformers.py:958-960 (metadata cues count too)
        max_time = max(max_time, e)
    samples = np.zeros(int(max_time * sample_rate) + 2, dtype=float)
    level = min(1.0 / ratio, 1.0)  # speech_transformers.py:977
    for i in range(n_cues):
        if not keep[i]:
            continue
        first = int(round((st[i] - start_seconds) * sample_rate))
        last = first + int(round((en[i] - st[i]) * sample_rate))
        lo, hi = _py_slice_bounds(first, last, len(samples))
        samples[lo:hi] = level
    start_frame, end_frame = frame_boundaries(samples)
    return samples, max_time - start_seconds, start_frame, end_frame


def frame_boundaries(speech_frames: np.ndarray) -> Tuple[Optional[int], Optional[int]]:
    """speech_transformers.py:310-317: first / last index with value > 0.5 (None, None if none)."""
    nz = np.flatnonzero(np.asarray(speech_frames) > 0.5)
    if len(nz) == 0:
        return None, None
    return int(nz[0]), int(nz[-1])


def synthetic_cues(seed: int, duration_s: float = 7200.0) -> Tuple[np.ndarray, np.ndarray]:
    """SURVEY.md section 8d generator: t=5; repeat d~U(1,5), cue (t,t+d), t += d + Exp(3)
    until t >= duration-10.  Times are rounded to milliseconds like an SRT file."""
    rng = np.random.RandomState(seed)
    t = 5.0
    starts: List[float] = []
    ends: List[float] = []
    while t < duration_s - 10.0:
        d = rng.uniform(1.0, 5.0)
        starts.append(round(t, 3))
        ends.append(round(t + d, 3))
        t += d + rng.exponential(3.0)
    return np.array(starts), np.array(ends)

"""Synthetic (video, subtitle) pairs for tests and bench.

A pair is a seeded cue list plus a reference speech mask that equals the subtitle mask at one
of the framerate ratios, delayed by a known number of 10 ms frames, with a fraction of the
frames flipped.  The reference PCM is synthesised ON THE DEVICE from the mask by the library's
counter-hash generator (b2_synth_pcm), so a 256-pair batch (59 GB of 16 kHz PCM) never crosses
PCIe; numpy can replay any window of it (oracle/vad_oracle.py:synth_pcm).
"""
from typing import List, NamedTuple, Sequence

import numpy as np

from . import _native

BENCH_RATIOS = [1.0, 24.0 / 23.976, 25.0 / 24.0, 23.976 / 24.0, 24.0 / 25.0]  # config 3 (K = 5)


class PairBatch(NamedTuple):
    window_class: np.ndarray   # uint8 per 10 ms reference window, all pairs back to back
    win_off: np.ndarray        # [B+1] window offsets
    cue_start: np.ndarray      # float64 seconds, all pairs back to back
    cue_end: np.ndarray
    cue_off: np.ndarray        # [B+1]
    true_k: np.ndarray         # index into the ratio list the reference was built from
    true_offset: np.ndarray    # frames the subtitles must move (positive = later)


def synthetic_cues(seed: int, duration_s: float):
    """t=5; repeat d~U(1,5): cue (t, t+d); t += d + Exp(mean 3) until t >= duration-10.
    Millisecond-rounded like an SRT file (~1.2 k cues, ~50 % duty for 2 h)."""
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


def make_pairs(seeds: Sequence[int], duration_s: float, ratios: Sequence[float], sample_rate: int = 100,
               max_shift: int = 3000, flip_fraction: float = 0.10, hiss_fraction: float = 0.05,
               handle=None) -> PairBatch:
    handle = handle or _native.get_handle()
    B = len(seeds)
    cues = [synthetic_cues(int(s), duration_s) for s in seeds]
    cue_off = np.concatenate([[0], np.cumsum([len(c[0]) for c in cues])]).astype(np.int64)
    cue_start = np.concatenate([c[0] for c in cues])
    cue_end = np.concatenate([c[1] for c in cues])
    rngs = [np.random.RandomState(int(s) + 100003) for s in seeds]
    true_k = np.array([r.randint(0, len(ratios)) for r in rngs])
    true_offset = np.array([r.randint(-max_shift, max_shift + 1) for r in rngs])
    pair_ratio = np.array([ratios[k] for k in true_k], dtype=np.float64)
    masks, mask_off = handle.rasterize(cue_start, cue_end, None, cue_off, pair_ratio, 1, True, sample_rate, 0.0)
    n = int(duration_s * sample_rate)
    cls = np.zeros(B * n, dtype=np.uint8)
    for b in range(B):
        m = masks[mask_off[b]:mask_off[b + 1]] != 0
        ref = np.zeros(n, dtype=bool)
        src = np.arange(n) - true_offset[b]
        ok = (src >= 0) & (src < len(m))
        ref[ok] = m[src[ok]]
        u = rngs[b].rand(n)
        ref ^= u < flip_fraction
        hiss = rngs[b].rand(n) < hiss_fraction
        cls[b * n:(b + 1) * n] = np.where(ref, 1, np.where(hiss, 2, 0))
    win_off = np.arange(B + 1, dtype=np.int64) * n
    return PairBatch(cls, win_off, cue_start, cue_end, cue_off, true_k, true_offset)

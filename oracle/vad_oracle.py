"""numpy restatement of the energy / zero-crossing detector (TEST INFRASTRUCTURE ONLY).

PARITY UNPINNED for the detector arithmetic (see oracle/__init__.py): the reference has no
in-tree energy/ZCR VAD; its detectors are third-party wheels that are not installed here.
What IS pinned to the reference is the contract every detector obeys:

  * input  = raw s16le mono PCM bytes (or a uint8 view of them) for one <=100 s chunk
             (speech_transformers.py:710-746),
  * window = int(frame_rate / sample_rate + 0.5) samples (speech_transformers.py:163-164),
  * output = 1-D float array, one value per window start in range(0, n_samples, window)
             -> ceil(n_samples / window) values, 1.0 for speech and ``non_speech_label``
             otherwise (speech_transformers.py:169-181),
  * a trailing partial window is non-speech (webrtcvad raises on it and the reference
    maps the exception to non-speech, speech_transformers.py:176-178).

The rule itself (this repo's definition, integer-exact so CPU == GPU bit for bit):

    E = sum x[i]^2                         (int64, over the window's samples)
    Z = #{ i in 1..window-1 : (x[i] < 0) != (x[i-1] < 0) }   (crossings inside the window)
    speech  <=>  E >= window * energy_threshold  and  z_lo <= Z <= z_hi

``energy_threshold = 10**5`` is auditok's ``energy_threshold=50`` (speech_transformers.py:125:
10*log10(mean x^2) >= 50) without the logarithm.
"""
from typing import Optional, Tuple

import numpy as np

DEFAULT_ENERGY_THRESHOLD = 100000


def frames_per_window(frame_rate: int, sample_rate: int) -> int:
    return int((1.0 / sample_rate) * frame_rate + 0.5)


def default_zcr_band(fpw: int) -> Tuple[int, int]:
    return 0, (3 * fpw) // 8


def window_features(pcm: np.ndarray, fpw: int) -> Tuple[np.ndarray, np.ndarray]:
    """(E int64[n_full], Z int64[n_full]) for the full windows of an int16 array."""
    n_full = len(pcm) // fpw
    x = pcm[: n_full * fpw].astype(np.int64).reshape(n_full, fpw)
    energy = (x * x).sum(axis=1)
    neg = x < 0
    crossings = (neg[:, 1:] != neg[:, :-1]).sum(axis=1).astype(np.int64)
    return energy, crossings


def energy_zcr_detect(
    asegment,
    sample_rate: int = 100,
    frame_rate: int = 16000,
    non_speech_label: float = 0.0,
    energy_threshold: int = DEFAULT_ENERGY_THRESHOLD,
    z_lo: Optional[int] = None,
    z_hi: Optional[int] = None,
) -> np.ndarray:
    """One detector call on one chunk of s16le bytes / uint8 array / int16 array."""
    if isinstance(asegment, (bytes, bytearray, memoryview)):
        raw = np.frombuffer(asegment, dtype=np.uint8)
    else:
        raw = np.asarray(asegment)
    if raw.dtype == np.int16:
        pcm = raw
    else:
        raw = raw.astype(np.uint8, copy=False)
        pcm = raw[: (len(raw) // 2) * 2].view("<i2")
    fpw = frames_per_window(frame_rate, sample_rate)
    dlo, dhi = default_zcr_band(fpw)
    z_lo = dlo if z_lo is None else z_lo
    z_hi = dhi if z_hi is None else z_hi
    n_windows = (len(pcm) + fpw - 1) // fpw
    out = np.full(n_windows, float(non_speech_label))
    energy, crossings = window_features(pcm, fpw)
    speech = (energy >= fpw * int(energy_threshold)) & (crossings >= z_lo) & (crossings <= z_hi)
    out[: len(speech)][speech] = 1.0
    return out


# ------------------------------------------------------------------ synthetic PCM
# Counter-based generator shared (bit for bit) with the CUDA synthesiser
# (ffsubsync_b200/csrc/synth.cu): every sample depends only on (seed, sample index,
# class of its window), so any window can be replayed on the host for spot checks.

def lowbias32(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint32)
    x ^= x >> np.uint32(16)
    x = (x * np.uint32(0x7FEB352D)).astype(np.uint32)
    x ^= x >> np.uint32(15)
    x = (x * np.uint32(0x846CA68B)).astype(np.uint32)
    x ^= x >> np.uint32(16)
    return x


def synth_pcm(window_class: np.ndarray, fpw: int, seed: int, first_sample: int = 0) -> np.ndarray:
    """int16 PCM for consecutive windows whose classes are given (uint8):
         0 = silence      : noise in [-32, 31]
         1 = voiced speech: +-6000 square wave, half-period 40 samples, plus noise in [-2048, 2047]
         2 = loud hiss    : noise in [-8192, 8191]  (high energy, high zero-crossing count)
    ``first_sample`` is the global index of the first generated sample (for the hash counter)."""
    n = len(window_class) * fpw
    idx = (np.arange(n, dtype=np.uint64) + np.uint64(first_sample)).astype(np.uint32)
    h = lowbias32(idx ^ np.uint32(seed & 0xFFFFFFFF))
    cls = np.repeat(np.asarray(window_class, dtype=np.uint8), fpw)
    noise13 = (h & np.uint32(0x3FFF)).astype(np.int32) - 8192       # [-8192, 8191]
    noise11 = ((h >> np.uint32(14)) & np.uint32(0xFFF)).astype(np.int32) - 2048  # [-2048, 2047]
    noise5 = ((h >> np.uint32(26)) & np.uint32(0x3F)).astype(np.int32) - 32       # [-32, 31]
    pos = (np.arange(n, dtype=np.int64) % fpw)
    square = np.where(((pos // 40) & 1) == 0, 6000, -6000).astype(np.int32)
    out = np.where(cls == 0, noise5, np.where(cls == 1, square + noise11, noise13))
    return out.astype(np.int16)

"""ctypes binding of the C ABI in include/ffsubsync_b200.h.

There is deliberately NO CPU fallback: if the CUDA library is missing or no B200 is visible,
every compute entry point raises.  Build the library with ``python __graft_entry__.py``.
"""
import ctypes
import os
import threading
from typing import Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libffsubsync_b200.so")

B2_HOST, B2_DEVICE = 0, 1
B2_DEVICE_RESIDENT = 2   # b2_sync_batch only: device inputs that nothing queued before the call still writes
B2_MAX_OFFSET_NONE = -(1 << 63)   # INT64_MIN: FFTAligner(max_offset_samples=None)
ALIGN_OK, ALIGN_EMPTY, ALIGN_ALL_MASKED, ALIGN_CAND_OVERFLOW = 0, 1, 2, 4
STATUS_NAMES = {0: "B2_OK", -1: "B2_ERR_BAD_ARG", -2: "B2_ERR_CUDA", -3: "B2_ERR_EMPTY_INPUT",
                -4: "B2_ERR_NO_ALIGNMENT", -5: "B2_ERR_NOMEM", -6: "B2_ERR_UNSUPPORTED"}

# every symbol declared in include/ffsubsync_b200.h (tests check the header against this list)
EXPORTS = [
    "b2_version", "b2_create", "b2_destroy", "b2_set_stream", "b2_synchronize", "b2_last_error",
    "b2_launch_count", "b2_vad_frames_per_window", "b2_vad_num_windows", "b2_vad_energy_zcr",
    "b2_rasterize_lengths", "b2_rasterize", "b2_blend_signals", "b2_first_last_nonzero", "b2_align_batch",
    "b2_reduce_ratios", "b2_sync_batch", "b2_synth_pcm", "b2_vad_stream_begin", "b2_vad_stream_push",
    "b2_vad_stream_windows", "b2_vad_stream_end", "b2_auditok_block_size", "b2_auditok_energy_floor",
    "b2_vad_auditok",
]


class NativeError(RuntimeError):
    def __init__(self, status: int, where: str, detail: str = ""):
        self.status = status
        super().__init__("%s failed with %s%s" % (where, STATUS_NAMES.get(status, status),
                                                  (": " + detail) if detail else ""))


_lib = None
_lib_lock = threading.Lock()
_tls = threading.local()

_vp, _i32, _i64, _f32, _f64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_double


def load() -> ctypes.CDLL:
    """Load the shared library (no CUDA initialisation happens here)."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "ffsubsync_b200: CUDA library %s not found - build it with "
                "`python __graft_entry__.py` (there is no CPU fallback)" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        lib.b2_version.restype = ctypes.c_int
        lib.b2_create.argtypes = [ctypes.c_int, ctypes.POINTER(_vp)]
        lib.b2_destroy.argtypes = [_vp]
        lib.b2_set_stream.argtypes = [_vp, _vp]
        lib.b2_synchronize.argtypes = [_vp]
        lib.b2_last_error.argtypes = [_vp]
        lib.b2_last_error.restype = ctypes.c_char_p
        lib.b2_launch_count.argtypes = [_vp]
        lib.b2_launch_count.restype = _i64
        lib.b2_vad_frames_per_window.argtypes = [ctypes.c_int, ctypes.c_int]
        lib.b2_vad_num_windows.argtypes = [_i64, ctypes.c_int, ctypes.c_int]
        lib.b2_vad_num_windows.restype = _i64
        lib.b2_vad_energy_zcr.argtypes = [_vp, _vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          _f32, _i64, ctypes.c_int, ctypes.c_int, _vp, _vp, ctypes.c_int]
        lib.b2_rasterize_lengths.argtypes = [_vp, _vp, ctypes.c_int, _vp, ctypes.c_int, ctypes.c_int,
                                             ctypes.c_int, _vp]
        lib.b2_rasterize.argtypes = [_vp, _vp, _vp, _vp, _vp, ctypes.c_int, _vp, ctypes.c_int,
                                     ctypes.c_int, _vp, ctypes.c_int, _f64, _vp, _vp, ctypes.c_int]
        lib.b2_first_last_nonzero.argtypes = [_vp, _vp, _vp, ctypes.c_int, _vp, _vp, ctypes.c_int]
        lib.b2_blend_signals.argtypes = [_vp, _vp, _vp, _i64, ctypes.c_int, _f64, _f64, _vp, ctypes.c_int]
        lib.b2_align_batch.argtypes = [_vp, _vp, _vp, _vp, _vp, ctypes.c_int, ctypes.c_int, _i64,
                                       _vp, _vp, _vp, ctypes.c_int]
        lib.b2_reduce_ratios.argtypes = [_vp, _vp, _vp, _vp, ctypes.c_int, ctypes.c_int, _i64,
                                         _vp, _vp, _vp, ctypes.c_int]
        lib.b2_sync_batch.argtypes = [_vp, _vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, _f32,
                                      _i64, ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, _vp, _vp,
                                      ctypes.c_int, _f64, _i64, _vp, _vp, _vp, _vp, _vp, ctypes.c_int]
        lib.b2_synth_pcm.argtypes = [_vp, _vp, _i64, ctypes.c_int, ctypes.c_uint32, _vp, ctypes.c_int]
        lib.b2_vad_stream_begin.argtypes = [_vp, ctypes.c_int, ctypes.c_int, _f32, _i64, ctypes.c_int,
                                            ctypes.c_int]
        lib.b2_vad_stream_push.argtypes = [_vp, _vp, _i64]
        lib.b2_vad_stream_windows.argtypes = [_vp]
        lib.b2_vad_stream_windows.restype = _i64
        lib.b2_vad_stream_end.argtypes = [_vp, _vp, _i64, ctypes.POINTER(_i64)]
        lib.b2_auditok_block_size.argtypes = [ctypes.c_int, ctypes.c_int]
        lib.b2_auditok_energy_floor.argtypes = [ctypes.c_int, _f64]
        lib.b2_auditok_energy_floor.restype = _i64
        lib.b2_vad_auditok.argtypes = [_vp, _vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, _f64, _f64,
                                       _f64, _i64, _f64, _i64, _vp, _vp, ctypes.c_int]
        _lib = lib
        return lib


def _ptr(a) -> Optional[int]:
    """Pointer of a numpy array (kept alive by the caller), an int device pointer, or None."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    return a.ctypes.data


def _mask_width(max_offset_samples) -> int:
    """None -> B2_MAX_OFFSET_NONE; any Python int -> an int64 the library feeds to the reference's
    slice arithmetic (aligners.py:31-43).  Widths beyond +-2^62 behave like every width larger
    than the padded length, so clamping them keeps the result and avoids ctypes wrap-around."""
    if max_offset_samples is None:
        return B2_MAX_OFFSET_NONE
    lim = 1 << 62
    return max(-lim, min(lim, int(max_offset_samples)))


def _i64a(x) -> np.ndarray:
    return np.ascontiguousarray(x, dtype=np.int64)


class Handle:
    """One library handle = one CUDA stream + workspace.  Not thread-safe; see get_handle()."""

    def __init__(self, device: int = 0):
        self.lib = load()
        h = _vp()
        st = self.lib.b2_create(int(device), ctypes.byref(h))
        if st != 0:
            raise NativeError(st, "b2_create(device=%d)" % device,
                              "no usable sm_100 CUDA device; this package has no CPU path")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.b2_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st: int, where: str):
        if st != 0:
            raise NativeError(st, where, self.lib.b2_last_error(self.h).decode("utf-8", "replace"))

    # -- plumbing ---------------------------------------------------------------------------
    def set_stream(self, cuda_stream: Optional[int]):
        """cuda_stream: a cudaStream_t handle as int (e.g. torch.cuda.Stream().cuda_stream);
        0 means the legacy default stream (torch's default), None gives the handle its own stream."""
        if cuda_stream is not None and int(cuda_stream) == 0:
            cuda_stream = 1  # cudaStreamLegacy: NULL would mean "own stream" in the C ABI
        self._check(self.lib.b2_set_stream(self.h, cuda_stream), "b2_set_stream")

    def synchronize(self):
        self._check(self.lib.b2_synchronize(self.h), "b2_synchronize")

    @property
    def launch_count(self) -> int:
        return int(self.lib.b2_launch_count(self.h))

    # -- VAD ----------------------------------------------------------------------------------
    def frames_per_window(self, frame_rate: int, sample_rate: int) -> int:
        return int(self.lib.b2_vad_frames_per_window(frame_rate, sample_rate))

    def vad_energy_zcr(self, pcm, pcm_off, frame_rate: int, sample_rate: int, non_speech_label: float,
                       energy_threshold: int, z_lo: int = -1, z_hi: int = -1, out=None,
                       memspace: int = B2_HOST):
        """pcm: int16 numpy array (host) or device pointer; pcm_off: [B+1] sample offsets."""
        pcm_off = _i64a(pcm_off)
        B = len(pcm_off) - 1
        fpw = self.frames_per_window(frame_rate, sample_rate)
        if fpw <= 0:
            raise ValueError("bad frame_rate / sample_rate")
        nwin = (np.diff(pcm_off) + fpw - 1) // fpw
        out_off = np.concatenate([[0], np.cumsum(nwin)]).astype(np.int64)
        if memspace == B2_HOST:
            pcm = np.ascontiguousarray(pcm, dtype=np.int16)
            out = np.empty(int(out_off[-1]), dtype=np.float32)
        st = self.lib.b2_vad_energy_zcr(self.h, _ptr(pcm), _ptr(pcm_off), B, frame_rate, sample_rate,
                                        float(non_speech_label), int(energy_threshold), int(z_lo),
                                        int(z_hi), _ptr(out), _ptr(out_off), memspace)
        self._check(st, "b2_vad_energy_zcr")
        return out, out_off

    def vad_auditok(self, pcm, pcm_off, frame_rate: int, sample_rate: int, non_speech_label: float,
                    energy_threshold_db: float = 50.0, min_length: Optional[float] = None,
                    max_length: Optional[int] = None, max_continuous_silence: Optional[float] = None,
                    chunk_samples: int = 0, out=None, memspace: int = B2_HOST):
        """auditok detector over B signals (b2_vad_auditok); tokenizer defaults are the reference's
        (speech_transformers.py:126-131).  Returns (float64 per block, out_off[B+1])."""
        pcm_off = _i64a(pcm_off)
        B = len(pcm_off) - 1
        fpw = int(self.lib.b2_auditok_block_size(frame_rate, sample_rate))
        if fpw <= 0:
            raise ValueError("auditok detector: unsupported frame_rate=%r / sample_rate=%r" % (frame_rate, sample_rate))
        min_length = 0.2 * sample_rate if min_length is None else min_length
        max_length = int(5 * sample_rate) if max_length is None else max_length
        max_continuous_silence = 0.25 * sample_rate if max_continuous_silence is None else max_continuous_silence
        n = np.diff(pcm_off)
        if chunk_samples > 0:
            full, rem = n // chunk_samples, n % chunk_samples
            nwin = full * ((chunk_samples + fpw - 1) // fpw) + (rem + fpw - 1) // fpw
        else:
            nwin = (n + fpw - 1) // fpw
        out_off = np.concatenate([[0], np.cumsum(nwin)]).astype(np.int64)
        if memspace == B2_HOST:
            pcm = np.ascontiguousarray(pcm, dtype=np.int16)
            out = np.empty(int(out_off[-1]), dtype=np.float64)
        st = self.lib.b2_vad_auditok(self.h, _ptr(pcm), _ptr(pcm_off), B, frame_rate, sample_rate,
                                     float(non_speech_label), float(energy_threshold_db), float(min_length),
                                     int(max_length), float(max_continuous_silence), int(chunk_samples),
                                     _ptr(out), _ptr(out_off), memspace)
        self._check(st, "b2_vad_auditok")
        return out, out_off

    # streaming detector (b2_vad_stream_*): push() returns before the chunk is processed
    def vad_stream_begin(self, frame_rate: int, sample_rate: int, non_speech_label: float,
                         energy_threshold: int, z_lo: int = -1, z_hi: int = -1) -> None:
        self._check(self.lib.b2_vad_stream_begin(self.h, frame_rate, sample_rate, float(non_speech_label),
                                                 int(energy_threshold), int(z_lo), int(z_hi)),
                    "b2_vad_stream_begin")

    def vad_stream_push(self, chunk) -> None:
        """chunk: bytes-like or uint8/int16 array (host); copied before the call returns."""
        if isinstance(chunk, np.ndarray):
            buf = np.ascontiguousarray(chunk)
            ptr, n = buf.ctypes.data, buf.nbytes
        else:
            buf = np.frombuffer(chunk, dtype=np.uint8)
            ptr, n = (buf.ctypes.data if len(buf) else None), len(buf)
        self._check(self.lib.b2_vad_stream_push(self.h, ptr, n), "b2_vad_stream_push")

    def vad_stream_end(self) -> np.ndarray:
        n = int(self.lib.b2_vad_stream_windows(self.h))
        out = np.empty(max(n, 0), dtype=np.float32)
        got = _i64(0)
        self._check(self.lib.b2_vad_stream_end(self.h, _ptr(out) if n > 0 else None, max(n, 0),
                                               ctypes.byref(got)), "b2_vad_stream_end")
        return out[: got.value]

    def synth_pcm(self, window_class, n_windows: int, fpw: int, seed: int, out=None,
                  memspace: int = B2_HOST):
        if memspace == B2_HOST:
            window_class = np.ascontiguousarray(window_class, dtype=np.uint8)
            n_windows = len(window_class)
            out = np.empty(n_windows * fpw, dtype=np.int16)
        st = self.lib.b2_synth_pcm(self.h, _ptr(window_class), int(n_windows), int(fpw),
                                   int(seed) & 0xFFFFFFFF, _ptr(out), memspace)
        self._check(st, "b2_synth_pcm")
        return out

    # -- subtitle side ------------------------------------------------------------------------
    def rasterize_lengths(self, cue_end_s, cue_off, ratios, K: int, per_pair: bool, sample_rate: int):
        cue_end_s = np.ascontiguousarray(cue_end_s, dtype=np.float64)
        cue_off = _i64a(cue_off)
        ratios = np.ascontiguousarray(ratios, dtype=np.float64)
        B = len(cue_off) - 1
        lengths = np.empty(B * K, dtype=np.int64)
        st = self.lib.b2_rasterize_lengths(_ptr(cue_end_s), _ptr(cue_off), B, _ptr(ratios), K,
                                           int(per_pair), sample_rate, _ptr(lengths))
        if st != 0:
            raise NativeError(st, "b2_rasterize_lengths")
        return lengths

    def rasterize(self, cue_start_s, cue_end_s, cue_keep, cue_off, ratios, K: int, per_pair: bool,
                  sample_rate: int, start_seconds: float, levels=None, out=None, out_off=None,
                  memspace: int = B2_HOST):
        cue_start_s = np.ascontiguousarray(cue_start_s, dtype=np.float64)
        cue_end_s = np.ascontiguousarray(cue_end_s, dtype=np.float64)
        cue_keep = None if cue_keep is None else np.ascontiguousarray(cue_keep, dtype=np.uint8)
        cue_off = _i64a(cue_off)
        ratios = np.ascontiguousarray(ratios, dtype=np.float64)
        levels = None if levels is None else np.ascontiguousarray(levels, dtype=np.float64)
        B = len(cue_off) - 1
        if out_off is None:
            lengths = self.rasterize_lengths(cue_end_s, cue_off, ratios, K, per_pair, sample_rate)
            out_off = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
        out_off = _i64a(out_off)
        if memspace == B2_HOST:
            out = np.empty(int(out_off[-1]), dtype=np.float32)
        st = self.lib.b2_rasterize(self.h, _ptr(cue_start_s), _ptr(cue_end_s), _ptr(cue_keep),
                                   _ptr(cue_off), B, _ptr(ratios), K, int(per_pair), _ptr(levels),
                                   sample_rate, float(start_seconds), _ptr(out), _ptr(out_off), memspace)
        self._check(st, "b2_rasterize")
        return out, out_off

    def blend_signals(self, a, b, mode: int, wa: float = 0.6, wb: float = 0.4, out=None, n=None,
                      memspace: int = B2_HOST):
        """mode 0 = min, 1 = max, 2 = wa*a + wb*b; host arrays are clipped to their common length."""
        if memspace == B2_HOST:
            n = min(len(a), len(b))
            a = np.ascontiguousarray(a[:n], dtype=np.float32)
            b = np.ascontiguousarray(b[:n], dtype=np.float32)
            out = np.empty(n, dtype=np.float32)
        st = self.lib.b2_blend_signals(self.h, _ptr(a), _ptr(b), int(n), int(mode), float(wa), float(wb),
                                       _ptr(out), memspace)
        self._check(st, "b2_blend_signals")
        return out

    def first_last_nonzero(self, sig, sig_off, memspace: int = B2_HOST, first=None, last=None):
        sig_off = _i64a(sig_off)
        n = len(sig_off) - 1
        if memspace == B2_HOST:
            sig = np.ascontiguousarray(sig, dtype=np.float32)
            first = np.empty(n, dtype=np.int64)
            last = np.empty(n, dtype=np.int64)
        st = self.lib.b2_first_last_nonzero(self.h, _ptr(sig), _ptr(sig_off), n, _ptr(first),
                                            _ptr(last), memspace)
        self._check(st, "b2_first_last_nonzero")
        return first, last

    # -- aligner ------------------------------------------------------------------------------
    def align_batch(self, ref, ref_off, sub, sub_off, B: int, K: int, max_offset_samples: Optional[int],
                    score=None, offset=None, status=None, memspace: int = B2_HOST):
        ref_off, sub_off = _i64a(ref_off), _i64a(sub_off)
        mos = _mask_width(max_offset_samples)
        if memspace == B2_HOST:
            ref = np.ascontiguousarray(ref, dtype=np.float32)
            sub = np.ascontiguousarray(sub, dtype=np.float32)
            score = np.empty(B * K, dtype=np.float64)
            offset = np.empty(B * K, dtype=np.int32)
            status = np.empty(B * K, dtype=np.int32)
        st = self.lib.b2_align_batch(self.h, _ptr(ref), _ptr(ref_off), _ptr(sub), _ptr(sub_off), B, K,
                                     mos, _ptr(score), _ptr(offset), _ptr(status), memspace)
        self._check(st, "b2_align_batch")
        return score, offset, status

    def reduce_ratios(self, score, offset, status, B: int, K: int, max_offset_samples: Optional[int],
                      best_score=None, best_offset=None, best_k=None, memspace: int = B2_HOST):
        mos = _mask_width(max_offset_samples)
        if memspace == B2_HOST:
            score = np.ascontiguousarray(score, dtype=np.float64)
            offset = np.ascontiguousarray(offset, dtype=np.int32)
            status = None if status is None else np.ascontiguousarray(status, dtype=np.int32)
            best_score = np.empty(B, dtype=np.float64)
            best_offset = np.empty(B, dtype=np.int32)
            best_k = np.empty(B, dtype=np.int32)
        st = self.lib.b2_reduce_ratios(self.h, _ptr(score), _ptr(offset), _ptr(status), B, K, mos,
                                       _ptr(best_score), _ptr(best_offset), _ptr(best_k), memspace)
        self._check(st, "b2_reduce_ratios")
        return best_score, best_offset, best_k

    def sync_batch(self, pcm, pcm_off, frame_rate: int, sample_rate: int, non_speech_label: float,
                   energy_threshold: int, z_lo: int, z_hi: int, cue_start_s, cue_end_s, cue_keep,
                   cue_off, ratios, start_seconds: float, max_offset_samples: Optional[int],
                   best_score=None, best_offset=None, best_k=None, all_score=None, all_offset=None,
                   want_all: bool = False, memspace: int = B2_HOST):
        pcm_off, cue_off = _i64a(pcm_off), _i64a(cue_off)
        B = len(pcm_off) - 1
        ratios = np.ascontiguousarray(ratios, dtype=np.float64)
        K = len(ratios)
        cue_start_s = np.ascontiguousarray(cue_start_s, dtype=np.float64)
        cue_end_s = np.ascontiguousarray(cue_end_s, dtype=np.float64)
        cue_keep = None if cue_keep is None else np.ascontiguousarray(cue_keep, dtype=np.uint8)
        mos = _mask_width(max_offset_samples)
        if memspace == B2_HOST:
            pcm = np.ascontiguousarray(pcm, dtype=np.int16)
            best_score = np.empty(B, dtype=np.float64)
            best_offset = np.empty(B, dtype=np.int32)
            best_k = np.empty(B, dtype=np.int32)
            if want_all:
                all_score = np.empty(B * K, dtype=np.float64)
                all_offset = np.empty(B * K, dtype=np.int32)
        st = self.lib.b2_sync_batch(self.h, _ptr(pcm), _ptr(pcm_off), B, frame_rate, sample_rate,
                                    float(non_speech_label), int(energy_threshold), int(z_lo), int(z_hi),
                                    _ptr(cue_start_s), _ptr(cue_end_s), _ptr(cue_keep), _ptr(cue_off),
                                    _ptr(ratios), K, float(start_seconds), mos, _ptr(best_score),
                                    _ptr(best_offset), _ptr(best_k), _ptr(all_score), _ptr(all_offset),
                                    memspace)
        self._check(st, "b2_sync_batch")
        return best_score, best_offset, best_k, all_score, all_offset


def _default_device() -> int:
    """The device a handle is created on when the caller names none: torch's current CUDA device
    when torch is loaded and has a CUDA context (so kernels launch where the caller's ``.cuda()``
    tensors live, also on rank > 0), else LOCAL_RANK when B2_DEVICE_FROM_RANK is set, else 0."""
    import sys
    torch = sys.modules.get("torch")
    if torch is not None:
        try:
            if torch.cuda.is_available() and torch.cuda.is_initialized():
                return int(torch.cuda.current_device())
        except Exception:
            pass
    if os.environ.get("B2_DEVICE_FROM_RANK"):
        return int(os.environ.get("LOCAL_RANK", "0"))
    return 0


def get_handle(device: Optional[int] = None) -> Handle:
    """Per-thread handle (the reference runs several VideoSpeechTransformer.fit on threads)."""
    if device is None:
        device = _default_device()
    handles = getattr(_tls, "handles", None)
    if handles is None:
        handles = _tls.handles = {}
    if device not in handles:
        handles[device] = Handle(device)
    return handles[device]

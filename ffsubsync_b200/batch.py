"""Batch front end of the hot path: many (video, subtitle) pairs per call.

``BatchSynchronizer.sync_device`` keeps everything resident in HBM (PCM in, per-pair
``(score, offset, ratio index)`` out) and launches on torch's current CUDA stream, so callers
can time it with torch.cuda.Event and chain it with NCCL collectives (ffsubsync_b200.distributed).
``sync_host`` is the same call with host buffers (numpy / pinned tensors): the C library copies
the PCM in, runs the identical kernels and copies the results out.

This is what ``ffs ref.mkv -i in.srt`` does per pair in the reference - VideoSpeechTransformer.fit
(ffsubsync/ffsubsync.py:637) followed by MaxScoreAligner over the ratio grid (:196-235) - for a
batch, with the subtitle scaling + rasterisation fused into one kernel.
"""
from typing import Optional, Sequence

import numpy as np

from . import _native
from .constants import DEFAULT_ENERGY_THRESHOLD, DEFAULT_MAX_OFFSET_SECONDS, SAMPLE_RATE


def load_serialized_speech(paths, non_speech_label: float = 0.0):
    """Batch ingest of ``--serialize-speech`` / ``--make-test-case`` artefacts (``.npz`` with key
    "speech", or ``.npy``): each file goes through DeserializeSpeechTransformer
    (ffsubsync/speech_transformers.py:987-1009: values < 1 become ``non_speech_label``).
    Returns (signals float32 back to back, offsets int64[B+1]) ready for ``sync_signals``."""
    from .speech_transformers import DeserializeSpeechTransformer
    sigs = [np.asarray(DeserializeSpeechTransformer(non_speech_label).fit(p).transform(), dtype=np.float32)
            for p in paths]
    off = np.concatenate([[0], np.cumsum([len(s) for s in sigs])]).astype(np.int64)
    return (np.concatenate(sigs) if sigs else np.zeros(0, np.float32)), off


class BatchSynchronizer:
    def __init__(self, ratios: Sequence[float], frame_rate: int = 16000, sample_rate: int = SAMPLE_RATE,
                 non_speech_label: float = 0.0, energy_threshold: int = DEFAULT_ENERGY_THRESHOLD,
                 z_lo: int = -1, z_hi: int = -1, start_seconds: float = 0.0,
                 max_offset_seconds: Optional[float] = DEFAULT_MAX_OFFSET_SECONDS,
                 device: Optional[int] = None) -> None:
        self.ratios = np.ascontiguousarray(ratios, dtype=np.float64)
        self.frame_rate = frame_rate
        self.sample_rate = sample_rate
        self.non_speech_label = non_speech_label
        self.energy_threshold = energy_threshold
        self.z_lo, self.z_hi = z_lo, z_hi
        self.start_seconds = start_seconds
        # MaxScoreAligner.__init__ (ffsubsync/aligners.py:98-101)
        self.max_offset_samples = None if max_offset_seconds is None else abs(int(max_offset_seconds * sample_rate))
        self.handle = _native.get_handle(device)

    def use_torch_stream(self) -> None:
        """Launch on torch's current stream (so torch events / NCCL ops order against our kernels)."""
        import torch
        self.handle.set_stream(torch.cuda.current_stream().cuda_stream)

    def sync_device(self, pcm, pcm_off, cue_start, cue_end, cue_off, cue_keep=None, out=None, all_out=None,
                    inputs_resident: bool = False):
        """pcm: int16 CUDA tensor with all pairs back to back; pcm_off: [B+1] sample offsets (host).
        out: optional dict of preallocated CUDA tensors best_score f64[B], best_offset i32[B],
        best_k i32[B].  Returns that dict; nothing is synchronised.
        inputs_resident=True (B2_DEVICE_RESIDENT): the caller promises that nothing queued on the handle's
        stream before this call still writes ``pcm`` (a corpus that sits in HBM).  Back-to-back calls then
        overlap: the VAD of this batch starts while the last correlation chain of the previous batch is
        still running.  Results are identical; outputs stay ordered on the stream."""
        import torch
        B = len(pcm_off) - 1
        K = len(self.ratios)
        dev = pcm.device
        if out is None:
            out = {"best_score": torch.empty(B, dtype=torch.float64, device=dev),
                   "best_offset": torch.empty(B, dtype=torch.int32, device=dev),
                   "best_k": torch.empty(B, dtype=torch.int32, device=dev)}
        a_s = all_out["score"].data_ptr() if all_out else None
        a_o = all_out["offset"].data_ptr() if all_out else None
        self.handle.sync_batch(
            pcm.data_ptr(), pcm_off, self.frame_rate, self.sample_rate, self.non_speech_label,
            self.energy_threshold, self.z_lo, self.z_hi, cue_start, cue_end, cue_keep, cue_off, self.ratios,
            self.start_seconds, self.max_offset_samples, out["best_score"].data_ptr(),
            out["best_offset"].data_ptr(), out["best_k"].data_ptr(), a_s, a_o,
            memspace=_native.B2_DEVICE_RESIDENT if inputs_resident else _native.B2_DEVICE)
        assert K == len(self.ratios)
        return out

    def sync_signals(self, ref, ref_off, cue_start, cue_end, cue_off, cue_keep=None):
        """Same as sync_device but starting from reference speech SIGNALS (100 Hz float32, all pairs
        back to back; numpy array or CUDA tensor) instead of PCM - the replay path of the
        reference's test-case bundles: ``ref.npz{"speech"}`` + ``in.srt``
        (ffsubsync/ffsubsync.py:338-343,639-644; DeserializeSpeechTransformer).
        Returns (best_score f64[B], best_offset i32[B], best_k i32[B]) as numpy arrays."""
        import torch
        h = self.handle
        ref_off = np.ascontiguousarray(ref_off, dtype=np.int64)
        B, K = len(ref_off) - 1, len(self.ratios)
        if isinstance(ref, np.ndarray):
            ref = torch.from_numpy(np.ascontiguousarray(ref, dtype=np.float32)).cuda()
        dev = ref.device
        lengths = h.rasterize_lengths(cue_end, cue_off, self.ratios, K, False, self.sample_rate)
        sub_off = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
        sub = torch.empty(int(sub_off[-1]), dtype=torch.float32, device=dev)
        h.rasterize(cue_start, cue_end, cue_keep, cue_off, self.ratios, K, False, self.sample_rate,
                    self.start_seconds, out=sub.data_ptr(), out_off=sub_off, memspace=_native.B2_DEVICE)
        score = torch.empty(B * K, dtype=torch.float64, device=dev)
        offset = torch.empty(B * K, dtype=torch.int32, device=dev)
        status = torch.empty(B * K, dtype=torch.int32, device=dev)
        h.align_batch(ref.data_ptr(), ref_off, sub.data_ptr(), sub_off, B, K, self.max_offset_samples,
                      score=score.data_ptr(), offset=offset.data_ptr(), status=status.data_ptr(),
                      memspace=_native.B2_DEVICE)
        bs = torch.empty(B, dtype=torch.float64, device=dev)
        bo = torch.empty(B, dtype=torch.int32, device=dev)
        bk = torch.empty(B, dtype=torch.int32, device=dev)
        h.reduce_ratios(score.data_ptr(), offset.data_ptr(), status.data_ptr(), B, K, self.max_offset_samples,
                        best_score=bs.data_ptr(), best_offset=bo.data_ptr(), best_k=bk.data_ptr(),
                        memspace=_native.B2_DEVICE)
        h.synchronize()
        return bs.cpu().numpy(), bo.cpu().numpy(), bk.cpu().numpy()

    def sync_device_candidate_sharded(self, pcm, pcm_off, cue_start, cue_end, cue_off, cue_keep=None,
                                      rank: int = 0, world: int = 1, group=None):
        """Secondary multi-GPU mode ("B < G": a few pairs on many GPUs): the K
        ratio candidates of every pair are dealt round-robin over the ranks.  Every rank runs the
        VAD of the (replicated) reference PCM - 33 us per 2 h signal, cheaper than shipping the
        signal - rasterises and aligns only ITS candidates (b2_align_batch with K_local ratios),
        the per-candidate (score, offset, status) triples are all-gathered back into list order
        (NCCL, 24 B per candidate) and b2_reduce_ratios applies the |offset| filter and the
        first-in-list tie rule on every rank, so all ranks hold the same (score, offset, ratio index)
        as a single-GPU run.  pcm: int16 CUDA tensor, same on every rank.  Returns CUDA tensors
        (best_score f64[B], best_offset i32[B], best_k i32[B])."""
        import torch
        from . import distributed
        h = self.handle
        self.use_torch_stream()   # torch ops and NCCL below are ordered against our kernels by the stream
        pcm_off = np.ascontiguousarray(pcm_off, dtype=np.int64)
        B, K = len(pcm_off) - 1, len(self.ratios)
        dev = pcm.device
        mine = distributed.shard_candidates(K, rank, world)
        fpw = h.frames_per_window(self.frame_rate, self.sample_rate)
        ref_off = np.concatenate([[0], np.cumsum((np.diff(pcm_off) + fpw - 1) // fpw)]).astype(np.int64)
        ref = torch.empty(int(ref_off[-1]), dtype=torch.float32, device=dev)
        h.vad_energy_zcr(pcm.data_ptr(), pcm_off, self.frame_rate, self.sample_rate, self.non_speech_label,
                         self.energy_threshold, self.z_lo, self.z_hi, out=ref.data_ptr(), memspace=_native.B2_DEVICE)
        k_loc = len(mine)
        local = torch.zeros((B, max(k_loc, 1), 3), dtype=torch.float64, device=dev)
        if k_loc:
            my_ratios = self.ratios[mine]
            lengths = h.rasterize_lengths(cue_end, cue_off, my_ratios, k_loc, False, self.sample_rate)
            sub_off = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
            sub = torch.empty(int(sub_off[-1]), dtype=torch.float32, device=dev)
            h.rasterize(cue_start, cue_end, cue_keep, cue_off, my_ratios, k_loc, False, self.sample_rate,
                        self.start_seconds, out=sub.data_ptr(), out_off=sub_off, memspace=_native.B2_DEVICE)
            score = torch.empty(B * k_loc, dtype=torch.float64, device=dev)
            offset = torch.empty(B * k_loc, dtype=torch.int32, device=dev)
            status = torch.empty(B * k_loc, dtype=torch.int32, device=dev)
            h.align_batch(ref.data_ptr(), ref_off, sub.data_ptr(), sub_off, B, k_loc, self.max_offset_samples,
                          score=score.data_ptr(), offset=offset.data_ptr(), status=status.data_ptr(),
                          memspace=_native.B2_DEVICE)
            local[:, :, 0] = score.view(B, k_loc)
            local[:, :, 1] = offset.view(B, k_loc).to(torch.float64)
            local[:, :, 2] = status.view(B, k_loc).to(torch.float64)
        full = distributed.allgather_candidate_results(local[:, :k_loc], K, rank, world, group=group)
        score_all = full[:, :, 0].contiguous().view(-1)
        offset_all = full[:, :, 1].to(torch.int32).contiguous().view(-1)
        status_all = full[:, :, 2].to(torch.int32).contiguous().view(-1)
        bs = torch.empty(B, dtype=torch.float64, device=dev)
        bo = torch.empty(B, dtype=torch.int32, device=dev)
        bk = torch.empty(B, dtype=torch.int32, device=dev)
        h.reduce_ratios(score_all.data_ptr(), offset_all.data_ptr(), status_all.data_ptr(), B, K,
                        self.max_offset_samples, best_score=bs.data_ptr(), best_offset=bo.data_ptr(),
                        best_k=bk.data_ptr(), memspace=_native.B2_DEVICE)
        return bs, bo, bk

    def sync_host(self, pcm, pcm_off, cue_start, cue_end, cue_off, cue_keep=None, want_all=False):
        """pcm: int16 numpy array (ideally backed by pinned memory).  Blocks until results are on
        the host.  Returns (best_score, best_offset, best_k[, all_score, all_offset])."""
        res = self.handle.sync_batch(
            pcm, pcm_off, self.frame_rate, self.sample_rate, self.non_speech_label, self.energy_threshold,
            self.z_lo, self.z_hi, cue_start, cue_end, cue_keep, cue_off, self.ratios, self.start_seconds,
            self.max_offset_samples, want_all=want_all, memspace=_native.B2_HOST)
        return res if want_all else res[:3]

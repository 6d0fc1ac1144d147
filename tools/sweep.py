#!/usr/bin/env python
"""BASELINE config 5: signal-length x batch sweep on one GPU (run under gpurun).

    python tools/sweep.py > gpurun_out/sweep_r1.md

For each duration T and batch B: whole-path alignments/s (PCM resident, K = 5, +-60 s window) and
the VAD kernel's GB/s of algorithmic bytes against the measured HBM peak.  Every cell is checked
(recovered offsets == ground truth).  CUDA events on the launching stream, 3 warm-up + 5 timed.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ffsubsync_b200 import _native  # noqa: E402
from ffsubsync_b200.batch import BatchSynchronizer  # noqa: E402
from ffsubsync_b200.synth import BENCH_RATIOS, make_pairs  # noqa: E402

FPW, FR = 160, 16000


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return json.load(open(p))["hbm_gbs"] if os.path.exists(p) else 6650.0


def cell(minutes, B, bs, stream):
    dev = torch.device("cuda", 0)
    dur = minutes * 60.0
    pairs = make_pairs([1000 * minutes + b for b in range(B)], dur, BENCH_RATIOS, handle=bs.handle)
    n_win = int(pairs.win_off[-1])
    cls_d = torch.from_numpy(pairs.window_class).to(dev)
    pcm = torch.empty(n_win * FPW, dtype=torch.int16, device=dev)
    bs.handle.synth_pcm(cls_d.data_ptr(), n_win, FPW, 99, out=pcm.data_ptr(), memspace=_native.B2_DEVICE)
    pcm_off = pairs.win_off * FPW
    out = None

    def step():
        return bs.sync_device(pcm, pcm_off, pairs.cue_start, pairs.cue_end, pairs.cue_off, out=out)

    for _ in range(3):
        out = step()
    torch.cuda.synchronize()
    ok = bool((out["best_offset"].cpu().numpy() == pairs.true_offset).all()
              and (out["best_k"].cpu().numpy() == pairs.true_k).all())
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(5):
        step()
    b.record(stream)
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    sig = torch.empty(n_win, dtype=torch.float32, device=dev)
    a.record(stream)
    for _ in range(5):
        bs.handle.vad_energy_zcr(pcm.data_ptr(), pcm_off, FR, 100, 0.0, 100000, out=sig.data_ptr(),
                                 memspace=_native.B2_DEVICE)
    b.record(stream)
    torch.cuda.synchronize()
    vad_ms = a.elapsed_time(b) / 5
    vad_gbs = (n_win * (2 * FPW + 4)) / (vad_ms * 1e-3) / 1e9
    return B / (ms * 1e-3), ms, vad_gbs, ok


def main():
    bs = BatchSynchronizer(BENCH_RATIOS, FR, 100, 0.0, max_offset_seconds=60, device=0)
    torch.cuda.set_device(0)
    bs.use_torch_stream()
    stream = torch.cuda.current_stream()
    pk = peak()
    print("| T | N (reference FFT size) | B | alignments/s | ms/step | VAD GB/s | VAD frac of %.0f GB/s | offsets ok |" % pk)
    print("|---|---|---|---|---|---|---|---|")
    budget_bytes = 60e9
    for minutes in (10, 30, 60, 120, 240):
        n_fft = 1 << int(np.ceil(np.log2(2 * minutes * 6000)))
        for B in (1, 8, 64, 512):
            if B * minutes * 60 * FR * 2 > budget_bytes:
                continue
            rate, ms, gbs, ok = cell(minutes, B, bs, stream)
            print("| %d min | 2^%d | %d | %.0f | %.3f | %.0f | %.2f | %s |"
                  % (minutes, int(np.log2(n_fft)), B, rate, ms, gbs, gbs / pk, ok), flush=True)
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Lane-per-window VAD kernel (csrc/vad_lane.cuh) against the lane-group kernel, and the SM-partitioned
overlap it is meant to enable: VAD on X SMs (one CTA per SM) next to the correlation kernels on the other
148 - X SMs, two streams.

    python tools/lane_probe.py [pairs]
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ffsubsync_b200 import _native  # noqa: E402
from ffsubsync_b200.synth import BENCH_RATIOS, make_pairs  # noqa: E402

FPW, FR = 160, 16000
KNOBS = ("B2_VAD_BATCH", "B2_VAD_WPL", "B2_VAD_LAYOUT", "B2_VAD_GRID", "B2_VAD_STAGES", "B2_VAD_CTAS_FORCE", "B2_CORR_MAX_CTAS", "B2_VAD_CONSUMERS")


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 74
    dev = torch.device("cuda", 0)
    h1, h2 = _native.Handle(0), _native.Handle(0)
    s1, s2 = torch.cuda.Stream(priority=-1), torch.cuda.Stream()
    h1.set_stream(s1.cuda_stream)
    h2.set_stream(s2.cuda_stream)
    pairs = make_pairs(list(range(50, 50 + B)), 7200.0, BENCH_RATIOS, handle=h1)
    n_win = int(pairs.win_off[-1])
    cls_d = torch.from_numpy(pairs.window_class).to(dev)
    pcm = torch.empty(n_win * FPW, dtype=torch.int16, device=dev)
    h1.synth_pcm(cls_d.data_ptr(), n_win, FPW, 5, out=pcm.data_ptr(), memspace=_native.B2_DEVICE)
    pcm_off = pairs.win_off * FPW
    ref = torch.empty(n_win, dtype=torch.float32, device=dev)
    ref2 = torch.empty(n_win, dtype=torch.float32, device=dev)
    K = len(BENCH_RATIOS)
    lengths = h1.rasterize_lengths(pairs.cue_end, pairs.cue_off, BENCH_RATIOS, K, False, 100)
    sub_off = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
    sub = torch.empty(int(sub_off[-1]), dtype=torch.float32, device=dev)
    sc = torch.empty(B * K, dtype=torch.float64, device=dev)
    of = torch.empty(B * K, dtype=torch.int32, device=dev)
    st = torch.empty(B * K, dtype=torch.int32, device=dev)
    gb = (pcm.numel() * 2 + n_win * 4) / 1e9

    def vad(dst=ref2):
        h1.vad_energy_zcr(pcm.data_ptr(), pcm_off, FR, 100, 0.0, 100000, out=dst.data_ptr(),
                          memspace=_native.B2_DEVICE)

    def align():
        h2.align_batch(ref.data_ptr(), pairs.win_off, sub.data_ptr(), sub_off, B, K, 6000, score=sc.data_ptr(),
                       offset=of.data_ptr(), status=st.data_ptr(), memspace=_native.B2_DEVICE)

    def clear():
        for k in KNOBS:
            os.environ.pop(k, None)

    clear()
    os.environ["B2_VAD_LAYOUT"] = "group"
    vad(ref)
    h1.synchronize()
    clear()
    vad(ref2)
    h1.synchronize()
    same = bool(torch.equal(ref, ref2))
    print("lane kernel == lane-group kernel on %d windows: %s (speech fraction %.3f)"
          % (n_win, same, float(ref.mean())), flush=True)
    h2.rasterize(pairs.cue_start, pairs.cue_end, None, pairs.cue_off, BENCH_RATIOS, K, False, 100, 0.0,
                 out=sub.data_ptr(), out_off=sub_off, memspace=_native.B2_DEVICE)
    for _ in range(2):
        align()
    torch.cuda.synchronize()

    def timed(fns, reps=4):
        for f in fns:
            f()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            for f in fns:
                f()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e3

    os.environ["B2_VAD_LAYOUT"] = "group"
    tg = timed([vad])
    clear()
    tv0, ta0 = timed([vad]), timed([align])
    print("pairs=%d: lane-group vad %.3f ms (%.0f GB/s); lane vad %.3f ms (%.0f GB/s); align alone %.3f ms; "
          "lane + align back to back %.3f ms" % (B, tg, gb / tg * 1e3, tv0, gb / tv0 * 1e3, ta0, tv0 + ta0), flush=True)
    print("--- lane VAD alone, one CTA per SM on X SMs")
    for wpl, batch in (("1", "2"), ("1", "3"), ("1", "4"), ("1", "5"), ("1", "6"), ("2", "1"), ("2", "2")):
        for X in (148, 90, 74):
            os.environ.update(B2_VAD_WPL=wpl, B2_VAD_BATCH=batch, B2_VAD_GRID=str(X))
            tv = timed([vad])
            print("wpl=%s batch=%s X=%3d: %.3f ms = %.0f GB/s = %.1f GB/s per SM"
                  % (wpl, batch, X, tv, gb / tv * 1e3, gb / tv * 1e3 / X), flush=True)
    clear()
    print("--- both together: lane VAD on X SMs, correlation on 148 - X CTAs")
    for X in (74, 86, 100):
        os.environ.update(B2_VAD_WPL="1", B2_VAD_BATCH=os.environ.get("LANE_PROBE_BATCH", "4"), B2_VAD_GRID=str(X), B2_CORR_MAX_CTAS=str(148 - X))
        tv = timed([vad])
        ta = timed([align])
        tb = timed([vad, align])
        print("X=%3d: vad alone %.3f ms, align alone on %d CTAs %.3f ms, both %.3f ms  (unpartitioned sum %.3f)"
              % (X, tv, 148 - X, ta, tb, tv0 + ta0), flush=True)
    clear()


if __name__ == "__main__":
    main()

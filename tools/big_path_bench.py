#!/usr/bin/env python
"""Timing of the wide-window (large-FFT) alignment path against the masked overlap-save path and
against the tiled path it replaces.

    python tools/big_path_bench.py [pairs] [minutes]
Signals resident on the device (b2_align_batch, float signals, K = 5); CUDA events.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ffsubsync_b200 import _native  # noqa: E402
from ffsubsync_b200.synth import BENCH_RATIOS, make_pairs  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    minutes = int(sys.argv[2]) if len(sys.argv) > 2 else 120
    h = _native.Handle(0)
    s = torch.cuda.Stream()
    torch.cuda.set_stream(s)
    h.set_stream(s.cuda_stream)
    K = len(BENCH_RATIOS)
    pairs = make_pairs([900 + b for b in range(B)], minutes * 60.0, BENCH_RATIOS, handle=h)
    n_win = int(pairs.win_off[-1])
    ref = torch.from_numpy((pairs.window_class == 1).astype(np.float32)).cuda()
    lengths = h.rasterize_lengths(pairs.cue_end, pairs.cue_off, BENCH_RATIOS, K, False, 100)
    sub_off = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
    sub = torch.empty(int(sub_off[-1]), dtype=torch.float32, device="cuda")
    h.rasterize(pairs.cue_start, pairs.cue_end, None, pairs.cue_off, BENCH_RATIOS, K, False, 100, 0.0,
                out=sub.data_ptr(), out_off=sub_off, memspace=_native.B2_DEVICE)
    sc = torch.empty(B * K, dtype=torch.float64, device="cuda")
    of = torch.empty(B * K, dtype=torch.int32, device="cuda")
    st = torch.empty(B * K, dtype=torch.int32, device="cuda")

    def run(mos, nb=B):
        h.align_batch(ref.data_ptr(), pairs.win_off[: nb + 1], sub.data_ptr(), sub_off[: nb * K + 1], nb, K, mos,
                      score=sc.data_ptr(), offset=of.data_ptr(), status=st.data_ptr(), memspace=_native.B2_DEVICE)

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s)
        for _ in range(reps):
            fn()
        b.record(s)
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    out = {"pairs": B, "minutes": minutes, "ratios": K}
    out["masked_60s_ms"] = timed(lambda: run(6000))
    run(6000)
    torch.cuda.synchronize()
    masked = of.cpu().numpy().copy()
    for path in os.environ.get("BIG_BENCH_PATHS", "big,tiled").split(","):
        os.environ["B2_ALIGN_PATH"] = path
        nb = B if path == "big" else min(B, 4)      # the tiled path is ~30x slower: time fewer pairs
        ms = timed(lambda: run(None, nb), reps=3 if path == "big" else 1)
        torch.cuda.synchronize()
        got = of.cpu().numpy()[: nb * K]
        best = (np.arange(nb) * K + pairs.true_k[:nb])
        out["unmasked_%s_ms_per_pair" % path] = ms / nb
        out["unmasked_%s_ok" % path] = bool((got[best] == pairs.true_offset[:nb]).all())
    os.environ.pop("B2_ALIGN_PATH")
    out["masked_60s_ms_per_pair"] = out["masked_60s_ms"] / B
    if "unmasked_big_ms_per_pair" in out:
        out["big_over_masked"] = out["unmasked_big_ms_per_pair"] / out["masked_60s_ms_per_pair"]
    if "unmasked_tiled_ms_per_pair" in out:
        out["tiled_over_masked"] = out["unmasked_tiled_ms_per_pair"] / out["masked_60s_ms_per_pair"]
    for mb in os.environ.get("BIG_BENCH_WS", "256,1024,4096").split(","):     # group size (workspace budget) sensitivity
        if not mb:
            continue
        os.environ["B2_BIG_WS_MB"] = mb
        out["unmasked_big_ws%s_ms_per_pair" % mb] = timed(lambda: run(None)) / B
    os.environ.pop("B2_BIG_WS_MB", None)
    print(json.dumps(out))


if __name__ == "__main__":
    main()

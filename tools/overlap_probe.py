#!/usr/bin/env python
"""Does the VAD kernel (HBM-bound) overlap with the correlation kernels (FP32/smem-bound) when they
are launched on two streams?  Times each alone and both together (CUDA events).

    python tools/overlap_probe.py [pairs]
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


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 74
    dev = torch.device("cuda", 0)
    h1, h2 = _native.Handle(0), _native.Handle(0)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    h1.set_stream(s1.cuda_stream)
    h2.set_stream(s2.cuda_stream)
    pairs = make_pairs(list(range(50, 50 + B)), 7200.0, BENCH_RATIOS, handle=h1)
    n_win = int(pairs.win_off[-1])
    cls_d = torch.from_numpy(pairs.window_class).to(dev)
    pcm = torch.empty(n_win * FPW, dtype=torch.int16, device=dev)
    h1.synth_pcm(cls_d.data_ptr(), n_win, FPW, 5, out=pcm.data_ptr(), memspace=_native.B2_DEVICE)
    pcm_off = pairs.win_off * FPW
    ref = torch.empty(n_win, dtype=torch.float32, device=dev)
    K = len(BENCH_RATIOS)
    lengths = h1.rasterize_lengths(pairs.cue_end, pairs.cue_off, BENCH_RATIOS, K, False, 100)
    sub_off = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
    sub = torch.empty(int(sub_off[-1]), dtype=torch.float32, device=dev)
    sc = torch.empty(B * K, dtype=torch.float64, device=dev)
    of = torch.empty(B * K, dtype=torch.int32, device=dev)
    st = torch.empty(B * K, dtype=torch.int32, device=dev)

    def vad():
        h1.vad_energy_zcr(pcm.data_ptr(), pcm_off, FR, 100, 0.0, 100000, out=ref.data_ptr(),
                          memspace=_native.B2_DEVICE)

    def align():
        h2.align_batch(ref.data_ptr(), pairs.win_off, sub.data_ptr(), sub_off, B, K, 6000, score=sc.data_ptr(),
                       offset=of.data_ptr(), status=st.data_ptr(), memspace=_native.B2_DEVICE)

    vad()
    h1.synchronize()
    h2.rasterize(pairs.cue_start, pairs.cue_end, None, pairs.cue_off, BENCH_RATIOS, K, False, 100, 0.0,
                 out=sub.data_ptr(), out_off=sub_off, memspace=_native.B2_DEVICE)
    for _ in range(2):
        align()
    torch.cuda.synchronize()

    def timed(fns, reps=4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            for f in fns:
                f()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e3

    for stages, ctas in (("4", "2"), ("4", "1"), ("3", "1"), ("2", "1")):
        os.environ["B2_VAD_STAGES"] = stages
        os.environ["B2_VAD_CTAS_FORCE"] = ctas
        tv = timed([vad])
        ta = timed([align])
        tb = timed([vad, align])
        print("pairs=%d vad stages=%s ctas/SM=%s: vad alone %.3f ms, align alone %.3f ms, sum %.3f, both (2 streams) %.3f ms"
              % (B, stages, ctas, tv, ta, tv + ta, tb), flush=True)


if __name__ == "__main__":
    main()

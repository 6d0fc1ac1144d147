#!/usr/bin/env python
"""VAD kernel tuning sweep (stages x CTAs/SM) + a plain read-bandwidth yardstick (torch reduction
over the same buffer).  python tools/vad_tune.py [pairs]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ffsubsync_b200 import _native  # noqa: E402

FR = int(os.environ.get("VAD_TUNE_FR", "16000"))   # e.g. 48000: ffsubsync's default frame rate
FPW = FR // 100


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    dev = torch.device("cuda", 0)
    h = _native.Handle(0)
    s = torch.cuda.Stream()
    torch.cuda.set_stream(s)
    h.set_stream(s.cuda_stream)
    n_win = B * 720000
    cls = torch.from_numpy(np.random.RandomState(0).randint(0, 3, n_win).astype(np.uint8)).to(dev)
    pcm = torch.empty(n_win * FPW, dtype=torch.int16, device=dev)
    h.synth_pcm(cls.data_ptr(), n_win, FPW, 5, out=pcm.data_ptr(), memspace=_native.B2_DEVICE)
    pcm_off = np.arange(B + 1, dtype=np.int64) * 720000 * FPW
    out = torch.empty(n_win, dtype=torch.float32, device=dev)
    gb = pcm.numel() * 2 / 1e9

    def timed(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s)
        for _ in range(reps):
            fn()
        b.record(s)
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    as_f32 = pcm.view(torch.float32)
    ms = timed(lambda: torch.max(as_f32))
    print("torch.max over the PCM bytes (read-only yardstick): %.3f ms = %.0f GB/s" % (ms, gb / ms * 1e3))
    as_i32 = pcm.view(torch.int32)
    ms = timed(lambda: torch.bitwise_xor(as_i32[: as_i32.numel() // 2], as_i32[as_i32.numel() // 2:]).max())
    print("torch xor+max (read + half write): %.3f ms" % ms)
    for wpt in ("1", "2", "4"):
        for stages in ("4", "3", "2"):
            for ctas in ("1", "2", "3", "4"):
                smem = int(stages) * (int(wpt) * 64 * FPW * 2 + 128) + 512
                if int(ctas) * smem > 227 * 1024 or (int(ctas) > 1 and (int(ctas) - 1) * smem > 227 * 1024):
                    continue
                os.environ["B2_VAD_WPT"] = wpt
                os.environ["B2_VAD_STAGES"] = stages
                os.environ["B2_VAD_CTAS_FORCE"] = ctas
                ms = timed(lambda: h.vad_energy_zcr(pcm.data_ptr(), pcm_off, FR, 100, 0.0, 100000,
                                                    out=out.data_ptr(), memspace=_native.B2_DEVICE))
                print("vad wpt=%s stages=%s ctas/SM=%s (%d KB/CTA): %.3f ms = %.0f GB/s"
                      % (wpt, stages, ctas, smem // 1024, ms, (gb + n_win * 4 / 1e9) / ms * 1e3), flush=True)


if __name__ == "__main__":
    main()

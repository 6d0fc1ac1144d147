#!/usr/bin/env python
"""Size ladder for the lane-per-window VAD kernel: lane vs lane-group output, one line per case."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ffsubsync_b200 import _native  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    h = _native.Handle(0)
    cases = [(16000, 1, 64), (16000, 1, 65), (16000, 1, 640), (16000, 1, 5000), (16000, 3, 20000),
             (16000, 2, 200000), (16000, 8, 720000), (48000, 1, 5000), (48000, 2, 100000), (8000, 2, 100000),
             (32000, 2, 100000), (16000, 40, 720000)]
    if len(sys.argv) > 1:
        cases = cases[int(sys.argv[1]):]
    for fr, B, nw in cases:
        fpw = fr // 100
        n_win = B * nw
        cls = torch.from_numpy(np.random.RandomState(nw).randint(0, 3, n_win).astype(np.uint8)).to(dev)
        pcm = torch.empty(n_win * fpw, dtype=torch.int16, device=dev)
        h.synth_pcm(cls.data_ptr(), n_win, fpw, 5, out=pcm.data_ptr(), memspace=_native.B2_DEVICE)
        pcm_off = np.arange(B + 1, dtype=np.int64) * nw * fpw
        a = torch.empty(n_win, dtype=torch.float32, device=dev)
        b = torch.empty(n_win, dtype=torch.float32, device=dev)
        os.environ["B2_VAD_LAYOUT"] = "group"
        h.vad_energy_zcr(pcm.data_ptr(), pcm_off, fr, 100, 0.0, 100000, out=a.data_ptr(), memspace=_native.B2_DEVICE)
        h.synchronize()
        os.environ.pop("B2_VAD_LAYOUT")
        print("fr=%d B=%d windows=%d: group ok, lane ..." % (fr, B, nw), end="", flush=True)
        h.vad_energy_zcr(pcm.data_ptr(), pcm_off, fr, 100, 0.0, 100000, out=b.data_ptr(), memspace=_native.B2_DEVICE)
        h.synchronize()
        msg = " equal=%s" % bool(torch.equal(a, b))
        if n_win * fpw <= 120_000_000:   # independent yardstick: plain torch ops on the same PCM
            x = pcm.view(n_win, fpw).to(torch.int64)
            e = (x * x).sum(1)
            neg = pcm.view(n_win, fpw) < 0
            z = (neg[:, 1:] != neg[:, :-1]).sum(1)
            want = ((e >= fpw * 100000) & (z >= 0) & (z <= (3 * fpw) // 8)).to(torch.float32)
            want = torch.where(want > 0, torch.ones_like(want), torch.zeros_like(want))
            msg += " group==torch %s lane==torch %s" % (bool(torch.equal(a, want)), bool(torch.equal(b, want)))
            del x, e, neg, z, want
        print(msg, flush=True)
        del cls, pcm, a, b


if __name__ == "__main__":
    main()

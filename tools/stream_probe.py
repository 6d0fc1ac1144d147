#!/usr/bin/env python
"""Host chunk loop (VideoSpeechTransformer.fit over 2 h of raw PCM in host memory): streaming
detector (b2_vad_stream_*, no per-chunk synchronisation) vs one synchronous detector call per chunk.

    python tools/stream_probe.py [hours]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ffsubsync_b200 import speech_transformers as st  # noqa: E402
from ffsubsync_b200 import _native  # noqa: E402


def main():
    hours = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
    n_win = int(hours * 360000)
    h = _native.get_handle()
    cls = np.random.RandomState(0).randint(0, 2, n_win).astype(np.uint8)
    pcm = h.synth_pcm(cls, n_win, 160, 7).tobytes()

    def run(streaming):
        def factory(sr, fr, label):
            det = st._make_energy_zcr_detector(sr, fr, label)
            return det if streaming else (lambda seg: det(seg))
        st.DETECTOR_FACTORIES["probe"] = factory
        t = st.VideoSpeechTransformer("probe", 100, 16000, 0.0)
        t.fit(pcm)
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            t.fit(pcm)
            best = min(best, time.perf_counter() - t0)
        return best, t.transform()

    ts, a = run(True)
    tn, b = run(False)
    assert np.array_equal(a, b)
    gb = len(pcm) / 1e9
    print("%.1f h PCM (%.1f MB, pageable host memory), 100 s chunks: streaming %.1f ms (%.1f GB/s), "
          "per-chunk synchronous %.1f ms (%.1f GB/s)" % (hours, gb * 1e3, ts * 1e3, gb / ts, tn * 1e3, gb / tn))


if __name__ == "__main__":
    main()

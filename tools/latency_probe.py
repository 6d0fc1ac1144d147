#!/usr/bin/env python
"""Single-pair host-to-host latency of b2_sync_batch (the actual ffsubsync use case: one video, one
subtitle file): 2 h of 16 kHz PCM in host memory -> (score, offset, ratio) on the host.

    python tools/latency_probe.py > gpurun_out/latency.json
Reports pinned and pageable input buffers, and the bare H2D copy of the same 230 MB for comparison.
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ffsubsync_b200 import _native, distributed  # noqa: E402
from ffsubsync_b200.batch import BatchSynchronizer  # noqa: E402
from ffsubsync_b200.synth import BENCH_RATIOS, make_pairs  # noqa: E402


def main():
    numa = distributed.bind_to_gpu_numa(0)
    bs = BatchSynchronizer(BENCH_RATIOS, 16000, 100, 0.0, max_offset_seconds=60, device=0)
    pairs = make_pairs([4242], 7200.0, BENCH_RATIOS, handle=bs.handle)
    cls_d = torch.from_numpy(pairs.window_class).cuda()
    pcm_d = torch.empty(len(pairs.window_class) * 160, dtype=torch.int16, device="cuda")
    bs.handle.synth_pcm(cls_d.data_ptr(), len(pairs.window_class), 160, 7, out=pcm_d.data_ptr(),
                        memspace=_native.B2_DEVICE)
    bs.handle.synchronize()
    pinned = torch.empty(pcm_d.numel(), dtype=torch.int16, pin_memory=True)
    pinned.copy_(pcm_d)
    torch.cuda.synchronize()
    pageable = pinned.numpy().copy()
    args = (pairs.win_off * 160, pairs.cue_start, pairs.cue_end, pairs.cue_off)

    def lat(buf, n=15):
        ts = []
        for _ in range(n + 3):
            t0 = time.perf_counter()
            res = bs.sync_host(buf, *args)
            ts.append((time.perf_counter() - t0) * 1e3)
            assert res[1][0] == pairs.true_offset[0] and res[2][0] == pairs.true_k[0]
        ts = sorted(ts[3:])
        return {"min_ms": ts[0], "median_ms": ts[len(ts) // 2], "max_ms": ts[-1]}

    def copy_only(src, n=15):
        dst = torch.empty_like(pcm_d)
        ts = []
        for _ in range(n + 3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
        ts = sorted(ts[3:])
        return {"min_ms": ts[0], "median_ms": ts[len(ts) // 2]}

    dev = bs.sync_device(pcm_d, *args)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    bs.use_torch_stream()
    a.record()
    for _ in range(20):
        bs.sync_device(pcm_d, *args)
    b.record()
    torch.cuda.synchronize()
    out = {"what": "one 2 h pair, K = 5, +-60 s: b2_sync_batch(B2_HOST) wall clock, results on the host",
           "pcm_bytes": int(pcm_d.numel() * 2), "numa": numa,
           "pinned_input": lat(pinned.numpy()), "pageable_input": lat(pageable),
           "h2d_copy_only_pinned": copy_only(pinned), "h2d_copy_only_pageable": copy_only(torch.from_numpy(pageable)),
           "device_resident_compute_ms": a.elapsed_time(b) / 20}
    out["pinned_gbs"] = out["pcm_bytes"] / out["pinned_input"]["median_ms"] / 1e6
    print(json.dumps(out))


if __name__ == "__main__":
    main()

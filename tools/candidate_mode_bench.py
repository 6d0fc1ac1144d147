#!/usr/bin/env python
"""Secondary multi-GPU mode (B < G): the K ratio candidates of a few pairs dealt over the ranks,
NCCL all-gather of the per-candidate results, local reduce (BatchSynchronizer.sync_device_candidate_sharded).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/candidate_mode_bench.py --pairs 1 --ratios 7

Prints one JSON line on rank 0: latency per call (device time, max over ranks) for the sharded mode
and for the same pairs on ONE GPU (rank 0, all K candidates), and whether all ranks agree with the
single-GPU result.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ffsubsync_b200 import _native, distributed  # noqa: E402
from ffsubsync_b200.batch import BatchSynchronizer  # noqa: E402
from ffsubsync_b200.constants import FRAMERATE_RATIOS  # noqa: E402
from ffsubsync_b200.synth import make_pairs  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=1)
    ap.add_argument("--ratios", type=int, default=7)
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    rank, world, local_rank = distributed.init_from_env("nccl")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    r = np.array(FRAMERATE_RATIOS)
    ratios = ([1.0] + list(np.concatenate([r, 1.0 / r])))[: args.ratios]
    bs = BatchSynchronizer(ratios, 16000, 100, 0.0, max_offset_seconds=60, device=local_rank)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    bs.use_torch_stream()
    B = args.pairs
    pairs = make_pairs([77 + b for b in range(B)], 7200.0, ratios, handle=bs.handle)   # same on every rank
    n_win = int(pairs.win_off[-1])
    cls_d = torch.from_numpy(pairs.window_class).to(dev)
    pcm = torch.empty(n_win * 160, dtype=torch.int16, device=dev)
    bs.handle.synth_pcm(cls_d.data_ptr(), n_win, 160, 4321, out=pcm.data_ptr(), memspace=_native.B2_DEVICE)
    pcm_off = pairs.win_off * 160
    call = (pcm, pcm_off, pairs.cue_start, pairs.cue_end, pairs.cue_off)

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        stream.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(args.steps):
            out = fn()
        b.record(stream)
        torch.cuda.synchronize()
        return distributed.max_over_ranks(a.elapsed_time(b) / args.steps, dev), out

    ms_sharded, got = timed(lambda: bs.sync_device_candidate_sharded(*call, rank=rank, world=world))
    single = bs.sync_device(*call)
    ms_single, single = timed(lambda: bs.sync_device(*call))
    same = bool(torch.equal(got[1], single["best_offset"]) and torch.equal(got[2], single["best_k"])
                and torch.equal(got[0], single["best_score"]))
    planted = bool((got[1].cpu().numpy() == pairs.true_offset).all() and (got[2].cpu().numpy() == pairs.true_k).all())
    agree = distributed.max_over_ranks(0.0 if (same and planted) else 1.0, dev) == 0.0
    if rank == 0:
        print(json.dumps({"mode": "candidates sharded over ranks (B < G)", "n_gpus": world, "pairs": B, "ratios": len(ratios),
                          "ms_per_call_sharded": ms_sharded, "ms_per_call_one_gpu": ms_single,
                          "all_ranks_equal_single_gpu_and_planted": agree,
                          "note": "every rank: VAD of the replicated PCM + its K/G candidates; one NCCL all-gather "
                                  "of 24 B per candidate; local reduce"}))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()

"""CPU suite: the oracle (oracle/*.py) against fixtures produced by the real reference
(tests/golden/make_golden.py).  This is what "pins" the oracle."""
import math
from datetime import timedelta

import numpy as np
import pytest

import cases
from oracle import aligner_oracle as ao
from oracle import gss_oracle as go
from oracle import raster_oracle as ro
from oracle import vad_oracle as vo


def _close(a, b, rel=1e-9, abs_=1e-7):
    if math.isinf(a) or math.isinf(b):
        return a == b
    return abs(a - b) <= max(abs_, rel * abs(b))


def test_alignment_kats(golden, gf):
    for c in golden["kats"]:
        score, off = ao.fft_align(c["ref"], c["sub"], c["mos"])
        assert off == c["offset"], c
        assert _close(score, gf(c["score"])), c


def test_alignment_small_cases_match_reference(golden, gf):
    assert len(golden["small"]) == 240
    for c in golden["small"]:
        ref, sub, mos = cases.small_align_case(c["seed"])
        assert mos == c["mos"]
        score, off = ao.fft_align(ref, sub, mos)
        assert off == c["offset"], c
        assert _close(score, gf(c["score"])), c


def test_exact_form_agrees_with_fft_form(golden, gf):
    """score(o) = sum_j s'[j] r'[j+o]: the closed form reproduces the reference's score at the
    reference's offset, and exact_align picks the same offset whenever the maximum is unique."""
    agree = 0
    for c in golden["small"]:
        ref, sub, mos = cases.small_align_case(c["seed"])
        want = gf(c["score"])
        if math.isinf(want):
            s, o = ao.exact_align(ref, sub, mos)
            assert s == want and o == c["offset"]
            continue
        assert _close(ao.exact_score(ref, sub, c["offset"]), want, abs_=1e-9)
        s, o = ao.exact_align(ref, sub, mos)
        o_lo, o_hi = ao.offset_range(len(ref), len(sub), mos)
        scores = ao.exact_scores_window(ref, sub, o_lo, o_hi)
        unique = np.sum(scores >= scores.max() - 1e-9) == 1
        if unique:
            assert o == c["offset"], c
            agree += 1
        assert _close(s, want, abs_=1e-9)
    assert agree > 150


@pytest.mark.parametrize("n", [6000, 60000])
def test_shifted_pairs(golden, gf, n):
    for c in golden["shifted"]:
        if c["n"] != n:
            continue
        ref, sub = cases.shifted_pair(n)
        score, off = ao.fft_align(ref, sub, c["mos"])
        assert off == c["offset"] == -1234
        assert _close(score, gf(c["score"]))
        assert _close(ao.exact_score(ref, sub, off), n - 1234, abs_=0)


def test_wide_window_cases(golden, gf):
    """Oracle vs the reference on the wide-window fixtures (unmasked / very wide masks; lopsided
    lengths, float levels, R + S at and just above a power of two) that pin the large-FFT path."""
    for c in golden["wide"]:
        kw = dict(c["case"])
        kw.pop("mos_list")
        ref, sub = cases.wide_pair(**kw)
        score, off = ao.fft_align(ref, sub, c["mos"])
        assert off == c["offset"], c
        assert _close(score, gf(c["score"]))
        assert _close(ao.exact_score(ref, sub, off), gf(c["score"]), rel=1e-9)


def test_empty_inputs_raise(golden):
    for c in golden["empty"]:
        assert c["raises"] and "empty speech data" in c["raises"]
        with pytest.raises(ao.OracleAlignmentFailure, match="empty speech data"):
            ao.fft_align(np.array(c["ref"]), np.array(c["sub"]))


def test_mask_range_matches_survey_probe():
    # max_offset_samples=5 leaves offsets -4..5
    assert ao.offset_range(50, 40, 5) == (-4, 5)
    assert ao.offset_range(50, 40, None) == (-40, 127 - 40)


def test_multi_segment(golden, golden_arrays, gf):
    grid = cases.ratio_grid()
    for ci, c in enumerate(golden["multi_segment"]):
        n = int(golden_arrays["ms_sparse_len_%d" % ci])
        sparse = np.unpackbits(golden_arrays["ms_sparse_%d" % ci])[:n].astype(float)
        _, sub = cases.multi_segment_case(c["scale"], c["shift"])
        subs = [cases.scaled_signal(sub, sf) for sf in grid]
        for sf, s, want in zip(grid, subs, c["per_ratio"]):
            score, off = ao.fft_align(sparse, s, 6000)
            assert off == want["offset"] and _close(score, gf(want["score"]))
        (score, off), k = ao.max_score_align(sparse, subs, 100, 60)
        assert k == c["best"]["index"] and off == c["best"]["offset"]
        # the reference test's own acceptance criterion (tests/test_multi_segment.py:166-167)
        assert grid[k] == pytest.approx(c["scale"], abs=1e-3)
        assert off / 100.0 == pytest.approx(c["shift"], abs=0.05)


def test_maxscore_grid(golden, gf):
    grid = cases.ratio_grid()
    for c in golden["maxscore"]:
        rng = np.random.RandomState(300 + c["seed"])
        n = 30000
        true_k = int(rng.randint(0, len(grid)))
        shift = int(rng.randint(-3000, 3001))
        base = (rng.rand(n) > 0.55).astype(float)
        ref = np.roll(cases.scaled_signal(base, grid[true_k]), shift)
        ref = np.where(rng.rand(len(ref)) < 0.10, 1.0 - ref, ref)
        subs = [cases.scaled_signal(base, r) * min(1.0 / r, 1.0) for r in grid]
        assert (true_k, shift) == (c["true_k"], c["shift"])
        for s, want in zip(subs, c["per_ratio"]):
            score, off = ao.fft_align(ref, s, 6000)
            assert off == want["offset"] and _close(score, gf(want["score"]))
        (score, off), k = ao.max_score_align(ref, subs, 100, 60)
        assert (k, off) == (c["best"]["index"], c["best"]["offset"])
    assert "max-offset-seconds" in golden["maxscore_fail"]
    with pytest.raises(ao.OracleAlignmentFailure):
        ao.max_score_select([(1.0, 300)], 1)


def test_first_ratio_wins_ties():
    assert ao.max_score_select([(5.0, 1), (5.0, 2), (4.0, 0)], None) == 0
    assert ao.max_score_select([(5.0, 100), (5.0, 2), (7.0, 50)], 10) == 1


# ------------------------------------------------------------------ rasteriser

def test_scale_roundtrip_closed_form(golden):
    for c in golden["scale_roundtrip"]:
        x = c["t"] * c["r"]
        assert ro.scale_seconds(c["t"], c["r"]) == c["scaled"]
        assert ro.seconds_via_timedelta_closed_form(x) == c["scaled"]
    rng = np.random.RandomState(11)
    for x in np.concatenate([rng.uniform(0, 20000, 3000), rng.randint(0, 10**7, 500) / 1000.0 + 0.0005,
                             [0.0, 0.9999995, 1.0000005, 2.5e-7, 1.5e-6]]):
        assert ro.seconds_via_timedelta_closed_form(float(x)) == timedelta(seconds=float(x)).total_seconds()


def test_raster_matches_reference(golden, gf):
    for c in golden["raster"]:
        starts, ends = cases.synthetic_cues(c["seed"], c["duration"])
        x, max_time, sf, ef = ro.rasterize(starts, ends, None, 100, c["start_seconds"], c["ratio"])
        assert len(x) == c["length"]
        levels, rs, re_ = cases.run_lengths(x)
        assert levels == c["levels"] and rs == c["run_starts"] and re_ == c["run_stops"], c["ratio"]
        assert max_time == gf(c["max_time"])
        assert (sf, ef) == (c["start_frame"], c["end_frame"])


def test_raster_kat_fake_srt(golden, gf):
    k = golden["raster_kat"]
    for c in k["cases"]:
        st = [k["starts"][i] for i in c["cue_idx"]]
        en = [k["ends"][i] for i in c["cue_idx"]]
        keep = [not ro.is_metadata(k["contents"][i], j == 0 or j + 1 == len(c["cue_idx"]))
                for j, i in enumerate(c["cue_idx"])]
        x, max_time, sf, ef = ro.rasterize(st, en, keep, c["sample_rate"], c["start_seconds"], 1.0,
                                           scale=False)
        assert len(x) == c["length"]
        assert cases.run_lengths(x) == (c["levels"], c["run_starts"], c["run_stops"])
        assert max_time == gf(c["max_time"])
        assert (sf, ef) == (c["start_frame"], c["end_frame"])
    # tests/test_subtitles.py:118-123
    full = [c for c in k["cases"] if c["sample_rate"] == 100 and c["start_seconds"] == 0][0]
    assert gf(full["max_time"]) == 6.062


def test_metadata_filter(golden):
    for c in golden["metadata"]:
        assert ro.is_metadata(c["content"], c["edge"]) == c["is_metadata"], c


def test_boundaries(golden):
    for c in golden["boundaries"]:
        assert ro.frame_boundaries(np.array(c["x"], dtype=float)) == (c["start"], c["end"])


# ------------------------------------------------------------------ golden-section search

def test_gss_trace(golden):
    want = golden["gss_quadratic"]
    interval, calls = go.golden_section_trace(lambda x, last: (x - 1.0417) ** 2, 0.9, 1.1)
    assert len(calls) == 17 == len(want["calls"])
    for (x, last), (wx, wl) in zip(calls, want["calls"]):
        assert x == wx and last == wl
    assert list(interval) == want["interval"]
    assert abs(calls[0][0] - 0.976393) < 1e-6 and abs(calls[1][0] - 1.023607) < 1e-6


def test_gss_fit(golden, gf):
    want = golden["gss_fit"]
    ref_full, sub = cases.multi_segment_case(25.0 / 24.0, 3.0)
    record = {}

    def f(ratio, last):
        score, off = ao.fft_align(ref_full, cases.scaled_signal(sub, ratio), 6000)
        if last:
            record.update(score=score, offset=off, ratio=ratio)
        return -score

    _, calls = go.golden_section_trace(f, 0.9, 1.1)
    assert [c[0] for c in calls] == want["evals"]
    assert record["offset"] == want["offset"] and record["ratio"] == want["ratio"]
    assert _close(record["score"], gf(want["score"]))


# ------------------------------------------------------------------ VAD contract (shape/labels)

def test_vad_contract_and_synthetic_classes():
    fpw = vo.frames_per_window(16000, 100)
    assert fpw == 160 and vo.frames_per_window(48000, 100) == 480 and vo.frames_per_window(44100, 100) == 441
    cls = np.array([0, 1, 2, 1, 1, 0, 2, 0, 1], dtype=np.uint8)
    pcm = vo.synth_pcm(cls, fpw, seed=3)
    out = vo.energy_zcr_detect(pcm.tobytes(), 100, 16000, 0.0)
    assert out.dtype == np.float64 and out.tolist() == [float(c == 1) for c in cls]
    # trailing partial window -> one extra, non-speech value; label respected
    out2 = vo.energy_zcr_detect(np.frombuffer(pcm.tobytes() + pcm.tobytes()[:100], np.uint8), 100, 16000, 0.5)
    assert len(out2) == len(cls) + 1 and out2[-1] == 0.5 and set(out2) == {0.5, 1.0}
    # chunking invariance at window-aligned boundaries (speech_transformers.py:710-746)
    a = vo.energy_zcr_detect(pcm[: 4 * fpw].tobytes(), 100, 16000, 0.0)
    b = vo.energy_zcr_detect(pcm[4 * fpw:].tobytes(), 100, 16000, 0.0)
    assert np.array_equal(np.concatenate([a, b]), out)
    e, z = vo.window_features(pcm, fpw)
    assert z[1] == 3 and e[1] > fpw * 10**5 and e[0] < fpw * 10**5 and z[2] > 60


# ===================================================================== auditok detector (V3) pins
# The auditok wheel is absent: the restatement in oracle/auditok_oracle.py is pinned to the
# known-answer examples of auditok 0.1.5's StreamTokenizer documentation / test-suite (quoted from
# memory, see the oracle header) and to the arithmetic identity of its energy rule.

def _upper(s):
    return [c.isupper() for c in s]


@pytest.mark.parametrize("text, min_len, max_len, max_sil, want", [
    ("aaaAAAABBbbb", 1, 9999, 0, [(3, 8)]),
    ("aaaAAAABBbbb", 3, 4, 0, [(3, 6), (7, 8)]),                 # short token kept: contiguous with a truncated one
    ("aaaAAAaaaBBbbbb", 3, 6, 3, [(3, 8), (9, 13)]),            # trailing silence is part of the token
    ("aAaaaAaAaaAaAaaaaaaaAAAAAAAA", 5, 20, 4, [(1, 16), (20, 27)]),
])
def test_auditok_tokenizer_published_examples(text, min_len, max_len, max_sil, want):
    from oracle import auditok_oracle as au
    assert au.tokenize(_upper(text), min_len, max_len, max_sil) == want


def test_auditok_energy_rule_is_an_integer_threshold():
    """10*log10(dot(x,x)/n) >= 50  <=>  sum x^2 >= n * 10^5 for int16 blocks, exact-edge blocks included."""
    from oracle import auditok_oracle as au
    rng = np.random.RandomState(3)
    for n in (1, 7, 80, 159, 160, 441, 480):
        assert au.energy_floor(n, 50) == n * 10 ** 5
        for _ in range(60):   # rms around the 50 dB edge (sqrt(1e5) = 316.2)
            x = np.round(rng.randn(n) * rng.uniform(250, 400)).astype(np.int16)
            e = int((x.astype(np.int64) ** 2).sum())
            assert au.block_is_valid(x, 50) == (e >= n * 10 ** 5)
    # exact edge: 100 samples of +-400 and 60 zeros give 16 000 000 = 160 * 10^5; one LSB less is invalid
    x = np.zeros(160, np.int16)
    x[:100] = 400
    assert au.block_is_valid(x, 50) and au.block_log_energy(x) == 50.0
    x[0] = 399
    assert not au.block_is_valid(x, 50)
    assert not au.block_is_valid(np.zeros(160, np.int16), 50)   # log energy -200


def test_auditok_detect_wrapper_semantics():
    """speech_transformers.py:143-150: length formula, impulse ASSIGNMENT and the cumsum/clip."""
    from oracle import auditok_oracle as au
    fpw = 160
    loud = (np.tile([400, -400], fpw // 2)).astype(np.int16)       # E = 160 * 160000 >= 160 * 1e5
    quiet = np.zeros(fpw, np.int16)

    def build(flags, tail=0):
        pcm = np.concatenate([loud if f else quiet for f in flags] + [loud[:tail]])
        return pcm.tobytes()

    # 30 valid blocks, 40 silent: token = 30 + 25 tolerated silent frames
    out = au.auditok_detect(build([1] * 30 + [0] * 40), 100, 16000, 0.0)
    assert len(out) == 70 and out[:55].tolist() == [1.0] * 55 and out[55:].tolist() == [0.0] * 15
    # a 10-frame blip is shorter than min_length = 20 (even with its tolerated silence it is judged on
    # len(data) = 10 + 25): 35 >= 20 -> delivered; a blip with the stream ending right after is not
    out = au.auditok_detect(build([0] * 5 + [1] * 10 + [0] * 50), 100, 16000, 0.0)
    assert out[5:40].tolist() == [1.0] * 35 and out[40:].sum() == 0
    out = au.auditok_detect(build([0] * 5 + [1] * 10), 100, 16000, 0.0)
    assert out.sum() == 0   # post-processing: 10 < min_length and not contiguous
    # 600 valid blocks: truncated at max_length = 500; the next token starts where the end impulse of
    # the first one was written and OVERWRITES it, so the level stays at 1 after the second token ends
    out = au.auditok_detect(build([1] * 600 + [0] * 100), 100, 16000, 0.0)
    assert len(out) == 700 and out.tolist() == [1.0] * 700
    out = au.auditok_detect(build([1] * 600 + [0] * 100), 100, 16000, 0.25)   # cumsum 1, 2, 1.25 -> clip
    assert out[:625].tolist() == [1.0] * 625 and out[625:].tolist() == [1.0] * 75
    # partial last block: judged on the samples it has; output length = ceil(n / fpw)
    out = au.auditok_detect(build([1] * 40, tail=7), 100, 16000, 0.0)
    assert len(out) == 41 and out.tolist() == [1.0] * 41
    with pytest.raises(ValueError):
        au.auditok_detect(b"\x00\x01\x02", 100, 16000, 0.0)


def test_auditok_fast_path_equals_literal_restatement():
    from oracle import auditok_oracle as au
    rng = np.random.RandomState(8)
    for frame_rate, nsl in ((16000, 0.0), (8000, 0.3), (44100, 0.0)):
        fpw = frame_rate // 100
        runs, valid = [], True
        while sum(runs) < 4000:
            runs.append(int(rng.choice([2, 8, 19, 20, 24, 25, 26, 60, 300, 520])))
        amp = np.concatenate([np.full(r, 1.15 if (i % 2 == 0) else 0.85) for i, r in enumerate(runs)])
        amp = amp * rng.uniform(0.93, 1.07, len(amp))
        pcm = np.round(rng.randn(len(amp) * fpw) * 316.2 * np.repeat(amp, fpw)).astype(np.int16)[: len(amp) * fpw - 13]
        a = au.auditok_detect(pcm.tobytes(), 100, frame_rate, nsl)
        b = au.auditok_detect_fast(pcm.tobytes(), 100, frame_rate, nsl)
        assert np.array_equal(a, b) and 0 < a.sum() < len(a)

// Bit-exact device restatement of the reference's cue arithmetic, shared by the stand-alone
// rasteriser (raster.cu) and the correlation kernel that rasterises subtitle blocks on the fly
// (corr.cu).  Reference: SubtitleScaler.fit (ffsubsync/subtitle_transformers.py:35-47) +
// SubtitleSpeechTransformer.fit (ffsubsync/speech_transformers.py:957-980).
#pragma once
#include <cuda_runtime.h>

// timedelta(seconds=t*ratio).total_seconds(): whole seconds exact, fractional part * 1e6 rounded
// half-to-even to integer microseconds, then one correctly rounded division (no FMA contraction).
__device__ __forceinline__ double b2_scaled_seconds(double t, double ratio) {
  const double x = __dmul_rn(t, ratio);
  double whole;
  const double frac = modf(x, &whole);
  const long long us = (long long)whole * 1000000LL + __double2ll_rn(__dmul_rn(frac, 1e6));
  return __ddiv_rn((double)us, 1e6);
}

// samples[first:last] of a length-n array with Python slice semantics (negative index wraps once).
__device__ __forceinline__ void b2_cue_bounds(double start_s, double end_s, double ratio,
                                              double start_seconds, int sample_rate, long long n,
                                              long long& first, long long& last) {
  const double st = b2_scaled_seconds(start_s, ratio);
  const double en = b2_scaled_seconds(end_s, ratio);
  first = __double2ll_rn(__dmul_rn(__dsub_rn(st, start_seconds), (double)sample_rate));
  last = first + __double2ll_rn(__dmul_rn(__dsub_rn(en, st), (double)sample_rate));
  if (first < 0) { first += n; if (first < 0) first = 0; } else if (first > n) first = n;
  if (last < 0) { last += n; if (last < 0) last = 0; } else if (last > n) last = n;
}

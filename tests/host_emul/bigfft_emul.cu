// CPU emulation of the large-window correlation kernels (ffsubsync_b200/csrc/bigfft.cu): the same
// __host__ __device__ phase functions (bigfft.cuh) run for tid = 0..511, one loop per barrier-separated
// phase.  Test infrastructure; tests/test_host_cpu.py drives it against np.fft.
//
// usage: bigfft_emul in.bin out.bin
// in.bin : int32 R, S, q1, mode ; float ref[R] ; float sub[S]     (mode 1: sub as a bit mask, level = max)
// out.bin: float scores[N]  (scores[m] ~ sum_j sub'[j] ref'[j + m - S], N = 2^(q1 + 11)) ; float Es, Er, ||c||^2
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../../ffsubsync_b200/csrc/bigfft.cuh"

using namespace bigfft;

template <int Q1>
static void cols_all(float2* buf, const Tables& t, bool inv) {
  if (!inv) {
    for (int tid = 0; tid < kThreads; ++tid) col_step<Q1, 0, false>(buf, t, tid);
    for (int tid = 0; tid < kThreads; ++tid) col_step<Q1, 1, false>(buf, t, tid);
    for (int tid = 0; tid < kThreads; ++tid) col_step<Q1, 2, false>(buf, t, tid);
  } else {
    for (int tid = 0; tid < kThreads; ++tid) col_step<Q1, 2, true>(buf, t, tid);
    for (int tid = 0; tid < kThreads; ++tid) col_step<Q1, 1, true>(buf, t, tid);
    for (int tid = 0; tid < kThreads; ++tid) col_step<Q1, 0, true>(buf, t, tid);
  }
}
static void cols_dispatch(int q1, float2* buf, const Tables& t, bool inv) {
  switch (q1) {
    case 6: cols_all<6>(buf, t, inv); break;
    case 7: cols_all<7>(buf, t, inv); break;
    case 8: cols_all<8>(buf, t, inv); break;
    case 9: cols_all<9>(buf, t, inv); break;
    case 10: cols_all<10>(buf, t, inv); break;
    case 11: cols_all<11>(buf, t, inv); break;
    case 12: cols_all<12>(buf, t, inv); break;
    default: exit(7);
  }
}
static void rows_all(float2* buf, const Tables& t, bool inv) {
  if (!inv) {
    for (int tid = 0; tid < kThreads; ++tid) row_step<0, false>(buf, t, tid);
    for (int tid = 0; tid < kThreads; ++tid) row_step<1, false>(buf, t, tid);
    for (int tid = 0; tid < kThreads; ++tid) row_step<2, false>(buf, t, tid);
  } else {
    for (int tid = 0; tid < kThreads; ++tid) row_step<2, true>(buf, t, tid);
    for (int tid = 0; tid < kThreads; ++tid) row_step<1, true>(buf, t, tid);
    for (int tid = 0; tid < kThreads; ++tid) row_step<0, true>(buf, t, tid);
  }
}

int main(int argc, char** argv) {
  if (argc != 3) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 3;
  int hdr[4];
  if (fread(hdr, 4, 4, f) != 4) return 4;
  const int R = hdr[0], S = hdr[1], q1 = hdr[2], mode = hdr[3];
  std::vector<float> ref(R), sub(S);
  if (fread(ref.data(), 4, R, f) != (size_t)R) return 4;
  if (fread(sub.data(), 4, S, f) != (size_t)S) return 4;
  fclose(f);
  const int M = 1 << (q1 + 10), N = 2 * M, tiles = (1 << q1) / 16;
  if (R + S > N) return 5;

  std::vector<float2> buf(kM), tw(1024), fine32(32), half(1024), coarse(kSkew1024), finem(skew(1 << q1) + 4), row_tw(16);
  for (int tid = 0; tid < kThreads; ++tid) init_tables(tw.data(), fine32.data(), tid);
  for (int tid = 0; tid < kThreads; ++tid) init_big_tables(half.data(), coarse.data(), finem.data(), q1, tid);
  const Tables t{tw.data(), fine32.data()};
  const BigTables bt{tw.data(), fine32.data(), half.data(), coarse.data(), finem.data()};

  float level = 0.f;
  for (float v : sub) level = v > level ? v : level;
  std::vector<uint32_t> words((S + 63) / 32 + 1, 0u);
  for (int i = 0; i < S; ++i)
    if (sub[i] != 0.f) words[i >> 5] |= 1u << (i & 31);
  const BigSource src_ref{ref.data(), nullptr, R, 0.f};
  const BigSource src_sub = mode == 1 ? BigSource{nullptr, words.data(), S, 2.f * level - 1.f}
                                      : BigSource{sub.data(), nullptr, S, 0.f};

  std::vector<float2> g_ref(M), g_sub(M), spec(M);
  float es = 0.f, er = 0.f;
  // reference: F1, F2 (store spectrum)
  for (int cg = 0; cg < tiles; ++cg) {
    for (int tid = 0; tid < kThreads; ++tid) er += f1_load(buf.data(), src_ref, q1, cg, tid);
    cols_dispatch(q1, buf.data(), t, false);
    for (int tid = 0; tid < kThreads; ++tid) f1_store(buf.data(), bt, q1, cg, tid, g_ref.data());
  }
  for (int g = 0; g < tiles; ++g) {
    for (int tid = 0; tid < kThreads; ++tid) f2_load(buf.data(), q1, g, tid, g_ref.data());
    for (int tid = 0; tid < kThreads; ++tid) f2_row_twiddles(row_tw.data(), q1, g, tid);
    rows_all(buf.data(), t, false);
    for (int tid = 0; tid < kThreads; ++tid) f2_untangle_inplace(buf.data(), bt, q1, g, row_tw.data(), tid);
    for (int tid = 0; tid < kThreads; ++tid) f2_store(buf.data(), q1, g, tid, spec.data());
  }
  // subtitles: F1, F2 (product + inverse rows), F3
  for (int cg = 0; cg < tiles; ++cg) {
    for (int tid = 0; tid < kThreads; ++tid) es += f1_load(buf.data(), src_sub, q1, cg, tid);
    cols_dispatch(q1, buf.data(), t, false);
    for (int tid = 0; tid < kThreads; ++tid) f1_store(buf.data(), bt, q1, cg, tid, g_sub.data());
  }
  for (int g = 0; g < tiles; ++g) {
    for (int tid = 0; tid < kThreads; ++tid) f2_load(buf.data(), q1, g, tid, g_sub.data());
    for (int tid = 0; tid < kThreads; ++tid) f2_row_twiddles(row_tw.data(), q1, g, tid);
    rows_all(buf.data(), t, false);
    for (int tid = 0; tid < kThreads; ++tid)
      f2_product_inplace(buf.data(), bt, q1, g, row_tw.data(), tid, spec.data());
    rows_all(buf.data(), t, true);
    for (int tid = 0; tid < kThreads; ++tid) f2_store_twiddled(buf.data(), bt, q1, g, tid, g_sub.data());
  }
  std::vector<float> out(N + 3);
  float mx = -INFINITY, cn = 0.f;
  for (int cg = 0; cg < tiles; ++cg) {
    for (int tid = 0; tid < kThreads; ++tid) f3_load(buf.data(), q1, cg, tid, g_sub.data());
    cols_dispatch(q1, buf.data(), t, true);
    for (int tid = 0; tid < kThreads; ++tid) f3_store(buf.data(), q1, cg, tid, out.data(), S, 0, N - 1, mx, cn);
  }
  out[N] = es;
  out[N + 1] = er;
  out[N + 2] = cn;
  f = fopen(argv[2], "wb");
  fwrite(out.data(), 4, N + 3, f);
  fclose(f);
  return 0;
}

// CPU emulation of the two correlation kernels of ffsubsync_b200/csrc/corr.cu: the very same
// __host__ __device__ phase functions (corr.cuh) are run for tid = 0..511 with a loop standing
// in for each __syncthreads-separated phase.  Test infrastructure (the build container has no
// GPU); tests/test_host_cpu.py drives it.
//
// usage: corr_emul in.bin out.bin
// in.bin : int32 R, S, o_t, W, L, mode ; float ref[R] ; float sub[S]
//          mode 0: float subtitle signal (sub_correlate_kernel);  mode 1: two-level subtitle signal
//          {0, level} fed as a bit mask (sub_correlate_bits_kernel; L must be a multiple of 32)
// out.bin: float c[W]  (c[m] ~ sum_j sub'[j] ref'[j + o_t + m]) ; float Es, Er, ||c||_2^2 of the whole
//          inverse-transform output (what sub_correlate_body stores for tau), blocks accumulated
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../../ffsubsync_b200/csrc/corr.cuh"

using namespace corr;

static void rest_all(float2* buf, const Tables& t) {
  for (int tid = 0; tid < kThreads; ++tid) dif16_pass_smem<6>(buf, t, tid);
  for (int tid = 0; tid < kThreads; ++tid) dif16_pass_smem<2>(buf, t, tid);
  for (int tid = 0; tid < kThreads; ++tid) r4_pass_smem<false>(buf, tid);
}

int main(int argc, char** argv) {
  if (argc != 3) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 3;
  int hdr[6];
  if (fread(hdr, 4, 6, f) != 6) return 4;
  const int R = hdr[0], S = hdr[1], o_t = hdr[2], W = hdr[3], L = hdr[4], mode = hdr[5];
  std::vector<float> ref(R), sub(S);
  if (fread(ref.data(), 4, R, f) != (size_t)R) return 4;
  if (fread(sub.data(), 4, S, f) != (size_t)S) return 4;
  fclose(f);
  if (mode == 1 && (L & 31)) return 5;

  std::vector<float2> buf(kM), tw(1024), fine(32);
  for (int tid = 0; tid < kThreads; ++tid) init_tables(tw.data(), fine.data(), tid);
  Tables t{tw.data(), fine.data()};
  std::vector<SubState> st(kThreads);
  for (auto& s : st) sub_state_clear(s);
  std::vector<float4> spec(kPairs);
  std::vector<float> ss_ref(kThreads, 0.f), ss_sub(kThreads, 0.f);
  float level = 0.f;
  for (float v : sub) level = v > level ? v : level;
  std::vector<uint32_t> words(kP / 32);

  const int nblk = (S + L - 1) / L;
  int n_acc = 0;
  for (int blk = 0; blk < nblk; ++blk) {
    const int j0 = blk * L, i0 = j0 + o_t;
    if (i0 >= R || i0 + kP <= 0) continue;  // block pruning, as the host planner does
    ++n_acc;
    // reference block (ref_spectra_kernel)
    const int r_lo = i0 < 0 ? -i0 : 0, r_hi = (R - i0) < kP ? (R - i0) : kP;
    for (int tid = 0; tid < kThreads; ++tid)
      ss_ref[tid] += float_pass1(buf.data(), t, tid, ref.data() + i0, r_lo, r_hi);
    rest_all(buf.data(), t);
    for (int tid = 0; tid < kThreads; ++tid) spec_store(buf.data(), t, pair_ctx(t, tid), tid, spec.data());
    // subtitle block
    const int t_hi = (S - j0) < L ? (S - j0) : L;
    if (mode == 1) {
      for (auto& w : words) w = 0xdeadbeefu;  // stale words beyond the block must not matter
      for (int w = 0; w < L / 32; ++w) {
        uint32_t bits = 0;
        for (int b = 0; b < 32; ++b) {
          const int tt = 32 * w + b;
          if (tt < t_hi && sub[j0 + tt] != 0.f) bits |= 1u << b;
        }
        words[w] = bits;
      }
      for (int tid = 0; tid < kThreads; ++tid)
        ss_sub[tid] += bits_pass1(buf.data(), t, tid, words.data(), t_hi, L, 2.f * level - 1.f);
    } else {
      for (int tid = 0; tid < kThreads; ++tid)
        ss_sub[tid] += float_pass1(buf.data(), t, tid, sub.data() + j0, 0, t_hi);
    }
    rest_all(buf.data(), t);
    for (int tid = 0; tid < kThreads; ++tid) sub_accumulate(st[tid], buf.data(), t, pair_ctx(t, tid), tid, spec.data());
  }
  for (int tid = 0; tid < kThreads; ++tid) sub_retangle_store(st[tid], buf.data(), t, pair_ctx(t, tid), tid);
  for (int tid = 0; tid < kThreads; ++tid) inverse_passes_1(buf.data(), tid);
  for (int tid = 0; tid < kThreads; ++tid) inverse_passes_2(buf.data(), t, tid);
  for (int tid = 0; tid < kThreads; ++tid) inverse_passes_3(buf.data(), t, tid);
  for (int tid = 0; tid < kThreads; ++tid) inverse_passes_4(buf.data(), t, tid);
  std::vector<float> out(W + 4);
  for (int m = 0; m < W; ++m) out[m] = window_value(buf.data(), m);
  float es = 0.f, er = 0.f, cn = 0.f;
  for (int tid = 0; tid < kThreads; ++tid) { es += ss_sub[tid]; er += ss_ref[tid]; }
  for (int i = 0; i < kM; ++i) cn += buf[i].x * buf[i].x + buf[i].y * buf[i].y;
  out[W] = es;
  out[W + 1] = er;
  out[W + 2] = cn * kOutScale * kOutScale;
  out[W + 3] = (float)n_acc;
  f = fopen(argv[2], "wb");
  fwrite(out.data(), 4, W + 4, f);
  fclose(f);
  return 0;
}

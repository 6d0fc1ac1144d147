// Large-window correlation: FFTAligner.fit with max_offset_samples=None (the reference's default,
// ffsubsync/aligners.py:25,67-80) or a mask wider than a few overlap-save tiles.
//
// Instead of tiling the N candidate offsets in 16 385-wide windows (corr.cu; every tile re-transforms
// every block) each signal gets ONE real FFT of the padded length N = 2^k >= R + S, computed as a
// four-step complex FFT of M = N/2 = M1 x 1024 points (bigfft.cuh): columns (F1), rows + untangle
// [+ product + inverse rows] (F2), inverse columns (F3).  The reference spectrum of a pair is computed
// once and reused by its K ratio candidates.  The N fp32 scores only nominate candidates (same
// worst-case round-off bound tau, same exact float64 re-score and argmax as the windowed path);
// selection over the N-sized score arrays is parallel: per-tile maxima from F3, a counting pass over
// 32 768-offset chunks and an ordered compaction of the chunks that hold candidates.
//
// HBM traffic per (pair, ratio) at N = 2^21: F1 3 + 8 MB, F2 16 + 8 MB, F3 8 + 8 MB, selection 8 MB
// (+ 27 MB / K for the reference) against ~105 MFLOP per transform: memory and FP32 work are balanced,
// groups of pairs are sized so that a group's work arrays stay L2-resident between the steps.
#include <math.h>

#include <algorithm>
#include <map>

#include "common.cuh"
#include "bigfft.cuh"
#include "corr_jobs.cuh"

int bigfft_min_log2n() { return bigfft::kMinQ1 + 11; }
int bigfft_max_log2n() { return bigfft::kMaxQ1 + 11; }

namespace {

using namespace bigfft;

struct BigXform {        // one transform = one real signal padded to N = 2 M
  long long src_off;     // element offset of the float signal, or word offset of the bit mask
  long long g_off;       // float2 offset of its M-point work array
  long long spec_off;    // reference: where its spectrum is stored (in place: == g_off);
                         // subtitles: the spectrum to multiply with
  long long score_off;   // subtitles: float offset of the N scores
  int len, S, m_lo, m_hi;
  int is_bits;
  float hi;
};

struct BigJob {          // one (pair, ratio) of the group
  int j;                 // global job index b * K + k
  int x_sub, x_ref;      // transform indices inside the group
};

constexpr int kChunk = 4096;    // offsets per counting CTA
constexpr int kFineCap = (1 << kMaxQ1) + (1 << (kMaxQ1 - 4)) + (1 << (kMaxQ1 - 8)) + 4;   // skewed size
constexpr size_t kBigSmemBytes = kSmemBytes + 1024 * 8 + (size_t)kSkew1024 * 8 + (size_t)kFineCap * 8 + 16 * 8 + 64;

struct Smem {
  float2 *buf, *tw1024, *fine32, *half_pos, *coarse, *fine, *row_tw;
};
__device__ __forceinline__ Smem carve(unsigned char* raw) {
  Smem s;
  s.buf = reinterpret_cast<float2*>(raw);
  s.tw1024 = s.buf + kM;
  s.fine32 = s.tw1024 + 1024;
  s.half_pos = s.fine32 + 32;
  s.coarse = s.half_pos + 1024;
  s.fine = s.coarse + kSkew1024;
  s.row_tw = s.fine + kFineCap;
  return s;
}

// L2 prefetch of what the NEXT tile of this CTA will load (one 512-thread CTA per SM cannot hide HBM
// latency with occupancy; the tile's own loads then hit L2).
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, uint32_t bytes) {   // 16-byte aligned, multiple of 16
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void l2_prefetch_line(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
// rows of an F2 tile (16 x 8 KB)
__device__ __forceinline__ void prefetch_f2_rows(const float2* base, int q1, int g, int tid) {
  if (tid < 16) l2_prefetch_bulk(base + ((size_t)f2_row_of_slot(q1, g, tid) << 10), kRow * 8);
}
// column group cg of a k1-major [M1][1024] array: M1 segments of cols * 8 bytes
__device__ __forceinline__ void prefetch_cols(const float2* base, int q1, int cg, int tid) {
  const int cl = 14 - q1, c0 = cg << cl, seg = 8 << cl;   // bytes per row segment
  for (int r = tid; r < (1 << q1); r += kThreads) {
    const char* p = reinterpret_cast<const char*>(base + ((size_t)r << 10) + c0);
    for (int b = 0; b < seg; b += 128) l2_prefetch_line(p + b);
  }
}

__device__ __forceinline__ float block_sum(float v, float* red) {   // all threads call; result on every thread
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  for (int w = 0; w < kThreads / 32; ++w) r += red[w];   // fixed order: deterministic
  return r;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = -INFINITY;
  for (int w = 0; w < kThreads / 32; ++w) r = fmaxf(r, red[w]);
  return r;
}

// F1: forward column transforms + four-step twiddle.  Persistent over (transform, column group) tiles.
template <int Q1>
__global__ void __launch_bounds__(kThreads, 1)
    big_cols_forward_kernel(const float* __restrict__ sig, const uint32_t* __restrict__ bits,
                            const BigXform* __restrict__ xf, int n_tiles, float2* __restrict__ G,
                            float* __restrict__ tile_energy) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ float red[kThreads / 32];
  const Smem sm = carve(smem_raw);
  const int tid = threadIdx.x;
  init_tables(sm.tw1024, sm.fine32, tid);
  init_big_tables(sm.half_pos, sm.coarse, sm.fine, Q1, tid);
  __syncthreads();
  const Tables t{sm.tw1024, sm.fine32};
  const BigTables bt{sm.tw1024, sm.fine32, sm.half_pos, sm.coarse, sm.fine};
  constexpr int tiles_per = (1 << Q1) / 16;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const BigXform X = xf[tile / tiles_per];
    const int cg = tile % tiles_per;
    const BigSource src{X.is_bits ? nullptr : sig + X.src_off, X.is_bits ? bits + X.src_off : nullptr, X.len, X.hi};
    const float ss = block_sum(f1_load(sm.buf, src, Q1, cg, tid), red);   // barriers inside
    if (tid == 0) tile_energy[tile] = ss;
    col_forward<Q1>(sm.buf, t, tid);
    __syncthreads();
    f1_store(sm.buf, bt, Q1, cg, tid, G + X.g_off);
    __syncthreads();
  }
}

// F2: row transforms.  MODE 0 (reference): untangle, store the packed real spectrum in place.
// MODE 1 (subtitles): untangle, conj(A) * B with the pair's stored spectrum, retangle, inverse row
// transforms, conjugate four-step twiddle, store in place.
template <int MODE>
__global__ void __launch_bounds__(kThreads, 1)
    big_rows_kernel(const BigXform* __restrict__ xf, int n_tiles, int q1, float2* __restrict__ G) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const Smem sm = carve(smem_raw);
  const int tid = threadIdx.x;
  init_tables(sm.tw1024, sm.fine32, tid);
  init_big_tables(sm.half_pos, sm.coarse, sm.fine, q1, tid);
  __syncthreads();
  const Tables t{sm.tw1024, sm.fine32};
  const BigTables bt{sm.tw1024, sm.fine32, sm.half_pos, sm.coarse, sm.fine};
  const int tiles_per = (1 << q1) / 16;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const BigXform X = xf[tile / tiles_per];
    const int g = tile % tiles_per;
    if (tile + (int)gridDim.x < n_tiles) {
      const int nt = tile + gridDim.x;
      const BigXform Xn = xf[nt / tiles_per];
      prefetch_f2_rows(G + Xn.g_off, q1, nt % tiles_per, tid);
      if (MODE == 1) prefetch_f2_rows(G + Xn.spec_off, q1, nt % tiles_per, tid);
    }
    f2_load(sm.buf, q1, g, tid, G + X.g_off);
    f2_row_twiddles(sm.row_tw, q1, g, tid);
    __syncthreads();
    rows_forward(sm.buf, t, tid);
    __syncthreads();
    if (MODE == 0) {
      f2_untangle_inplace(sm.buf, bt, q1, g, sm.row_tw, tid);
      __syncthreads();
      f2_store(sm.buf, q1, g, tid, G + X.spec_off);
    } else {
      f2_product_inplace(sm.buf, bt, q1, g, sm.row_tw, tid, G + X.spec_off);
      __syncthreads();
      rows_inverse(sm.buf, t, tid);
      __syncthreads();
      f2_store_twiddled(sm.buf, bt, q1, g, tid, G + X.g_off);
    }
    __syncthreads();
  }
}

// F3: inverse column transforms; scores, their maximum over the surviving window, sum of squares.
template <int Q1>
__global__ void __launch_bounds__(kThreads, 1)
    big_cols_inverse_kernel(const BigXform* __restrict__ xf, int n_tiles, const float2* __restrict__ G,
                            float* __restrict__ scores, float* __restrict__ tile_max,
                            float* __restrict__ tile_cn) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ float red[kThreads / 32];
  const Smem sm = carve(smem_raw);
  const int tid = threadIdx.x;
  init_tables(sm.tw1024, sm.fine32, tid);
  __syncthreads();
  const Tables t{sm.tw1024, sm.fine32};
  constexpr int tiles_per = (1 << Q1) / 16;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const BigXform X = xf[tile / tiles_per];
    const int cg = tile % tiles_per;
    if (tile + (int)gridDim.x < n_tiles) {
      const int nt = tile + gridDim.x;
      prefetch_cols(G + xf[nt / tiles_per].g_off, Q1, nt % tiles_per, tid);
    }
    f3_load(sm.buf, Q1, cg, tid, G + X.g_off);
    __syncthreads();
    col_inverse<Q1>(sm.buf, t, tid);
    __syncthreads();
    float mx = -INFINITY, cn = 0.f;
    f3_store(sm.buf, Q1, cg, tid, scores + X.score_off, X.S, X.m_lo, X.m_hi, mx, cn);
    mx = block_max(mx, red);
    cn = block_sum(cn, red);
    if (tid == 0) {
      tile_max[tile] = mx;
      tile_cn[tile] = cn;
    }
    __syncthreads();
  }
}

// ---- selection over N-sized score arrays -----------------------------------------------------------
// per job: fp32 maximum over the surviving window and the round-off bound tau (corr_jobs.cuh)
__global__ void __launch_bounds__(256) big_stat_kernel(const BigJob* __restrict__ jobs, int tiles_per,
                                                        const float* __restrict__ ref_energy,
                                                        const float* __restrict__ sub_energy,
                                                        const float* __restrict__ tile_max,
                                                        const float* __restrict__ tile_cn,
                                                        float2* __restrict__ job_stat) {
  const BigJob jb = jobs[blockIdx.x];
  __shared__ float s_es[256], s_er[256], s_cn[256], s_mx[256];
  float es = 0.f, er = 0.f, cn = 0.f, mx = -INFINITY;
  for (int i = threadIdx.x; i < tiles_per; i += 256) {
    es += sub_energy[(size_t)jb.x_sub * tiles_per + i];
    er += ref_energy[(size_t)jb.x_ref * tiles_per + i];
    cn += tile_cn[(size_t)jb.x_sub * tiles_per + i];
    mx = fmaxf(mx, tile_max[(size_t)jb.x_sub * tiles_per + i]);
  }
  s_es[threadIdx.x] = es;
  s_er[threadIdx.x] = er;
  s_cn[threadIdx.x] = cn;
  s_mx[threadIdx.x] = mx;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w) {
      s_es[threadIdx.x] += s_es[threadIdx.x + w];
      s_er[threadIdx.x] += s_er[threadIdx.x + w];
      s_cn[threadIdx.x] += s_cn[threadIdx.x + w];
      s_mx[threadIdx.x] = fmaxf(s_mx[threadIdx.x], s_mx[threadIdx.x + w]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float tau = kU * (kTauFwd * sqrtf(s_es[0] * s_er[0]) + kTauInv * sqrtf(s_cn[0]));
    job_stat[jb.j] = make_float2(s_mx[0], tau * 1.0001f + 1e-30f);
  }
}

// jobs that never reach big_stat_kernel (empty input, everything masked) must not look like winners
// to their siblings' winner-only test
__global__ void big_init_stat_kernel(float2* __restrict__ job_stat, int n) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) job_stat[j] = make_float2(-INFINITY, 0.f);
}

// The cut of a job: fp32 maximum minus tau; with winner_only a ratio that by the bound tau cannot be the
// pair's best keeps only its fp32 argmax (same rule as select_candidates_kernel in corr.cu).
__device__ __forceinline__ float job_cut(const float2* __restrict__ job_stat, int j, int K, int winner_only,
                                         bool& approx_only) {
  const float2 stat = job_stat[j];
  float cut = stat.x - stat.y;
  approx_only = false;
  if (winner_only) {   // callers pass winner_only = 0 for a job with no_prune set
    const int b0 = (j / K) * K;
    float best_floor = -INFINITY;
    for (int k = 0; k < K; ++k) {
      const float2 s = job_stat[b0 + k];
      best_floor = fmaxf(best_floor, s.x - s.y);
    }
    if (stat.x + stat.y < best_floor) {
      approx_only = true;
      cut = stat.x;
    }
  }
  return cut;
}

__global__ void __launch_bounds__(256) big_count_kernel(const SelJob* __restrict__ sel,
                                                         const BigJob* __restrict__ jobs,
                                                         const float* __restrict__ scores,
                                                         const float2* __restrict__ job_stat, int K,
                                                         int winner_only, int n_chunks,
                                                         int* __restrict__ chunk_cnt) {
  const BigJob jb = jobs[blockIdx.y];
  const SelJob job = sel[jb.j];
  __shared__ int s_cnt[8];
  int cnt = 0;
  if (job.kind == 0 && job.m_lo <= job.m_hi) {
    bool approx;
    const float cut = job_cut(job_stat, jb.j, K, winner_only && !job.no_prune, approx);
    const float* c = scores + job.score_off;
    const int lo = max(job.m_lo, (int)blockIdx.x * kChunk), hi = min(job.m_hi, (int)(blockIdx.x + 1) * kChunk - 1);
    for (int m = lo + threadIdx.x; m <= hi; m += 256) cnt += c[m] >= cut;
  }
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
    for (int w = 0; w < 8; ++w) tot += s_cnt[w];
    chunk_cnt[(size_t)blockIdx.y * n_chunks + blockIdx.x] = tot;
  }
}

// Ordered compaction: candidates from the LARGEST offset down (np.argmax keeps the lowest index =
// largest offset among equal values), at most kCandMax; only chunks that hold candidates are walked.
__global__ void __launch_bounds__(256) big_select_kernel(const SelJob* __restrict__ sel,
                                                          const BigJob* __restrict__ jobs,
                                                          const float* __restrict__ scores,
                                                          const float2* __restrict__ job_stat, int K,
                                                          int winner_only, int n_chunks,
                                                          const int* __restrict__ chunk_cnt,
                                                          int* __restrict__ cand_off, int* __restrict__ cand_cnt,
                                                          int* __restrict__ work_list,
                                                          int* __restrict__ work_count) {
  const BigJob jb = jobs[blockIdx.x];
  const SelJob job = sel[jb.j];
  const int tid = threadIdx.x;
  __shared__ int scount;
  __shared__ int swarp[8];
  if (job.kind != 0 || job.m_lo > job.m_hi) {
    if (tid == 0) cand_cnt[jb.j] = 0;
    return;
  }
  bool approx_only;
  const float cut = job_cut(job_stat, jb.j, K, winner_only && !job.no_prune, approx_only);
  const float* c = scores + job.score_off;
  // 1. the (at most kCandMax) highest chunks that hold candidates, in descending order, and the total
  __shared__ int hit_chunk[kCandMax];
  __shared__ int n_hit, s_total;
  if (tid == 0) {
    scount = 0;
    n_hit = 0;
    s_total = 0;
  }
  __syncthreads();
  const int* cnt = chunk_cnt + (size_t)blockIdx.x * n_chunks;
  for (int top = n_chunks - 1; top >= 0; top -= 256) {
    const int ch = top - tid;
    const int here = ch >= 0 ? cnt[ch] : 0;
    const unsigned ball = __ballot_sync(0xffffffffu, here > 0);
    int wsum = here;
    for (int o = 16; o > 0; o >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
    if ((tid & 31) == 0) {
      swarp[tid >> 5] = __popc(ball);
      atomicAdd(&s_total, wsum);
    }
    __syncthreads();
    int before = n_hit;
    for (int w = 0; w < (tid >> 5); ++w) before += swarp[w];
    before += __popc(ball & ((1u << (tid & 31)) - 1u));
    if (here > 0 && before < kCandMax) hit_chunk[before] = ch;
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int w = 0; w < 8; ++w) tot += swarp[w];
      n_hit += tot;
    }
    __syncthreads();
  }
  const int total = s_total;
  const int n_walk = min(n_hit, kCandMax);
  // 2. ordered compaction inside those chunks
  for (int h = 0; h < n_walk; ++h) {
    if (scount >= kCandMax) break;   // uniform: scount is read after a barrier
    const int ch = hit_chunk[h];
    const int lo = max(job.m_lo, ch * kChunk), hi = min(job.m_hi, (ch + 1) * kChunk - 1);
    for (int top = hi; top >= lo; top -= 256) {
      const int m = top - tid;
      const bool hit = (m >= lo) && (c[m] >= cut);
      const unsigned ball = __ballot_sync(0xffffffffu, hit);
      if ((tid & 31) == 0) swarp[tid >> 5] = __popc(ball);
      __syncthreads();
      int before = scount;
      for (int w = 0; w < (tid >> 5); ++w) before += swarp[w];
      before += __popc(ball & ((1u << (tid & 31)) - 1u));
      if (hit && before < kCandMax) cand_off[(size_t)jb.j * kCandMax + before] = job.o_first + m;
      __syncthreads();
      if (tid == 0) {
        int tot = 0;
        for (int w = 0; w < 8; ++w) tot += swarp[w];
        scount += tot;
      }
      __syncthreads();
    }
  }
  if (approx_only) {  // slot 0 holds the largest offset attaining the fp32 maximum
    if (tid == 0) cand_cnt[jb.j] = -1;
    return;
  }
  if (tid == 0) {
    cand_cnt[jb.j] = total;
    scount = min(total, kCandMax);
    swarp[0] = atomicAdd(work_count, scount);
  }
  __syncthreads();
  if (tid < scount) work_list[swarp[0] + tid] = (jb.j << 5) | tid;
}

template <int Q1>
int launch_cols(b2_ctx* h, bool inverse, const float* d_sig, const uint32_t* d_bits, const BigXform* d_xf,
                int n_tiles, float2* G, float* scores, float* a0, float* a1) {
  const unsigned grid = (unsigned)std::min(n_tiles, h->sm_count);
  if (!inverse) {
    B2_CUDA(h, cudaFuncSetAttribute(big_cols_forward_kernel<Q1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)kBigSmemBytes));
    big_cols_forward_kernel<Q1><<<grid, kThreads, kBigSmemBytes, h->stream>>>(d_sig, d_bits, d_xf, n_tiles, G, a0);
    B2_CHECK_LAUNCH(h, "big_cols_forward_kernel");
  } else {
    B2_CUDA(h, cudaFuncSetAttribute(big_cols_inverse_kernel<Q1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)kBigSmemBytes));
    big_cols_inverse_kernel<Q1><<<grid, kThreads, kBigSmemBytes, h->stream>>>(d_xf, n_tiles, G, scores, a0, a1);
    B2_CHECK_LAUNCH(h, "big_cols_inverse_kernel");
  }
  return B2_OK;
}

int launch_cols_q(b2_ctx* h, int q1, bool inverse, const float* d_sig, const uint32_t* d_bits,
                  const BigXform* d_xf, int n_tiles, float2* G, float* scores, float* a0, float* a1) {
  switch (q1) {
    case 6: return launch_cols<6>(h, inverse, d_sig, d_bits, d_xf, n_tiles, G, scores, a0, a1);
    case 7: return launch_cols<7>(h, inverse, d_sig, d_bits, d_xf, n_tiles, G, scores, a0, a1);
    case 8: return launch_cols<8>(h, inverse, d_sig, d_bits, d_xf, n_tiles, G, scores, a0, a1);
    case 9: return launch_cols<9>(h, inverse, d_sig, d_bits, d_xf, n_tiles, G, scores, a0, a1);
    case 10: return launch_cols<10>(h, inverse, d_sig, d_bits, d_xf, n_tiles, G, scores, a0, a1);
    case 11: return launch_cols<11>(h, inverse, d_sig, d_bits, d_xf, n_tiles, G, scores, a0, a1);
    case 12: return launch_cols<12>(h, inverse, d_sig, d_bits, d_xf, n_tiles, G, scores, a0, a1);
    default: B2_FAIL(h, B2_ERR_UNSUPPORTED, "big align: unsupported transform size 2^%d", q1 + 11);
  }
}

}  // namespace

int b2i_align_big(b2_ctx* h, const float* d_ref, const float* d_sub, const uint32_t* d_bits, int B, int K,
                  std::vector<SelJob>& sel, const std::vector<long long>& idx_lo,
                  const std::vector<long long>& idx_hi, const std::vector<long long>& n_pad, int winner_only,
                  const B2CandBuffers& cb, const SelJob** d_sel_out) {
  B2Range range("b2:align_big (four-step FFT per signal)");
  const size_t J = (size_t)B * K;
  // transform size of a pair = the largest padded length among its live ratios (more zero padding
  // changes nothing: each job's offsets and surviving window come from ITS OWN padded length)
  std::map<int, std::vector<int>> pairs_by_q1;
  for (int b = 0; b < B; ++b) {
    long long n_max = 0;
    for (int k = 0; k < K; ++k) {
      const size_t j = (size_t)b * K + k;
      SelJob& s = sel[j];
      if (s.kind != 0) continue;
      n_max = std::max(n_max, n_pad[j]);
      s.o_first = -s.S;
      s.m_lo = (int)(n_pad[j] - idx_hi[j]);       // m = offset + S = N - 1 - idx
      s.m_hi = (int)(n_pad[j] - 1 - idx_lo[j]);
      s.n_tiles = 1;
      s.n_split = 1;
    }
    if (n_max == 0) continue;
    int lg = 0;
    while ((1LL << lg) < n_max) ++lg;
    pairs_by_q1[lg - 11].push_back(b);
  }
  // group = as many pairs as keep the work arrays within the workspace budget
  size_t budget = (size_t)4 << 30;
  if (const char* e = getenv("B2_BIG_WS_MB")) budget = (size_t)std::max(64, atoi(e)) << 20;
  struct Group { int q1; std::vector<int> pairs; };
  std::vector<Group> groups;
  for (auto& kv : pairs_by_q1) {
    const size_t m = (size_t)1 << (kv.first + 10);
    const size_t per_pair = m * 8 + (size_t)K * (m * 8 + m * 2 * 4);
    const size_t cap = std::max<size_t>(1, budget / per_pair);
    for (size_t i = 0; i < kv.second.size(); i += cap)
      groups.push_back({kv.first, std::vector<int>(kv.second.begin() + i,
                                                   kv.second.begin() + std::min(kv.second.size(), i + cap))});
  }
  // score_off is group-local (the score workspace is reused by the next group)
  size_t max_g = 0, max_scores = 0, max_tiles = 0, max_cnt = 0;
  for (auto& g : groups) {
    const size_t m = (size_t)1 << (g.q1 + 10), n = 2 * m;
    size_t n_sub = 0;
    for (int b : g.pairs)
      for (int k = 0; k < K; ++k) {
        SelJob& s = sel[(size_t)b * K + k];
        if (s.kind != 0) continue;
        s.score_off = (long long)(n_sub * n);
        ++n_sub;
      }
    const size_t tiles_per = ((size_t)1 << g.q1) / 16;
    max_g = std::max(max_g, (g.pairs.size() + n_sub) * m);
    max_scores = std::max(max_scores, n_sub * n);
    max_tiles = std::max(max_tiles, (g.pairs.size() + 3 * n_sub) * tiles_per);
    max_cnt = std::max(max_cnt, n_sub * ((n + kChunk - 1) / kChunk));
  }
  // The job table is read by kernels of EVERY group and by the common tail, i.e. long after later
  // metadata arenas have been committed - outside the reuse contract of the arena ring (a slot may be
  // recycled 8 arenas later).  It lives in its own workspace.
  void *d_selv, *h_selv;
  B2_TRY(b2i_ws(h, b2_ctx::WS_META, J * sizeof(SelJob) + 256, &d_selv));
  // staged through a pinned buffer of the handle (a copy from pageable memory would first wait for
  // the stream to drain); the event guards the buffer against the next call
  if (!h->pinned_ev[0]) B2_CUDA(h, cudaEventCreateWithFlags(&h->pinned_ev[0], cudaEventDisableTiming));
  B2_CUDA(h, cudaEventSynchronize(h->pinned_ev[0]));
  B2_TRY(b2i_pinned(h, 0, J * sizeof(SelJob) + 256, &h_selv));
  memcpy(h_selv, sel.data(), J * sizeof(SelJob));
  B2_CUDA(h, cudaMemcpyAsync(d_selv, h_selv, J * sizeof(SelJob), cudaMemcpyHostToDevice, h->stream));
  B2_CUDA(h, cudaEventRecord(h->pinned_ev[0], h->stream));
  const SelJob* d_sel = (const SelJob*)d_selv;
  *d_sel_out = d_sel;
  B2_CUDA(h, cudaMemsetAsync(cb.work_count, 0, sizeof(int), h->stream));
  B2_CUDA(h, cudaMemsetAsync(cb.cand_cnt, 0, J * sizeof(int), h->stream));   // jobs that are not live: no candidates
  big_init_stat_kernel<<<(unsigned)((J + 255) / 256), 256, 0, h->stream>>>(cb.job_stat, (int)J);
  B2_CHECK_LAUNCH(h, "big_init_stat_kernel");
  if (groups.empty()) return B2_OK;

  void *d_g, *d_s;
  B2_TRY(b2i_ws(h, b2_ctx::WS_SPEC, max_g * 8 + 256, &d_g));
  B2_TRY(b2i_ws(h, b2_ctx::WS_SCORES, max_scores * 4 + max_tiles * 4 + max_cnt * 4 + 1024, &d_s));
  float2* G = (float2*)d_g;
  float* scores = (float*)d_s;
  float* tile_arr = scores + max_scores;                 // ref_energy | sub_energy | tile_max | tile_cn
  int* chunk_cnt = (int*)(tile_arr + max_tiles);
  B2_CUDA(h, cudaFuncSetAttribute(big_rows_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)kBigSmemBytes));
  B2_CUDA(h, cudaFuncSetAttribute(big_rows_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)kBigSmemBytes));

  for (auto& g : groups) {
    const int q1 = g.q1;
    const size_t m = (size_t)1 << (q1 + 10), n = 2 * m;
    const int tiles_per = (1 << q1) / 16;
    std::vector<BigXform> xr, xs;
    std::vector<BigJob> jobs;
    for (size_t pi = 0; pi < g.pairs.size(); ++pi) {
      const int b = g.pairs[pi];
      BigXform r;
      memset(&r, 0, sizeof(r));
      bool have_ref = false;
      for (int k = 0; k < K; ++k) {
        const size_t j = (size_t)b * K + k;
        const SelJob& s = sel[j];
        if (s.kind != 0) continue;
        if (!have_ref) {
          r.src_off = s.ref_off;
          r.len = s.R;
          r.g_off = (long long)(pi * m);
          r.spec_off = r.g_off;
          have_ref = true;
        }
        BigXform x;
        memset(&x, 0, sizeof(x));
        x.is_bits = s.bits_off >= 0 ? 1 : 0;
        x.src_off = x.is_bits ? s.bits_off : s.sub_off;
        x.hi = 2.f * s.sub_level - 1.f;
        x.len = s.S;
        x.S = s.S;
        x.m_lo = s.m_lo;
        x.m_hi = s.m_hi;
        x.g_off = (long long)((g.pairs.size() + xs.size()) * m);
        x.spec_off = r.g_off;
        x.score_off = s.score_off;
        jobs.push_back({(int)j, (int)xs.size(), (int)pi});
        xs.push_back(x);
      }
      xr.push_back(r);   // a pair without live jobs keeps a zero-length dummy (never referenced)
    }
    const int n_ref = (int)xr.size(), n_sub = (int)xs.size();
    if (n_sub == 0) continue;
    MetaArena ga;
    B2_TRY(b2i_meta_begin(h, &ga, (xr.size() + xs.size()) * sizeof(BigXform) + jobs.size() * sizeof(BigJob) + 256));
    const BigXform* d_xr = (const BigXform*)b2i_meta_put(&ga, xr.data(), xr.size() * sizeof(BigXform));
    const BigXform* d_xs = (const BigXform*)b2i_meta_put(&ga, xs.data(), xs.size() * sizeof(BigXform));
    const BigJob* d_jobs = (const BigJob*)b2i_meta_put(&ga, jobs.data(), jobs.size() * sizeof(BigJob));
    B2_TRY(b2i_meta_commit(&ga));
    float* ref_energy = tile_arr;
    float* sub_energy = ref_energy + (size_t)n_ref * tiles_per;
    float* tile_max = sub_energy + (size_t)n_sub * tiles_per;
    float* tile_cn = tile_max + (size_t)n_sub * tiles_per;
    const int n_chunks = (int)((n + kChunk - 1) / kChunk);
    const unsigned rows_grid_r = (unsigned)std::min(n_ref * tiles_per, h->sm_count);
    const unsigned rows_grid_s = (unsigned)std::min(n_sub * tiles_per, h->sm_count);

    B2_TRY(launch_cols_q(h, q1, false, d_ref, nullptr, d_xr, n_ref * tiles_per, G, nullptr, ref_energy, nullptr));
    big_rows_kernel<0><<<rows_grid_r, kThreads, kBigSmemBytes, h->stream>>>(d_xr, n_ref * tiles_per, q1, G);
    B2_CHECK_LAUNCH(h, "big_rows_kernel<ref>");
    B2_TRY(launch_cols_q(h, q1, false, d_sub, d_bits, d_xs, n_sub * tiles_per, G, nullptr, sub_energy, nullptr));
    big_rows_kernel<1><<<rows_grid_s, kThreads, kBigSmemBytes, h->stream>>>(d_xs, n_sub * tiles_per, q1, G);
    B2_CHECK_LAUNCH(h, "big_rows_kernel<sub>");
    B2_TRY(launch_cols_q(h, q1, true, nullptr, nullptr, d_xs, n_sub * tiles_per, G, scores, tile_max, tile_cn));
    big_stat_kernel<<<n_sub, 256, 0, h->stream>>>(d_jobs, tiles_per, ref_energy, sub_energy, tile_max, tile_cn,
                                                  cb.job_stat);
    B2_CHECK_LAUNCH(h, "big_stat_kernel");
    big_count_kernel<<<dim3(n_chunks, n_sub), 256, 0, h->stream>>>(d_sel, d_jobs, scores, cb.job_stat, K,
                                                                   winner_only, n_chunks, chunk_cnt);
    B2_CHECK_LAUNCH(h, "big_count_kernel");
    big_select_kernel<<<n_sub, 256, 0, h->stream>>>(d_sel, d_jobs, scores, cb.job_stat, K, winner_only, n_chunks,
                                                    chunk_cnt, cb.cand_off, cb.cand_cnt, cb.work_list,
                                                    cb.work_count);
    B2_CHECK_LAUNCH(h, "big_select_kernel");
  }
  return B2_OK;
}

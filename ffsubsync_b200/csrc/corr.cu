// K3-K5: batched cross-correlation of reference / subtitle speech signals over a window of
// candidate offsets, + exact re-scoring and argmax.  Replaces FFTAligner.fit
// (ffsubsync/aligners.py:50-80, mask :31-43, argmax :45-48).
//
// What the reference computes, in closed form:
//     score(o) = sum_j s'[j] * r'[j + o],   x' = 2x - 1,  terms outside either signal = 0
// for the offsets o = N-1-idx-S that survive the max_offset mask, and returns the maximum
// (first index = largest offset among equals).
//
// How it is computed here.  Only a window [o_lo, o_hi] of offsets is wanted (12 000 of them for
// the default --max-offset-seconds 60), so the subtitle signal is cut into blocks of
// L = P - W + 1 samples; each block and the P reference samples it can meet are transformed with
// a P = 2^15 point real FFT that lives entirely in shared memory, conj(A)*B is accumulated over
// the blocks in registers, and ONE inverse transform yields the W scores (overlap-save
// correlation).  The reference-side spectra are computed once per pair and reused by all K
// ratio candidates.  The fp32 scores only nominate candidates: every offset within a first-order
// worst-case round-off bound (tau, below) of the maximum is re-scored exactly (float64 direct sum), so the returned
// offset and score do not depend on FFT round-off.  Offset ranges wider than P/2 are tiled.
#include <math.h>

#include <algorithm>

#include "common.cuh"
#include "corr.cuh"
#include "corr_jobs.cuh"

namespace {

using namespace corr;

struct SpecItem {      // one reference block to transform
  long long ref_off;   // element offset of the pair's reference signal
  int R;
  int i0;              // reference index of sample 0 of the block (may be negative)
};

struct SubJob {        // one (pair, ratio, offset tile)
  long long sub_off;
  long long score_off; // where the tile's Wt scores go
  long long spec_base; // index of the spectrum of block blk_lo
  int S, blk_lo, blk_hi, n_out, energy_slot;
  // bit-mask mode (b2_sync_batch): the subtitle signal is one bit per frame (raster_bits_kernel)
  long long bits_off;  // word offset of this (pair, ratio)'s speech bit mask
  float hi;            // 2*min(1/ratio, 1) - 1: value of a frame inside a cue after x -> 2x-1
};

// bit-mask mode: the speech bits of the current and the next block (kP/32 words each) sit behind
// the twiddle tables in shared memory
constexpr size_t kSmemBytesBits = kSmemBytes + 2 * (size_t)(kP / 32) * 4;

// Bulk L2 prefetch (16-byte aligned address, size a multiple of 16).
__device__ __forceinline__ void l2_prefetch(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// ------------------------------------------------------------------------------------------
// Bulk L2 prefetch of [p, p + n floats), trimmed to 16-byte granules inside the range.
__device__ __forceinline__ void l2_prefetch_floats(const float* p, int n) {
  if (n <= 0) return;
  const uintptr_t a0 = (reinterpret_cast<uintptr_t>(p) + 15) & ~uintptr_t(15);
  const uintptr_t a1 = reinterpret_cast<uintptr_t>(p + n) & ~uintptr_t(15);
  if (a1 > a0) l2_prefetch(reinterpret_cast<const void*>(a0), (uint32_t)(a1 - a0));
}

// Persistent over the reference blocks (one CTA per SM): twiddle tables are built once per CTA
// and the next item's samples are pulled into L2 while the current one is transformed.
__global__ void __maxnreg__(96)  // leaves registers for a co-resident VAD CTA (see sub_correlate_kernel)
    ref_spectra_kernel(const float* __restrict__ ref, const SpecItem* __restrict__ items, int n_items,
                       float4* __restrict__ spec, float* __restrict__ spec_energy) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* buf = reinterpret_cast<float2*>(smem_raw);
  float2* tw1024 = buf + kM;
  float2* fine32 = tw1024 + 1024;
  __shared__ float red[kThreads / 32];
  const int tid = threadIdx.x;
  init_tables(tw1024, fine32, tid);
  __syncthreads();
  const Tables t{tw1024, fine32};
  const PairCtx pc = pair_ctx(t, tid);
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const SpecItem it = items[item];
    if (tid == 0 && item + (int)gridDim.x < n_items) {
      const SpecItem nx = items[item + gridDim.x];
      const int lo = nx.i0 < 0 ? -nx.i0 : 0, hi = min(nx.R - nx.i0, kP);
      l2_prefetch_floats(ref + nx.ref_off + nx.i0 + lo, hi - lo);
    }
    float ss = float_pass1(buf, t, tid, ref + it.ref_off + it.i0, it.i0 < 0 ? -it.i0 : 0,
                           min(it.R - it.i0, kP));
    forward_rest(buf, t, tid);
    spec_store(buf, t, pc, tid, spec + (size_t)item * kPairs);
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if ((tid & 31) == 0) red[tid >> 5] = ss;
    __syncthreads();  // also: every thread is done reading buf before the next item overwrites it
    if (tid == 0) {
      float e = 0.f;
      for (int w = 0; w < kThreads / 32; ++w) e += red[w];
      spec_energy[item] = e;
    }
  }
}

// ---- tensor memory as accumulator storage ------------------------------------------------------
// The 64 accumulator floats of each thread (conj(A)*B summed over the blocks) do not fit next to
// the butterfly registers without pinning the kernel at 128 registers x 512 threads = the whole
// register file.  Blackwell's tensor memory (256 KB per SM, unused by this path otherwise) holds
// them instead: warp w owns lanes 32*(w%4).., columns 64*(w/4)..; each thread moves 16 columns at
// a time with tcgen05.ld / tcgen05.st (32x32b shape = one 32-bit column per register).
constexpr int kTmemCols = 256;

__device__ __forceinline__ uint32_t smem_addr_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_addr_u32(smem_dst)),
               "r"(kTmemCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // one full warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(kTmemCols)
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_wait_ld() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
      "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
      "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])),
      "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
      "r"(__float_as_uint(v[15]))
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// acc (tensor memory) += conj(A) * B for one block.  Column group g holds the thread's pair slots
// 4g..4g+3 as (cp.x, cp.y, cq.x, cq.y) each; the load of group g+1 is in flight while group g is
// updated.  FIRST: nothing accumulated yet, start from zero instead of loading.
// DEPTH = how many slots ahead the stored reference spectrum is read (measured on the 148-pair
// bench: 2, 4 and 8 give 7.20 / 7.23 / 7.23 ms per step - the loads are not the limiter).
template <bool FIRST, int DEPTH>
__device__ __forceinline__ void accumulate_block_tmem(uint32_t taddr, const float2* buf,
                                                      const Tables& t, const PairCtx& pc, int tid,
                                                      const float4* __restrict__ spec) {
  float4 bq[DEPTH];
#pragma unroll
  for (int d = 0; d < DEPTH; ++d) bq[d] = __ldg(spec + tid + d * kThreads);
  float cur[16], nxt[16];
  if (!FIRST) {
    tmem_ld16(taddr, cur);
    tmem_wait_ld();
  }
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    if (FIRST) {
#pragma unroll
      for (int i = 0; i < 16; ++i) cur[i] = 0.f;
    } else if (g + 1 < 4) {
      tmem_ld16(taddr + 16 * (g + 1), nxt);
    }
#pragma unroll
    for (int uu = 0; uu < 4; ++uu) {
      const int u = 4 * g + uu;
      const float4 b = bq[u % DEPTH];
      if (u + DEPTH < 16) bq[u % DEPTH] = __ldg(spec + tid + (u + DEPTH) * kThreads);
      float2 dp, dq;
      product_terms(buf, t, pc, tid, u, b, dp, dq);
      cur[4 * uu + 0] += dp.x;
      cur[4 * uu + 1] += dp.y;
      cur[4 * uu + 2] += dq.x;
      cur[4 * uu + 3] += dq.y;
    }
    tmem_st16(taddr + 16 * g, cur);
    if (!FIRST && g + 1 < 4) {
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 16; ++i) cur[i] = nxt[i];
    }
  }
  tmem_wait_st();
}

// BITS = false: the subtitle signal is a float array in global memory (b2_align_batch).
// BITS = true : it is a bit mask, one bit per frame (b2_sync_batch: raster_bits_kernel writes the
//   K masks of a pair straight from the cue list, 1/32 of the bytes of the float signals, and the
//   exact re-score reads the same masks).  The words of block blk+1 are fetched while block blk is
//   transformed (registers -> shared memory, double buffered).
template <bool TMEM, bool BITS, int DEPTH = 2>
__device__ __forceinline__ void sub_correlate_body(
    const float* __restrict__ sub, const SubJob* __restrict__ jobs, const float4* __restrict__ spec,
    const float* __restrict__ spec_energy, int L, float* __restrict__ scores,
    float4* __restrict__ job_energy, const uint32_t* __restrict__ sub_bits) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* buf = reinterpret_cast<float2*>(smem_raw);
  float2* tw1024 = buf + kM;
  float2* fine32 = tw1024 + 1024;
  uint32_t* bit_words = reinterpret_cast<uint32_t*>(smem_raw + kSmemBytes);  // [2][kP/32] (BITS only)
  __shared__ float red[kThreads / 32];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x;
  const SubJob job = jobs[blockIdx.x];
  float* out = scores + job.score_off;
  if (job.blk_lo >= job.blk_hi) {  // no subtitle block meets the reference at these offsets
    for (int m = tid; m < job.n_out; m += kThreads) out[m] = 0.f;
    if (tid == 0) job_energy[job.energy_slot] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  init_tables(tw1024, fine32, tid);
  const int wpb = L >> 5;  // mask words per block (L is a multiple of 32)
  const uint32_t* bits = BITS ? sub_bits + job.bits_off : nullptr;
  if (BITS) {
    const uint32_t* src = bits + (long long)job.blk_lo * wpb;
    uint32_t* dst = bit_words + (job.blk_lo & 1) * (kP / 32);
    for (int w = tid; w < wpb; w += kThreads) dst[w] = __ldg(src + w);
  }
  uint32_t taddr = 0;
  if (TMEM) {
    if (tid < 32) tmem_alloc(&tmem_base_s);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (TMEM) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int warp = tid >> 5;
    taddr = tmem_base_s + (uint32_t)(((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64);
  }
  const Tables t{tw1024, fine32};
  const PairCtx pc = pair_ctx(t, tid);
  SubState st;
  sub_state_clear(st);
  float er = 0.f;
  for (int blk = job.blk_lo; blk < job.blk_hi; ++blk) {
    const int j0 = blk * L;
    const bool more = blk + 1 < job.blk_hi;
    if (tid == 0 && more) {
      // pull the next block's samples and reference spectrum into L2 while this block computes
      const int jn = j0 + L;
      if (!BITS) l2_prefetch_floats(sub + job.sub_off + jn, min(job.S - jn, L));
      l2_prefetch(spec + (size_t)(job.spec_base + (blk + 1 - job.blk_lo)) * kPairs, kPairs * 16);
    }
    uint32_t nw0 = 0, nw1 = 0;  // next block's mask words: loaded now, parked in smem after the math
    if (BITS) {
      const uint32_t* src = bits + (long long)(blk + 1) * wpb;
      if (more && tid < wpb) nw0 = __ldg(src + tid);
      if (more && tid + kThreads < wpb) nw1 = __ldg(src + tid + kThreads);
      st.ss += bits_pass1(buf, t, tid, bit_words + (blk & 1) * (kP / 32), min(job.S - j0, L), L, job.hi);
      forward_rest(buf, t, tid);
    } else {
      BlockSource s;
      s.src = sub + job.sub_off + j0;
      s.t_lo = 0;
      s.t_hi = min(job.S - j0, L);
      st.ss += forward_block(buf, t, tid, s);
    }
    const size_t item = (size_t)(job.spec_base + (blk - job.blk_lo));
    if (TMEM) {
      if (blk == job.blk_lo) accumulate_block_tmem<true, DEPTH>(taddr, buf, t, pc, tid, spec + item * kPairs);
      else accumulate_block_tmem<false, DEPTH>(taddr, buf, t, pc, tid, spec + item * kPairs);
    } else {
      sub_accumulate(st, buf, t, pc, tid, spec + item * kPairs);
    }
    er += spec_energy[item];
    if (BITS) {  // the other buffer was last read in block blk-1's first pass
      uint32_t* dst = bit_words + ((blk + 1) & 1) * (kP / 32);
      dst[tid] = nw0;
      dst[tid + kThreads] = nw1;
    }
    __syncthreads();  // buf is rewritten by the next block's first pass
  }
  if (TMEM) {  // accumulators back into registers for the retangle
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float v[16];
      tmem_ld16(taddr + 16 * g, v);
      tmem_wait_ld();
#pragma unroll
      for (int uu = 0; uu < 4; ++uu) {
        st.cp[4 * g + uu] = make_float2(v[4 * uu + 0], v[4 * uu + 1]);
        st.cq[4 * g + uu] = make_float2(v[4 * uu + 2], v[4 * uu + 3]);
      }
    }
  }
  sub_retangle_store(st, buf, t, pc, tid);
  __syncthreads();
  inverse_passes_1(buf, tid);
  __syncthreads();
  inverse_passes_2(buf, t, tid);
  __syncthreads();
  inverse_passes_3(buf, t, tid);
  __syncthreads();
  inverse_passes_4(buf, t, tid);
  __syncthreads();
  for (int m = tid; m < job.n_out; m += kThreads) out[m] = window_value(buf, m);
  // ||c||_2^2 of the whole inverse-transform output (all kP real values, in score units): the
  // quantity the inverse transform's round-off is relative to (tau, window_max_kernel)
  float cn = 0.f;
  for (int i = tid; i < kM; i += kThreads) {
    const float2 z = buf[i];
    cn = fmaf(z.x, z.x, fmaf(z.y, z.y, cn));
  }
  cn *= kOutScale * kOutScale;
  float ss = st.ss;
  for (int o = 16; o > 0; o >>= 1) {
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
    cn += __shfl_xor_sync(0xffffffffu, cn, o);
  }
  __shared__ float red2[kThreads / 32];
  if ((tid & 31) == 0) {
    red[tid >> 5] = ss;
    red2[tid >> 5] = cn;
  }
  __syncthreads();
  if (tid == 0) {
    float e = 0.f, c2 = 0.f;
    for (int w = 0; w < kThreads / 32; ++w) {
      e += red[w];
      c2 += red2[w];
    }
    job_energy[job.energy_slot] = make_float4(e, er, c2, (float)(job.blk_hi - job.blk_lo));
  }
  if (TMEM) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) tmem_dealloc(tmem_base_s);
  }
}

// Product kernel: accumulators in tensor memory, 96 registers per thread so that 16 K registers
// and 88 KB of shared memory per SM stay free for a co-resident VAD CTA of the next sub-batch
// (b2_sync_batch's optional sub-batch pipeline, B2_SUBBATCHES).  Measured: the co-resident VAD CTA
// cannot keep enough bytes in flight in that space, the overlap gains 2 % (profiles/README.md), so
// the pipeline is off by default; a 128-register variant of this kernel is not faster either.
constexpr int kSubRegs = 96;
__global__ void __maxnreg__(kSubRegs)
    sub_correlate_kernel(const float* __restrict__ sub, const SubJob* __restrict__ jobs,
                         const float4* __restrict__ spec, const float* __restrict__ spec_energy,
                         int L, float* __restrict__ scores, float4* __restrict__ job_energy) {
  sub_correlate_body<true, false>(sub, jobs, spec, spec_energy, L, scores, job_energy, nullptr);
}

// b2_sync_batch: subtitle signals as bit masks (no float subtitle signal in HBM).
__global__ void __maxnreg__(kSubRegs)
    sub_correlate_bits_kernel(const SubJob* __restrict__ jobs, const float4* __restrict__ spec,
                              const float* __restrict__ spec_energy, int L,
                              float* __restrict__ scores, float4* __restrict__ job_energy,
                              const uint32_t* __restrict__ sub_bits) {
  sub_correlate_body<true, true>(nullptr, jobs, spec, spec_energy, L, scores, job_energy, sub_bits);
}

// A/B variant (B2_ACC=reg): accumulators in registers, 128 registers per thread.
__global__ void __launch_bounds__(kThreads, 1)
    sub_correlate_regacc_kernel(const float* __restrict__ sub, const SubJob* __restrict__ jobs,
                                const float4* __restrict__ spec,
                                const float* __restrict__ spec_energy, int L,
                                float* __restrict__ scores, float4* __restrict__ job_energy) {
  sub_correlate_body<false, false>(sub, jobs, spec, spec_energy, L, scores, job_energy, nullptr);
}

// Same for the bit-mask path (compute-sanitizer's synccheck does not model tcgen05.alloc; the
// sanitizer runs use B2_ACC=reg).
__global__ void __launch_bounds__(kThreads, 1)
    sub_correlate_bits_regacc_kernel(const SubJob* __restrict__ jobs, const float4* __restrict__ spec,
                                     const float* __restrict__ spec_energy, int L,
                                     float* __restrict__ scores, float4* __restrict__ job_energy,
                                     const uint32_t* __restrict__ sub_bits) {
  sub_correlate_body<false, true>(nullptr, jobs, spec, spec_energy, L, scores, job_energy, sub_bits);
}

// ---- candidate selection ---------------------------------------------------------------------
// Per (pair, ratio): approximate maximum over the surviving window, then every offset whose
// fp32 score is within tau of it, taken from the LARGEST offset down (np.argmax returns the
// lowest index = largest offset among equal values), at most kCandMax of them.
// Phase 1: fp32 maximum of the surviving window and the round-off bound tau, per (pair, ratio).
__global__ void __launch_bounds__(256) window_max_kernel(const SelJob* __restrict__ jobs,
                                                          float* __restrict__ scores,
                                                          const float4* __restrict__ job_energy,
                                                          float2* __restrict__ job_stat, int wt) {
  const SelJob job = jobs[blockIdx.x];
  const int tid = threadIdx.x;
  __shared__ float smax[256];
  if (job.kind != 0 || job.m_lo > job.m_hi) {
    if (tid == 0) job_stat[blockIdx.x] = make_float2(-INFINITY, 0.f);
    return;
  }
  float* c = scores + job.score_off;
  if (job.n_split > 1) {  // add the partial score arrays (fixed order: deterministic) into the first
    const int n = job.n_tiles * wt;
    for (int m = tid; m < n; m += 256) {
      float v = c[m];
      for (int sp = 1; sp < job.n_split; ++sp) v += c[(size_t)sp * n + m];
      c[m] = v;
    }
    __syncthreads();
  }
  float mx = -INFINITY;
  for (int m = job.m_lo + tid; m <= job.m_hi; m += 256) mx = fmaxf(mx, c[m]);
  smax[tid] = mx;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (tid < w) smax[tid] = fmaxf(smax[tid], smax[tid + w]);
    __syncthreads();
  }
  if (tid == 0) {
    float tau = 0.f;
    for (int i = 0; i < job.n_tiles; ++i) {
      // energies add over the chunks of a tile (Cauchy-Schwarz); the norms of the partial inverse
      // transforms add (triangle inequality); blocks beyond kTauBlocks add u each to the
      // accumulation term
      float ex = 0.f, ey = 0.f, cn = 0.f, blocks = 0.f;
      for (int sp = 0; sp < job.n_split; ++sp) {
        const float4 e = job_energy[job.energy_slot + i * job.n_split + sp];
        ex += e.x;
        ey += e.y;
        cn += sqrtf(e.z);
        blocks = fmaxf(blocks, e.w);
      }
      const float fwd = kTauFwd + fmaxf(0.f, blocks - (float)kTauBlocks);
      tau = fmaxf(tau, kU * (fwd * sqrtf(ex * ey) + (kTauInv + (float)(job.n_split - 1)) * cn));
    }
    job_stat[blockIdx.x] = make_float2(smax[0], tau * 1.0001f + 1e-30f);
  }
}

// Phase 2.  With winner_only (b2_sync_batch when only the best ratio is wanted) a (pair, ratio)
// whose fp32 maximum cannot reach the best ratio's even after round-off (mx + tau < max_k (mx_k -
// tau_k)) is not re-scored: it reports its fp32 maximum and is flagged B2_ALIGN_APPROX.
__global__ void __launch_bounds__(256) select_candidates_kernel(
    const SelJob* __restrict__ jobs, const float* __restrict__ scores,
    const float2* __restrict__ job_stat, int K, int winner_only, int* __restrict__ cand_off,
    int* __restrict__ cand_cnt, int* __restrict__ work_list, int* __restrict__ work_count) {
  const SelJob job = jobs[blockIdx.x];
  const int tid = threadIdx.x;
  __shared__ int scount;
  __shared__ int swarp[8];
  if (job.kind != 0 || job.m_lo > job.m_hi) {
    if (tid == 0) cand_cnt[blockIdx.x] = 0;
    return;
  }
  const float* c = scores + job.score_off;
  const float2 stat = job_stat[blockIdx.x];
  float cut = stat.x - stat.y;
  bool approx_only = false;
  if (winner_only && !job.no_prune) {
    const int b0 = (blockIdx.x / K) * K;
    float best_floor = -INFINITY;
    for (int k = 0; k < K; ++k) {
      const float2 s = job_stat[b0 + k];
      best_floor = fmaxf(best_floor, s.x - s.y);
    }
    if (stat.x + stat.y < best_floor) {  // cannot win: keep only the fp32 argmax (largest offset)
      approx_only = true;
      cut = stat.x;
    }
  }
  if (tid == 0) scount = 0;
  __syncthreads();
  // walk m from high to low so that slots fill in order of increasing index m_hi - m ... wait:
  // offset o = o_first + m, so the largest offset is the largest m.
  for (int top = job.m_hi; top >= job.m_lo; top -= 256) {
    const int m = top - tid;
    const bool hit = (m >= job.m_lo) && (c[m] >= cut);
    const unsigned ball = __ballot_sync(0xffffffffu, hit);
    if ((tid & 31) == 0) swarp[tid >> 5] = __popc(ball);
    __syncthreads();
    int before = scount;
    for (int w = 0; w < (tid >> 5); ++w) before += swarp[w];
    before += __popc(ball & ((1u << (tid & 31)) - 1u));
    if (hit && before < kCandMax) cand_off[(size_t)blockIdx.x * kCandMax + before] = job.o_first + m;
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int w = 0; w < 8; ++w) tot += swarp[w];
      scount += tot;
    }
    __syncthreads();
  }
  if (approx_only) {  // slot 0 holds the largest offset attaining the fp32 maximum
    if (tid == 0) cand_cnt[blockIdx.x] = -1;
    return;
  }
  if (tid == 0) {
    cand_cnt[blockIdx.x] = scount;
    scount = min(scount, kCandMax);
    swarp[0] = atomicAdd(work_count, scount);  // slots in the global re-score work list
  }
  __syncthreads();
  if (tid < scount) work_list[swarp[0] + tid] = ((int)blockIdx.x << 5) | tid;
}

// ---- exact re-score ----------------------------------------------------------------------------
// score(o) = sum over the overlap of (2 s[j] - 1)(2 r[j+o] - 1) in float64.  Each candidate's
// overlap is cut into kRescoreSeg segments handled by different CTAs (persistent grid over the
// work list); every partial sum has a fixed summation order and pick_kernel adds the partials in
// segment order, so the result is deterministic.
__global__ void __launch_bounds__(256) rescore_kernel(const SelJob* __restrict__ jobs,
                                                       const float* __restrict__ ref,
                                                       const float* __restrict__ sub,
                                                       const int* __restrict__ cand_off,
                                                       const int* __restrict__ work_list,
                                                       const int* __restrict__ work_count,
                                                       const uint32_t* __restrict__ sub_bits,
                                                       double* __restrict__ cand_partial) {
  __shared__ double sh[256];
  const int total = *work_count * kRescoreSeg;
  for (int w = blockIdx.x; w < total; w += gridDim.x) {
    const int item = work_list[w / kRescoreSeg], seg = w % kRescoreSeg;
    const int j = item >> 5, ci = item & 31;
    const SelJob job = jobs[j];
    const int o = cand_off[(size_t)j * kCandMax + ci];
    const float* r = ref + job.ref_off;
    const float* s = sub + job.sub_off;
    const int j_lo = max(0, -o), j_hi = min(job.S, job.R - o);
    const int len = max(0, j_hi - j_lo);
    const int per = (len + kRescoreSeg - 1) / kRescoreSeg;
    const int a0 = j_lo + seg * per, a1 = min(j_hi, a0 + per);
    double acc = 0.0;
    int i = a0 + threadIdx.x;
    if (job.bits_off >= 0) {
      // bit-mask mode: subtitle frame i is bit i of the mask written by raster_bits_kernel;
      // its value after x -> 2x-1 is (2*level - 1) inside a cue and -1 outside
      const uint32_t* bits = sub_bits + job.bits_off;
      const double hi = 2.0 * (double)job.sub_level - 1.0;
      for (; i + 7 * 256 < a1; i += 8 * 256) {  // 16 independent loads in flight per thread
        uint32_t bw[8];
        float rv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          bw[u] = __ldg(bits + ((i + u * 256) >> 5));
          rv[u] = __ldg(r + i + u * 256 + o);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)  // 256 = 0 (mod 32): the bit position is the same for all u
          acc = fma(((bw[u] >> (i & 31)) & 1u) ? hi : -1.0, 2.0 * (double)rv[u] - 1.0, acc);
      }
      for (; i < a1; i += 256) {
        const double a = ((__ldg(bits + (i >> 5)) >> (i & 31)) & 1u) ? hi : -1.0;
        const double b = 2.0 * (double)__ldg(r + i + o) - 1.0;
        acc = fma(a, b, acc);
      }
    }
    for (; i + 3 * 256 < a1; i += 4 * 256) {  // 8 independent loads in flight per thread
      float sv[4], rv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        sv[u] = __ldg(s + i + u * 256);
        rv[u] = __ldg(r + i + u * 256 + o);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        acc = fma(2.0 * (double)sv[u] - 1.0, 2.0 * (double)rv[u] - 1.0, acc);
    }
    for (; i < a1; i += 256) {
      const double a = 2.0 * (double)__ldg(s + i) - 1.0;
      const double b = 2.0 * (double)__ldg(r + i + o) - 1.0;
      acc = fma(a, b, acc);
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int h = 128; h > 0; h >>= 1) {
      if (threadIdx.x < h) sh[threadIdx.x] += sh[threadIdx.x + h];
      __syncthreads();
    }
    if (threadIdx.x == 0) cand_partial[((size_t)j * kCandMax + ci) * kRescoreSeg + seg] = sh[0];
    __syncthreads();
  }
}

__global__ void __launch_bounds__(128) pick_kernel(const SelJob* __restrict__ jobs, int n_jobs,
                                                    const int* __restrict__ cand_off,
                                                    const int* __restrict__ cand_cnt,
                                                    const double* __restrict__ cand_partial,
                                                    const float2* __restrict__ job_stat,
                                                    double* __restrict__ score,
                                                    int32_t* __restrict__ offset,
                                                    int32_t* __restrict__ status) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_jobs) return;
  const SelJob job = jobs[j];
  if (job.kind == 1) {
    score[job.out_index] = 0.0;
    offset[job.out_index] = 0;
    status[job.out_index] = B2_ALIGN_EMPTY;
    return;
  }
  if (job.kind == 2 || job.m_lo > job.m_hi) {
    score[job.out_index] = -INFINITY;  // np.argmax of an all -inf array: index 0
    offset[job.out_index] = job.masked_offset;
    status[job.out_index] = B2_ALIGN_ALL_MASKED;
    return;
  }
  const int cnt = cand_cnt[j];
  if (cnt < 0) {  // winner_only: this ratio cannot win, fp32 result reported
    score[job.out_index] = (double)job_stat[j].x;
    offset[job.out_index] = cand_off[(size_t)j * kCandMax];
    status[job.out_index] = B2_ALIGN_APPROX;
    return;
  }
  const int n = min(cnt, kCandMax);
  double bs = -INFINITY;
  int bo = 0;
  for (int c = 0; c < n; ++c) {
    double s = 0.0;
    for (int g = 0; g < kRescoreSeg; ++g) s += cand_partial[((size_t)j * kCandMax + c) * kRescoreSeg + g];
    const int o = cand_off[(size_t)j * kCandMax + c];
    if (c == 0 || s > bs || (s == bs && o > bo)) {
      bs = s;
      bo = o;
    }
  }
  score[job.out_index] = bs;
  offset[job.out_index] = bo;
  status[job.out_index] = cnt > kCandMax ? B2_ALIGN_CAND_OVERFLOW : B2_ALIGN_OK;
}

// ---- host planning ----------------------------------------------------------------------------
long long padded_length(const b2_ctx* h, long long n) {
  // int(2 ** math.ceil(math.log(n, 2))), aligners.py:67-68, libm quirks included
  int k = 0;
  while ((1LL << k) < n) ++k;
  if ((1LL << k) == n && (h->log2_quirk_mask >> k) & 1ULL) ++k;
  return 1LL << k;
}

long long floor_div(long long a, long long b) {
  long long q = a / b;
  if ((a % b != 0) && ((a < 0) != (b < 0))) --q;
  return q;
}

}  // namespace

int b2i_align_launch(b2_ctx* h, const float* d_ref, const int64_t* ref_off, const float* d_sub,
                     const int64_t* sub_off, int B, int K, int64_t max_offset_samples,
                     double* d_score, int32_t* d_offset, int32_t* d_status, int winner_only,
                     const B2CueSource* cue_src) {
  B2Range range("b2:align (ref_spectra, sub_correlate, select, rescore, pick)");
  const size_t J = (size_t)B * K;
  // cue mode (b2_sync_batch): the subtitle signals exist only as bit masks, rasterised from the cue
  // list by raster_bits_kernel below (sub_off then only carries the signal lengths)
  const bool cue_mode = cue_src != nullptr;
  std::vector<long long> bits_off(cue_mode ? J + 1 : 1, 0);
  if (cue_mode)
    for (size_t j = 0; j < J; ++j)
      bits_off[j + 1] = bits_off[j] + ((sub_off[j + 1] - sub_off[j]) + kP) / 32 + 1;
  std::vector<SelJob> sel(J);
  std::vector<long long> idx_lo(J, 0), idx_hi(J, 0), n_pad(J, 0);
  struct PairPlan { long long o_min, o_max; int n_tiles; bool any; };
  std::vector<PairPlan> pp(B);
  long long max_w = 1;
  bool big_ok = true;   // every live job's padded length suits the large-window path
  for (int b = 0; b < B; ++b) {
    const long long R = ref_off[b + 1] - ref_off[b];
    if (R < 0 || R > 0x3fffffff) B2_FAIL(h, B2_ERR_BAD_ARG, "align: bad reference length at %d", b);
    PairPlan& p = pp[b];
    p.any = false;
    p.o_min = 0;
    p.o_max = -1;
    for (int k = 0; k < K; ++k) {
      const size_t j = (size_t)b * K + k;
      const long long S = sub_off[j + 1] - sub_off[j];
      if (S < 0 || S > 0x3fffffff) B2_FAIL(h, B2_ERR_BAD_ARG, "align: bad subtitle length at %zu", j);
      SelJob& s = sel[j];
      memset(&s, 0, sizeof(s));
      s.ref_off = ref_off[b];
      s.sub_off = sub_off[j];
      s.R = (int)R;
      s.S = (int)S;
      s.out_index = (int)j;
      s.bits_off = cue_mode ? bits_off[j] : -1;
      s.sub_level = cue_mode ? (float)std::min(1.0 / cue_src->ratios[k], 1.0) : 0.f;  // speech_transformers.py:977
      if (R == 0 || S == 0) {  // aligners.py:58-66
        s.kind = 1;
        continue;
      }
      const long long N = padded_length(h, R + S);
      long long lo = 0, hi = N;  // surviving index range, aligners.py:31-43 with slice semantics
      if (max_offset_samples != B2_MAX_OFFSET_NONE) {
        // any int64 width, negative ones included, through the reference's slice arithmetic; widths
        // are clamped to +-2^40 first (beyond every padded length, so the result is unchanged)
        const long long mo = std::max<long long>(-(1LL << 40), std::min<long long>(1LL << 40, max_offset_samples));
        const long long a = N - 1 - mo - S;
        const long long bb = N - 1 + mo - S;
        lo = a >= 0 ? std::min(a, N) : std::max(a + N, 0LL);
        hi = bb >= 0 ? std::min(bb, N) : std::max(bb + N, 0LL);
      }
      if (lo >= hi) {
        s.kind = 2;
        s.masked_offset = (int)(N - 1 - S);
        continue;
      }
      const long long o_lo = N - S - hi, o_hi = N - 1 - S - lo;  // aligners.py:47
      idx_lo[j] = lo;
      idx_hi[j] = hi;
      n_pad[j] = N;
      if (N < (1LL << (bigfft_min_log2n())) || N > (1LL << bigfft_max_log2n())) big_ok = false;
      s.kind = 0;
      s.m_lo = (int)o_lo;  // temporarily absolute offsets; rebased below
      s.m_hi = (int)o_hi;
      if (!p.any) {
        p.o_min = o_lo;
        p.o_max = o_hi;
        p.any = true;
      } else {
        p.o_min = std::min(p.o_min, o_lo);
        p.o_max = std::max(p.o_max, o_hi);
      }
    }
    if (p.any) max_w = std::max(max_w, p.o_max - p.o_min + 1);
    if (p.any && max_offset_samples != B2_MAX_OFFSET_NONE) {
      const long long mo = std::max<long long>(-(1LL << 40), std::min<long long>(1LL << 40, max_offset_samples));
      if (std::max(llabs(p.o_min), llabs(p.o_max)) > mo)
        for (int k = 0; k < K; ++k) sel[(size_t)b * K + k].no_prune = 1;
    }
  }
  // offsets per tile: Wt = 1 (mod 32) so that L = P - Wt + 1 is a multiple of 32 (vector loads,
  // whole words of the speech bit mask per block), at most P/2 + 1
  const int Wt = (int)(max_w <= kP / 2 + 1 ? 32 * ((max_w + 30) / 32) + 1 : (kP / 2 + 1));
  const int L = kP - Wt + 1;
  uint32_t* d_bits = nullptr;
  if (cue_mode) {
    void* db;
    B2_TRY(b2i_ws(h, b2_ctx::WS_SIG_SUB, (size_t)bits_off[J] * 4 + 64, &db));
    d_bits = (uint32_t*)db;
    B2_TRY(b2i_raster_bits_launch(h, cue_src, B, K, sub_off, bits_off.data(), d_bits));
    B2_CUDA(h, cudaFuncSetAttribute(sub_correlate_bits_kernel,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytesBits));
    B2_CUDA(h, cudaFuncSetAttribute(sub_correlate_bits_regacc_kernel,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytesBits));
  }
  void* d_cand;
  B2_TRY(b2i_ws(h, b2_ctx::WS_CAND, J * kCandMax * (8 * kRescoreSeg + 8) + J * 12 + 64, &d_cand));
  B2CandBuffers cb;
  cb.cand_partial = (double*)d_cand;
  cb.job_stat = (float2*)(cb.cand_partial + J * kCandMax * kRescoreSeg);
  cb.cand_off = (int*)(cb.job_stat + J);
  cb.cand_cnt = cb.cand_off + J * kCandMax;
  cb.work_list = cb.cand_cnt + J;
  cb.work_count = cb.work_list + J * kCandMax;
  if (J >= (1u << 26)) B2_FAIL(h, B2_ERR_UNSUPPORTED, "align: B*K too large for one call");

  // Large windows (FFTAligner's default max_offset_samples=None, or a mask wider than a few tiles):
  // the overlap-save path would recompute every block for every 16 385-offset tile; one padded-length
  // FFT per signal (four-step, bigfft.cu) is cheaper from kBigMinTiles tiles on.
  // B2_ALIGN_PATH=tiled|big: test / A-B knob.
  bool use_big = big_ok && max_w > (long long)kBigMinTiles * (kP / 2 + 1);
  if (const char* e = getenv("B2_ALIGN_PATH")) {
    if (!strcmp(e, "tiled")) use_big = false;
    if (!strcmp(e, "big") && big_ok) use_big = true;
  }
  if (use_big) {
    const SelJob* d_sel_big = nullptr;
    B2_TRY(b2i_align_big(h, d_ref, d_sub, d_bits, B, K, sel, idx_lo, idx_hi, n_pad, winner_only, cb, &d_sel_big));
    return b2i_rescore_pick(h, d_sel_big, J, d_ref, d_sub, d_bits, cb, d_score, d_offset, d_status);
  }

  // score buffers + per-(pair,ratio) bookkeeping
  // Small batches: with fewer (pair, ratio, tile) jobs than SMs the block loop of a job (35 blocks
  // for a 2 h signal) would run on a handful of SMs.  The block range of every job is then cut into
  // n_split chunks, each CTA inverse-transforms its own partial accumulator (the inverse FFT is
  // linear) into its own partial score array, and window_max_kernel adds the partial arrays in a
  // fixed order.  Costs one extra inverse transform per chunk, so chunks keep >= 4 blocks.
  long long n_jobs_total = 0, max_blocks = 1;
  for (int b = 0; b < B; ++b) {
    PairPlan& p = pp[b];
    p.n_tiles = p.any ? (int)ceil_div64(p.o_max - p.o_min + 1, Wt) : 0;
    for (int k = 0; k < K; ++k) {
      const SelJob& s = sel[(size_t)b * K + k];
      if (s.kind != 0) continue;
      n_jobs_total += p.n_tiles;
      max_blocks = std::max<long long>(max_blocks, ceil_div64(s.S, L));
    }
  }
  int n_split = 1;
  if (n_jobs_total > 0 && n_jobs_total < 2LL * h->sm_count)
    n_split = (int)std::max<long long>(1, std::min<long long>(ceil_div64(2LL * h->sm_count, n_jobs_total),
                                                               max_blocks / 4));
  if (const char* e = getenv("B2_ALIGN_SPLIT")) n_split = std::max(1, atoi(e));  // test / tuning knob
  long long score_total = 0, energy_total = 0;
  for (int b = 0; b < B; ++b) {
    PairPlan& p = pp[b];
    for (int k = 0; k < K; ++k) {
      SelJob& s = sel[(size_t)b * K + k];
      if (s.kind != 0) continue;
      s.o_first = (int)p.o_min;
      s.m_lo -= (int)p.o_min;
      s.m_hi -= (int)p.o_min;
      s.score_off = score_total;
      s.energy_slot = (int)energy_total;
      s.n_tiles = p.n_tiles;
      s.n_split = n_split;
      score_total += (long long)p.n_tiles * Wt * n_split;
      energy_total += (long long)p.n_tiles * n_split;
    }
  }

  void* d_scores;
  B2_TRY(b2i_ws(h, b2_ctx::WS_SCORES, (size_t)(score_total + 16) * 4 + (size_t)(energy_total + 2) * 16,
                &d_scores));
  float* scores = (float*)d_scores;
  float4* job_energy = (float4*)((char*)d_scores + (((size_t)(score_total + 16) * 4 + 15) & ~size_t(15)));

  B2_CUDA(h, cudaFuncSetAttribute(ref_spectra_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)kSmemBytes));
  B2_CUDA(h, cudaFuncSetAttribute(sub_correlate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)kSmemBytes));
  B2_CUDA(h, cudaFuncSetAttribute(sub_correlate_regacc_kernel,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));

  // Work is issued in groups so that the reference spectra of a group fit the workspace cap.
  const size_t kSpecBytes = (size_t)kPairs * 16;
  const size_t cap_items = std::max<size_t>(64, ((size_t)6 << 30) / kSpecBytes);
  std::vector<SpecItem> items;
  std::vector<SubJob> jobs;
  auto flush = [&]() -> int {
    if (jobs.empty()) {
      items.clear();
      return B2_OK;
    }
    void* d_spec;
    B2_TRY(b2i_ws(h, b2_ctx::WS_SPEC, items.size() * kSpecBytes + items.size() * 4 + 256, &d_spec));
    float4* spec = (float4*)d_spec;
    float* spec_energy = (float*)((char*)d_spec + items.size() * kSpecBytes);
    MetaArena a;
    B2_TRY(b2i_meta_begin(h, &a, items.size() * sizeof(SpecItem) + jobs.size() * sizeof(SubJob) + 256));
    const SpecItem* d_items = (const SpecItem*)b2i_meta_put(&a, items.data(), items.size() * sizeof(SpecItem));
    const SubJob* d_jobs = (const SubJob*)b2i_meta_put(&a, jobs.data(), jobs.size() * sizeof(SubJob));
    B2_TRY(b2i_meta_commit(&a));
    if (!items.empty()) {
      // persistent: one CTA per SM - or per SM the pipeline leaves to the correlation stage
      int max_ctas = h->corr_max_ctas > 0 ? std::min(h->corr_max_ctas, h->sm_count) : h->sm_count;
      if (const char* e = getenv("B2_CORR_MAX_CTAS")) max_ctas = std::max(1, std::min(h->sm_count, atoi(e)));  // probe knob
      const unsigned grid = (unsigned)std::min<size_t>(items.size(), (size_t)max_ctas);
      ref_spectra_kernel<<<grid, kThreads, kSmemBytes, h->stream>>>(d_ref, d_items, (int)items.size(),
                                                                    spec, spec_energy);
      B2_CHECK_LAUNCH(h, "ref_spectra_kernel");
    }
    if (cue_mode && h->acc_in_tmem)
      sub_correlate_bits_kernel<<<(unsigned)jobs.size(), kThreads, kSmemBytesBits, h->stream>>>(
          d_jobs, spec, spec_energy, L, scores, job_energy, d_bits);
    else if (cue_mode)
      sub_correlate_bits_regacc_kernel<<<(unsigned)jobs.size(), kThreads, kSmemBytesBits, h->stream>>>(
          d_jobs, spec, spec_energy, L, scores, job_energy, d_bits);
    else if (h->acc_in_tmem)
      sub_correlate_kernel<<<(unsigned)jobs.size(), kThreads, kSmemBytes, h->stream>>>(
          d_sub, d_jobs, spec, spec_energy, L, scores, job_energy);
    else
      sub_correlate_regacc_kernel<<<(unsigned)jobs.size(), kThreads, kSmemBytes, h->stream>>>(
          d_sub, d_jobs, spec, spec_energy, L, scores, job_energy);
    B2_CHECK_LAUNCH(h, "sub_correlate_kernel");
    items.clear();
    jobs.clear();
    return B2_OK;
  };

  for (int b = 0; b < B; ++b) {
    const PairPlan& p = pp[b];
    if (!p.any) continue;
    const long long R = ref_off[b + 1] - ref_off[b];
    long long nblk_max = 0;
    for (int k = 0; k < K; ++k) {
      const SelJob& s = sel[(size_t)b * K + k];
      if (s.kind == 0) nblk_max = std::max<long long>(nblk_max, ceil_div64(s.S, L));
    }
    for (int tile = 0; tile < p.n_tiles; ++tile) {
      const long long o_t = p.o_min + (long long)tile * Wt;
      const long long blk_lo = std::max(0LL, floor_div(-o_t - kP, L) + 1);
      const long long blk_hi = std::min<long long>(nblk_max, R - o_t > 0 ? ceil_div64(R - o_t, L) : 0LL);
      const long long n_items = std::max(0LL, blk_hi - blk_lo);
      if (items.size() + (size_t)n_items > cap_items) B2_TRY(flush());
      const long long spec_base = (long long)items.size();
      for (long long blk = blk_lo; blk < blk_hi; ++blk) {
        SpecItem it;
        it.ref_off = ref_off[b];
        it.R = (int)R;
        it.i0 = (int)(blk * L + o_t);
        items.push_back(it);
      }
      for (int k = 0; k < K; ++k) {
        const SelJob& s = sel[(size_t)b * K + k];
        if (s.kind != 0) continue;
        const long long job_hi = std::min<long long>(blk_hi, ceil_div64(s.S, L));
        const long long n_blk = std::max(0LL, job_hi - blk_lo);
        for (int sp = 0; sp < n_split; ++sp) {
          const long long c_lo = blk_lo + n_blk * sp / n_split, c_hi = blk_lo + n_blk * (sp + 1) / n_split;
          SubJob jb;
          jb.sub_off = s.sub_off;
          jb.S = s.S;
          jb.score_off = s.score_off + ((long long)sp * p.n_tiles + tile) * Wt;
          jb.spec_base = spec_base + (c_lo - blk_lo);
          jb.blk_lo = (int)c_lo;
          jb.blk_hi = (int)c_hi;
          jb.n_out = Wt;
          jb.energy_slot = s.energy_slot + tile * n_split + sp;
          jb.bits_off = 0;
          jb.hi = 0.f;
          if (cue_mode) {
            jb.bits_off = s.bits_off;
            jb.hi = 2.f * s.sub_level - 1.f;  // the value load_block16 gives the float signal
          }
          jobs.push_back(jb);
        }
      }
    }
  }
  B2_TRY(flush());

  B2Range range_sel("b2:select+rescore+pick");

  MetaArena a;
  B2_TRY(b2i_meta_begin(h, &a, J * sizeof(SelJob) + 256));
  const SelJob* d_sel = (const SelJob*)b2i_meta_put(&a, sel.data(), J * sizeof(SelJob));
  B2_TRY(b2i_meta_commit(&a));
  B2_CUDA(h, cudaMemsetAsync(cb.work_count, 0, sizeof(int), h->stream));
  window_max_kernel<<<(unsigned)J, 256, 0, h->stream>>>(d_sel, scores, job_energy, cb.job_stat, Wt);
  B2_CHECK_LAUNCH(h, "window_max_kernel");
  select_candidates_kernel<<<(unsigned)J, 256, 0, h->stream>>>(d_sel, scores, cb.job_stat, K, winner_only,
                                                                cb.cand_off, cb.cand_cnt, cb.work_list,
                                                                cb.work_count);
  B2_CHECK_LAUNCH(h, "select_candidates_kernel");
  return b2i_rescore_pick(h, d_sel, J, d_ref, d_sub, d_bits, cb, d_score, d_offset, d_status);
}

int b2i_rescore_pick(b2_ctx* h, const SelJob* d_sel, size_t J, const float* d_ref, const float* d_sub,
                     const uint32_t* d_bits, const B2CandBuffers& cb, double* d_score, int32_t* d_offset,
                     int32_t* d_status) {
  rescore_kernel<<<(unsigned)(h->sm_count * 8), 256, 0, h->stream>>>(
      d_sel, d_ref, d_sub, cb.cand_off, cb.work_list, cb.work_count, d_bits, cb.cand_partial);
  B2_CHECK_LAUNCH(h, "rescore_kernel");
  pick_kernel<<<(unsigned)((J + 127) / 128), 128, 0, h->stream>>>(d_sel, (int)J, cb.cand_off, cb.cand_cnt,
                                                                   cb.cand_partial, cb.job_stat, d_score,
                                                                   d_offset, d_status);
  B2_CHECK_LAUNCH(h, "pick_kernel");
  return B2_OK;
}

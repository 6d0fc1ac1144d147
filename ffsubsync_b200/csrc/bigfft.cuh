// Device building blocks of the LARGE-WINDOW correlation (FFTAligner with max_offset_samples=None
// or a mask wider than a few overlap-save tiles; ffsubsync/aligners.py:67-80): one real FFT of the
// padded length N = 2 M per signal, done as a four-step complex FFT of M = M1 x 1024 points over
// HBM / L2, every step a batched in-shared-memory transform of one 16 384-point tile:
//
//   F1  columns: tile = all M1 rows x (16384 / M1) adjacent columns, row-major exactly as in global
//       memory (coalesced, no transpose): the column FFT over n1 is the TOP log2(M1) levels of the
//       16 384-point decimation-in-frequency network (passes at strides 1024, 64 and "cols", twiddles
//       taken at the column-cleared index), then the four-step twiddle w_M^(n2 k1), stored k1-major.
//   F2  rows:    tile = 8 rows + their 8 partner rows (k1 <-> M1 - k1), each a contiguous 1024-point
//       transform (the passes corr.cuh already has); the real-FFT untangle pairs element (k1, k2)
//       with (M1 - k1, 1023 - k2) = the mirrored POSITION in the partner row.  Reference: store the
//       spectrum.  Subtitles: conj(A) * B, retangle, inverse row transform and the conjugate
//       four-step twiddle in the same kernel (one global round trip saved).
//   F3  columns, inverse: the mirror of F1; writes the N scores, their maximum over the surviving
//       window and the sum of squares (the norm the round-off bound needs).
//
// Like corr.cuh everything is __host__ __device__: tests/host_emul/bigfft_emul.cu runs the same
// code thread by thread on the CPU against np.fft (the build container has no GPU).
#pragma once
#include "corr.cuh"

namespace bigfft {

using namespace corr;

constexpr int kRow = 1024;      // M2: row length of the four-step factorisation
constexpr int kMinQ1 = 6;       // M1 = 2^q1 rows, 64 ... 4096  ->  N = 2^17 ... 2^23
constexpr int kMaxQ1 = 12;

// ---- small DFTs without twiddles ---------------------------------------------------------------
template <bool INV>
CORR_HD void r2(float2& a0, float2& a1) {
  const float2 s = cadd(a0, a1), d = csub(a0, a1);
  a0 = s;
  a1 = d;
}

// multiply by exp(-+ 2 pi i n / 8)
template <bool INV, int N>
CORR_HD float2 rot8(float2 v) {
  constexpr float r = 0.70710678118654752f;
  constexpr int n = N & 7;
  if (n == 0) return v;
  if (n == 2) return INV ? make_float2(-v.y, v.x) : make_float2(v.y, -v.x);
  float cr = (n == 1) ? r : -r, ci = -r;   // n = 1: (r, -r), n = 3: (-r, -r)
  if (INV) ci = -ci;
  return make_float2(v.x * cr - v.y * ci, v.x * ci + v.y * cr);
}

// v[q] = x[q] in, v[k] = X[k] out (natural order both sides)
template <bool INV>
CORR_HD void dft8(float2 (&v)[8]) {
  // decimation in frequency: t_j = x_j + x_{j+4},  u_j = (x_j - x_{j+4}) w8^j
  float2 t[4], u[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    t[j] = cadd(v[j], v[j + 4]);
    u[j] = csub(v[j], v[j + 4]);
  }
  u[1] = rot8<INV, 1>(u[1]);
  u[2] = rot8<INV, 2>(u[2]);
  u[3] = rot8<INV, 3>(u[3]);
  r4<INV>(t[0], t[1], t[2], t[3]);
  r4<INV>(u[0], u[1], u[2], u[3]);
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    v[2 * m] = t[m];
    v[2 * m + 1] = u[m];
  }
}

// v[q] = x[q] in, v[k] = X[k] out
template <bool INV>
CORR_HD void dft16(float2 (&v)[16]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) r4<INV>(v[q], v[q + 4], v[q + 8], v[q + 12]);
  // now v[q + 4 a] = sum over the q-residue class, index a; twiddle w16^(q a)
  v[5] = rot16<INV, 1>(v[5]);
  v[6] = rot16<INV, 2>(v[6]);
  v[7] = rot16<INV, 3>(v[7]);
  v[9] = rot16<INV, 2>(v[9]);
  v[10] = rot16<INV, 4>(v[10]);
  v[11] = rot16<INV, 6>(v[11]);
  v[13] = rot16<INV, 3>(v[13]);
  v[14] = rot16<INV, 6>(v[14]);
  v[15] = rot16<INV, 9>(v[15]);
#pragma unroll
  for (int a = 0; a < 4; ++a) r4<INV>(v[4 * a], v[4 * a + 1], v[4 * a + 2], v[4 * a + 3]);
  // v[4 a + b] = X[a + 4 b]  ->  natural order
  float2 w[16];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) w[a + 4 * b] = v[4 * a + b];
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = w[k];
}

// ---- column passes on a row-major tile [M1][cols], cols = 2^CL ---------------------------------
// The two upper passes are corr.cuh's radix-16 passes with the twiddle taken at the column-cleared
// index (the twiddle depends on the row part of j only).
template <int SUB_LOG2, int CL>
CORR_HD void dif16_pass_cols(float2* buf, const Tables& t, int tid) {
#pragma unroll 1
  for (int rep = 0; rep < 2; ++rep) {
    const int u = tid + rep * kThreads;
    const int j = u & ((1 << SUB_LOG2) - 1);
    int reg[4];
    pass_addr_init<SUB_LOG2>(u, reg);
    float2 v[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] = buf[pass_addr<SUB_LOG2>(reg, q)];
    bfly16_dif(v, pass_twiddle<SUB_LOG2>(t, j & ~((1 << CL) - 1)));
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) buf[pass_addr<SUB_LOG2>(reg, a + 4 * b)] = v[4 * a + b];
  }
}

template <int SUB_LOG2, int CL>
CORR_HD void dit16_pass_cols(float2* buf, const Tables& t, int tid) {
#pragma unroll 1
  for (int rep = 0; rep < 2; ++rep) {
    const int u = tid + rep * kThreads;
    const int j = u & ((1 << SUB_LOG2) - 1);
    int reg[4];
    pass_addr_init<SUB_LOG2>(u, reg);
    float2 v[16];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) v[4 * a + b] = buf[pass_addr<SUB_LOG2>(reg, a + 4 * b)];
    bfly16_dit(v, cconj(pass_twiddle<SUB_LOG2>(t, j & ~((1 << CL) - 1))));
#pragma unroll
    for (int q = 0; q < 16; ++q) buf[pass_addr<SUB_LOG2>(reg, q)] = v[q];
  }
}

// Last forward / first inverse column pass: radix R at stride cols, no twiddles.
template <bool INV, int R, int CL>
CORR_HD void last_pass_cols(float2* buf, int tid) {
  constexpr int n_bfly = kM / R;
#pragma unroll 1
  for (int u = tid; u < n_bfly; u += kThreads) {
    const int c = u & ((1 << CL) - 1);
    const int base = (((u >> CL) * R) << CL) + c;
    float2 v[R];
#pragma unroll
    for (int q = 0; q < R; ++q) v[q] = buf[swz(base + (q << CL))];
    if constexpr (R == 2) r2<INV>(v[0], v[1]);
    if constexpr (R == 4) r4<INV>(v[0], v[1], v[2], v[3]);
    if constexpr (R == 8) dft8<INV>(v);
    if constexpr (R == 16) dft16<INV>(v);
#pragma unroll
    for (int q = 0; q < R; ++q) buf[swz(base + (q << CL))] = v[q];
  }
}

// The column transform of a tile with M1 = 2^Q1 rows (cols = 2^(14 - Q1)) as three steps with a
// barrier between them (a step may be empty).  Forward: step 0, 1, 2; inverse: step 2, 1, 0 with
// INV = true.  Output of the forward transform: row position p1 (tile index >> CL) holds frequency
// col_freq_of_pos(Q1, p1).
template <int Q1, int STEP, bool INV>
CORR_HD void col_step(float2* buf, const Tables& t, int tid) {
  constexpr int CL = 14 - Q1;
  if constexpr (STEP == 0) {
    if constexpr (INV) dit16_pass_cols<10, CL>(buf, t, tid);
    else dif16_pass_cols<10, CL>(buf, t, tid);
  } else if constexpr (STEP == 1) {
    if constexpr (Q1 >= 8) {
      if constexpr (INV) dit16_pass_cols<6, (CL < 6 ? CL : 6)>(buf, t, tid);
      else dif16_pass_cols<6, (CL < 6 ? CL : 6)>(buf, t, tid);
    } else {
      last_pass_cols<INV, (1 << (Q1 - 4)), CL>(buf, tid);   // 64 rows: radix 4, 128 rows: radix 8
    }
  } else {
    if constexpr (Q1 > 8) last_pass_cols<INV, (1 << (Q1 - 8)), CL>(buf, tid);
  }
}
template <int Q1>
CORR_HD void col_forward(float2* buf, const Tables& t, int tid) {
  col_step<Q1, 0, false>(buf, t, tid);
  CORR_SYNC();
  col_step<Q1, 1, false>(buf, t, tid);
  CORR_SYNC();
  col_step<Q1, 2, false>(buf, t, tid);
}
template <int Q1>
CORR_HD void col_inverse(float2* buf, const Tables& t, int tid) {
  col_step<Q1, 2, true>(buf, t, tid);
  CORR_SYNC();
  col_step<Q1, 1, true>(buf, t, tid);
  CORR_SYNC();
  col_step<Q1, 0, true>(buf, t, tid);
}

// frequency held at row position p1 after col_forward (digits of the passes, first pass = lowest)
CORR_HD int col_freq_of_pos(int q1, int p1) {
  const int d0 = p1 >> (q1 - 4);
  if (q1 >= 8) {
    const int d1 = (p1 >> (q1 - 8)) & 15;
    const int d2 = p1 & ((1 << (q1 - 8)) - 1);
    return d0 | (d1 << 4) | (d2 << 8);
  }
  const int d2 = p1 & ((1 << (q1 - 4)) - 1);
  return d0 | (d2 << 4);
}
CORR_HD int col_pos_of_freq(int q1, int f) {
  const int d0 = f & 15;
  if (q1 >= 8) {
    const int d1 = (f >> 4) & 15, d2 = f >> 8;
    return (d0 << (q1 - 4)) | (d1 << (q1 - 8)) | d2;
  }
  return (d0 << (q1 - 4)) | (f >> 4);
}

// ---- rows: 16 contiguous 1024-point transforms per tile (corr.cuh's passes 2..4) -----------------
template <int STEP, bool INV>
CORR_HD void row_step(float2* buf, const Tables& t, int tid) {
  if constexpr (STEP == 0) {
    if constexpr (INV) dit16_pass_smem<6>(buf, t, tid);
    else dif16_pass_smem<6>(buf, t, tid);
  } else if constexpr (STEP == 1) {
    if constexpr (INV) dit16_pass_smem<2>(buf, t, tid);
    else dif16_pass_smem<2>(buf, t, tid);
  } else {
    r4_pass_smem<INV>(buf, tid);
  }
}
CORR_HD void rows_forward(float2* buf, const Tables& t, int tid) {
  row_step<0, false>(buf, t, tid);
  CORR_SYNC();
  row_step<1, false>(buf, t, tid);
  CORR_SYNC();
  row_step<2, false>(buf, t, tid);
}
CORR_HD void rows_inverse(float2* buf, const Tables& t, int tid) {
  row_step<2, true>(buf, t, tid);
  CORR_SYNC();
  row_step<1, true>(buf, t, tid);
  CORR_SYNC();
  row_step<0, true>(buf, t, tid);
}
// position inside a row <-> frequency k2 (radices 16, 16, 4)
CORR_HD int row_freq_of_pos(int p) { return ((p >> 6) & 15) | (((p >> 2) & 15) << 4) | ((p & 3) << 8); }
CORR_HD int row_pos_of_freq(int f) { return ((f & 15) << 6) | (((f >> 4) & 15) << 2) | (f >> 8); }

// ---- tables -------------------------------------------------------------------------------------
// half_pos[p] = exp(-i pi k2 / 1024), k2 = row_freq_of_pos(p): the k2 part of the untangle twiddle
//   w_N^k, stored in POSITION order (consecutive lanes read consecutive entries; in frequency order
//   the digit-reversed k2 of 32 consecutive positions all fall into one bank);
// coarse[skew(t)] = exp(-2 pi i t / 1024), fine[skew(t)] = exp(-2 pi i t / M), t < M1: four-step
//   twiddle w_M^x = coarse[x >> q1] * fine[x & (M1 - 1)].  x = n2 * k1 runs over a warp with stride k1,
//   often a multiple of a power of two: the skew t + t/16 + t/256 spreads such strides over the banks.
CORR_HD int skew(int t) { return t + (t >> 4) + (t >> 8); }
constexpr int kSkew1024 = 1024 + 64 + 4;
struct BigTables {
  const float2* tw1024;
  const float2* fine32;
  const float2* half_pos;
  const float2* coarse;
  const float2* fine;
};
CORR_HD int row_freq_of_pos(int p);
CORR_HD void init_big_tables(float2* half_pos, float2* coarse, float2* fine, int q1, int tid) {
  for (int t = tid; t < 1024; t += kThreads) {
    float s, c;
    sincospif(-(float)row_freq_of_pos(t) * (1.0f / 1024.0f), &s, &c);
    half_pos[t] = make_float2(c, s);
    sincospif(-(float)t * (1.0f / 512.0f), &s, &c);
    coarse[skew(t)] = make_float2(c, s);
  }
  const int m1 = 1 << q1;
  const float inv_half_m = 2.0f / (float)(m1 << 10);   // exp(-2 pi i t / M) = sincospi(-2 t / M)
  for (int t = tid; t < m1; t += kThreads) {
    float s, c;
    sincospif(-(float)t * inv_half_m, &s, &c);
    fine[skew(t)] = make_float2(c, s);
  }
}
// w_M^x, x < M = 2^(q1 + 10)
CORR_HD float2 step_twiddle(const BigTables& bt, int q1, int x) {
  return cmul(bt.coarse[skew(x >> q1)], bt.fine[skew(x & ((1 << q1) - 1))]);
}

// ---- tile geometry --------------------------------------------------------------------------------
// F2 tile g of a transform with M1 rows: slot s < 8 holds row 8 g + s, slot 8 + s its partner row
// M1 - (8 g + s); in tile 0 the self-paired rows 0 and M1/2 share the pair (slot 0, slot 8).
CORR_HD int f2_row_of_slot(int q1, int g, int slot) {
  const int m1 = 1 << q1;
  const int k1 = 8 * g + (slot & 7);
  if (slot < 8) return k1;
  return k1 == 0 ? (m1 >> 1) : m1 - k1;
}

// Untangle of one pair of bins (k, M - k): zp = Z[k], zq = Z[M - k], w = w_N^k  ->  2 A[k], 2 A[M - k]
CORR_HD void untangle_bins(float2 zp, float2 zq, float2 w, float2& hp, float2& hq) {
  const float2 e = make_float2(zp.x + zq.x, zp.y - zq.y);
  const float2 d = make_float2(zp.x - zq.x, zp.y + zq.y);
  const float2 tt = cmul(w, d);
  hp = make_float2(e.x + tt.y, e.y - tt.x);
  hq = make_float2(e.x - tt.y, -e.y - tt.x);
}
// Inverse packing of the product spectrum: cp = C[k], cq = C[M - k]  ->  Zc[k], Zc[M - k]
CORR_HD void retangle_bins(float2 cp, float2 cq, float2 w, float2& zp, float2& zq) {
  const float2 e = make_float2(cp.x + cq.x, cp.y - cq.y);
  const float2 d = make_float2(cp.x - cq.x, cp.y + cq.y);
  const float2 tt = cmul(cconj(w), d);
  zp = make_float2(e.x - tt.y, e.y + tt.x);
  zq = make_float2(e.x + tt.y, -e.y + tt.x);
}

// The pairs of an F2 tile, enumerated so that every bin is visited exactly once: pair slot
// e = s * 1024 + p, s < 8, p < 1024 (16 slots per thread) plus ONE extra visit in tile 0:
//   regular slot pairs (s, 8 + s): element (s, p) <-> (8 + s, 1023 - p)
//   tile 0, s = 0:   row 0     : k2 <-> 1024 - k2 for 0 < k2 < 512 (p = k2), k2 = 0 = DC + Nyquist of
//                                the whole transform (p = 0); its own partner k2 = 512 is the extra visit
//                    row M1/2  : k2 <-> 1023 - k2 for k2 < 512 (p = 512 + k2)
// ea, eb: tile element indices (slot * 1024 + position, NOT swizzled) of the two bins (ea == eb for the
// self-paired bin); w = w_N^k of the first; kind 0 = regular pair, 1 = DC/Nyquist (ea only),
// 2 = self-paired bin.
struct PairGeo {
  int ea, eb, kind;
  float2 w;
};
CORR_HD PairGeo f2_pair(const BigTables& bt, int g, const float2* row_tw, int e) {
  const int s = e >> 10, p = e & 1023;
  PairGeo r;
  r.kind = 0;
  if (g == 0 && s == 0) {
    const int k2 = p & 511;
    if (p >= 512) {
      const int pa = row_pos_of_freq(k2);
      r.ea = 8 * 1024 + pa;
      r.eb = 8 * 1024 + 1023 - pa;
      r.w = cmul(row_tw[8], bt.half_pos[pa]);
    } else if (k2 == 0) {
      r.ea = r.eb = row_pos_of_freq(0);
      r.w = make_float2(1.f, 0.f);
      r.kind = 1;
    } else {
      r.ea = row_pos_of_freq(k2);
      r.eb = row_pos_of_freq(1024 - k2);
      r.w = cmul(row_tw[0], bt.half_pos[r.ea]);
    }
  } else {
    r.ea = s * 1024 + p;
    r.eb = (8 + s) * 1024 + 1023 - p;
    r.w = cmul(row_tw[s], bt.half_pos[p]);
  }
  return r;
}
CORR_HD PairGeo f2_extra_pair(const BigTables& bt, const float2* row_tw) {   // tile 0 only: bin k = M/2
  PairGeo r;
  r.ea = r.eb = row_pos_of_freq(512);
  r.w = cmul(row_tw[0], bt.half_pos[r.ea]);
  r.kind = 2;
  return r;
}
constexpr int kPairSlotsPerThread = 8 * 1024 / kThreads;   // 16

// scale of the real outputs: unnormalised transforms (M) and factors 2 (A), 2 (B), 2 (retangle)
CORR_HD float out_scale(int q1) { return 1.0f / (8.0f * (float)(1 << (q1 + 10))); }


// ---- per-thread phases of the three kernels (also driven by tests/host_emul/bigfft_emul.cu) ------

// A real signal as the transforms see it: value(t) = 2 x[t] - 1 for t < len (x from a float array
// or, bit-mask subtitles, hi / -1 per bit), 0 beyond (zero padding up to N).
struct BigSource {
  const float* f;
  const uint32_t* bits;
  int len;
  float hi;
};
// F1 load: tile cg = columns [cg * cols, (cg + 1) * cols) of the [M1][1024] array z[n] = x[2n] + i x[2n+1].
// Returns the thread's partial sum of squares.  Global loads are issued in batches of kLoadBatch per
// thread with clamped (always valid) addresses and masked afterwards, so that the loads of a batch are
// in flight together (one 512-thread CTA per SM covers the latency with loads in flight, not with
// warps); batches that lie entirely in the zero padding issue no loads.
constexpr int kLoadBatch = 16;
CORR_HD float f1_load(float2* buf, const BigSource& src, int q1, int cg, int tid) {
  const int cl = 14 - q1;
  const int last = src.len - 1;   // len >= 1 (empty signals never reach the transforms)
  // sample index of element i of this thread: t(i) = t_base + i * dt (see the note at f1_store)
  const int t_base = 2 * (((tid >> cl) << 10) + (cg << cl) + (tid & ((1 << cl) - 1)));
  const int dt = 2 * ((kThreads >> cl) << 10);
  const int s0 = swz(tid);
  float ss = 0.f;
#pragma unroll 1
  for (int i0 = 0; i0 < kM / kThreads; i0 += kLoadBatch) {
    const int t_first = t_base + i0 * dt;
    if (t_first >= src.len) {   // t grows with i: the rest of the tile is zero padding
#pragma unroll
      for (int i = 0; i < kLoadBatch; ++i) buf[s0 + (i0 + i) * kThreads] = make_float2(0.f, 0.f);
      continue;
    }
    float2 v[kLoadBatch];
    if (src.bits) {
      uint32_t w[kLoadBatch];
#pragma unroll
      for (int i = 0; i < kLoadBatch; ++i) {
        const int t0 = t_first + i * dt;
        w[i] = CORR_LDG(src.bits + ((t0 < last ? t0 : last) >> 5));
      }
#pragma unroll
      for (int i = 0; i < kLoadBatch; ++i) {
        const int t0 = t_first + i * dt;
        const uint32_t m = w[i] >> (t0 & 31);   // t0 even: both bits in one word
        v[i].x = t0 < src.len ? ((m & 1u) ? src.hi : -1.f) : 0.f;
        v[i].y = t0 + 1 < src.len ? ((m & 2u) ? src.hi : -1.f) : 0.f;
      }
    } else {
      float a[kLoadBatch], b[kLoadBatch];
#pragma unroll
      for (int i = 0; i < kLoadBatch; ++i) {
        const int t0 = t_first + i * dt;
        a[i] = CORR_LDG(src.f + (t0 < last ? t0 : last));
        b[i] = CORR_LDG(src.f + (t0 + 1 < last ? t0 + 1 : last));
      }
#pragma unroll
      for (int i = 0; i < kLoadBatch; ++i) {
        const int t0 = t_first + i * dt;
        v[i].x = t0 < src.len ? 2.f * a[i] - 1.f : 0.f;
        v[i].y = t0 + 1 < src.len ? 2.f * b[i] - 1.f : 0.f;
      }
    }
#pragma unroll
    for (int i = 0; i < kLoadBatch; ++i) {
      ss += v[i].x * v[i].x + v[i].y * v[i].y;
      buf[s0 + (i0 + i) * kThreads] = v[i];
    }
  }
  return ss;
}
// F1 store: four-step twiddle w_M^(n2 k1), rows in natural k1 order.
// (Element e = tid + 512 i of a thread: swz(e) = swz(tid) + 512 i - the swizzle reads bits 4..7 and
// writes bits 0..3 only - its column e & (cols - 1) is constant and its row e >> cl advances by
// 512 >> cl per step: the loops below carry these instead of recomputing them.)
CORR_HD void f1_store(const float2* buf, const BigTables& bt, int q1, int cg, int tid, float2* g) {
  const int cl = 14 - q1, mmask = (1 << (q1 + 10)) - 1;
  const int n2 = (cg << cl) + (tid & ((1 << cl) - 1)), dp = kThreads >> cl;
  int p1 = tid >> cl, idx = swz(tid);
#pragma unroll 4
  for (int i = 0; i < kM / kThreads; ++i, p1 += dp, idx += kThreads) {
    const int k1 = col_freq_of_pos(q1, p1);
    g[((size_t)k1 << 10) + n2] = cmul(buf[idx], step_twiddle(bt, q1, (n2 * k1) & mmask));
  }
}

// F2 load / store of the 16 rows of tile g (slot order, see f2_row_of_slot).
CORR_HD void f2_load(float2* buf, int q1, int g, int tid, const float2* src) {
  const int s0 = swz(tid);
#pragma unroll 1
  for (int slot0 = 0; slot0 < 16; slot0 += kLoadBatch / 2) {
    float2 v[kLoadBatch];
#pragma unroll
    for (int i = 0; i < kLoadBatch / 2; ++i) {
      const float2* row = src + ((size_t)f2_row_of_slot(q1, g, slot0 + i) << 10);
      v[2 * i] = CORR_LDG(row + tid);
      v[2 * i + 1] = CORR_LDG(row + tid + kThreads);
    }
#pragma unroll
    for (int i = 0; i < kLoadBatch / 2; ++i) {
      buf[s0 + (slot0 + i) * 1024] = v[2 * i];
      buf[s0 + (slot0 + i) * 1024 + kThreads] = v[2 * i + 1];
    }
  }
}
CORR_HD void f2_store(const float2* buf, int q1, int g, int tid, float2* dst) {
  const int s0 = swz(tid);
#pragma unroll 4
  for (int slot = 0; slot < 16; ++slot) {
    float2* row = dst + ((size_t)f2_row_of_slot(q1, g, slot) << 10);
    row[tid] = buf[s0 + slot * 1024];
    row[tid + kThreads] = buf[s0 + slot * 1024 + kThreads];
  }
}
// row_tw[slot] = exp(-i pi k1 / M): threads 0..15
CORR_HD void f2_row_twiddles(float2* row_tw, int q1, int g, int tid) {
  if (tid < 16) {
    float sn, cs;
    sincospif(-(float)f2_row_of_slot(q1, g, tid) / (float)(1 << (q1 + 10)), &sn, &cs);
    row_tw[tid] = make_float2(cs, sn);
  }
}
// Reference: rows (already transformed, position order) -> packed real spectrum 2 B[k], in place.
CORR_HD void untangle_visit(float2* buf, const PairGeo& pg) {
  const int ia = swz(pg.ea), ib = swz(pg.eb);
  const float2 zp = buf[ia], zq = buf[ib];
  if (pg.kind == 1) {
    buf[ia] = make_float2(2.f * (zp.x + zp.y), 2.f * (zp.x - zp.y));   // DC, Nyquist
    return;
  }
  float2 hp, hq;
  untangle_bins(zp, zq, pg.w, hp, hq);
  buf[ia] = hp;
  if (pg.kind == 0) buf[ib] = hq;
}
CORR_HD void f2_untangle_inplace(float2* buf, const BigTables& bt, int q1, int g, const float2* row_tw, int tid) {
#pragma unroll 1
  for (int it = 0; it < kPairSlotsPerThread; ++it) untangle_visit(buf, f2_pair(bt, g, row_tw, tid + it * kThreads));
  if (g == 0 && tid == 0) untangle_visit(buf, f2_extra_pair(bt, row_tw));
}
// Subtitles: rows (transformed) -> conj(2 A) * (2 B) -> packed for the inverse transform, in place.
// spec = the stored reference spectrum of the same transform geometry (tile layout [k1][position]).
CORR_HD size_t spec_index(int q1, int g, int e) { return ((size_t)f2_row_of_slot(q1, g, e >> 10) << 10) + (e & 1023); }
CORR_HD void product_visit(float2* buf, const PairGeo& pg, float2 bp, float2 bq) {
  const int ia = swz(pg.ea), ib = swz(pg.eb);
  const float2 zp = buf[ia], zq = buf[ib];
  if (pg.kind == 1) {
    const float c0 = 2.f * (zp.x + zp.y) * bp.x, cm = 2.f * (zp.x - zp.y) * bp.y;
    buf[ia] = make_float2(c0 + cm, c0 - cm);
    return;
  }
  float2 hp, hq;
  untangle_bins(zp, zq, pg.w, hp, hq);
  const float2 cp = cmul_conj_a(hp, bp);
  const float2 cq = pg.kind == 0 ? cmul_conj_a(hq, bq) : cp;
  float2 np, nq;
  retangle_bins(cp, cq, pg.w, np, nq);
  buf[ia] = np;
  if (pg.kind == 0) buf[ib] = nq;
}
// Regular slot pair (s, 8 + s) of a thread: its two pairs are (s, p) <-> (8 + s, 1023 - p) for p = tid and
// p = tid + 512.  Everything that depends on the thread only (swizzled offsets inside a row, the k2
// part of the twiddle) is computed once per call, everything that depends on the slot (row base
// pointers, the k1 part of the twiddle) once per slot.
struct PairLanes {
  int ia0, ib0, ia1, ib1;   // swizzled in-row indices of the two pairs (add slot * 1024 / (8 + slot) * 1024)
  float2 h0, h1;            // half_pos[tid], half_pos[tid + 512]
};
CORR_HD PairLanes pair_lanes(const BigTables& bt, int tid) {
  PairLanes l;
  l.ia0 = swz(tid);
  l.ib0 = swz(1023 - tid);
  l.ia1 = swz(tid + kThreads);
  l.ib1 = swz(1023 - tid - kThreads);
  l.h0 = bt.half_pos[tid];
  l.h1 = bt.half_pos[tid + kThreads];
  return l;
}
CORR_HD void product_regular(float2* buf, int ia, int ib, float2 w, float2 bp, float2 bq) {
  const float2 zp = buf[ia], zq = buf[ib];
  float2 hp, hq;
  untangle_bins(zp, zq, w, hp, hq);
  float2 np, nq;
  retangle_bins(cmul_conj_a(hp, bp), cmul_conj_a(hq, bq), w, np, nq);
  buf[ia] = np;
  buf[ib] = nq;
}
CORR_HD void f2_product_inplace(float2* buf, const BigTables& bt, int q1, int g, const float2* row_tw, int tid,
                                const float2* spec) {
  const PairLanes l = pair_lanes(bt, tid);
  constexpr int kS = 4;   // slot pairs per batch: 16 spectrum loads in flight per thread
#pragma unroll 1
  for (int s0 = 0; s0 < 8; s0 += kS) {
    float2 bp0[kS], bq0[kS], bp1[kS], bq1[kS];
#pragma unroll
    for (int i = 0; i < kS; ++i) {
      const int s = s0 + i;
      if (g == 0 && s == 0) continue;   // the two self-paired rows: general path below
      const float2* ra = spec + ((size_t)f2_row_of_slot(q1, g, s) << 10);
      const float2* rb = spec + ((size_t)f2_row_of_slot(q1, g, 8 + s) << 10);
      bp0[i] = CORR_LDG(ra + tid);
      bq0[i] = CORR_LDG(rb + 1023 - tid);
      bp1[i] = CORR_LDG(ra + tid + kThreads);
      bq1[i] = CORR_LDG(rb + 1023 - tid - kThreads);
    }
#pragma unroll
    for (int i = 0; i < kS; ++i) {
      const int s = s0 + i;
      if (g == 0 && s == 0) continue;
      const float2 rt = row_tw[s];
      product_regular(buf, s * 1024 + l.ia0, (8 + s) * 1024 + l.ib0, cmul(rt, l.h0), bp0[i], bq0[i]);
      product_regular(buf, s * 1024 + l.ia1, (8 + s) * 1024 + l.ib1, cmul(rt, l.h1), bp1[i], bq1[i]);
    }
  }
  if (g == 0) {   // tile 0: rows 0 and M1/2 pair with themselves (slots 0 and 8), plus the bin k = M/2
    for (int h = 0; h < 2; ++h) {
      const PairGeo pg = f2_pair(bt, g, row_tw, tid + h * kThreads);
      product_visit(buf, pg, CORR_LDG(spec + spec_index(q1, g, pg.ea)), CORR_LDG(spec + spec_index(q1, g, pg.eb)));
    }
    if (tid == 0) {
      const PairGeo pg = f2_extra_pair(bt, row_tw);
      const float2 b = CORR_LDG(spec + spec_index(q1, g, pg.ea));
      product_visit(buf, pg, b, b);
    }
  }
}
// F2 (subtitles) store after the inverse row transform: conjugate four-step twiddle.
CORR_HD void f2_store_twiddled(const float2* buf, const BigTables& bt, int q1, int g, int tid, float2* dst) {
  const int mmask = (1 << (q1 + 10)) - 1;
  const int s0 = swz(tid);
#pragma unroll 2
  for (int slot = 0; slot < 16; ++slot) {
    const int k1 = f2_row_of_slot(q1, g, slot);
    float2* row = dst + ((size_t)k1 << 10);
    const int x0 = (tid * k1) & mmask, x1 = (x0 + kThreads * k1) & mmask;   // n2 = tid, tid + 512
    row[tid] = cmul(buf[s0 + slot * 1024], cconj(step_twiddle(bt, q1, x0)));
    row[tid + kThreads] = cmul(buf[s0 + slot * 1024 + kThreads], cconj(step_twiddle(bt, q1, x1)));
  }
}

// F3 load: columns of the k1-major array into row positions (digit-reversed for the inverse passes).
CORR_HD void f3_load(float2* buf, int q1, int cg, int tid, const float2* g) {
  const int cl = 14 - q1, dp = kThreads >> cl;
  const float2* col = g + (cg << cl) + (tid & ((1 << cl) - 1));
  const int s0 = swz(tid), p0 = tid >> cl;
#pragma unroll 1
  for (int i0 = 0; i0 < kM / kThreads; i0 += kLoadBatch) {
    float2 v[kLoadBatch];
#pragma unroll
    for (int i = 0; i < kLoadBatch; ++i)
      v[i] = CORR_LDG(col + ((size_t)col_freq_of_pos(q1, p0 + (i0 + i) * dp) << 10));
#pragma unroll
    for (int i = 0; i < kLoadBatch; ++i) buf[s0 + (i0 + i) * kThreads] = v[i];
  }
}
// F3 store: c[2n], c[2n+1] = re, im of z[n]; score index m = (lag + S) mod N (offset o = m - S).
// mx: maximum over the surviving window [m_lo, m_hi]; cn: sum of squares of everything written.
CORR_HD void f3_store(const float2* buf, int q1, int cg, int tid, float* scores, int S, int m_lo, int m_hi,
                      float& mx, float& cn) {
  const int cl = 14 - q1;
  const int nmask = (2 << (q1 + 10)) - 1;
  const float sc = out_scale(q1);
  const int n0 = ((tid >> cl) << 10) + (cg << cl) + (tid & ((1 << cl) - 1));
  const int dm = 2 * ((kThreads >> cl) << 10);   // score-index step between consecutive elements of a thread
  int m = (2 * n0 + S) & nmask, idx = swz(tid);
#pragma unroll 4
  for (int i = 0; i < kM / kThreads; ++i, m = (m + dm) & nmask, idx += kThreads) {
    const float2 z = buf[idx];
    const float v0 = z.x * sc, v1 = z.y * sc;
    const int m1 = (m + 1) & nmask;
    scores[m] = v0;
    scores[m1] = v1;
    cn += v0 * v0 + v1 * v1;
    if (m >= m_lo && m <= m_hi) mx = fmaxf(mx, v0);
    if (m1 >= m_lo && m1 <= m_hi) mx = fmaxf(mx, v1);
  }
}

}  // namespace bigfft

// Types and constants shared by the two correlation paths (corr.cu: overlap-save windows;
// bigfft.cu: one large FFT per signal) and their common tail (exact re-score, pick).
#pragma once
#include <stdint.h>

constexpr int kCandMax = 32;
constexpr int kRescoreSeg = 16;
// Nomination threshold tau ("Round-off bound"): every offset whose fp32 score
// is within tau of the fp32 maximum is re-scored exactly, with
//     tau = u * (kTauFwd * sqrt(Es*Er) + (kTauInv + n_split - 1) * ||c||_2),   u = 2^-24,
// a first-order WORST-CASE bound on |fp32 score - exact score| (all rounding errors aligned):
//   kTauFwd: both forward transforms (first pass 95u: the twiddle w^k of the depth-4 product chain
//            carries k <= 15 times the 5.7u error of the two-table base twiddle; passes 2-4
//            32u + 22u + 3u; untangle 16u) = 2 x 168u, spectral product 3u, accumulation over
//            <= 64 blocks in fp32 <= 63u (more blocks: see window_max_kernel), retangle 16u -> <= 418u,
//            rounded up to 512 for the second-order terms;
//   kTauInv: the inverse transform is backward stable in the 2-norm, |error[m]| <= eps_inv*||c||_2
//            with eps_inv = 16u + 3u + 22u + 32u + 95u = 168u -> 192; ||c||_2 is the norm of the
//            tile's whole inverse-transform output, computed by the kernel (it exceeds sqrt(Es*Er)
//            only for signals with a large mean, whose correlation is a broad ramp);
//   n_split - 1: fp32 addition of the partial score arrays of a split job.
// Measured on random, constant, periodic, sparse and wide-dynamic-range inputs the error stays
// below tau / 400 (at most 26 u sqrt(Es*Er), tests/test_host_cpu.py::test_roundoff_bound_*); the slack
// costs nothing on real data, where the runner-up is thousands of units below the peak.
constexpr float kU = 5.9604645e-8f;
constexpr float kTauFwd = 512.0f;
constexpr float kTauInv = 192.0f;
constexpr int kTauBlocks = 64;  // block count covered by kTauFwd

struct SelJob {        // one (pair, ratio)
  long long ref_off, sub_off, score_off;
  int R, S, o_first;   // offset of scores[score_off]
  int m_lo, m_hi;      // valid window of m (inclusive); m_lo > m_hi: nothing survives
  int energy_slot, n_tiles;
  int n_split;         // partial score arrays per tile (small batches split the block range over CTAs)
  int out_index;       // b*K + k
  int kind;            // 0 normal, 1 empty input, 2 everything masked
  int masked_offset;   // offset reported when kind == 2
  int no_prune;        // 1: some surviving offset of this pair exceeds max_offset_samples in magnitude (the
                       // negative-slice corners of aligners.py:31-43), so MaxScoreAligner.transform's
                       // |offset| filter (:160) may drop a ratio's winner - winner-only pruning is off
  long long bits_off;  // >= 0: the subtitle signal is a bit mask (one bit per frame)
  float sub_level;     // value of a frame inside a cue, min(1/ratio, 1) as float32 (bit-mask mode)
};


// The large-window path takes over from kBigMinTiles overlap-save tiles on; padded lengths it handles.
constexpr int kBigMinTiles = 4;
int bigfft_min_log2n();
int bigfft_max_log2n();

// Common tail (corr.cu): exact float64 re-score of the nominated candidates and the argmax.
// cand_off / cand_cnt / work_list / work_count / job_stat as filled by either path's selection.
struct B2CandBuffers {
  double* cand_partial;
  float2* job_stat;
  int* cand_off;
  int* cand_cnt;
  int* work_list;
  int* work_count;
};
int b2i_rescore_pick(b2_ctx* h, const SelJob* d_sel, size_t J, const float* d_ref, const float* d_sub,
                     const uint32_t* d_bits, const B2CandBuffers& cb, double* d_score, int32_t* d_offset,
                     int32_t* d_status);
// Large-window path (bigfft.cu).  sel: host copy of the jobs (kind / R / S / offsets filled in by the
// planner; this call sets o_first, m_lo, m_hi, score_off), surviving index range per job in idx_lo /
// idx_hi (half open, in the reference's conv[] index space), padded lengths n_pad per pair.
int b2i_align_big(b2_ctx* h, const float* d_ref, const float* d_sub, const uint32_t* d_bits, int B, int K,
                  std::vector<SelJob>& sel, const std::vector<long long>& idx_lo,
                  const std::vector<long long>& idx_hi, const std::vector<long long>& n_pad, int winner_only,
                  const B2CandBuffers& cb, const SelJob** d_sel_out);

// CPU check of the lane-per-window VAD arithmetic (ffsubsync_b200/csrc/vad_lane.cuh) against the plain
// definition: E = sum x^2, Z = sign changes inside the window.  Test infrastructure (the build
// container has no GPU); prints "ok <cases>" or the first mismatches.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../ffsubsync_b200/csrc/vad_lane.cuh"

using namespace vadlane;

static uint32_t rng_state = 12345u;
static uint32_t rnd() {
  rng_state ^= rng_state << 13;
  rng_state ^= rng_state >> 17;
  rng_state ^= rng_state << 5;
  return rng_state;
}

template <int C>
static int run(int n_cases, long long* total) {
  constexpr int RMAX = rotation_max(C);
  constexpr int n = 8 * C;
  int bad = 0;
  // tile of 32 windows like the kernel sees it (+ one chunk of slack either side)
  std::vector<unsigned char> raw(16 * C * 32 + 64);
  unsigned char* tile = raw.data() + (16 - ((uintptr_t)raw.data() & 15));
  for (int it = 0; it < n_cases; ++it) {
    short* x = reinterpret_cast<short*>(tile);
    const int mode = it % 7;
    for (int i = 0; i < n * 32; ++i) {
      short v;
      switch (mode) {
        case 0: v = (short)rnd(); break;                                  // full range
        case 1: v = (short)((rnd() % 7) - 3); break;                      // around zero: many crossings, zeros
        case 2: v = (i & 1) ? 32767 : -32768; break;                      // extremes, crossing every sample
        case 3: v = -32768; break;                                        // largest energy
        case 4: v = (short)(((i / 3) & 1) ? -(int)(rnd() % 200) : (int)(rnd() % 200)); break;
        case 5: v = (rnd() & 15) ? 0 : (short)rnd(); break;               // sparse
        default: v = (short)((rnd() & 1) ? 255 : -256); break;            // byte boundaries
      }
      x[i] = v;
    }
    for (int lane = 0; lane < 32; ++lane) {
      const unsigned char* wbase = tile + (size_t)lane * 16 * C;
      const short* xs = reinterpret_cast<const short*>(wbase);
      long long e_ref = 0;
      int z_ref = 0;
      for (int i = 0; i < n; ++i) {
        e_ref += (long long)xs[i] * xs[i];
        if (i > 0) z_ref += ((xs[i] < 0) != (xs[i - 1] < 0));
      }
      for (int r = 0; r <= RMAX; ++r) {   // every start chunk, not only the lane's own
        long long e;
        int z;
        lane_window<C, RMAX>(wbase, r, e, z);
        ++*total;
        if (e != e_ref || z != z_ref) {
          if (bad < 5)
            printf("mismatch C=%d mode=%d lane=%d r=%d: e %lld vs %lld, z %d vs %d\n", C, mode, lane, r, e, e_ref, z, z_ref);
          ++bad;
        }
      }
      if (lane_rotation(C, lane) > RMAX) {
        printf("rotation out of range C=%d lane=%d\n", C, lane);
        ++bad;
      }
    }
    // bank groups: the 8 lanes of every quarter-warp must hit 8 distinct 16-byte bank groups at every step
    for (int q = 0; q < 4; ++q)
      for (int c = 0; c < C; ++c) {
        unsigned seen = 0;
        for (int l = 0; l < 8; ++l) {
          const int lane = 8 * q + l, r = lane_rotation(C, lane);
          const int k = (c + r) % C;
          const int grp = (lane * C + k) & 7;
          seen |= 1u << grp;
        }
        if (seen != 0xffu) {
          if (bad < 5) printf("bank conflict C=%d quarter=%d step=%d mask=%02x\n", C, q, c, seen);
          ++bad;
        }
      }
  }
  return bad;
}

int main(int argc, char** argv) {
  const int n_cases = argc > 1 ? atoi(argv[1]) : 70;
  long long total = 0;
  int bad = 0;
  bad += run<5>(n_cases, &total);     //  4 kHz
  bad += run<10>(n_cases, &total);    //  8 kHz
  bad += run<15>(n_cases, &total);    // 12 kHz
  bad += run<20>(n_cases, &total);    // 16 kHz
  bad += run<30>(n_cases, &total);    // 24 kHz
  bad += run<40>(n_cases, &total);    // 32 kHz
  bad += run<60>(n_cases, &total);    // 48 kHz
  if (bad) {
    printf("FAILED %d\n", bad);
    return 1;
  }
  printf("ok %lld\n", total);
  return 0;
}

// C-ABI entry points (include/ffsubsync_b200.h): handle lifecycle, workspace management and
// the host<->device staging that turns B2_HOST calls into the device path.  All arithmetic
// lives in the kernels (vad.cu, raster.cu, corr.cu, select.cu); nothing here computes results
// on the CPU.
#include <math.h>

#include <chrono>
#include <algorithm>
#include <atomic>
#include <thread>

#include "common.cuh"

// Entry guard: the handle's device is current for the duration of the call and the caller
// thread's previous device is restored on every return path.
struct DeviceScope {
  int prev = -1, want;
  bool ok = true;
  explicit DeviceScope(int dev) : want(dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != want) ok = cudaSetDevice(want) == cudaSuccess;
  }
  ~DeviceScope() {
    if (prev >= 0 && prev != want) cudaSetDevice(prev);
  }
};
extern "C" int b2_version(void) { return 200; }

static uint64_t compute_log2_quirk_mask() {
  // CPython: total_bits = math.log(n, 2) == log(n)/log(2) in double; ceil() of that is k+1 for
  // some exact powers of two (k = 29, 31, 39, ... with glibc).  Reproduced with the same libm.
  uint64_t mask = 0;
  for (int k = 0; k < 63; ++k) {
    double v = log((double)(1ULL << k)) / log(2.0);
    if (ceil(v) > (double)k) mask |= (1ULL << k);
  }
  return mask;
}

extern "C" int b2_create(int device, b2_handle* out) {
  if (!out) return B2_ERR_BAD_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count)
    return B2_ERR_CUDA;
  b2_ctx* h = new b2_ctx();
  h->device = device;
  DeviceScope scope(device);
  if (!scope.ok) { delete h; return B2_ERR_CUDA; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete h; return B2_ERR_CUDA; }
  h->sm_count = prop.multiProcessorCount;
  if (prop.major != 10) {  // built for sm_100a only: fail loudly instead of falling back
    delete h;
    return B2_ERR_UNSUPPORTED;
  }
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete h;
    return B2_ERR_CUDA;
  }
  h->own_stream = true;
  // internal second stream at the highest priority: when b2_sync_batch pipelines sub-batches, the VAD CTAs
  // queued on it take the SMs that the (lower-priority) correlation CTAs of the caller's stream give up
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  if (cudaStreamCreateWithPriority(&h->stream2, cudaStreamNonBlocking, prio_hi) != cudaSuccess) {
    cudaStreamDestroy(h->stream);
    delete h;
    return B2_ERR_CUDA;
  }
  for (auto& e : h->ev_pool)
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) {
      delete h;
      return B2_ERR_CUDA;
    }
  if (cudaEventCreateWithFlags(&h->resident_fence, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->resident_done, cudaEventDisableTiming) != cudaSuccess) {
    delete h;
    return B2_ERR_CUDA;
  }
  h->log2_quirk_mask = compute_log2_quirk_mask();
  const char* acc = getenv("B2_ACC");  // A/B switch for profiling: "reg" keeps the accumulators in registers
  h->acc_in_tmem = !(acc && strcmp(acc, "reg") == 0);
  *out = h;
  return B2_OK;
}

extern "C" int b2_destroy(b2_handle h) {
  if (!h) return B2_OK;
  DeviceScope scope(h->device);
  if (h->stream2) cudaStreamSynchronize(h->stream2);
  cudaStreamSynchronize(h->stream);
  for (auto& w : h->ws)
    if (w.p) cudaFree(w.p);
  for (auto& p : h->pinned)
    if (p.p) cudaFreeHost(p.p);
  for (auto& e : h->pinned_ev)
    if (e) cudaEventDestroy(e);
  for (auto& b : h->bounce) {
    if (b.p) cudaFreeHost(b.p);
    if (b.ev) cudaEventDestroy(b.ev);
  }
  for (auto& sl : h->vs.slot) {
    if (sl.hp) cudaFreeHost(sl.hp);
    if (sl.hout) cudaFreeHost(sl.hout);
    if (sl.dp) cudaFree(sl.dp);
    if (sl.dout) cudaFree(sl.dout);
    if (sl.ev) cudaEventDestroy(sl.ev);
  }
  if (h->stream2) cudaStreamSynchronize(h->stream2);
  for (auto& ring : h->meta)
    for (auto& m : ring) {
      if (m.d) cudaFree(m.d);
      if (m.p) cudaFreeHost(m.p);
      if (m.ev) cudaEventDestroy(m.ev);
    }
  for (auto& e : h->ev_pool)
    if (e) cudaEventDestroy(e);
  if (h->resident_fence) cudaEventDestroy(h->resident_fence);
  if (h->resident_done) cudaEventDestroy(h->resident_done);
  if (h->stream2) cudaStreamDestroy(h->stream2);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return B2_OK;
}

extern "C" int b2_set_stream(b2_handle h, void* s) {
  if (!h) return B2_ERR_BAD_ARG;
  DeviceScope scope(h->device);
  if (h->own_stream && h->stream) {
    cudaStreamSynchronize(h->stream);
    cudaStreamDestroy(h->stream);
  }
  if (s) {
    h->stream = (cudaStream_t)s;
    h->own_stream = false;
  } else {
    B2_CUDA(h, cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->own_stream = true;
  }
  return B2_OK;
}

extern "C" int b2_synchronize(b2_handle h) {
  if (!h) return B2_ERR_BAD_ARG;
  DeviceScope scope(h->device);
  B2_CUDA(h, cudaStreamSynchronize(h->stream));
  return B2_OK;
}

extern "C" const char* b2_last_error(b2_handle h) { return h ? h->err.c_str() : "null handle"; }
extern "C" int64_t b2_launch_count(b2_handle h) { return h ? h->launches : 0; }

// ---- workspaces --------------------------------------------------------------------------
int b2i_ws(b2_ctx* h, int which, size_t bytes, void** out) {
  DeviceBuf& w = h->ws[which];
  if (bytes > w.cap) {
    // stream-ordered reuse: earlier kernels may still read the old block
    B2_CUDA(h, cudaStreamSynchronize(h->stream));
    if (w.p) B2_CUDA(h, cudaFree(w.p));
    w.p = nullptr;
    w.cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&w.p, want);
    if (e != cudaSuccess) {
      cudaGetLastError();
      B2_FAIL(h, B2_ERR_NOMEM, "cudaMalloc(%zu) for workspace %d failed: %s", want, which,
              cudaGetErrorString(e));
    }
    w.cap = want;
  }
  *out = w.p;
  return B2_OK;
}

int b2i_pinned(b2_ctx* h, int which, size_t bytes, void** out) {
  HostBuf& w = h->pinned[which];
  if (bytes > w.cap) {
    B2_CUDA(h, cudaStreamSynchronize(h->stream));
    if (w.p) B2_CUDA(h, cudaFreeHost(w.p));
    w.p = nullptr;
    w.cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    B2_CUDA(h, cudaMallocHost(&w.p, want));
    w.cap = want;
  }
  *out = w.p;
  return B2_OK;
}

// Metadata arena: host-side tables (offsets, per-job parameters) are packed into one pinned
// block and uploaded with a single async copy per call.
// Arenas come from a ring of kMetaSlots (pinned, device) buffer pairs so that the host can plan
// and upload several launches ahead of the GPU without synchronising the stream.  A slot is
// reused kMetaSlots arenas later; the event recorded at the commit of the arena that FOLLOWED its
// previous use guarantees (in-order stream) that both its upload and every kernel that read its
// device copy have finished.
int b2i_meta_begin(b2_ctx* h, MetaArena* a, size_t bytes) {
  a->h = h;
  bytes = (bytes + 255) & ~size_t(255);
  b2_ctx::MetaSlot* ring = h->meta[h->ring];
  uint64_t& seq = h->meta_seq[h->ring];
  const int slot = (int)(seq % b2_ctx::kMetaSlots);
  b2_ctx::MetaSlot& next = ring[(slot + 1) % b2_ctx::kMetaSlots];
  if (seq >= (uint64_t)b2_ctx::kMetaSlots && next.ev) B2_CUDA(h, cudaEventSynchronize(next.ev));
  b2_ctx::MetaSlot& s = ring[slot];
  if (!s.ev) B2_CUDA(h, cudaEventCreateWithFlags(&s.ev, cudaEventDisableTiming));
  if (bytes > s.cap) {
    // grow every slot of this ring at once (pinned allocations cost milliseconds): after the
    // first large call no later arena, whichever slot it lands on, allocates again
    B2_CUDA(h, cudaStreamSynchronize(h->stream));
    const size_t want = bytes + bytes / 4 + 65536;
    for (int i = 0; i < b2_ctx::kMetaSlots; ++i) {
      b2_ctx::MetaSlot& m = ring[i];
      if (m.cap >= want) continue;
      if (m.d) B2_CUDA(h, cudaFree(m.d));
      if (m.p) B2_CUDA(h, cudaFreeHost(m.p));
      m.d = m.p = nullptr;
      m.cap = 0;
      B2_CUDA(h, cudaMalloc(&m.d, want));
      B2_CUDA(h, cudaMallocHost(&m.p, want));
      m.cap = want;
    }
  }
  a->slot = slot;
  a->ring = h->ring;
  a->dbase = (char*)s.d;
  a->hbase = (char*)s.p;
  a->cap = bytes;
  a->used = 0;
  seq++;
  return B2_OK;
}

void* b2i_meta_reserve(MetaArena* a, size_t bytes, void** host_view) {
  size_t at = (a->used + 15) & ~size_t(15);
  if (at + bytes > a->cap) return nullptr;
  a->used = at + bytes;
  if (host_view) *host_view = a->hbase + at;
  return a->dbase + at;
}

void* b2i_meta_put(MetaArena* a, const void* src, size_t bytes) {
  void* hv;
  void* d = b2i_meta_reserve(a, bytes, &hv);
  if (d && bytes) memcpy(hv, src, bytes);
  return d;
}

int b2i_meta_commit(MetaArena* a) {
  b2_ctx* h = a->h;
  if (a->used)
    B2_CUDA(h, cudaMemcpyAsync(a->dbase, a->hbase, a->used, cudaMemcpyHostToDevice, h->stream));
  B2_CUDA(h, cudaEventRecord(h->meta[a->ring][a->slot].ev, h->stream));
  return B2_OK;
}

// Run `body` with the handle temporarily launching on its internal second stream.
struct Stream2Scope {
  b2_ctx* h;
  cudaStream_t saved;
  int saved_ring;
  explicit Stream2Scope(b2_ctx* hh) : h(hh), saved(hh->stream), saved_ring(hh->ring) {
    h->stream = h->stream2;
    h->ring = 1;
  }
  ~Stream2Scope() {
    h->stream = saved;
    h->ring = saved_ring;
  }
};

static cudaEvent_t next_event(b2_ctx* h) {
  cudaEvent_t e = h->ev_pool[h->ev_next];
  h->ev_next = (h->ev_next + 1) % b2_ctx::kEvents;
  return e;
}

// ---- helpers for B2_HOST calls -------------------------------------------------------------
// Large PAGEABLE inputs (a numpy array of PCM: 230 MB per 2 h signal): cudaMemcpyAsync stages them
// through the driver's bounce buffer with one thread at ~11 GB/s (measured: 19.9 ms per 230 MB against
// 4.2 ms from pinned memory, profiles/r2e_latency_single_pair.json).  Here the copy goes through two
// pinned 32 MB buffers filled by kCopyThreads host threads while the previous buffer is on the bus.
static const size_t kBounceBytes = (size_t)32 << 20;
static const int kCopyThreads = 6;

static bool is_pageable(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeUnregistered;
}

// Copy workers live for one staged transfer: worker t copies slice t of every chunk into the bounce
// buffer the main thread announces (`ready`), and reports through `done`; the main thread turns each
// filled buffer into an async H2D copy.  No thread is created per chunk; the only blocking wait is
// on the event of the H2D copy that last read the buffer.
struct CopyCrew {
  std::atomic<int> ready{-1};           // index of the chunk whose bounce buffer may be filled
  std::atomic<int> done{0};             // slices finished, over all chunks
  std::atomic<bool> abort{false};
};

static void copy_slice(char* dst, const char* src, size_t n, int t, int nthreads) {
  const size_t per = ((n / nthreads) + 4095) & ~(size_t)4095;
  const size_t lo = std::min(n, per * (size_t)t), hi = t == nthreads - 1 ? n : std::min(n, per * (size_t)(t + 1));
  if (hi > lo) memcpy(dst + lo, src + lo, hi - lo);
}

static int staged_h2d(b2_ctx* h, void* dev, const void* host, size_t bytes) {
  for (auto& b : h->bounce)
    if (!b.p) {
      B2_CUDA(h, cudaMallocHost(&b.p, kBounceBytes));
      B2_CUDA(h, cudaEventCreateWithFlags(&b.ev, cudaEventDisableTiming));
    }
  const int n_chunks = (int)((bytes + kBounceBytes - 1) / kBounceBytes);
  CopyCrew crew;
  std::thread workers[kCopyThreads];
  int n_workers = 0;   // helpers besides this thread (slice 0 is copied here)
  auto work = [&](int t, int nthreads) {
    for (int c = 0; c < n_chunks; ++c) {
      while (crew.ready.load(std::memory_order_acquire) < c) {
        if (crew.abort.load(std::memory_order_relaxed)) return;
        std::this_thread::yield();
      }
      const size_t off = (size_t)c * kBounceBytes, n = std::min(kBounceBytes, bytes - off);
      copy_slice((char*)h->bounce[c & 1].p, (const char*)host + off, n, t, nthreads);
      crew.done.fetch_add(1, std::memory_order_release);
    }
  };
  try {
    for (; n_workers < kCopyThreads - 1; ++n_workers) workers[n_workers] = std::thread(work, n_workers + 1, kCopyThreads);
  } catch (...) {   // no more threads to be had: nothing may escape through the C ABI
    crew.abort.store(true);
    for (int i = 0; i < n_workers; ++i) workers[i].join();
    n_workers = 0;
    crew.abort.store(false);
  }
  const int nthreads = n_workers == kCopyThreads - 1 ? kCopyThreads : 1;   // all helpers or none
  int status = B2_OK;
  for (int c = 0; c < n_chunks && status == B2_OK; ++c) {
    b2_ctx::Bounce& b = h->bounce[c & 1];
    const size_t off = (size_t)c * kBounceBytes, n = std::min(kBounceBytes, bytes - off);
    if (cudaEventSynchronize(b.ev) != cudaSuccess) {   // the copy that last read this buffer
      status = B2_ERR_CUDA;
      break;
    }
    crew.ready.store(c, std::memory_order_release);
    copy_slice((char*)b.p, (const char*)host + off, n, 0, nthreads);
    while (crew.done.load(std::memory_order_acquire) < (c + 1) * (nthreads - 1)) std::this_thread::yield();
    if (cudaMemcpyAsync((char*)dev + off, b.p, n, cudaMemcpyHostToDevice, h->stream) != cudaSuccess ||
        cudaEventRecord(b.ev, h->stream) != cudaSuccess)
      status = B2_ERR_CUDA;
  }
  if (status != B2_OK) crew.abort.store(true);
  crew.ready.store(n_chunks, std::memory_order_release);
  for (int i = 0; i < n_workers; ++i) workers[i].join();
  if (status != B2_OK) B2_FAIL(h, status, "staged host-to-device copy failed: %s", cudaGetErrorString(cudaGetLastError()));
  return B2_OK;
}

static int stage_in(b2_ctx* h, int which, const void* host, size_t bytes, void** dev) {
  B2_TRY(b2i_ws(h, which, bytes ? bytes : 16, dev));
  if (!bytes) return B2_OK;
  if (bytes >= kBounceBytes / 2 && is_pageable(host)) return staged_h2d(h, *dev, host, bytes);
  B2_CUDA(h, cudaMemcpyAsync(*dev, host, bytes, cudaMemcpyHostToDevice, h->stream));
  return B2_OK;
}

static int copy_out(b2_ctx* h, void* host, const void* dev, size_t bytes) {
  if (bytes)
    B2_CUDA(h, cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, h->stream));
  return B2_OK;
}

#define B2_ENTER(h)                                              \
  if (!(h)) return B2_ERR_BAD_ARG;                               \
  (h)->err.clear();                                              \
  const bool _b2_fence_was_valid = (h)->resident_fence_valid;    \
  (void)_b2_fence_was_valid;                                     \
  (h)->resident_fence_valid = false; /* any entry point breaks a chain of resident b2_sync_batch calls */ \
  DeviceScope _b2_scope((h)->device);                            \
  if (!_b2_scope.ok) return B2_ERR_CUDA;

// B2_DEVICE calls: a bulk pointer must be device (or managed) memory of the handle's device.
static int check_device_ptr(b2_ctx* h, const void* p, const char* what) {
  if (!p) return B2_OK;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    B2_FAIL(h, B2_ERR_BAD_ARG, "%s: not a CUDA pointer", what);
  }
  if (at.type == cudaMemoryTypeManaged) return B2_OK;
  if (at.type != cudaMemoryTypeDevice)
    B2_FAIL(h, B2_ERR_BAD_ARG, "%s: B2_DEVICE call with a host pointer", what);
  if (at.device != h->device)
    B2_FAIL(h, B2_ERR_BAD_ARG, "%s: pointer lives on device %d, the handle on device %d", what, at.device,
            h->device);
  return B2_OK;
}
#define B2_CHECK_DEV(h, memspace, p, what)                                                       \
  do {                                                                                           \
    if ((memspace) != B2_HOST && (memspace) != B2_DEVICE) /* b2_sync_batch maps RESIDENT first */ \
      B2_FAIL(h, B2_ERR_BAD_ARG, "%s: memspace must be B2_HOST or B2_DEVICE", what);             \
    if ((memspace) == B2_DEVICE) B2_TRY(check_device_ptr(h, p, what));                           \
  } while (0)

// ---- VAD -----------------------------------------------------------------------------------
extern "C" int b2_vad_frames_per_window(int frame_rate, int sample_rate) {
  if (frame_rate <= 0 || sample_rate <= 0) return 0;
  // speech_transformers.py:163-164: int(window_duration * frame_rate + 0.5)
  return (int)((1.0 / (double)sample_rate) * (double)frame_rate + 0.5);
}

extern "C" int64_t b2_vad_num_windows(int64_t n_samples, int frame_rate, int sample_rate) {
  int fpw = b2_vad_frames_per_window(frame_rate, sample_rate);
  if (fpw <= 0 || n_samples < 0) return -1;
  return (n_samples + fpw - 1) / fpw;  // len(range(0, n, fpw)), speech_transformers.py:169
}

extern "C" int b2_vad_energy_zcr(b2_handle h, const int16_t* pcm, const int64_t* pcm_off, int B,
                                 int frame_rate, int sample_rate, float non_speech_label,
                                 int64_t energy_threshold, int z_lo, int z_hi, float* out,
                                 const int64_t* out_off, int memspace) {
  B2_ENTER(h);
  if (B < 0 || !pcm_off || !out_off) B2_FAIL(h, B2_ERR_BAD_ARG, "vad: null offset table / B<0");
  int fpw = b2_vad_frames_per_window(frame_rate, sample_rate);
  if (fpw <= 0) B2_FAIL(h, B2_ERR_BAD_ARG, "vad: bad frame_rate/sample_rate");
  if (energy_threshold < 0) B2_FAIL(h, B2_ERR_BAD_ARG, "vad: negative energy threshold");
  if (z_lo < 0) z_lo = 0;
  if (z_hi < 0) z_hi = (3 * fpw) / 8;
  for (int b = 0; b < B; ++b) {
    int64_t n = pcm_off[b + 1] - pcm_off[b];
    if (n < 0) B2_FAIL(h, B2_ERR_BAD_ARG, "vad: pcm_off not monotone at %d", b);
    if (out_off[b + 1] - out_off[b] != (n + fpw - 1) / fpw)
      B2_FAIL(h, B2_ERR_BAD_ARG, "vad: out_off[%d] span must be ceil(n/fpw)", b);
  }
  if (B == 0) return B2_OK;
  int64_t n_total = pcm_off[B], w_total = out_off[B];
  if ((n_total && !pcm) || (w_total && !out)) B2_FAIL(h, B2_ERR_BAD_ARG, "vad: null data pointer");
  B2_CHECK_DEV(h, memspace, pcm, "vad: pcm");
  B2_CHECK_DEV(h, memspace, out, "vad: out");
  const int16_t* d_pcm = pcm;
  float* d_out = out;
  if (memspace == B2_HOST) {
    void *dp, *dq;
    B2_TRY(stage_in(h, b2_ctx::WS_STAGE_IN0, pcm, (size_t)n_total * 2, &dp));
    B2_TRY(b2i_ws(h, b2_ctx::WS_STAGE_OUT, (size_t)w_total * 4 + 16, &dq));
    d_pcm = (const int16_t*)dp;
    d_out = (float*)dq;
  }
  B2_TRY(b2i_vad_launch(h, d_pcm, pcm_off, B, fpw, non_speech_label, (int64_t)fpw * energy_threshold,
                        z_lo, z_hi, d_out, out_off));
  if (memspace == B2_HOST) {
    B2_TRY(copy_out(h, out, d_out, (size_t)w_total * 4));
    B2_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  return B2_OK;
}

// ---- auditok detector (speech_transformers.py:101-152) ---------------------------------------
extern "C" int b2_auditok_block_size(int frame_rate, int sample_rate) {
  if (frame_rate <= 0 || sample_rate <= 0) return 0;
  // ADSFactory.ads(block_dur=1.0/sample_rate): int(sampling_rate * block_dur), speech_transformers.py:140;
  // the output length formula uses frame_rate // sample_rate (:122,143-145): the two must agree
  volatile double dur = 1.0 / (double)sample_rate;
  volatile double prod = (double)frame_rate * dur;
  const int block = (int)prod;
  return block == frame_rate / sample_rate ? block : 0;
}

extern "C" int64_t b2_auditok_energy_floor(int n_samples, double energy_threshold_db) {
  // smallest integer sum of squares E with 10*log10(E/n) >= threshold, evaluated with the same
  // float64 expression auditok's AudioEnergyValidator uses (log energy -200 for E = 0)
  if (n_samples <= 0) return INT64_MAX;
  if (-200.0 >= energy_threshold_db) return 0;
  auto valid = [&](int64_t e) {
    volatile double energy = (double)e / (double)n_samples;
    volatile double le = 10.0 * log10(energy);
    return le >= energy_threshold_db;
  };
  const double guess = (double)n_samples * pow(10.0, energy_threshold_db / 10.0);
  const double e_max = (double)n_samples * 32768.0 * 32768.0;   // int16 blocks cannot exceed this
  if (!(guess <= 2.0 * e_max)) return INT64_MAX;
  int64_t lo = 0, hi = (int64_t)guess + 1;                       // !valid(0) holds: log energy -200
  while (!valid(hi)) {
    if ((double)hi > 4.0 * e_max) return INT64_MAX;
    hi *= 2;
  }
  while (hi - lo > 1) {
    const int64_t mid = lo + (hi - lo) / 2;
    if (valid(mid)) hi = mid;
    else lo = mid;
  }
  return hi;
}

extern "C" int b2_vad_auditok(b2_handle h, const int16_t* pcm, const int64_t* pcm_off, int B,
                              int frame_rate, int sample_rate, double non_speech_label,
                              double energy_threshold_db, double min_length, int64_t max_length,
                              double max_continuous_silence, int64_t chunk_samples, double* out,
                              const int64_t* out_off, int memspace) {
  B2_ENTER(h);
  if (B < 0 || !pcm_off || !out_off) B2_FAIL(h, B2_ERR_BAD_ARG, "auditok: null offset table / B<0");
  const int fpw = b2_auditok_block_size(frame_rate, sample_rate);
  if (fpw <= 0)
    B2_FAIL(h, B2_ERR_UNSUPPORTED, "auditok: int(frame_rate/sample_rate) block size and frame_rate//sample_rate "
                                   "window size differ (or are 0) for %d / %d", frame_rate, sample_rate);
  // StreamTokenizer.__init__ argument checks
  if (max_length <= 0 || !(min_length > 0) || min_length > (double)max_length ||
      !(max_continuous_silence < (double)max_length) || chunk_samples < 0)
    B2_FAIL(h, B2_ERR_BAD_ARG, "auditok: bad tokenizer parameters");
  // one detector call per chunk: chunk c of signal b becomes "signal" n of the batched energy kernel
  std::vector<int64_t> c_pcm(1, 0), c_out(1, 0), tail;
  for (int b = 0; b < B; ++b) {
    const int64_t n = pcm_off[b + 1] - pcm_off[b];
    if (n < 0) B2_FAIL(h, B2_ERR_BAD_ARG, "auditok: pcm_off not monotone at %d", b);
    if (c_pcm.back() != pcm_off[b] - pcm_off[0])
      B2_FAIL(h, B2_ERR_BAD_ARG, "auditok: internal chunk table mismatch");
    int64_t nwin = 0;
    const int64_t step = chunk_samples > 0 ? chunk_samples : (n > 0 ? n : 1);
    for (int64_t s = 0; s < n; s += step) {
      const int64_t len = std::min(step, n - s);
      const int64_t w = (len + fpw - 1) / fpw;
      c_pcm.push_back(c_pcm.back() + len);
      c_out.push_back(c_out.back() + w);
      const int rem = (int)(len % fpw);
      tail.push_back(rem ? b2_auditok_energy_floor(rem, energy_threshold_db) : 0);
      nwin += w;
    }
    if (out_off[b + 1] - out_off[b] != nwin)
      B2_FAIL(h, B2_ERR_BAD_ARG, "auditok: out_off[%d] span must be the sum of ceil(chunk/fpw)", b);
  }
  const int n_chunks = (int)tail.size();
  if (n_chunks == 0) return B2_OK;
  const int64_t n_total = pcm_off[B] - pcm_off[0], w_total = c_out.back();
  if ((n_total && !pcm) || (w_total && !out)) B2_FAIL(h, B2_ERR_BAD_ARG, "auditok: null data pointer");
  B2_CHECK_DEV(h, memspace, pcm, "auditok: pcm");
  B2_CHECK_DEV(h, memspace, out, "auditok: out");
  const int16_t* d_pcm = pcm + pcm_off[0];
  double* d_out = out + out_off[0];
  void *d_flags, *dp, *dq;
  if (memspace == B2_HOST) {
    B2_TRY(stage_in(h, b2_ctx::WS_STAGE_IN0, pcm + pcm_off[0], (size_t)n_total * 2, &dp));
    B2_TRY(b2i_ws(h, b2_ctx::WS_STAGE_OUT, (size_t)w_total * 8 + 16, &dq));
    d_pcm = (const int16_t*)dp;
    d_out = (double*)dq;
  }
  B2_TRY(b2i_ws(h, b2_ctx::WS_SIG_REF, (size_t)w_total * 4 + 64, &d_flags));
  B2_TRY(b2i_vad_launch(h, d_pcm, c_pcm.data(), n_chunks, fpw, 0.0f,
                        b2_auditok_energy_floor(fpw, energy_threshold_db), 0, fpw, (float*)d_flags,
                        c_out.data(), tail.data()));
  B2TokenizerParams tp{min_length, max_continuous_silence, non_speech_label, (long long)max_length};
  B2_TRY(b2i_tokenize_launch(h, (const float*)d_flags, c_out.data(), n_chunks, tp, d_out));
  if (memspace == B2_HOST) {
    B2_TRY(copy_out(h, out + out_off[0], d_out, (size_t)w_total * 8));
    B2_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  return B2_OK;
}

// ---- streaming VAD (chunk loop of speech_transformers.py:710-746) -----------------------------
static int vs_collect(b2_ctx* h, b2_ctx::VadStream::Slot& sl) {
  if (!sl.busy) return B2_OK;
  B2_CUDA(h, cudaEventSynchronize(sl.ev));
  h->vs.results.insert(h->vs.results.end(), sl.hout, sl.hout + sl.n_out);
  sl.busy = false;
  return B2_OK;
}

extern "C" int b2_vad_stream_begin(b2_handle h, int frame_rate, int sample_rate, float non_speech_label,
                                   int64_t energy_threshold, int z_lo, int z_hi) {
  B2_ENTER(h);
  auto& vs = h->vs;
  if (vs.active) B2_FAIL(h, B2_ERR_BAD_ARG, "vad stream: already open on this handle");
  const int fpw = b2_vad_frames_per_window(frame_rate, sample_rate);
  if (fpw <= 0) B2_FAIL(h, B2_ERR_BAD_ARG, "vad stream: bad frame_rate/sample_rate");
  if (energy_threshold < 0) B2_FAIL(h, B2_ERR_BAD_ARG, "vad stream: negative energy threshold");
  vs.frame_rate = frame_rate;
  vs.sample_rate = sample_rate;
  vs.fpw = fpw;
  vs.label = non_speech_label;
  vs.thr = energy_threshold;
  vs.z_lo = z_lo < 0 ? 0 : z_lo;
  vs.z_hi = z_hi < 0 ? (3 * fpw) / 8 : z_hi;
  vs.windows = 0;
  vs.seq = 0;
  vs.results.clear();
  vs.active = true;
  return B2_OK;
}

extern "C" int b2_vad_stream_push(b2_handle h, const void* pcm_bytes, int64_t n_bytes) {
  B2_ENTER(h);
  auto& vs = h->vs;
  if (!vs.active) B2_FAIL(h, B2_ERR_BAD_ARG, "vad stream: not open");
  if (n_bytes < 0 || (n_bytes && !pcm_bytes)) B2_FAIL(h, B2_ERR_BAD_ARG, "vad stream: bad chunk");
  const int64_t n = n_bytes / 2;  // an odd trailing byte is ignored (speech_transformers.py:745)
  if (n == 0) return B2_OK;
  const int64_t nwin = (n + vs.fpw - 1) / vs.fpw;
  auto& sl = vs.slot[vs.seq % b2_ctx::VadStream::kSlots];
  B2_TRY(vs_collect(h, sl));  // the slot's previous chunk (kSlots pushes ago) must be done
  if ((size_t)n * 2 > sl.cap) {
    if (sl.hp) B2_CUDA(h, cudaFreeHost(sl.hp));
    if (sl.dp) B2_CUDA(h, cudaFree(sl.dp));
    sl.hp = sl.dp = nullptr;
    sl.cap = 0;
    const size_t want = (size_t)n * 2 + 256;
    B2_CUDA(h, cudaMallocHost(&sl.hp, want));
    B2_CUDA(h, cudaMalloc(&sl.dp, want));
    sl.cap = want;
  }
  if ((size_t)nwin * 4 > sl.out_cap) {
    if (sl.hout) B2_CUDA(h, cudaFreeHost(sl.hout));
    if (sl.dout) B2_CUDA(h, cudaFree(sl.dout));
    sl.hout = sl.dout = nullptr;
    sl.out_cap = 0;
    const size_t want = (size_t)nwin * 4 + 256;
    B2_CUDA(h, cudaMallocHost((void**)&sl.hout, want));
    B2_CUDA(h, cudaMalloc((void**)&sl.dout, want));
    sl.out_cap = want;
  }
  if (!sl.ev) B2_CUDA(h, cudaEventCreateWithFlags(&sl.ev, cudaEventDisableTiming));
  memcpy(sl.hp, pcm_bytes, (size_t)n * 2);
  B2_CUDA(h, cudaMemcpyAsync(sl.dp, sl.hp, (size_t)n * 2, cudaMemcpyHostToDevice, h->stream));
  const int64_t pcm_off[2] = {0, n}, out_off[2] = {0, nwin};
  B2_TRY(b2i_vad_launch(h, (const int16_t*)sl.dp, pcm_off, 1, vs.fpw, vs.label, (int64_t)vs.fpw * vs.thr,
                        vs.z_lo, vs.z_hi, sl.dout, out_off));
  B2_CUDA(h, cudaMemcpyAsync(sl.hout, sl.dout, (size_t)nwin * 4, cudaMemcpyDeviceToHost, h->stream));
  B2_CUDA(h, cudaEventRecord(sl.ev, h->stream));
  sl.n_out = nwin;
  sl.busy = true;
  vs.windows += nwin;
  ++vs.seq;
  return B2_OK;
}

extern "C" int64_t b2_vad_stream_windows(b2_handle h) { return (h && h->vs.active) ? h->vs.windows : -1; }

extern "C" int b2_vad_stream_end(b2_handle h, float* out, int64_t capacity, int64_t* n_out) {
  B2_ENTER(h);
  auto& vs = h->vs;
  if (!vs.active) B2_FAIL(h, B2_ERR_BAD_ARG, "vad stream: not open");
  const int ns = b2_ctx::VadStream::kSlots;
  for (uint64_t i = vs.seq > (uint64_t)ns ? vs.seq - ns : 0; i < vs.seq; ++i)  // oldest first
    B2_TRY(vs_collect(h, vs.slot[i % ns]));
  vs.active = false;
  const int64_t total = (int64_t)vs.results.size();
  if (n_out) *n_out = total;
  if (total > capacity || (total && !out)) {
    vs.results.clear();
    B2_FAIL(h, B2_ERR_BAD_ARG, "vad stream: output holds %lld windows, capacity %lld", (long long)total,
            (long long)capacity);
  }
  if (total) memcpy(out, vs.results.data(), (size_t)total * 4);
  vs.results.clear();
  return B2_OK;
}

extern "C" int b2_synth_pcm(b2_handle h, const uint8_t* window_class, int64_t n_windows, int fpw,
                            uint32_t seed, int16_t* pcm_out, int memspace) {
  B2_ENTER(h);
  if (n_windows < 0 || fpw <= 0) B2_FAIL(h, B2_ERR_BAD_ARG, "synth: bad sizes");
  if (n_windows == 0) return B2_OK;
  if (!window_class || !pcm_out) B2_FAIL(h, B2_ERR_BAD_ARG, "synth: null pointer");
  const uint8_t* d_cls = window_class;
  int16_t* d_out = pcm_out;
  size_t out_bytes = (size_t)n_windows * fpw * 2;
  if (memspace == B2_HOST) {
    void *dp, *dq;
    B2_TRY(stage_in(h, b2_ctx::WS_STAGE_IN0, window_class, (size_t)n_windows, &dp));
    B2_TRY(b2i_ws(h, b2_ctx::WS_STAGE_OUT, out_bytes, &dq));
    d_cls = (const uint8_t*)dp;
    d_out = (int16_t*)dq;
  }
  B2_TRY(b2i_synth_launch(h, d_cls, n_windows, fpw, seed, d_out));
  if (memspace == B2_HOST) {
    B2_TRY(copy_out(h, pcm_out, d_out, out_bytes));
    B2_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  return B2_OK;
}

// ---- rasteriser ----------------------------------------------------------------------------
static double scaled_seconds_host(double t, double ratio) {
  // timedelta(seconds=t*ratio).total_seconds(): see raster.cu (same arithmetic, host copy used
  // only to size the output arrays).
  volatile double x = t * ratio;
  double whole;
  double frac = modf(x, &whole);
  volatile double fus = frac * 1e6;
  long long us = (long long)whole * 1000000LL + (long long)nearbyint(fus);
  return (double)us / 1e6;
}

extern "C" int b2_rasterize_lengths(const double* cue_end_s, const int64_t* cue_off, int B,
                                    const double* ratios, int K, int per_pair_ratios,
                                    int sample_rate, int64_t* lengths) {
  if (B < 0 || K < 0 || !cue_off || (!ratios && K) || !lengths || sample_rate <= 0)
    return B2_ERR_BAD_ARG;
  for (int b = 0; b < B; ++b) {
    // max over cues of scaled(end) == scaled(max end) for ratio > 0: the product, the microsecond
    // rounding and the division are all monotone non-decreasing (speech_transformers.py:958-960)
    double max_end = 0.0;
    bool any = false;
    for (int64_t c = cue_off[b]; c < cue_off[b + 1]; ++c)
      if (!any || cue_end_s[c] > max_end) { max_end = cue_end_s[c]; any = true; }
    for (int k = 0; k < K; ++k) {
      double r = per_pair_ratios ? ratios[(size_t)b * K + k] : ratios[k];
      if (!(r > 0.0)) return B2_ERR_BAD_ARG;
      double max_time = 0.0;
      if (any) {
        double e = scaled_seconds_host(max_end, r);
        if (e > max_time) max_time = e;
      }
      volatile double prod = max_time * (double)sample_rate;
      lengths[(size_t)b * K + k] = (int64_t)prod + 2;  // speech_transformers.py:962
    }
  }
  return B2_OK;
}

extern "C" int b2_rasterize(b2_handle h, const double* cue_start_s, const double* cue_end_s,
                            const uint8_t* cue_keep, const int64_t* cue_off, int B,
                            const double* ratios, int K, int per_pair_ratios,
                            const double* levels, int sample_rate, double start_seconds,
                            float* out, const int64_t* out_off, int memspace) {
  B2_ENTER(h);
  if (B < 0 || K < 0 || !cue_off || !out_off || sample_rate <= 0)
    B2_FAIL(h, B2_ERR_BAD_ARG, "rasterize: bad arguments");
  if (B == 0 || K == 0) return B2_OK;
  if (!ratios || (cue_off[B] && (!cue_start_s || !cue_end_s)))
    B2_FAIL(h, B2_ERR_BAD_ARG, "rasterize: null cue/ratio arrays");
  int64_t total = out_off[(size_t)B * K];
  if (total && !out) B2_FAIL(h, B2_ERR_BAD_ARG, "rasterize: null output");
  float* d_out = out;
  if (memspace == B2_HOST) {
    void* dq;
    B2_TRY(b2i_ws(h, b2_ctx::WS_STAGE_OUT, (size_t)total * 4 + 16, &dq));
    d_out = (float*)dq;
  }
  B2_TRY(b2i_raster_launch(h, cue_start_s, cue_end_s, cue_keep, cue_off, B, ratios, K,
                           per_pair_ratios, levels, sample_rate, start_seconds, d_out, out_off));
  if (memspace == B2_HOST) {
    B2_TRY(copy_out(h, out, d_out, (size_t)total * 4));
    B2_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  return B2_OK;
}

// ---- fused-VAD blend -------------------------------------------------------------------------
extern "C" int b2_blend_signals(b2_handle h, const float* a, const float* b, int64_t n, int mode,
                                double wa, double wb, float* out, int memspace) {
  B2_ENTER(h);
  if (n < 0 || mode < 0 || mode > 2) B2_FAIL(h, B2_ERR_BAD_ARG, "blend: bad arguments");
  if (n == 0) return B2_OK;
  if (!a || !b || !out) B2_FAIL(h, B2_ERR_BAD_ARG, "blend: null pointer");
  const float *d_a = a, *d_b = b;
  float* d_out = out;
  if (memspace == B2_HOST) {
    void *da, *db, *dq;
    B2_TRY(stage_in(h, b2_ctx::WS_STAGE_IN0, a, (size_t)n * 4, &da));
    B2_TRY(stage_in(h, b2_ctx::WS_STAGE_IN1, b, (size_t)n * 4, &db));
    B2_TRY(b2i_ws(h, b2_ctx::WS_STAGE_OUT, (size_t)n * 4, &dq));
    d_a = (const float*)da;
    d_b = (const float*)db;
    d_out = (float*)dq;
  }
  B2_TRY(b2i_blend_launch(h, d_a, d_b, n, mode, wa, wb, d_out));
  if (memspace == B2_HOST) {
    B2_TRY(copy_out(h, out, d_out, (size_t)n * 4));
    B2_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  return B2_OK;
}

// ---- boundaries ----------------------------------------------------------------------------
extern "C" int b2_first_last_nonzero(b2_handle h, const float* sig, const int64_t* sig_off, int n,
                                     int64_t* first, int64_t* last, int memspace) {
  B2_ENTER(h);
  if (n < 0 || !sig_off || !first || !last) B2_FAIL(h, B2_ERR_BAD_ARG, "boundaries: bad arguments");
  if (n == 0) return B2_OK;
  const float* d_sig = sig;
  int64_t *d_first = first, *d_last = last;
  if (memspace == B2_HOST) {
    void *dp, *dq;
    B2_TRY(stage_in(h, b2_ctx::WS_STAGE_IN0, sig, (size_t)sig_off[n] * 4, &dp));
    B2_TRY(b2i_ws(h, b2_ctx::WS_STAGE_OUT, (size_t)n * 16, &dq));
    d_sig = (const float*)dp;
    d_first = (int64_t*)dq;
    d_last = d_first + n;
  }
  B2_TRY(b2i_bounds_launch(h, d_sig, sig_off, n, d_first, d_last));
  if (memspace == B2_HOST) {
    B2_TRY(copy_out(h, first, d_first, (size_t)n * 8));
    B2_TRY(copy_out(h, last, d_last, (size_t)n * 8));
    B2_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  return B2_OK;
}

// ---- aligner -------------------------------------------------------------------------------
extern "C" int b2_align_batch(b2_handle h, const float* ref, const int64_t* ref_off,
                              const float* sub, const int64_t* sub_off, int B, int K,
                              int64_t max_offset_samples, double* score, int32_t* offset,
                              int32_t* status, int memspace) {
  B2_ENTER(h);
  if (B < 0 || K < 0 || !ref_off || !sub_off) B2_FAIL(h, B2_ERR_BAD_ARG, "align: bad arguments");
  if (B == 0 || K == 0) return B2_OK;
  if (!score || !offset || !status) B2_FAIL(h, B2_ERR_BAD_ARG, "align: null output");
  size_t J = (size_t)B * K;
  B2_CHECK_DEV(h, memspace, ref, "align: ref");
  B2_CHECK_DEV(h, memspace, sub, "align: sub");
  B2_CHECK_DEV(h, memspace, score, "align: score");
  const float *d_ref = ref, *d_sub = sub;
  double* d_score = score;
  int32_t *d_offset = offset, *d_status = status;
  if (memspace == B2_HOST) {
    void *dr, *ds, *dq;
    B2_TRY(stage_in(h, b2_ctx::WS_STAGE_IN0, ref, (size_t)ref_off[B] * 4, &dr));
    B2_TRY(stage_in(h, b2_ctx::WS_STAGE_IN1, sub, (size_t)sub_off[J] * 4, &ds));
    B2_TRY(b2i_ws(h, b2_ctx::WS_STAGE_OUT, J * 16 + 64, &dq));
    d_ref = (const float*)dr;
    d_sub = (const float*)ds;
    d_score = (double*)dq;
    d_offset = (int32_t*)(d_score + J);
    d_status = d_offset + J;
  }
  B2_TRY(b2i_align_launch(h, d_ref, ref_off, d_sub, sub_off, B, K, max_offset_samples, d_score,
                          d_offset, d_status, /*winner_only=*/0, /*cue_src=*/nullptr));
  if (memspace == B2_HOST) {
    B2_TRY(copy_out(h, score, d_score, J * 8));
    B2_TRY(copy_out(h, offset, d_offset, J * 4));
    B2_TRY(copy_out(h, status, d_status, J * 4));
    B2_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  return B2_OK;
}

extern "C" int b2_reduce_ratios(b2_handle h, const double* score, const int32_t* offset,
                                const int32_t* status, int B, int K, int64_t max_offset_samples,
                                double* best_score, int32_t* best_offset, int32_t* best_k,
                                int memspace) {
  B2_ENTER(h);
  if (B < 0 || K <= 0) B2_FAIL(h, B2_ERR_BAD_ARG, "reduce: bad arguments");
  if (B == 0) return B2_OK;
  if (!score || !offset || !best_score || !best_offset || !best_k)
    B2_FAIL(h, B2_ERR_BAD_ARG, "reduce: null pointer");
  size_t J = (size_t)B * K;
  const double* d_score = score;
  const int32_t *d_offset = offset, *d_status = status;
  double* d_bs = best_score;
  int32_t *d_bo = best_offset, *d_bk = best_k;
  if (memspace == B2_HOST) {
    void *d0, *dq;
    B2_TRY(b2i_ws(h, b2_ctx::WS_STAGE_IN0, J * 16 + 64, &d0));
    B2_CUDA(h, cudaMemcpyAsync(d0, score, J * 8, cudaMemcpyHostToDevice, h->stream));
    int32_t* doff = (int32_t*)((char*)d0 + J * 8);
    B2_CUDA(h, cudaMemcpyAsync(doff, offset, J * 4, cudaMemcpyHostToDevice, h->stream));
    int32_t* dst = doff + J;
    if (status)
      B2_CUDA(h, cudaMemcpyAsync(dst, status, J * 4, cudaMemcpyHostToDevice, h->stream));
    B2_TRY(b2i_ws(h, b2_ctx::WS_STAGE_OUT, (size_t)B * 16 + 64, &dq));
    d_score = (const double*)d0;
    d_offset = doff;
    d_status = status ? dst : nullptr;
    d_bs = (double*)dq;
    d_bo = (int32_t*)(d_bs + B);
    d_bk = d_bo + B;
  }
  B2_TRY(b2i_reduce_launch(h, d_score, d_offset, d_status, B, K, max_offset_samples, d_bs, d_bo,
                           d_bk));
  if (memspace == B2_HOST) {
    B2_TRY(copy_out(h, best_score, d_bs, (size_t)B * 8));
    B2_TRY(copy_out(h, best_offset, d_bo, (size_t)B * 4));
    B2_TRY(copy_out(h, best_k, d_bk, (size_t)B * 4));
    B2_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  return B2_OK;
}

// ---- whole hot path --------------------------------------------------------------------------
// VideoSpeechTransformer.fit (VAD) -> K x (SubtitleScaler + SubtitleSpeechTransformer) ->
// MaxScoreAligner(FFTAligner).fit_transform, for B pairs (ffsubsync/ffsubsync.py:637,196-235).
extern "C" int b2_sync_batch(b2_handle h, const int16_t* pcm, const int64_t* pcm_off, int B,
                             int frame_rate, int sample_rate, float non_speech_label,
                             int64_t energy_threshold, int z_lo, int z_hi,
                             const double* cue_start_s, const double* cue_end_s,
                             const uint8_t* cue_keep, const int64_t* cue_off, const double* ratios,
                             int K, double start_seconds, int64_t max_offset_samples,
                             double* best_score, int32_t* best_offset, int32_t* best_k,
                             double* all_score, int32_t* all_offset, int memspace) {
  B2_ENTER(h);
  B2Range range("b2_sync_batch");
  const bool resident = memspace == B2_DEVICE_RESIDENT;
  if (resident) memspace = B2_DEVICE;
  if (B < 0 || K <= 0 || !pcm_off || !cue_off || !ratios)
    B2_FAIL(h, B2_ERR_BAD_ARG, "sync_batch: bad arguments");
  if (B == 0) return B2_OK;
  if (!best_score || !best_offset || !best_k) B2_FAIL(h, B2_ERR_BAD_ARG, "sync_batch: null output");
  const int fpw = b2_vad_frames_per_window(frame_rate, sample_rate);
  if (fpw <= 0) B2_FAIL(h, B2_ERR_BAD_ARG, "sync_batch: bad frame_rate/sample_rate");
  if (z_lo < 0) z_lo = 0;
  if (z_hi < 0) z_hi = (3 * fpw) / 8;
  const size_t J = (size_t)B * K;
  std::vector<int64_t> ref_off(B + 1), sub_off(J + 1), lengths(J);
  ref_off[0] = 0;
  for (int b = 0; b < B; ++b) {
    const int64_t n = pcm_off[b + 1] - pcm_off[b];
    if (n < 0) B2_FAIL(h, B2_ERR_BAD_ARG, "sync_batch: pcm_off not monotone");
    ref_off[b + 1] = ref_off[b] + (n + fpw - 1) / fpw;
  }
  if (b2_rasterize_lengths(cue_end_s, cue_off, B, ratios, K, 0, sample_rate, lengths.data()) != B2_OK)
    B2_FAIL(h, B2_ERR_BAD_ARG, "sync_batch: bad cue list / ratios");
  sub_off[0] = 0;
  for (size_t j = 0; j < J; ++j) sub_off[j + 1] = sub_off[j] + lengths[j];

  // Default: the K subtitle signals of a pair are never materialised as floats - the cue list is
  // rasterised into bit masks (1 bit per frame) that the correlation kernel and the exact re-score
  // read.  B2_FUSED_RASTER=0 (A/B and test knob): raster_cues_kernel writes float signals to HBM and
  // the generic aligner (the b2_align_batch path) reads them back.
  bool fused = true;
  if (const char* e = getenv("B2_FUSED_RASTER")) fused = atoi(e) != 0;
  B2CueSource cue_src{cue_start_s, cue_end_s, cue_keep, cue_off, ratios, sample_rate, start_seconds};

  void *d_refsig, *d_subsig = nullptr, *d_res;
  // a chained resident call (see below) writes the buffer the previous call is not reading any more
  const bool chained_candidate = resident && _b2_fence_was_valid;
  const int refsig_slot = chained_candidate && h->refsig_parity == 0 ? b2_ctx::WS_SIG_REF2 : b2_ctx::WS_SIG_REF;
  B2_TRY(b2i_ws(h, refsig_slot, (size_t)ref_off[B] * 4 + 64, &d_refsig));
  if (!fused) B2_TRY(b2i_ws(h, b2_ctx::WS_SIG_SUB, (size_t)sub_off[J] * 4 + 64, &d_subsig));
  B2_TRY(b2i_ws(h, b2_ctx::WS_MISC, J * 16 + (size_t)B * 16 + 256, &d_res));
  double* d_score = (double*)d_res;
  double* d_bs = d_score + J;
  int32_t* d_offset = (int32_t*)(d_bs + B);
  int32_t* d_status = d_offset + J;
  int32_t* d_bo = d_status + J;
  int32_t* d_bk = d_bo + B;

  B2_CHECK_DEV(h, memspace, pcm, "sync_batch: pcm");
  B2_CHECK_DEV(h, memspace, best_score, "sync_batch: best_score");
  const int16_t* d_pcm = pcm;
  if (memspace == B2_HOST) {
    void* dp;
    B2_TRY(stage_in(h, b2_ctx::WS_STAGE_IN0, pcm, (size_t)pcm_off[B] * 2, &dp));
    d_pcm = (const int16_t*)dp;
  }
  double* o_score = (memspace == B2_DEVICE && all_score) ? all_score : d_score;
  int32_t* o_offset = (memspace == B2_DEVICE && all_offset) ? all_offset : d_offset;
  double* o_bs = memspace == B2_DEVICE ? best_score : d_bs;
  int32_t* o_bo = memspace == B2_DEVICE ? best_offset : d_bo;
  int32_t* o_bk = memspace == B2_DEVICE ? best_k : d_bk;
  // only the best ratio of each pair is reported unless the per-ratio arrays are requested:
  // ratios that cannot win even after the round-off bound tau are then not re-scored exactly (B2_ALIGN_APPROX)
  const int winner_only = (!all_score && !all_offset) ? 1 : 0;

  // Software pipeline over sub-batches of pairs.  The VAD (HBM-bound) of every sub-batch is queued on the
  // internal high-priority stream: sub-batch 0 on the whole GPU, the later ones on `vad_sms` SMs only (one
  // lane-per-window CTA per SM, csrc/vad.cu); the rasterisation / correlation / reduction of sub-batch i
  // (FP32- and shared-memory bound) follows on the caller's stream as soon as its VAD is done and runs on
  // the SMs the VAD leaves free - a VAD CTA owns its SM's shared memory, so the block scheduler keeps the
  // two apart.  Needs the lane-per-window kernel (1.3 instructions per byte: ~80 GB/s per SM); with the
  // lane-group kernel (every SM's issue slots to reach the HBM roofline) partitioning never paid
  // (profiles/r2a_partition_probe.txt).  B2_SUBBATCHES / B2_VAD_SMS override the defaults; 1 / 0 = off.
  // Defaults (measured on 256 two-hour pairs, tools/pipeline_probe.py: 3 sub-batches x 80 of 148 SMs =
  // 11.16 ms per step against 12.49 unpipelined; 2-6 sub-batches and 74-86 SMs are within 4 %); small
  // batches stay unpipelined (the alignment of a third of a small batch is launch- and tail-bound).
  int n_sub = 1, vad_sms = 0;
  if (B >= 96 && b2i_vad_lane_eligible(pcm_off, B, fpw)) {
    n_sub = 3;
    vad_sms = (h->sm_count * 80 + 74) / 148;
  }
  if (const char* e = getenv("B2_SUBBATCHES")) n_sub = std::max(1, std::min(B, atoi(e)));
  if (const char* e = getenv("B2_VAD_SMS")) vad_sms = std::max(0, std::min(h->sm_count, atoi(e)));
  // probe knobs (tools/pipeline_probe.py): share of the pairs in the first sub-batch (its VAD has nothing to
  // overlap with; 0 = even split), and a cap on the persistent correlation grid while a VAD holds vad_sms SMs
  int head_pct = 0, corr_cap = 0;
  if (const char* e = getenv("B2_PIPE_HEAD_PCT")) head_pct = std::max(0, std::min(90, atoi(e)));
  if (const char* e = getenv("B2_PIPE_CORR_CAP")) corr_cap = atoi(e) != 0;
  if (n_sub == 1) {
    B2_TRY(b2i_vad_launch(h, d_pcm, pcm_off, B, fpw, non_speech_label, (int64_t)fpw * energy_threshold, z_lo,
                          z_hi, (float*)d_refsig, ref_off.data()));
    if (!fused)
      B2_TRY(b2i_raster_launch(h, cue_start_s, cue_end_s, cue_keep, cue_off, B, ratios, K, 0, nullptr,
                               sample_rate, start_seconds, (float*)d_subsig, sub_off.data()));
    B2_TRY(b2i_align_launch(h, (const float*)d_refsig, ref_off.data(), (const float*)d_subsig, sub_off.data(),
                            B, K, max_offset_samples, o_score, o_offset, d_status, winner_only,
                            fused ? &cue_src : nullptr));
    B2_TRY(b2i_reduce_launch(h, o_score, o_offset, d_status, B, K, max_offset_samples, o_bs, o_bo, o_bk));
  } else {
    if (n_sub > b2_ctx::kEvents - 2) n_sub = b2_ctx::kEvents - 2;
    std::vector<int> cut(n_sub + 1);
    for (int i = 0; i <= n_sub; ++i) cut[i] = (int)((int64_t)B * i / n_sub);
    if (head_pct > 0 && n_sub > 1) {
      const int head = std::max(1, std::min(B - (n_sub - 1), (int)((int64_t)B * head_pct / 100)));
      for (int i = 1; i <= n_sub; ++i) cut[i] = head + (int)((int64_t)(B - head) * (i - 1) / (n_sub - 1));
    }
    if (const char* e = getenv("B2_PIPE_CUTS")) {   // probe knob: explicit first pairs of sub-batches 1.., e.g. "54,135,197"
      std::vector<int> c{0};
      for (const char* q = e; *q;) {
        char* end = nullptr;
        const long v = strtol(q, &end, 10);
        if (end == q) break;
        if (v > c.back() && v < B && (int)c.size() < b2_ctx::kEvents - 2) c.push_back((int)v);
        q = *end ? end + 1 : end;
      }
      c.push_back(B);
      cut = c;
      n_sub = (int)cut.size() - 1;
    }
    std::vector<cudaEvent_t> vad_done(n_sub);
    cudaEvent_t inputs_ready = next_event(h);
    B2_CUDA(h, cudaEventRecord(inputs_ready, h->stream));
    // B2_PIPE_TRACE=1 (diagnostic; synchronises): device timeline of the sub-batches and host enqueue times
    const bool trace = getenv("B2_PIPE_TRACE") != nullptr;
    std::vector<cudaEvent_t> tev;   // t0, then per sub-batch: VAD start, VAD end, chain start, chain end
    std::vector<double> host_ms(n_sub + 1, 0.0);
    const auto host_t0 = std::chrono::steady_clock::now();
    auto host_now = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count(); };
    if (trace) {
      tev.resize(1 + 4 * (size_t)n_sub);
      for (auto& e : tev) B2_CUDA(h, cudaEventCreate(&e));
      B2_CUDA(h, cudaEventRecord(tev[0], h->stream));
    }
    // B2_DEVICE_RESIDENT, previous entry point on this handle = a pipelined resident b2_sync_batch: the VAD
    // starts behind that call's fence (everything on the caller's stream up to, not including, its last
    // correlation chain) and overlaps that chain like a further sub-batch - on vad_sms SMs if it is still running.
    // It writes the other reference-signal buffer (the chain still reads the previous one); the chains of the
    // call before that, which read this buffer, precede the fence.
    const bool chained = chained_candidate;
    bool prev_busy = false;
    if (chained) {
      prev_busy = cudaEventQuery(h->resident_done) == cudaErrorNotReady;
      cudaGetLastError();
    }
    {
      Stream2Scope on2(h);
      B2_CUDA(h, cudaStreamWaitEvent(h->stream, chained ? h->resident_fence : inputs_ready, 0));
      for (int i = 0; i < n_sub; ++i) {
        const int b0 = cut[i], b1 = cut[i + 1];
        h->vad_partition_sms = (i > 0 || prev_busy) ? vad_sms : 0;
        if (trace) B2_CUDA(h, cudaEventRecord(tev[1 + 4 * i], h->stream));
        const int st = b1 > b0 ? b2i_vad_launch(h, d_pcm, pcm_off + b0, b1 - b0, fpw, non_speech_label,
                                                (int64_t)fpw * energy_threshold, z_lo, z_hi, (float*)d_refsig,
                                                ref_off.data() + b0)
                               : B2_OK;
        h->vad_partition_sms = 0;
        if (st != B2_OK) return st;
        vad_done[i] = next_event(h);
        B2_CUDA(h, cudaEventRecord(vad_done[i], h->stream));
        if (trace) B2_CUDA(h, cudaEventRecord(tev[2 + 4 * i], h->stream));
      }
    }
    host_ms[0] = host_now();
    for (int i = 0; i < n_sub; ++i) {
      const int b0 = cut[i], b1 = cut[i + 1];
      const int nb = b1 - b0;
      if (resident && i == n_sub - 1) B2_CUDA(h, cudaEventRecord(h->resident_fence, h->stream));
      B2_CUDA(h, cudaStreamWaitEvent(h->stream, vad_done[i], 0));
      if (trace) B2_CUDA(h, cudaEventRecord(tev[3 + 4 * i], h->stream));
      if (nb == 0) {
        if (trace) B2_CUDA(h, cudaEventRecord(tev[4 + 4 * i], h->stream));
        continue;
      }
      struct CapScope {   // restores the handle's grid cap on every exit path
        b2_ctx* h;
        int saved;
        ~CapScope() { h->corr_max_ctas = saved; }
      } cap_scope{h, h->corr_max_ctas};
      if (corr_cap && i + 1 < n_sub && vad_sms > 0 && vad_sms < h->sm_count) h->corr_max_ctas = h->sm_count - vad_sms;
      const size_t j0 = (size_t)b0 * K;
      if (!fused)
        B2_TRY(b2i_raster_launch(h, cue_start_s, cue_end_s, cue_keep, cue_off + b0, nb, ratios, K, 0, nullptr,
                                 sample_rate, start_seconds, (float*)d_subsig, sub_off.data() + j0));
      B2CueSource sub_src = cue_src;
      sub_src.cue_off = cue_off + b0;
      B2_TRY(b2i_align_launch(h, (const float*)d_refsig, ref_off.data() + b0, (const float*)d_subsig,
                              sub_off.data() + j0, nb, K, max_offset_samples, o_score + j0, o_offset + j0,
                              d_status + j0, winner_only, fused ? &sub_src : nullptr));
      B2_TRY(b2i_reduce_launch(h, o_score + j0, o_offset + j0, d_status + j0, nb, K, max_offset_samples,
                               o_bs + b0, o_bo + b0, o_bk + b0));
      if (trace) B2_CUDA(h, cudaEventRecord(tev[4 + 4 * i], h->stream));
      host_ms[i + 1] = host_now();
    }
    if (resident) {
      B2_CUDA(h, cudaEventRecord(h->resident_done, h->stream));
      h->refsig_parity = refsig_slot == b2_ctx::WS_SIG_REF2 ? 1 : 0;
      h->resident_fence_valid = true;
    }
    if (trace) {
      B2_CUDA(h, cudaStreamSynchronize(h->stream2));
      B2_CUDA(h, cudaStreamSynchronize(h->stream));
      auto at = [&](size_t k) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, tev[0], tev[k]);
        return ms;
      };
      fprintf(stderr, "[b2 pipe] B=%d n_sub=%d vad_sms=%d head_pct=%d corr_cap=%d chained=%d prev_busy=%d; host: VADs enqueued at %.3f ms\n",
              B, n_sub, vad_sms, head_pct, corr_cap, (int)chained, (int)prev_busy, host_ms[0]);
      for (int i = 0; i < n_sub; ++i)
        fprintf(stderr, "[b2 pipe]  sub %d pairs %4d..%4d  VAD %7.3f -> %7.3f ms   chain %7.3f -> %7.3f ms   host enqueued chain at %.3f ms\n",
                i, cut[i], cut[i + 1], at(1 + 4 * i), at(2 + 4 * i), at(3 + 4 * i), at(4 + 4 * i), host_ms[i + 1]);
      for (auto& e : tev) cudaEventDestroy(e);
    }
  }
  if (memspace == B2_DEVICE) return B2_OK;
  B2_TRY(copy_out(h, best_score, d_bs, (size_t)B * 8));
  B2_TRY(copy_out(h, best_offset, d_bo, (size_t)B * 4));
  B2_TRY(copy_out(h, best_k, d_bk, (size_t)B * 4));
  if (all_score) B2_TRY(copy_out(h, all_score, d_score, J * 8));
  if (all_offset) B2_TRY(copy_out(h, all_offset, d_offset, J * 4));
  B2_CUDA(h, cudaStreamSynchronize(h->stream));
  return B2_OK;
}

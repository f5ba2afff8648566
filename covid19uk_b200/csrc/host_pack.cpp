// Host side of seir_log_prob_host (seir_api.cu): a small persistent thread pool that narrows the caller's float64
// event tensor (integer-valued counts, model_spec.py:118-126) to uint16 in pinned staging memory, so that the PCIe
// transfer carries 2 bytes per count instead of 8.  The narrowing is exact or refused: a block holding any value that
// is not an integer in [0, 65535] is reported and the caller ships that block as float64 instead (the device-side
// ingest then flags invalid events exactly as it does for device-resident input).
//
// Plain C++ (g++), no CUDA: compiled separately and linked into libseir_b200.so.
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

namespace {

// exact double -> uint16 narrowing of n values; returns false if any value is not representable
bool pack_block_scalar(const double* __restrict__ src, uint16_t* __restrict__ dst, size_t n) {
  int bad = 0;
  for (size_t i = 0; i < n; ++i) {
    const double v = src[i];
    const bool in_range = v >= 0.0 && v <= 65535.0;  // (false for NaN; the conversion below is defined only inside the range)
    const int32_t k = in_range ? (int32_t)v : 0;
    bad |= !in_range | ((double)k != v);
    dst[i] = (uint16_t)k;
  }
  return bad == 0;
}

#if defined(__x86_64__)
// AVX2: 16 counts per iteration -- truncate to int32, convert back and compare (exactness), check the 16-bit range, pack with
// unsigned saturation.  (The auto-vectorised scalar loop ran at ~4.5 GB/s per thread; this one is bound by the memory read:
// one thread's line-fill buffers, so the source is prefetched 4 KB ahead into L1/L2 -- measured on an 8-vCPU Xeon guest,
// 7 threads: 30 GB/s without the prefetch, 39 GB/s with it.)
__attribute__((target("avx2"))) bool pack_block_avx2(const double* __restrict__ src, uint16_t* __restrict__ dst, size_t n) {
  __m256d okacc = _mm256_castsi256_pd(_mm256_set1_epi64x(-1));
  __m128i hiacc = _mm_setzero_si128();
  size_t i = 0;
  // Streaming (non-temporal) stores when the destination is 16-byte aligned: the staging buffer is read next by the
  // GPU's DMA engine, never by a core; lines left Modified in the cores' private caches made the LAST chunks of a
  // batch travel at ~7 GB/s instead of ~50 GB/s (nothing evicts them once the pool goes idle; tools/e2e_probe.py).
  const bool nt = (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
  for (; i + 16 <= n; i += 16) {
    _mm_prefetch(reinterpret_cast<const char*>(src + i + 512), _MM_HINT_T0);
    _mm_prefetch(reinterpret_cast<const char*>(src + i + 520), _MM_HINT_T0);
    const __m256d a = _mm256_loadu_pd(src + i), b = _mm256_loadu_pd(src + i + 4), c = _mm256_loadu_pd(src + i + 8),
                  d = _mm256_loadu_pd(src + i + 12);
    const __m128i ia = _mm256_cvttpd_epi32(a), ib = _mm256_cvttpd_epi32(b), ic = _mm256_cvttpd_epi32(c), id = _mm256_cvttpd_epi32(d);
    okacc = _mm256_and_pd(okacc, _mm256_and_pd(_mm256_and_pd(_mm256_cmp_pd(_mm256_cvtepi32_pd(ia), a, _CMP_EQ_OQ),
                                                             _mm256_cmp_pd(_mm256_cvtepi32_pd(ib), b, _CMP_EQ_OQ)),
                                               _mm256_and_pd(_mm256_cmp_pd(_mm256_cvtepi32_pd(ic), c, _CMP_EQ_OQ),
                                                             _mm256_cmp_pd(_mm256_cvtepi32_pd(id), d, _CMP_EQ_OQ))));
    // negative or > 65535 => a non-zero high half
    hiacc = _mm_or_si128(hiacc, _mm_or_si128(_mm_or_si128(_mm_srli_epi32(ia, 16), _mm_srli_epi32(ib, 16)),
                                             _mm_or_si128(_mm_srli_epi32(ic, 16), _mm_srli_epi32(id, 16))));
    if (nt) {
      _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), _mm_packus_epi32(ia, ib));
      _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 8), _mm_packus_epi32(ic, id));
    } else {
      _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i), _mm_packus_epi32(ia, ib));
      _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i + 8), _mm_packus_epi32(ic, id));
    }
  }
  if (nt) _mm_sfence();  // the chunk's "done" flag is published after this
  bool ok = _mm256_movemask_pd(okacc) == 0xF && _mm_testz_si128(hiacc, hiacc);
  if (i < n) ok = pack_block_scalar(src + i, dst + i, n - i) && ok;
  return ok;
}
#endif

bool pack_block(const double* __restrict__ src, uint16_t* __restrict__ dst, size_t n) {
#if defined(__x86_64__)
  static const bool have_avx2 = __builtin_cpu_supports("avx2");
  if (have_avx2) return pack_block_avx2(src, dst, n);
#endif
  return pack_block_scalar(src, dst, n);
}

struct Pool {
  std::vector<std::thread> workers;
  std::mutex mu;
  std::condition_variable cv;
  uint64_t generation = 0;
  bool stop = false;
  // current batch
  const double* src = nullptr;
  uint16_t* dst = nullptr;
  size_t chunk_elems = 0, job_elems = 0, total_elems = 0;
  int njobs = 0, jobs_per_chunk = 1;
  std::atomic<int> next{0};
  std::atomic<int> active{0};  // workers inside run_jobs (a new batch waits for stragglers of the previous one)
  std::atomic<int>* chunk_done = nullptr;  // finished jobs per chunk
  std::atomic<int>* chunk_bad = nullptr;
  std::atomic<int>* chunk_owner = nullptr;  // 0 free, 1 pool (being narrowed), 2 caller (shipped as float64)

  void run_jobs() {
    for (;;) {
      const int j = next.fetch_add(1, std::memory_order_relaxed);
      if (j >= njobs) return;
      const int ch = j / jobs_per_chunk;
      int own = chunk_owner[ch].load(std::memory_order_acquire);
      if (own == 0) {
        int expect = 0;
        own = chunk_owner[ch].compare_exchange_strong(expect, 1, std::memory_order_acq_rel) ? 1 : expect;
      }
      if (own == 2) return;  // the caller claims chunks from the back: this one and all later ones go as float64
      const size_t within = (size_t)(j - ch * jobs_per_chunk) * job_elems;  // jobs never straddle chunks
      size_t n = within + job_elems <= chunk_elems ? job_elems : chunk_elems - within;
      const size_t o = (size_t)ch * chunk_elems + within;
      if (o >= total_elems) n = 0;
      else if (o + n > total_elems) n = total_elems - o;  // the last chunk may be short
      const bool ok = pack_block(src + o, dst + o, n);
      if (!ok) chunk_bad[ch].store(1, std::memory_order_relaxed);
      chunk_done[ch].fetch_add(1, std::memory_order_release);
    }
  }

  void worker() {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return stop || generation != seen; });
        if (stop) return;
        seen = generation;
        active.fetch_add(1, std::memory_order_acq_rel);
      }
      run_jobs();
      active.fetch_sub(1, std::memory_order_acq_rel);
    }
  }

  explicit Pool(int n) {
    for (int i = 0; i < n; ++i) workers.emplace_back([this] { worker(); });
  }
  ~Pool() {
    {
      std::lock_guard<std::mutex> lk(mu);
      stop = true;
    }
    cv.notify_all();
    for (auto& t : workers) t.join();
  }
};

Pool* g_pool = nullptr;
std::vector<std::atomic<int>>* g_done = nullptr;
std::vector<std::atomic<int>>* g_bad = nullptr;
std::vector<std::atomic<int>>* g_owner = nullptr;

}  // namespace

extern "C" {

int seir_pack_threads(void) {
  unsigned hc = std::thread::hardware_concurrency();
  if (hc == 0) hc = 4;
  if (const char* e = getenv("SEIR_PACK_THREADS")) {
    const int n = atoi(e);
    if (n >= 1) return n > 64 ? 64 : n;
  }
  // one process per GPU: the ranks of a node share its cores (torchrun exports LOCAL_WORLD_SIZE)
  if (const char* e = getenv("LOCAL_WORLD_SIZE")) {
    const int lw = atoi(e);
    if (lw > 1) hc = hc / (unsigned)lw > 2 ? hc / (unsigned)lw : 2;
  }
  int n = (int)hc - 1;  // the calling thread drives the copies
  if (n < 1) n = 1;
  if (n > 32) n = 32;
  return n;
}

// Start narrowing `nchunks` chunks of `chunk_elems` values each (the last one may be short: `total_elems` in all),
// src -> dst at the same element offsets, front to back.  Returns at once with the number of jobs per chunk.
//   seir_pack_poll(chunk, jpc)   -1 not finished, 1 narrowed exactly, 0 holds a value outside uint16
//   seir_pack_wait(chunk, jpc)   blocking form of poll
//   seir_pack_claim_raw(chunk)   1: the caller now owns the chunk (ships it as float64 itself, the pool will not touch
//                                it nor any later chunk); 0: the pool got there first
// One batch at a time (the chain-set handle is not thread-safe anyway).
int seir_pack_begin(const double* src, uint16_t* dst, size_t chunk_elems, size_t total_elems, int nchunks) {
  if (!g_pool) g_pool = new Pool(seir_pack_threads());
  Pool& p = *g_pool;
  // Workers enter run_jobs() only while holding `mu` (they raise `active` under it), so with the lock held and
  // active == 0 no worker reads the batch descriptor and none can start: everything below is rewritten under the lock.
  std::unique_lock<std::mutex> lk(p.mu);
  while (p.active.load(std::memory_order_acquire) != 0) {  // stragglers of the previous batch
    lk.unlock();
    std::this_thread::yield();
    lk.lock();
  }
  if (!g_done || (int)g_done->size() < nchunks) {
    delete g_done;
    delete g_bad;
    delete g_owner;
    g_done = new std::vector<std::atomic<int>>(nchunks);
    g_bad = new std::vector<std::atomic<int>>(nchunks);
    g_owner = new std::vector<std::atomic<int>>(nchunks);
  }
  const int nthreads = (int)p.workers.size();
  // Every thread works on the EARLIEST unfinished chunk: one job per thread per chunk (jobs of at least 8192 counts), so
  // that chunks complete one after the other at the pool's full rate and the link never waits for a burst of them.
  int jpc = nthreads;
  if (jpc < 1) jpc = 1;
  size_t job = (chunk_elems + jpc - 1) / jpc;
  if (job < 8192) job = 8192;
  job = (job + 63) / 64 * 64;
  jpc = (int)((chunk_elems + job - 1) / job);
  // jobs must not straddle chunks: lay them out per chunk
  for (int c = 0; c < nchunks; ++c) {
    (*g_done)[c].store(0, std::memory_order_relaxed);
    (*g_bad)[c].store(0, std::memory_order_relaxed);
    (*g_owner)[c].store(0, std::memory_order_relaxed);
  }
  {
    p.src = src;
    p.dst = dst;
    p.chunk_elems = chunk_elems;
    p.total_elems = total_elems;
    p.chunk_owner = g_owner->data();
    p.job_elems = job;
    p.jobs_per_chunk = jpc;
    p.njobs = jpc * nchunks;
    p.chunk_done = g_done->data();
    p.chunk_bad = g_bad->data();
    p.next.store(0, std::memory_order_relaxed);
    ++p.generation;
  }
  lk.unlock();
  p.cv.notify_all();
  return jpc;
}

int seir_pack_poll(int chunk, int jobs_per_chunk) {
  if ((*g_done)[chunk].load(std::memory_order_acquire) < jobs_per_chunk) return -1;
  return (*g_bad)[chunk].load(std::memory_order_relaxed) ? 0 : 1;
}

int seir_pack_claim_raw(int chunk) {
  int expect = 0;
  return (*g_owner)[chunk].compare_exchange_strong(expect, 2, std::memory_order_acq_rel) ? 1 : 0;
}

// Abandon the current batch (error path of the caller): no further job is handed out, and the call returns once no
// worker is reading the source buffer any more.
void seir_pack_cancel(void) {
  if (!g_pool) return;
  Pool& p = *g_pool;
  p.next.store(p.njobs, std::memory_order_release);
  while (p.active.load(std::memory_order_acquire) != 0) std::this_thread::yield();
}

int seir_pack_owner(int chunk) { return (*g_owner)[chunk].load(std::memory_order_acquire); }

int seir_pack_wait(int chunk, int jobs_per_chunk) {
  std::atomic<int>& d = (*g_done)[chunk];
  unsigned spins = 0;
  while (d.load(std::memory_order_acquire) < jobs_per_chunk) {
    if (++spins > 64) std::this_thread::yield();
  }
  return (*g_bad)[chunk].load(std::memory_order_relaxed) ? 0 : 1;
}

}  // extern "C"

// Host-side context of libvirgena_gpu.so.
#pragma once
#include <stdlib.h>
#include <stdarg.h>

#include <algorithm>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "common.cuh"

// A growable device buffer owned by the context.
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T>
  T* as() const {
    return (T*)p;
  }
};

struct SeedConfig {  // one k-mer counter (KMerCounter or KMerCounterMSA)
  int K = 0;
  float lowCoef = 0.f, highCoef = 0.f;
  std::vector<int32_t> cutoffs;
  DevBuf d_cutoffs;
};

struct vg_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[8] = {};
  int smCount = 148;
  char err[1024] = {0};
  int64_t required = 0;
  int64_t launches = 0;

  // parameters
  SeedConfig seed;      // single-reference KMerCounter
  SeedConfig seedMsa;   // KMerCounterMSA
  int insertLen = 1000;
  SwParams sw{2, -3, 5, 2};

  // reference (Reference.java fields)
  int G = 0;
  bool haveRef = false;
  std::vector<uint8_t> h_ref;
  std::vector<int32_t> contigEnds, fragmentEnds;
  bool isFragmented = false;
  DevBuf d_ref, d_contigEnds, d_fragmentEnds;
  // single-reference seeding tables (see seed.cuh)
  DevBuf d_poscode;   // u32 per reference position: k-mer id | dup flag
  DevBuf d_exoKeys;   // sorted 64-bit packed exotic k-mers (contain a non-ACGT byte)
  int nExo = 0;
  double seedHitsPerKmer = 0;  // expected raw hits of one read k-mer (sizes the fast path's per-item slot)
  // MSA index
  bool haveMsa = false;
  int msaLength = 0;
  DevBuf d_msaOff, d_msaCols, d_msaExoKeys;  // CSR: 4^K + nExo keys
  int nMsaExo = 0;
  double msaHitsPerKmer = 0;  // expected raw hits of one read k-mer against the MSA index

  // resident reads
  bool haveReads = false;
  int64_t nPairs = 0, nBases = 0;
  int maxReadLen = 0;
  int minReadLenGe[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // shortest read of at least k bases (sizes the window lists)
  DevBuf d_bases, d_off1, d_off2, d_len1, d_len2;
  DevBuf d_packed;  // staging of vg_upload_reads_packed (2-bit words + exceptions)
  std::vector<uint8_t> h_exist;  // per pair: bit0 mate1 exists, bit1 mate2 exists

  // results of the last map call
  int64_t nResults = 0, seq1Bytes = 0;
  bool resultsMsa = false;
  DevBuf d_results, d_seq1pool, d_pairIdx;
  // scratch
  DevBuf d_tasks, d_fill, d_alns, d_rev, d_rowck, d_colck, d_counter, d_scratch, d_seqOff, d_cand, d_ncand, d_misc;
  DevBuf d_s1pool, d_s2pool, d_s1off, d_s2off, d_rowpool, d_rowLen, d_scratchR;
  DevBuf d_swSorted, d_swBins;  // alignment tasks sorted by (R bin, window length), the counting sort's tables
  DevBuf d_swSeq1;  // sequence1 output of vg_sw_align_batch (its own buffer: d_seq1pool belongs to the last mapping)
  DevBuf d_bamN, d_bamOff, d_bamRec, d_bamOps;  // vg_bam_fields
  DevBuf d_rand;  // vg_random_model_counts: genomes, Markov table, offsets
  DevBuf d_seedS, d_seedNS, d_seedFb;  // fast seeding path: per-item LongHit slots, counts, fallback list
  int64_t seedFallbacks = 0;           // items of the last seeding call redone by the general kernel
  std::vector<cudaEvent_t> seedEv;     // 3 events per sub-chunk of the fast seeding path
  double seedStats[8] = {0};           // last map call: gen ms, window ms, fallback ms, LongHits in, LongHits kept, fallback items,
                                       // windows out, items
  int64_t lastTasks = 0;
  // pileup
  DevBuf d_counts;
  bool haveCounts = false;
  double timings[6] = {0, 0, 0, 0, 0, 0};
  // multi-GPU (multi.cuh): NCCL communicator of this context's rank, if any
  size_t budgetCkpt = 0, budgetSeed = 0, budgetWin = 0;  // work-buffer budgets, fixed at first use
  void* comm = nullptr;
  int commRank = -1, commSize = 0;
  double allreduceMs = 0;
};

static inline int vg_fail(vg_ctx* c, int code, const char* fmt, ...) {
  if (c) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(c->err, sizeof c->err, fmt, ap);
    va_end(ap);
  }
  return code;
}

#define VG_CUDA(c, call)                                                                                   \
  do {                                                                                                     \
    cudaError_t e__ = (call);                                                                              \
    if (e__ != cudaSuccess)                                                                                \
      return vg_fail((c), VG_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,                \
                     cudaGetErrorString(e__));                                                             \
  } while (0)

#define VG_TRY(expr)              \
  do {                            \
    int rc__ = (expr);            \
    if (rc__ != VG_OK) return rc__; \
  } while (0)

// Work-buffer budgets are sized for an idle B200 (180 GB); on a device that other contexts share they shrink to a fraction
// of what is free right now plus what this buffer already holds (the work is then cut into more chunks, never refused).
// The device is asked once per context and buffer kind (cudaMemGetInfo costs a driver round trip).
static inline size_t vg_mem_budget(size_t want, double fraction, size_t held) {
  size_t freeB = 0, totalB = 0;
  if (cudaMemGetInfo(&freeB, &totalB) != cudaSuccess) {
    cudaGetLastError();
    return want;
  }
  const size_t cap = (size_t)((double)(freeB + held) * fraction);
  return std::max<size_t>((size_t)64 << 20, std::min(want, cap));
}

// VG_SYNC_DEBUG=1 synchronises after every launch so that a faulting kernel is reported at its own launch site
static inline bool vg_sync_debug() {
  static const bool on = getenv("VG_SYNC_DEBUG") != nullptr;
  return on;
}
#define VG_LAUNCH_CHECK(c)                                                                              \
  do {                                                                                                  \
    (c)->launches++;                                                                                    \
    cudaError_t e__ = cudaGetLastError();                                                               \
    if (e__ == cudaSuccess && vg_sync_debug()) e__ = cudaStreamSynchronize((c)->stream);                \
    if (e__ != cudaSuccess)                                                                             \
      return vg_fail((c), VG_ERR_CUDA, "kernel launch failed at %s:%d: %s", __FILE__, __LINE__,         \
                     cudaGetErrorString(e__));                                                          \
  } while (0)

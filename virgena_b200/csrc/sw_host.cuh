// Host driver of the Smith-Waterman kernels (chunking, R dispatch, launches).
#pragma once
#include <stdlib.h>

#include "ctx.cuh"
#include "sw.cuh"

static inline int sw_pick_R(int mmax) {
  static const int Rs[] = {2, 3, 4, 5, 6, 8, 10, 12, 16, 24, 32};
  for (int r : Rs)
    if (32 * r >= mmax) return r;
  return -1;
}

static inline int sw_validate(vg_ctx* c, int mmax, int nmax) {
  const SwParams& p = c->sw;
  if (p.match <= 0 || p.mismatch >= 0 || p.gop <= 0 || p.gep < 0)
    return vg_fail(c, VG_ERR_INVALID, "aligner needs match>0, mismatch<0, gop>0, gep>=0 (got %d/%d/%d/%d)", p.match,
                   p.mismatch, p.gop, p.gep);
  if (p.match - p.mismatch > 250 || p.gep > 8000 || p.gop > 8000)
    return vg_fail(c, VG_ERR_INVALID, "aligner parameters out of the int16 kernel's range");
  if (sw_pick_R(mmax) < 0) return vg_fail(c, VG_ERR_INVALID, "sequence length %d exceeds the supported 1024", mmax);
  if ((int64_t)p.match * std::min(mmax, nmax) + p.match + 1 > 32000)
    return vg_fail(c, VG_ERR_INVALID, "alignment scores would exceed int16 (%d x %d)", mmax, nmax);
  return VG_OK;
}

template <int R, bool COMPACT>
static int sw_launch_R(vg_ctx* c, const SwTask* d_tasks, int ntasks, const uint8_t* d_s1, const uint8_t* d_s2,
                       int nmax, bool traceback, SwFillOut* d_fill, SwAln* d_alns, uint8_t* d_rev, int64_t revStride,
                       double* msFill, double* msTb) {
  using CkT = typename SwCk<COMPACT>::T;
  const int nW = nmax + 1;
  const int nT = (nmax + 32 + 3) & ~3;  // wavefront steps 1 .. nmax + 31, rounded to whole groups of 4
  const int nTb = nT / SW_TC + 1;     // column checkpoints at t = SW_TC, 2 SW_TC, ...
  const size_t perPair = ((size_t)nT * 32 + (size_t)nTb * 32 * R) * sizeof(CkT);
  const size_t perTask = perPair / 2;
  size_t budget = 0;
  if (const char* e = getenv("VG_CKPT_BYTES")) budget = (size_t)atoll(e);
  else if (traceback) {
    if (!c->budgetCkpt) c->budgetCkpt = vg_mem_budget((size_t)24 << 30, 0.6, c->d_rowck.cap + c->d_colck.cap);
    budget = c->budgetCkpt;
  }
  int64_t chunk = ntasks;
  if (traceback) {
    chunk = std::min<int64_t>(ntasks, std::max<int64_t>(2, (int64_t)(budget / perTask)));
    chunk &= ~(int64_t)1;
    if (chunk < 2) chunk = 2;
    VG_CUDA(c, c->d_rowck.reserve((size_t)(chunk / 2) * nT * 32 * sizeof(CkT)));
    VG_CUDA(c, c->d_colck.reserve((size_t)(chunk / 2) * nTb * 32 * R * sizeof(CkT)));
  }
  VG_CUDA(c, c->d_counter.reserve(sizeof(int)));
  const int threads = 256, warps = threads / 32;
  const size_t smem = (size_t)warps * nW * sizeof(uint32_t);
  int perSm = 1;
  if (traceback) {
    VG_CUDA(c, cudaFuncSetAttribute(sw_fill_kernel<R, true, COMPACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VG_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, sw_fill_kernel<R, true, COMPACT>, threads, smem));
  } else {
    VG_CUDA(c, cudaFuncSetAttribute(sw_fill_kernel<R, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VG_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, sw_fill_kernel<R, false, false>, threads, smem));
  }
  if (perSm < 1) return vg_fail(c, VG_ERR_INVALID, "window length %d needs too much shared memory", nmax);
  float fms = 0, tms = 0;
  for (int64_t s = 0; s < ntasks; s += chunk) {
    const int nt = (int)std::min<int64_t>(chunk, ntasks - s);
    const int npairs = (nt + 1) / 2;
    const int blocks = std::max(1, std::min((npairs + warps - 1) / warps, c->smCount * perSm));
    VG_CUDA(c, cudaMemsetAsync(c->d_counter.p, 0, sizeof(int), c->stream));
    VG_CUDA(c, cudaEventRecord(c->ev[4], c->stream));
    if (traceback)
      sw_fill_kernel<R, true, COMPACT><<<blocks, threads, smem, c->stream>>>(d_tasks + s, nt, d_s1, d_s2, c->sw,
                                                                             c->d_rowck.as<CkT>(), c->d_colck.as<CkT>(), nT,
                                                                             nTb, d_fill + s, c->d_counter.as<int>(), nW);
    else
      sw_fill_kernel<R, false, false><<<blocks, threads, smem, c->stream>>>(d_tasks + s, nt, d_s1, d_s2, c->sw, nullptr,
                                                                            nullptr, nT, nTb, d_fill + s,
                                                                            c->d_counter.as<int>(), nW);
    VG_LAUNCH_CHECK(c);
    VG_CUDA(c, cudaEventRecord(c->ev[5], c->stream));
    if (traceback) {
      sw_traceback_kernel<R, COMPACT><<<(nt + 127) / 128, 128, 0, c->stream>>>(d_tasks + s, nt, d_s1, d_s2, c->sw,
                                                                               c->d_rowck.as<CkT>(), c->d_colck.as<CkT>(), nT,
                                                                               nTb, d_fill + s, d_alns, d_rev, revStride);
      VG_LAUNCH_CHECK(c);
    }
    VG_CUDA(c, cudaEventRecord(c->ev[6], c->stream));
    VG_CUDA(c, cudaEventSynchronize(c->ev[6]));
    float a = 0, b = 0;
    cudaEventElapsedTime(&a, c->ev[4], c->ev[5]);
    cudaEventElapsedTime(&b, c->ev[5], c->ev[6]);
    fms += a;
    tms += b;
  }
  if (msFill) *msFill += fms;
  if (msTb) *msTb += tms;
  return VG_OK;
}

// Runs fill (+ traceback) over device-resident tasks. d_alns/d_rev are indexed by SwTask.out.
static int sw_run(vg_ctx* c, const SwTask* d_tasks, int64_t ntasks, const uint8_t* d_s1, const uint8_t* d_s2, int mmax,
                  int nmax, bool traceback, SwFillOut* d_fill, SwAln* d_alns, uint8_t* d_rev, int64_t revStride,
                  double* msFill, double* msTb) {
  if (ntasks == 0) return VG_OK;
  if (ntasks > 0x7fffffff) return vg_fail(c, VG_ERR_INVALID, "too many alignment tasks in one call");
  VG_TRY(sw_validate(c, mmax, nmax));
  const int R = sw_pick_R(mmax);
  // 4-byte checkpoints (V in 12 bits + the 4 bits of the gap state that matter, see SwCk) when the scores and the
  // penalties allow it, else 8-byte ones
  const bool compact = traceback && (int64_t)c->sw.match * std::min(mmax, nmax) + c->sw.match + 1 < 4096 && c->sw.gop >= c->sw.gep &&
                       c->sw.gop - c->sw.gep <= 15 && !getenv("VG_SW_WIDE");
#define SW_CASE(RR)                                                                                                   \
  case RR:                                                                                                            \
    return compact ? sw_launch_R<RR, true>(c, d_tasks, (int)ntasks, d_s1, d_s2, nmax, traceback, d_fill, d_alns, d_rev,  \
                                           revStride, msFill, msTb)                                                    \
                   : sw_launch_R<RR, false>(c, d_tasks, (int)ntasks, d_s1, d_s2, nmax, traceback, d_fill, d_alns, d_rev, \
                                            revStride, msFill, msTb);
  switch (R) {
    SW_CASE(2)
    SW_CASE(3)
    SW_CASE(4)
    SW_CASE(5)
    SW_CASE(6)
    SW_CASE(8)
    SW_CASE(10)
    SW_CASE(12)
    SW_CASE(16)
    SW_CASE(24)
    SW_CASE(32)
  }
#undef SW_CASE
  return vg_fail(c, VG_ERR_INVALID, "no kernel for read length %d", mmax);
}

// Host driver of the Smith-Waterman kernels (chunking, R dispatch, launches).
#pragma once
#include <stdlib.h>

#include "ctx.cuh"
#include "sw.cuh"

// Alignment geometry by read length: W lanes per pair of alignments (32 / W pairs per warp), R rows per lane; a bin holds the
// reads of at most W * R bases. Short reads get narrow groups (full lanes, a skew of W - 1 steps), never thin strips: the
// traceback crosses |s1| / R strips whatever the read length.
// Measured per read length (tools/sw_len_probe.py, profiles/sw_r2.md): (8, 8) 3.0x, (16, 6) 1.9x, (16, 8) 1.6x the fill rate of
// (32, 6) for their lengths; (32, 5) and (32, 7) were no better than (32, 6) and (32, 8): a step costs about the same for 5 rows
// as for 6.
#define SW_NBINS 10
static const int sw_binW_host[SW_NBINS] = {8, 16, 16, 32, 32, 32, 32, 32, 32, 32};
static const int sw_binR_host[SW_NBINS] = {8, 6, 8, 6, 8, 10, 12, 16, 24, 32};
static inline int sw_bin_maxm(int b) { return sw_binW_host[b] * sw_binR_host[b]; }
static inline int sw_pick_bin(int mmax, int minBin = 0) {
  for (int b = minBin; b < SW_NBINS; b++)
    if (sw_bin_maxm(b) >= mmax) return b;
  return -1;
}

static inline int sw_validate(vg_ctx* c, int mmax, int nmax) {
  const SwParams& p = c->sw;
  if (p.match <= 0 || p.mismatch >= 0 || p.gop <= 0 || p.gep < 0)
    return vg_fail(c, VG_ERR_INVALID, "aligner needs match>0, mismatch<0, gop>0, gep>=0 (got %d/%d/%d/%d)", p.match,
                   p.mismatch, p.gop, p.gep);
  if (p.match - p.mismatch > 250 || p.gep > 8000 || p.gop > 8000)
    return vg_fail(c, VG_ERR_INVALID, "aligner parameters out of the int16 kernel's range");
  if (sw_pick_bin(mmax) < 0) return vg_fail(c, VG_ERR_INVALID, "sequence length %d exceeds the supported 1024", mmax);
  if ((int64_t)p.match * std::min(mmax, nmax) + p.match + 1 > 32000)
    return vg_fail(c, VG_ERR_INVALID, "alignment scores would exceed int16 (%d x %d)", mmax, nmax);
  return VG_OK;
}

template <int R, bool COMPACT, int W>
static int sw_launch_R(vg_ctx* c, const SwTask* d_tasks, int ntasks, const uint8_t* d_s1, const uint8_t* d_s2,
                       int nmax, bool traceback, SwFillOut* d_fill, SwAln* d_alns, uint8_t* d_rev, int64_t revStride,
                       double* msFill, double* msTb) {
  using CkT = typename SwCk<COMPACT>::T;
  const int nW = nmax + 1;
  const int nT = (nmax + W + 3) & ~3;  // wavefront steps 1 .. nmax + W - 1, rounded to whole groups of 4
  const int nTb = nT / SW_TC + 1;     // column checkpoints at t = SW_TC, 2 SW_TC, ...
  const size_t perPair = ((size_t)nT * W + (size_t)nTb * W * R) * sizeof(CkT);
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
    VG_CUDA(c, c->d_rowck.reserve((size_t)(chunk / 2) * nT * W * sizeof(CkT)));
    VG_CUDA(c, c->d_colck.reserve((size_t)(chunk / 2) * nTb * W * R * sizeof(CkT)));
  }
  VG_CUDA(c, c->d_counter.reserve(sizeof(int)));
  constexpr int NG = 32 / W;  // pairs of alignments per warp
  const int threads = 256, warps = threads / 32;
  const size_t smem = (size_t)warps * NG * nW * sizeof(uint32_t);
  int perSm = 1;
  if (traceback) {
    VG_CUDA(c, cudaFuncSetAttribute(sw_fill_kernel<R, true, COMPACT, false, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VG_CUDA(c, cudaFuncSetAttribute(sw_fill_kernel<R, true, COMPACT, COMPACT, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VG_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, sw_fill_kernel<R, true, COMPACT, false, W>, threads, smem));
  } else {
    VG_CUDA(c, cudaFuncSetAttribute(sw_fill_kernel<R, false, false, false, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VG_CUDA(c, cudaFuncSetAttribute(sw_fill_kernel<R, false, false, true, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VG_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, sw_fill_kernel<R, false, false, false, W>, threads, smem));
  }
  if (perSm < 1) return vg_fail(c, VG_ERR_INVALID, "window length %d needs too much shared memory", nmax);
  // the diagonal term on the FMA pipe (sw.cuh, FASTF) needs mismatch + gop >= 0; instantiated for the compact-checkpoint
  // and the score-only kernels (the defaults)
  const bool fastf = c->sw.mismatch + c->sw.gop >= 0 && (COMPACT || !traceback) && !getenv("VG_SW_SLOWF");
  float fms = 0, tms = 0;
  for (int64_t s = 0; s < ntasks; s += chunk) {
    const int nt = (int)std::min<int64_t>(chunk, ntasks - s);
    const int npairs = (nt + 1) / 2;
    const int blocks = std::max(1, std::min((npairs + warps * NG - 1) / (warps * NG), c->smCount * perSm));
    VG_CUDA(c, cudaMemsetAsync(c->d_counter.p, 0, sizeof(int), c->stream));
    VG_CUDA(c, cudaEventRecord(c->ev[4], c->stream));
    if (traceback && fastf)  // FASTF is only instantiated with COMPACT (the last argument is false for wide checkpoints)
      sw_fill_kernel<R, true, COMPACT, COMPACT, W><<<blocks, threads, smem, c->stream>>>(d_tasks + s, nt, d_s1, d_s2, c->sw,
                                                                                      c->d_rowck.as<CkT>(), c->d_colck.as<CkT>(), nT,
                                                                                      nTb, d_fill + s, c->d_counter.as<int>(), nW);
    else if (traceback)
      sw_fill_kernel<R, true, COMPACT, false, W><<<blocks, threads, smem, c->stream>>>(d_tasks + s, nt, d_s1, d_s2, c->sw,
                                                                                    c->d_rowck.as<CkT>(), c->d_colck.as<CkT>(), nT,
                                                                                    nTb, d_fill + s, c->d_counter.as<int>(), nW);
    else if (fastf)
      sw_fill_kernel<R, false, false, true, W><<<blocks, threads, smem, c->stream>>>(d_tasks + s, nt, d_s1, d_s2, c->sw, nullptr,
                                                                                  nullptr, nT, nTb, d_fill + s,
                                                                                  c->d_counter.as<int>(), nW);
    else
      sw_fill_kernel<R, false, false, false, W><<<blocks, threads, smem, c->stream>>>(d_tasks + s, nt, d_s1, d_s2, c->sw, nullptr,
                                                                                   nullptr, nT, nTb, d_fill + s,
                                                                                   c->d_counter.as<int>(), nW);
    VG_LAUNCH_CHECK(c);
    VG_CUDA(c, cudaEventRecord(c->ev[5], c->stream));
    if (traceback) {
      sw_traceback_kernel<R, COMPACT, W><<<(nt + 127) / 128, 128, 0, c->stream>>>(d_tasks + s, nt, d_s1, d_s2, c->sw,
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

// ---- length binning -----------------------------------------------------------------------------------------------------
// The fill kernel gives every lane R consecutive rows, R from the longest read of the launch, and packs two alignments
// into one warp whose wavefront runs max(n_a, n_b) + lanes steps: a ragged batch (trimmed Illumina reads, windows cut
// by contig ends) wastes lanes and steps. The tasks are therefore counting-sorted on the device by (R bin of |s1|, |s2|):
// every bin is launched with its own R and checkpoint geometry, partners inside a bin have (nearly) equal |s2|, and the
// 32 alignments of a traceback warp walk paths of similar length. Results are addressed by SwTask.out, so the order in
// which tasks run is invisible to the caller.
#define SW_NKEYS 1024  // |s2| buckets per bin
__constant__ int sw_binMaxM[SW_NBINS] = {64, 96, 128, 192, 256, 320, 384, 512, 768, 1024};

__device__ __forceinline__ int sw_task_key(const SwTask& t, int nShift, int minBin) {
  int b = 0;
#pragma unroll
  for (int k = 0; k < SW_NBINS - 1; k++) b += (sw_binMaxM[k] < t.m) ? 1 : 0;
  b = max(b, minBin);
  // inside a bin: eight groups of |s1| (partners use the same lanes), then 128 buckets of |s2|
  const int mg = min(7, max(t.m - 1, 0) * 8 / sw_binMaxM[b]);
  return b * SW_NKEYS + mg * 128 + min(t.n >> nShift, 127);
}
// hist[key]++, binMaxN[bin] = max |s2|, binMaxM[bin] = max |s1|. The batches of the mapping pipeline have nearly uniform
// lengths (a handful of hot keys), so the atomics are aggregated per warp: one per distinct key of the warp.
__global__ void sw_bin_hist_kernel(const SwTask* __restrict__ tasks, int n, int nShift, int minBin, int* __restrict__ hist,
                                   int* __restrict__ binMaxN, int* __restrict__ binMaxM) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool ok = i < n;
  SwTask t;
  t.m = t.n = 0;
  if (ok) t = tasks[i];
  const int key = ok ? sw_task_key(t, nShift, minBin) : -1 - (int)(threadIdx.x & 31);
  const unsigned grp = __match_any_sync(VG_FULL, key);
  const int gn = __reduce_max_sync(grp, t.n), gm = __reduce_max_sync(grp, t.m);
  if (ok && (int)(threadIdx.x & 31) == __ffs(grp) - 1) {
    atomicAdd(hist + key, __popc(grp));
    atomicMax(binMaxN + key / SW_NKEYS, gn);
    atomicMax(binMaxM + key / SW_NKEYS, gm);
  }
}
// exclusive scan of hist[SW_NBINS * SW_NKEYS] by one block of 1024 threads (SW_NBINS keys per thread); binStart[b] = first task
// of bin b, binStart[SW_NBINS] = n
__global__ void sw_bin_scan_kernel(int* __restrict__ hist, int* __restrict__ binStart) {
  __shared__ int part[1024];
  const int t = threadIdx.x;
  int loc[SW_NBINS], sum = 0;
#pragma unroll
  for (int k = 0; k < SW_NBINS; k++) {
    loc[k] = hist[t * SW_NBINS + k];
    sum += loc[k];
  }
  part[t] = sum;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {
    const int v = (t >= d) ? part[t - d] : 0;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  int run = part[t] - sum;
#pragma unroll
  for (int k = 0; k < SW_NBINS; k++) {
    const int key = t * SW_NBINS + k;
    if (key % SW_NKEYS == 0) binStart[key / SW_NKEYS] = run;
    hist[key] = run;
    run += loc[k];
  }
  if (t == 1023) binStart[SW_NBINS] = run;
}
__global__ void sw_bin_scatter_kernel(const SwTask* __restrict__ tasks, int n, int nShift, int minBin, int* __restrict__ cursor,
                                      SwTask* __restrict__ sorted) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool ok = i < n;
  SwTask t;
  t.m = t.n = 0;
  if (ok) t = tasks[i];
  const int key = ok ? sw_task_key(t, nShift, minBin) : -1 - lane;
  const unsigned grp = __match_any_sync(VG_FULL, key);
  const int leader = __ffs(grp) - 1;
  int base = 0;
  if (ok && lane == leader) base = atomicAdd(cursor + key, __popc(grp));
  base = __shfl_sync(VG_FULL, base, leader);
  if (ok) sorted[base + __popc(grp & ((1u << lane) - 1u))] = t;
}

// Runs fill (+ traceback) over device-resident tasks. d_alns/d_rev are indexed by SwTask.out. With traceback the tasks
// run in length bins (see above); d_fill is scratch then. Score-only keeps the caller's order (d_fill[i] = task i).
static int sw_run_bin(vg_ctx* c, const SwTask* d_tasks, int64_t ntasks, const uint8_t* d_s1, const uint8_t* d_s2, int mmax,
                      int nmax, bool traceback, SwFillOut* d_fill, SwAln* d_alns, uint8_t* d_rev, int64_t revStride,
                      double* msFill, double* msTb, int minBin = 0) {
  int bin = sw_pick_bin(mmax, minBin);
  // a group of W lanes keeps its window's symbols in shared memory: 32 / W windows per warp. A short read against a very long
  // window (vg_sw_align_batch is free to ask for it) takes the next wider geometry instead of running out of shared memory.
  while (bin >= 0 && bin < SW_NBINS - 1 && sw_binW_host[bin] < 32 &&
         (size_t)8 * (32 / sw_binW_host[bin]) * (size_t)(nmax + 1) * 4 > (size_t)96 * 1024)
    bin++;
  // 4-byte checkpoints (V in 12 bits + the 4 bits of the gap state that matter, see SwCk) when the scores and the
  // penalties allow it, else 8-byte ones
  const bool compact = traceback && (int64_t)c->sw.match * std::min(mmax, nmax) + c->sw.match + 1 < 4096 && c->sw.gop >= c->sw.gep &&
                       c->sw.gop - c->sw.gep <= 15 && !getenv("VG_SW_WIDE");
#define SW_CASE(BIN, WW, RR)                                                                                                  \
  case BIN:                                                                                                                   \
    return compact ? sw_launch_R<RR, true, WW>(c, d_tasks, (int)ntasks, d_s1, d_s2, nmax, traceback, d_fill, d_alns, d_rev,     \
                                               revStride, msFill, msTb)                                                        \
                   : sw_launch_R<RR, false, WW>(c, d_tasks, (int)ntasks, d_s1, d_s2, nmax, traceback, d_fill, d_alns, d_rev,    \
                                                revStride, msFill, msTb);
  switch (bin) {  // (W, R) of sw_binW_host / sw_binR_host
    SW_CASE(0, 8, 8)
    SW_CASE(1, 16, 6)
    SW_CASE(2, 16, 8)
    SW_CASE(3, 32, 6)
    SW_CASE(4, 32, 8)
    SW_CASE(5, 32, 10)
    SW_CASE(6, 32, 12)
    SW_CASE(7, 32, 16)
    SW_CASE(8, 32, 24)
    SW_CASE(9, 32, 32)
  }
#undef SW_CASE
  return vg_fail(c, VG_ERR_INVALID, "no kernel for read length %d", mmax);
}

static int sw_run(vg_ctx* c, const SwTask* d_tasks, int64_t ntasks, const uint8_t* d_s1, const uint8_t* d_s2, int mmax,
                  int nmax, bool traceback, SwFillOut* d_fill, SwAln* d_alns, uint8_t* d_rev, int64_t revStride,
                  double* msFill, double* msTb) {
  if (ntasks == 0) return VG_OK;
  if (ntasks > 0x7fffffff) return vg_fail(c, VG_ERR_INVALID, "too many alignment tasks in one call");
  VG_TRY(sw_validate(c, mmax, nmax));
  if (!traceback || getenv("VG_SW_NOBIN"))
    return sw_run_bin(c, d_tasks, ntasks, d_s1, d_s2, mmax, nmax, traceback, d_fill, d_alns, d_rev, revStride, msFill, msTb);
  // counting sort by (R bin, |s2|)
  const int n = (int)ntasks;
  int nShift = 0;
  while ((nmax >> nShift) >= 128) nShift++;
  // VG_SW_MINR=r (experiments): no groups narrower than a warp and no strips thinner than r rows
  int minBin = 0;
  if (const char* e = getenv("VG_SW_MINR")) {
    while (minBin < SW_NBINS - 1 && (sw_binW_host[minBin] < 32 || sw_binR_host[minBin] < atoi(e))) minBin++;
  }
  const size_t histBytes = (size_t)SW_NBINS * SW_NKEYS * 4;
  VG_CUDA(c, c->d_swBins.reserve(histBytes + 256));
  VG_CUDA(c, c->d_swSorted.reserve((size_t)n * sizeof(SwTask)));
  int* d_hist = c->d_swBins.as<int>();
  int* d_binStart = d_hist + SW_NBINS * SW_NKEYS;  // [SW_NBINS + 1], then binMaxN[SW_NBINS], binMaxM[SW_NBINS]
  int* d_binMaxN = d_binStart + SW_NBINS + 1;
  int* d_binMaxM = d_binMaxN + SW_NBINS;
  VG_CUDA(c, cudaMemsetAsync(d_hist, 0, histBytes + 256, c->stream));
  sw_bin_hist_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(d_tasks, n, nShift, minBin, d_hist, d_binMaxN, d_binMaxM);
  VG_LAUNCH_CHECK(c);
  sw_bin_scan_kernel<<<1, 1024, 0, c->stream>>>(d_hist, d_binStart);
  VG_LAUNCH_CHECK(c);
  sw_bin_scatter_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(d_tasks, n, nShift, minBin, d_hist, c->d_swSorted.as<SwTask>());
  VG_LAUNCH_CHECK(c);
  int host[3 * SW_NBINS + 1];
  VG_CUDA(c, cudaMemcpyAsync(host, d_binStart, sizeof host, cudaMemcpyDeviceToHost, c->stream));
  VG_CUDA(c, cudaStreamSynchronize(c->stream));
  const int* binStart = host;
  const int* binMaxN = host + SW_NBINS + 1;
  const int* binMaxM = binMaxN + SW_NBINS;
  if (binStart[SW_NBINS] != n) return vg_fail(c, VG_ERR_CUDA, "alignment task binning lost tasks (%d of %d)", binStart[SW_NBINS], n);
  for (int b = 0; b < SW_NBINS; b++) {
    const int cnt = binStart[b + 1] - binStart[b];
    if (cnt <= 0) continue;
    // a bin's tasks all have |s1| <= 32 R of the bin; empty reads (m == 0) sit in the first bin
    const int mBin = std::max((b ? sw_bin_maxm(b - 1) : 0) + 1, std::min(binMaxM[b], sw_bin_maxm(b)));
    VG_TRY(sw_run_bin(c, c->d_swSorted.as<SwTask>() + binStart[b], cnt, d_s1, d_s2, mBin, std::max(1, binMaxN[b]), true,
                      d_fill + binStart[b], d_alns, d_rev, revStride, msFill, msTb, b));
  }
  return VG_OK;
}

// Host orchestration of the mapping pipeline: reference/reads upload, Mapper.mapReads,
// MapperMSA.mapAllReads, result retrieval, pileup and consensus.
#pragma once
#include "ctx.cuh"
#include "pileup.cuh"
#include "bam.cuh"
#include "identity.cuh"
#include "seed_host.cuh"
#include "sw_host.cuh"

// ---- small utility kernels ------------------------------------------------------------------------
__global__ void mh_iota_kernel(int64_t* a, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = i;
}

// Bytes the Java reference would throw on: >= 127 (Constants.complement has 127 entries, bytes are
// signed) and ('T','a') exclusive (Constants.nToI has 'T'+1 entries; such a byte aligned on the diagonal
// makes the pileup throw).
__global__ void mh_validate_kernel(const uint8_t* __restrict__ b, int64_t n, int* __restrict__ flag) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int bad = 0;
  for (; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint8_t c = b[i];
    bad |= (c >= 127) || (c > 'T' && c < 'a');
  }
  if (__any_sync(VG_FULL, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

// ---- packed reads (north star: 2-bit reads; SURVEY 8b "vg_upload_reads ... or /4 packed + mask") -------------------------
// Host format: 2 bits per base (A=0, C=1, G=2, T=3), base i in bits 2 (i & 3) of byte i >> 2, plus a list of exceptions
// (position, byte) for everything that is not ACGT ('N', '-', the '*' of an absent mate, IUPAC codes). The device
// unpacks into the 1 B/base pool the kernels read: the alphabet of the path is bytes, not 2-bit codes (any byte takes part
// in byte equality), so the packing is a transfer format — 4x fewer PCIe bytes — not a change of the kernels' input.
__global__ void mh_unpack_kernel(const uint32_t* __restrict__ packed, int64_t nWords, uint4* __restrict__ bases) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nWords) return;
  const uint32_t w = packed[i];
  uint32_t o[4];
#pragma unroll
  for (int q = 0; q < 4; q++) {
    uint32_t v = 0;
#pragma unroll
    for (int b = 0; b < 4; b++) {
      const uint32_t code = (w >> (8 * q + 2 * b)) & 3u;
      const uint32_t ch = (0x54474341u >> (8 * code)) & 0xFFu;  // "ACGT"
      v |= ch << (8 * b);
    }
    o[q] = v;
  }
  bases[i] = make_uint4(o[0], o[1], o[2], o[3]);
}
__global__ void mh_exceptions_kernel(const int64_t* __restrict__ pos, const uint8_t* __restrict__ byte, int64_t n,
                                     int64_t nBases, uint8_t* __restrict__ bases, int* __restrict__ flag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t p = pos[i];
  if (p < 0 || p >= nBases) atomicOr(flag, 2);
  else bases[p] = byte[i];
}

// ---- compact results: sequence1 as run-length edit operations -----------------------------------------------------------
// Alignment.sequence1 (SmithWatermanGotoh.java:157-222) holds, per alignment column, the read base (DIAGONAL), the lower-
// cased read base (UP: insertion in the read) or '-' (LEFT: deletion). Given the read, its strand and start1, the string is
// therefore a function of the run lengths of these three column classes: len << 4 | op with op 0 = aligned base, 1 =
// insertion, 2 = deletion (the BAM op codes). One warp per alignment; EMIT = false counts the runs.
template <bool EMIT>
__global__ void mh_ops_kernel(const vg_pair_result* __restrict__ res, int64_t n2, const uint8_t* __restrict__ pool,
                              int32_t* __restrict__ nOps, const int64_t* __restrict__ opOff, uint32_t* __restrict__ ops) {
  const int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (q >= n2) return;
  const vg_mate& m = res[q >> 1].m[q & 1];
  if (!m.present || m.seq1_offset < 0) {
    if (!EMIT && lane == 0) nOps[q] = 0;
    return;
  }
  const int len = m.length;
  const uint8_t* s = pool + m.seq1_offset;
  uint32_t* out = EMIT ? ops + opOff[q] : nullptr;
  const unsigned lt = (1u << lane) - 1u;
  int n = 0, openPos = 0, openCls = 0, prevCls = -1;
  for (int t0 = 0; t0 < len; t0 += 32) {
    const int t = t0 + lane;
    const uint8_t ch = (t < len) ? s[t] : 0;
    const int cls = (ch == '-') ? 2 : (ch < 'a') ? 0 : 1;
    int p = __shfl_up_sync(VG_FULL, cls, 1);
    if (lane == 0) p = prevCls;
    const bool st = (t < len) && cls != p;
    const unsigned mk = __ballot_sync(VG_FULL, st);
    prevCls = __shfl_sync(VG_FULL, cls, 31);
    if (mk) {
      if (EMIT) {
        if (n > 0 && lane == 0) out[n - 1] = ((uint32_t)(t0 + __ffs(mk) - 1 - openPos) << 4) | (uint32_t)openCls;
        const unsigned higher = mk & ~((2u << lane) - 1u);
        if (st && higher) out[n + __popc(mk & lt)] = ((uint32_t)(__ffs(higher) - 1 - lane) << 4) | (uint32_t)cls;
      }
      const int last = 31 - __clz(mk);
      openPos = t0 + last;
      openCls = __shfl_sync(VG_FULL, cls, last);
      n += __popc(mk);
    }
  }
  if (EMIT) {
    if (n > 0 && lane == 0) out[n - 1] = ((uint32_t)(len - openPos) << 4) | (uint32_t)openCls;
  } else if (lane == 0) {
    nOps[q] = n;
  }
}
// the caller's copy of the records: seq1_offset -> first operation, pad_ -> number of operations
__global__ void mh_compact_records_kernel(const vg_pair_result* __restrict__ res, int64_t n2, const int32_t* __restrict__ nOps,
                                          const int64_t* __restrict__ opOff, vg_pair_result* __restrict__ out) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n2) return;
  if ((q & 1) == 0) {
    out[q >> 1].cls = res[q >> 1].cls;
    out[q >> 1].pad_ = 0;
  }
  vg_mate m = res[q >> 1].m[q & 1];
  if (m.present && m.seq1_offset >= 0) {
    m.seq1_offset = opOff[q];
    m.pad_ = nOps[q];
  } else {
    m.pad_ = 0;
  }
  out[q >> 1].m[q & 1] = m;
}

// exclusive scan int32 -> int64, three kernels
__global__ void mh_scan1_kernel(const int32_t* __restrict__ in, int64_t* __restrict__ out, int64_t* __restrict__ blockSums, int64_t n) {
  __shared__ int64_t wsum[32];
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int64_t v = (i < n) ? in[i] : 0, x = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int64_t y = __shfl_up_sync(VG_FULL, x, d);
    if (lane >= d) x += y;
  }
  if (lane == 31) wsum[w] = x;
  __syncthreads();
  if (w == 0) {
    int64_t s = wsum[lane], t = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int64_t y = __shfl_up_sync(VG_FULL, t, d);
      if (lane >= d) t += y;
    }
    wsum[lane] = t - s;
    if (lane == 31) blockSums[blockIdx.x] = t;
  }
  __syncthreads();
  if (i < n) out[i] = wsum[w] + x - v;
}
__global__ void mh_scan2_kernel(int64_t* blockSums, int64_t nb, int64_t* total) {
  // single thread block of 1 warp: sequential over chunks of 32 with carry (nb is small)
  const int lane = threadIdx.x;
  int64_t carry = 0;
  for (int64_t b0 = 0; b0 < nb; b0 += 32) {
    int64_t v = (b0 + lane < nb) ? blockSums[b0 + lane] : 0, x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int64_t y = __shfl_up_sync(VG_FULL, x, d);
      if (lane >= d) x += y;
    }
    if (b0 + lane < nb) blockSums[b0 + lane] = carry + x - v;
    carry += __shfl_sync(VG_FULL, x, 31);
  }
  if (lane == 0) *total = carry;
}
__global__ void mh_scan3_kernel(int64_t* __restrict__ out, const int64_t* __restrict__ blockSums, const int32_t* __restrict__ in,
                                int64_t n, int plain) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (plain || in[i] > 0) ? out[i] + blockSums[blockIdx.x] : -1;  // -1: nothing to emit (unless plain)
}

// Mapper.java:422-426 (empty reference): every pair goes to `unmapped`, r1 = r2 = null
__global__ void mh_unmapped_kernel(vg_pair_result* __restrict__ res, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  vg_pair_result r;
  memset(&r, 0, sizeof r);
  r.cls = VG_CLS_UNMAPPED;
  r.m[0].seq1_offset = r.m[1].seq1_offset = -1;
  res[i] = r;
}

__global__ void mh_lengths_kernel(const vg_pair_result* __restrict__ res, const SwAln* __restrict__ alns, int64_t n2,
                                  int32_t* __restrict__ lens) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n2) return;
  const vg_mate& m = res[q >> 1].m[q & 1];
  lens[q] = m.present ? alns[q].length : 0;
}

__global__ void mh_finalize_kernel(vg_pair_result* __restrict__ res, const SwAln* __restrict__ alns,
                                   const int64_t* __restrict__ seqOff, int64_t n2) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n2) return;
  vg_mate& m = res[q >> 1].m[q & 1];
  if (!m.present) return;
  const SwAln a = alns[q];
  m.score = a.score, m.start1 = a.start1, m.end1 = a.end1, m.start2 = a.start2, m.end2 = a.end2;
  m.identity = a.identity, m.length = a.length;
  // an empty alignment still has a (zero-length) sequence1; give it a valid offset
  m.seq1_offset = seqOff[q] >= 0 ? seqOff[q] : 0;
}

// Mapping stats 1)-7) (Mapper.java:446-473 / MapperMSA.java:335-357)
__global__ void mh_stats_kernel(const vg_pair_result* __restrict__ res, int64_t n, const int64_t* __restrict__ pairIdx,
                                const uint8_t* __restrict__ bases, const int64_t* __restrict__ off1,
                                const int32_t* __restrict__ len1, const int64_t* __restrict__ off2,
                                const int32_t* __restrict__ len2, int msa, unsigned long long* __restrict__ st) {
  __shared__ unsigned long long acc[9];
  if (threadIdx.x < 9) acc[threadIdx.x] = 0;
  __syncthreads();
  const int64_t pi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pi < n) {
    const int64_t p = pairIdx[pi];
    const vg_pair_result& r = res[pi];
    const bool e1 = !(len1[p] == 1 && bases[off1[p]] == '*'), e2 = !(len2[p] == 1 && bases[off2[p]] == '*');
    unsigned long long v[9] = {1, 0, 0, 0, 0, 0, 0, 0, 0};
    if (e1 && e2) v[1] = 1, v[2] = 2;
    else if (e1 || e2) v[2] = 1;
    long long score = 0;
    for (int k = 0; k < 2; k++)
      if (r.m[k].present) {
        v[3]++;
        score += msa ? r.m[k].count : r.m[k].score;
      }
    const bool conc = r.cls == VG_CLS_CONCORDANT_FR || r.cls == VG_CLS_CONCORDANT_RF;
    if (conc) {
      v[4] = 1, v[5] = 1, v[8] = 1;
    } else {
      for (int k = 0; k < 2; k++)
        if (r.m[k].present) (r.m[k].reverse ? v[5] : v[4])++;
    }
    if (conc || r.cls == VG_CLS_DISCORDANT) v[6] = 1;
    v[7] = (unsigned long long)score;
    for (int k = 0; k < 9; k++)
      if (v[k]) atomicAdd(&acc[k], v[k]);
  }
  __syncthreads();
  if (threadIdx.x < 9 && acc[threadIdx.x]) atomicAdd(&st[threadIdx.x], acc[threadIdx.x]);
}

static int mh_exclusive_scan(vg_ctx* c, const int32_t* d_in, int64_t* d_out, int64_t n, int64_t* totalOut, bool plain = false) {
  const int threads = 1024;
  const int64_t nb = (n + threads - 1) / threads;
  VG_CUDA(c, c->d_misc.reserve((size_t)(nb + 2) * 8 + 64));
  int64_t* bs = c->d_misc.as<int64_t>() + 1;
  int64_t* tot = c->d_misc.as<int64_t>();
  mh_scan1_kernel<<<(unsigned)nb, threads, 0, c->stream>>>(d_in, d_out, bs, n);
  VG_LAUNCH_CHECK(c);
  mh_scan2_kernel<<<1, 32, 0, c->stream>>>(bs, nb, tot);
  VG_LAUNCH_CHECK(c);
  mh_scan3_kernel<<<(unsigned)nb, threads, 0, c->stream>>>(d_out, bs, d_in, n, plain ? 1 : 0);
  VG_LAUNCH_CHECK(c);
  VG_CUDA(c, cudaMemcpyAsync(totalOut, tot, 8, cudaMemcpyDeviceToHost, c->stream));
  VG_CUDA(c, cudaStreamSynchronize(c->stream));
  return VG_OK;
}

static int mh_map(vg_ctx* c, bool msa, const int64_t* pair_idx, int64_t n_idx, vg_map_stats* stats) {
  if (!c) return VG_ERR_INVALID;
  VG_CUDA(c, cudaSetDevice(c->device));
  if (!c->haveReads) return vg_fail(c, VG_ERR_STATE, "vg_upload_reads has not been called");
  if (msa ? !c->haveMsa : !c->haveRef) return vg_fail(c, VG_ERR_STATE, msa ? "vg_set_msa_index has not been called" : "vg_set_reference has not been called");
  const int64_t n = pair_idx ? n_idx : c->nPairs;
  if (n < 0) return vg_fail(c, VG_ERR_INVALID, "negative pair count");
  c->nResults = n;
  c->seq1Bytes = 0;
  c->resultsMsa = msa;
  for (double& t : c->timings) t = 0;
  for (double& t : c->seedStats) t = 0;
  if (stats) memset(stats, 0, sizeof *stats);
  VG_CUDA(c, c->d_pairIdx.reserve((size_t)std::max<int64_t>(n, 1) * 8));
  if (pair_idx) {
    for (int64_t i = 0; i < n; i++)
      if (pair_idx[i] < 0 || pair_idx[i] >= c->nPairs) return vg_fail(c, VG_ERR_INVALID, "pair index %lld out of range", (long long)pair_idx[i]);
    VG_CUDA(c, cudaMemcpyAsync(c->d_pairIdx.p, pair_idx, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
  } else if (n > 0) {
    mh_iota_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->d_pairIdx.as<int64_t>(), n);
    VG_LAUNCH_CHECK(c);
  }
  VG_CUDA(c, c->d_results.reserve((size_t)std::max<int64_t>(n, 1) * sizeof(vg_pair_result)));
  if (n == 0) return VG_OK;
  const bool emptyRef = !msa && c->G == 0;  // Mapper.java:422-426: everything unmapped
  VG_CUDA(c, c->d_tasks.reserve((size_t)n * 2 * sizeof(SwTask)));
  VG_CUDA(c, c->d_ncand.reserve(64));
  // task counter + max window length live in d_fill's head (two ints), zeroed here
  VG_CUDA(c, c->d_alns.reserve((size_t)n * 2 * sizeof(SwAln) + 64));
  int* d_nTasks = nullptr;
  {
    VG_CUDA(c, c->d_s1off.reserve(64));
    d_nTasks = c->d_s1off.as<int>();
    VG_CUDA(c, cudaMemsetAsync(d_nTasks, 0, 16, c->stream));
  }
  const SeedConfig& sc = msa ? c->seedMsa : c->seed;
  const int range = msa ? c->msaLength : c->G;
  VG_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
  // ---- seeding + pair selection, in chunks of pairs -------------------------------------------------
  int maxWin = 64;
  // pruned windows do not overlap and an unclipped window is at least as wide as its read: at most range / L + 2 per
  // contig. Sized for the shortest read that can be seeded (the batch's windows buffer is chunked accordingly).
  if (c->maxReadLen > 0) {
    const int K = sc.K;
    const int lmin = std::max(std::max(K, 1), std::min(c->minReadLenGe[std::min(std::max(K, 0), 8)], c->maxReadLen));
    const int nCont = msa ? 1 : (int)c->contigEnds.size();
    maxWin = std::max(8, std::min(8192, std::max(range / std::max(1, c->maxReadLen / 2), range / lmin) + 2 * nCont + 4));
  }
  if (!c->budgetWin) c->budgetWin = vg_mem_budget(3ull << 30, 0.2, c->d_cand.cap);
  int64_t chunkPairs = std::max<int64_t>(1, std::min<int64_t>(n, (int64_t)(c->budgetWin / ((size_t)maxWin * 48 + 64))));
  if (const char* e = getenv("VG_MAP_CHUNK_PAIRS")) chunkPairs = std::max<int64_t>(1, std::min<int64_t>(n, atoll(e)));  // tests
  const int nf = (int)c->fragmentEnds.size();
  if (emptyRef) {
    mh_unmapped_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->d_results.as<vg_pair_result>(), n);
    VG_LAUNCH_CHECK(c);
  }
  for (int64_t s = 0; s < n && !emptyRef; s += chunkPairs) {
    const int64_t np = std::min(chunkPairs, n - s);
    VG_TRY(seed_launch(c, msa, c->d_bases.as<uint8_t>(), c->d_off1.as<int64_t>(), c->d_len1.as<int32_t>(),
                       c->d_off2.as<int64_t>(), c->d_len2.as<int32_t>(), c->d_pairIdx.as<int64_t>() + s, np * 4, false,
                       c->maxReadLen, maxWin, nullptr, nullptr));
    PairArgs P;
    memset(&P, 0, sizeof P);
    P.windows = c->d_cand.as<int32_t>();
    P.nWindows = c->d_ncand.as<int32_t>();
    P.maxWin = maxWin;
    P.nPairs = np;
    P.pairIdx = c->d_pairIdx.as<int64_t>() + s;
    P.off1 = c->d_off1.as<int64_t>(), P.len1 = c->d_len1.as<int32_t>();
    P.off2 = c->d_off2.as<int64_t>(), P.len2 = c->d_len2.as<int32_t>();
    P.fragmentEnds = c->d_fragmentEnds.as<int32_t>();
    P.nFragments = nf;
    P.isFragmented = (c->isFragmented && !msa) ? 1 : 0;
    P.insertLen = c->insertLen;
    P.msa = msa ? 1 : 0;
    P.results = c->d_results.as<vg_pair_result>() + s;
    P.tasks = c->d_tasks.as<SwTask>();
    P.nTasks = d_nTasks;
    P.outBase = s * 2;
    if (P.isFragmented) {
      VG_CUDA(c, c->d_s2off.reserve((size_t)np * 6 * nf * 4));
      P.fragScratch = c->d_s2off.as<int32_t>();
    }
    pair_select_kernel<<<(unsigned)((np + 127) / 128), 128, 0, c->stream>>>(P);
    VG_LAUNCH_CHECK(c);
  }
  VG_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
  struct { int nTasks, maxN; unsigned long long cells; } hostCnt = {0, 0, 0};
  VG_CUDA(c, cudaMemcpyAsync(&hostCnt, d_nTasks, 16, cudaMemcpyDeviceToHost, c->stream));
  VG_CUDA(c, cudaStreamSynchronize(c->stream));
  const int64_t nTasks = hostCnt.nTasks;
  const int nmaxWinLen = hostCnt.maxN;
  c->timings[5] = (double)hostCnt.cells;
  (void)sc;
  // ---- alignment ------------------------------------------------------------------------------------
  if (!msa && !emptyRef) {
    const int mmax = c->maxReadLen;
    const int64_t revStride = mmax + nmaxWinLen + 1;
    VG_CUDA(c, cudaMemsetAsync(c->d_alns.p, 0, (size_t)n * 2 * sizeof(SwAln), c->stream));
    VG_CUDA(c, c->d_rev.reserve((size_t)n * 2 * revStride + 64));
    VG_CUDA(c, c->d_fill.reserve((size_t)std::max<int64_t>(nTasks, 1) * sizeof(SwFillOut)));
    VG_TRY(sw_run(c, c->d_tasks.as<SwTask>(), nTasks, c->d_bases.as<uint8_t>(), c->d_ref.as<uint8_t>(), mmax, nmaxWinLen,
                  true, c->d_fill.as<SwFillOut>(), c->d_alns.as<SwAln>(), c->d_rev.as<uint8_t>(), revStride,
                  &c->timings[1], &c->timings[2]));
    // sequence1 pool offsets = exclusive prefix sum of the alignment lengths in (pair, mate) order
    VG_CUDA(c, c->d_s2off.reserve((size_t)n * 2 * 4 + 64));
    int32_t* d_lens = c->d_s2off.as<int32_t>();
    VG_CUDA(c, c->d_seqOff.reserve((size_t)n * 2 * 8 + 64));
    mh_lengths_kernel<<<(unsigned)((n * 2 + 255) / 256), 256, 0, c->stream>>>(c->d_results.as<vg_pair_result>(),
                                                                              c->d_alns.as<SwAln>(), n * 2, d_lens);
    VG_LAUNCH_CHECK(c);
    int64_t total = 0;
    VG_TRY(mh_exclusive_scan(c, d_lens, c->d_seqOff.as<int64_t>(), n * 2, &total));
    c->seq1Bytes = total;
    VG_CUDA(c, c->d_seq1pool.reserve((size_t)total + 64));
    VG_CUDA(c, c->d_rowpool.reserve((size_t)total + 64));
    VG_CUDA(c, c->d_rowLen.reserve((size_t)n * 2 * 4 + 64));
    pu_emit_kernel<<<(unsigned)((n * 2 * 32 + 255) / 256), 256, 0, c->stream>>>(
        c->d_alns.as<SwAln>(), n * 2, c->d_rev.as<uint8_t>(), revStride, c->d_seqOff.as<int64_t>(),
        c->d_seq1pool.as<uint8_t>(), c->d_rowpool.as<uint8_t>(), c->d_rowLen.as<int32_t>());
    VG_LAUNCH_CHECK(c);
    mh_finalize_kernel<<<(unsigned)((n * 2 + 255) / 256), 256, 0, c->stream>>>(c->d_results.as<vg_pair_result>(),
                                                                               c->d_alns.as<SwAln>(), c->d_seqOff.as<int64_t>(),
                                                                               n * 2);
    VG_LAUNCH_CHECK(c);
  }
  VG_CUDA(c, cudaEventRecord(c->ev[2], c->stream));
  // ---- stats ----------------------------------------------------------------------------------------
  VG_CUDA(c, c->d_ncand.reserve(9 * 8 + 64));
  unsigned long long* d_st = c->d_ncand.as<unsigned long long>();
  VG_CUDA(c, cudaMemsetAsync(d_st, 0, 9 * 8, c->stream));
  mh_stats_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(
      c->d_results.as<vg_pair_result>(), n, c->d_pairIdx.as<int64_t>(), c->d_bases.as<uint8_t>(), c->d_off1.as<int64_t>(),
      c->d_len1.as<int32_t>(), c->d_off2.as<int64_t>(), c->d_len2.as<int32_t>(), msa ? 1 : 0, d_st);
  VG_LAUNCH_CHECK(c);
  unsigned long long st[9];
  VG_CUDA(c, cudaMemcpyAsync(st, d_st, sizeof st, cudaMemcpyDeviceToHost, c->stream));
  VG_CUDA(c, cudaEventRecord(c->ev[3], c->stream));
  VG_CUDA(c, cudaStreamSynchronize(c->stream));
  if (stats) {
    stats->total_pairs = (int64_t)st[0], stats->both_exist = (int64_t)st[1], stats->reads_total = (int64_t)st[2];
    stats->total_mapped = (int64_t)st[3], stats->mapped_forward = (int64_t)st[4], stats->mapped_reverse = (int64_t)st[5];
    stats->both_mapped = (int64_t)st[6], stats->total_score = (int64_t)st[7], stats->total_concordant = (int64_t)st[8];
  }
  float a = 0, b = 0;
  cudaEventElapsedTime(&a, c->ev[0], c->ev[1]);
  cudaEventElapsedTime(&b, c->ev[0], c->ev[3]);
  c->timings[0] = a;
  c->timings[4] = b;
  c->lastTasks = nTasks;
  return VG_OK;
}

extern "C" {

int vg_set_reference(vg_ctx* c, const uint8_t* seq, int32_t G, const int32_t* contig_ends, int32_t n_contigs,
                     const int32_t* fragment_ends, int32_t n_fragments, int32_t is_fragmented, const uint8_t* keys,
                     const int64_t* offsets, const int32_t* positions, int32_t n_keys, int32_t index_thread_num) {
  if (!c) return VG_ERR_INVALID;
  VG_CUDA(c, cudaSetDevice(c->device));
  if (G < 0 || (G > 0 && !seq)) return vg_fail(c, VG_ERR_INVALID, "bad reference");
  if (n_contigs < 1 || !contig_ends || contig_ends[n_contigs - 1] != G)
    return vg_fail(c, VG_ERR_INVALID, "contig_ends must be ascending and end at the reference length");
  for (int i = 0; i < n_contigs; i++)
    if (contig_ends[i] < 0 || (i && contig_ends[i] < contig_ends[i - 1])) return vg_fail(c, VG_ERR_INVALID, "contig_ends not ascending");
  if (n_fragments < 1 || !fragment_ends) return vg_fail(c, VG_ERR_INVALID, "fragment_ends missing");
  if (keys && (!offsets || !positions || n_keys < 0)) return vg_fail(c, VG_ERR_INVALID, "incomplete CSR index");
  for (int i = 0; i < G; i++)
    if (seq[i] >= 127) return vg_fail(c, VG_ERR_ALPHABET, "reference byte %d at %d (bytes >= 127 are not supported: the Java tables have 127 entries)", (int)seq[i], i);
  c->haveRef = false;
  c->haveCounts = false;
  c->G = G;
  c->h_ref.assign(seq, seq + G);
  c->contigEnds.assign(contig_ends, contig_ends + n_contigs);
  c->fragmentEnds.assign(fragment_ends, fragment_ends + n_fragments);
  c->isFragmented = is_fragmented != 0;
  VG_CUDA(c, c->d_ref.reserve((size_t)G + 64));
  if (G) VG_CUDA(c, cudaMemcpy(c->d_ref.p, seq, G, cudaMemcpyHostToDevice));
  VG_CUDA(c, c->d_contigEnds.reserve((size_t)n_contigs * 4));
  VG_CUDA(c, cudaMemcpy(c->d_contigEnds.p, contig_ends, (size_t)n_contigs * 4, cudaMemcpyHostToDevice));
  VG_CUDA(c, c->d_fragmentEnds.reserve((size_t)n_fragments * 4));
  VG_CUDA(c, cudaMemcpy(c->d_fragmentEnds.p, fragment_ends, (size_t)n_fragments * 4, cudaMemcpyHostToDevice));
  if (G >= c->seed.K) VG_TRY(seed_build_reference(c, keys, offsets, positions, n_keys, index_thread_num));
  else if (G > 0) {
    // no k-mer fits: nothing is ever seeded; keep an all-absent table
    VG_TRY(seed_build_reference(c, nullptr, nullptr, nullptr, 0, 1));
  }
  c->haveRef = true;
  return VG_OK;
}

int vg_set_msa_index(vg_ctx* c, int32_t K, float low_coef, float high_coef, const int32_t* cutoffs, int32_t n_cutoffs,
                     const uint8_t* keys, const int64_t* offsets, const int32_t* columns, int32_t n_keys,
                     int32_t msa_length) {
  if (!c) return VG_ERR_INVALID;
  VG_CUDA(c, cudaSetDevice(c->device));
  if (n_keys < 0 || (n_keys > 0 && (!keys || !offsets || !columns)) || msa_length < 0)
    return vg_fail(c, VG_ERR_INVALID, "bad MSA index");
  c->haveMsa = false;
  if (K < 1 || K > 7) return vg_fail(c, VG_ERR_INVALID, "K=%d outside the supported range 1..7", K);
  if (!cutoffs || n_cutoffs < 1) return vg_fail(c, VG_ERR_INVALID, "cutoffs missing");
  c->seedMsa.K = K;
  c->seedMsa.lowCoef = low_coef;
  c->seedMsa.highCoef = high_coef;
  c->seedMsa.cutoffs.assign(cutoffs, cutoffs + n_cutoffs);
  VG_CUDA(c, c->seedMsa.d_cutoffs.reserve((size_t)n_cutoffs * 4));
  VG_CUDA(c, cudaMemcpy(c->seedMsa.d_cutoffs.p, cutoffs, (size_t)n_cutoffs * 4, cudaMemcpyHostToDevice));
  c->msaLength = msa_length;
  static const int64_t zero = 0;
  VG_TRY(seed_build_msa(c, keys, n_keys ? offsets : &zero, columns, n_keys));
  c->haveMsa = true;
  return VG_OK;
}

static int mh_upload(vg_ctx* c, const uint8_t* bases, const uint8_t* packed, const int64_t* exc_pos, const uint8_t* exc_byte,
                     int64_t n_exc, int64_t n_bases, const int64_t* off1, const int32_t* len1, const int64_t* off2,
                     const int32_t* len2, int64_t n_pairs) {
  if (!c) return VG_ERR_INVALID;
  VG_CUDA(c, cudaSetDevice(c->device));
  if (n_pairs < 0 || n_bases < 0 || n_exc < 0 || (n_pairs > 0 && ((!bases && !packed) || !off1 || !len1 || !off2 || !len2)) ||
      (n_exc > 0 && (!exc_pos || !exc_byte)))
    return vg_fail(c, VG_ERR_INVALID, "bad read arrays");
  c->haveReads = false;
  int maxLen = 0;
  int minGe[9];  // shortest read of at least k bases, k = 0 .. 8 (a read shorter than K seeds nothing)
  for (int& m : minGe) m = 0x7fffffff;
  for (int64_t i = 0; i < n_pairs; i++) {
    if (len1[i] < 0 || len2[i] < 0 || off1[i] < 0 || off2[i] < 0 || off1[i] + len1[i] > n_bases || off2[i] + len2[i] > n_bases)
      return vg_fail(c, VG_ERR_INVALID, "read %lld lies outside the byte pool", (long long)i);
    maxLen = std::max(maxLen, std::max(len1[i], len2[i]));
    for (int l : {len1[i], len2[i]})
      for (int k = std::min(l, 8); k >= 0 && minGe[k] > l; k--) minGe[k] = l;
  }
  for (int k = 0; k < 9; k++) c->minReadLenGe[k] = minGe[k];
  const int64_t nWords = (n_bases + 15) / 16;  // 16 bases per packed 32-bit word
  VG_CUDA(c, c->d_bases.reserve((size_t)nWords * 16 + 64));
  VG_CUDA(c, c->d_off1.reserve((size_t)n_pairs * 8 + 8));
  VG_CUDA(c, c->d_off2.reserve((size_t)n_pairs * 8 + 8));
  VG_CUDA(c, c->d_len1.reserve((size_t)n_pairs * 4 + 8));
  VG_CUDA(c, c->d_len2.reserve((size_t)n_pairs * 4 + 8));
  if (n_pairs > 0) {
    VG_CUDA(c, c->d_counter.reserve(16));
    VG_CUDA(c, cudaMemsetAsync(c->d_counter.p, 0, 16, c->stream));
    if (packed) {
      VG_CUDA(c, c->d_packed.reserve((size_t)nWords * 4 + (size_t)n_exc * 9 + 64));
      uint8_t* d_pk = c->d_packed.as<uint8_t>();
      int64_t* d_ep = (int64_t*)(d_pk + (((size_t)nWords * 4 + 15) & ~(size_t)15));
      uint8_t* d_eb = (uint8_t*)(d_ep + n_exc);
      // the last word may be partly outside the caller's buffer: copy the bytes that exist
      VG_CUDA(c, cudaMemcpyAsync(d_pk, packed, (size_t)((n_bases + 3) / 4), cudaMemcpyHostToDevice, c->stream));
      if (n_exc > 0) {
        VG_CUDA(c, cudaMemcpyAsync(d_ep, exc_pos, (size_t)n_exc * 8, cudaMemcpyHostToDevice, c->stream));
        VG_CUDA(c, cudaMemcpyAsync(d_eb, exc_byte, (size_t)n_exc, cudaMemcpyHostToDevice, c->stream));
      }
      if (nWords > 0) {
        mh_unpack_kernel<<<(unsigned)((nWords + 255) / 256), 256, 0, c->stream>>>((const uint32_t*)d_pk, nWords, c->d_bases.as<uint4>());
        VG_LAUNCH_CHECK(c);
      }
      if (n_exc > 0) {
        mh_exceptions_kernel<<<(unsigned)((n_exc + 255) / 256), 256, 0, c->stream>>>(d_ep, d_eb, n_exc, n_bases, c->d_bases.as<uint8_t>(),
                                                                                     c->d_counter.as<int>());
        VG_LAUNCH_CHECK(c);
      }
    } else {
      VG_CUDA(c, cudaMemcpyAsync(c->d_bases.p, bases, (size_t)n_bases, cudaMemcpyHostToDevice, c->stream));
    }
    VG_CUDA(c, cudaMemcpyAsync(c->d_off1.p, off1, (size_t)n_pairs * 8, cudaMemcpyHostToDevice, c->stream));
    VG_CUDA(c, cudaMemcpyAsync(c->d_off2.p, off2, (size_t)n_pairs * 8, cudaMemcpyHostToDevice, c->stream));
    VG_CUDA(c, cudaMemcpyAsync(c->d_len1.p, len1, (size_t)n_pairs * 4, cudaMemcpyHostToDevice, c->stream));
    VG_CUDA(c, cudaMemcpyAsync(c->d_len2.p, len2, (size_t)n_pairs * 4, cudaMemcpyHostToDevice, c->stream));
    if (n_bases > 0) {
      mh_validate_kernel<<<c->smCount * 8, 256, 0, c->stream>>>(c->d_bases.as<uint8_t>(), n_bases, c->d_counter.as<int>());
      VG_LAUNCH_CHECK(c);
    }
    int flag = 0;
    VG_CUDA(c, cudaMemcpyAsync(&flag, c->d_counter.p, 4, cudaMemcpyDeviceToHost, c->stream));
    VG_CUDA(c, cudaStreamSynchronize(c->stream));
    if (flag & 2) return vg_fail(c, VG_ERR_INVALID, "an exception position lies outside the read pool");
    if (flag)
      return vg_fail(c, VG_ERR_ALPHABET, "a read contains a byte >= 127 or in ('T','a'): the Java reference throws on it");
  }
  c->nPairs = n_pairs;
  c->nBases = n_bases;
  c->maxReadLen = maxLen;
  c->haveReads = true;
  return VG_OK;
}

int vg_upload_reads(vg_ctx* c, const uint8_t* bases, int64_t n_bases, const int64_t* off1, const int32_t* len1,
                    const int64_t* off2, const int32_t* len2, int64_t n_pairs) {
  if (c && n_pairs > 0 && !bases) return vg_fail(c, VG_ERR_INVALID, "bad read arrays");
  return mh_upload(c, bases, nullptr, nullptr, nullptr, 0, n_bases, off1, len1, off2, len2, n_pairs);
}

int vg_upload_reads_packed(vg_ctx* c, const uint8_t* packed, int64_t n_bases, const int64_t* exc_pos, const uint8_t* exc_byte,
                           int64_t n_exc, const int64_t* off1, const int32_t* len1, const int64_t* off2, const int32_t* len2,
                           int64_t n_pairs) {
  if (c && n_pairs > 0 && !packed) return vg_fail(c, VG_ERR_INVALID, "bad read arrays");
  return mh_upload(c, nullptr, packed, exc_pos, exc_byte, n_exc, n_bases, off1, len1, off2, len2, n_pairs);
}

// Host-side helper (no device involved): bytes -> the packed format of vg_upload_reads_packed. packed: (n_bases + 3) / 4
// bytes; exceptions are written in ascending position order; VG_ERR_CAPACITY (and *n_exc = the number needed) when there are
// more than exc_cap of them.
int vg_pack_reads(const uint8_t* bases, int64_t n_bases, uint8_t* packed, int64_t* exc_pos, uint8_t* exc_byte, int64_t exc_cap,
                  int64_t* n_exc) {
  if (n_bases < 0 || (n_bases > 0 && (!bases || !packed)) || !n_exc) return VG_ERR_INVALID;
  int64_t ne = 0;
  for (int64_t i = 0; i < n_bases; i += 4) {
    uint8_t b = 0;
    for (int k = 0; k < 4 && i + k < n_bases; k++) {
      const uint8_t ch = bases[i + k];
      const int cd = ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : ch == 'T' ? 3 : -1;
      if (cd >= 0) b |= (uint8_t)(cd << (2 * k));
      else {
        if (ne < exc_cap && exc_pos && exc_byte) exc_pos[ne] = i + k, exc_byte[ne] = bases[i + k];
        ne++;
      }
    }
    packed[i >> 2] = b;
  }
  *n_exc = ne;
  return ne > exc_cap ? VG_ERR_CAPACITY : VG_OK;
}

int vg_map_pairs(vg_ctx* c, const int64_t* pair_idx, int64_t n_idx, vg_map_stats* stats) {
  return mh_map(c, false, pair_idx, n_idx, stats);
}

int vg_map_pairs_msa(vg_ctx* c, vg_map_stats* stats) { return mh_map(c, true, nullptr, 0, stats); }

int vg_result_sizes(vg_ctx* c, int64_t* n_results, int64_t* seq1_bytes) {
  if (!c) return VG_ERR_INVALID;
  if (n_results) *n_results = c->nResults;
  if (seq1_bytes) *seq1_bytes = c->seq1Bytes;
  return VG_OK;
}

int vg_get_results(vg_ctx* c, vg_pair_result* out, int64_t n_out, uint8_t* seq1_pool, int64_t pool_cap) {
  if (!c) return VG_ERR_INVALID;
  VG_CUDA(c, cudaSetDevice(c->device));
  if (n_out < c->nResults || pool_cap < c->seq1Bytes) {
    c->required = n_out < c->nResults ? c->nResults : c->seq1Bytes;
    return vg_fail(c, VG_ERR_CAPACITY, "output buffers too small (%lld results, %lld sequence1 bytes)",
                   (long long)c->nResults, (long long)c->seq1Bytes);
  }
  if (c->nResults > 0) {
    if (!out) return vg_fail(c, VG_ERR_INVALID, "null output");
    VG_CUDA(c, cudaMemcpyAsync(out, c->d_results.p, (size_t)c->nResults * sizeof(vg_pair_result), cudaMemcpyDeviceToHost, c->stream));
  }
  if (c->seq1Bytes > 0) {
    if (!seq1_pool) return vg_fail(c, VG_ERR_INVALID, "null sequence1 pool");
    VG_CUDA(c, cudaMemcpyAsync(seq1_pool, c->d_seq1pool.p, (size_t)c->seq1Bytes, cudaMemcpyDeviceToHost, c->stream));
  }
  VG_CUDA(c, cudaStreamSynchronize(c->stream));
  return VG_OK;
}

// vg_get_results without the sequence1 bytes: records whose seq1_offset / pad_ address run-length edit operations
// (mh_ops_kernel) from which Alignment.sequence1 is a function of the read (SURVEY 8b "an edit string from which sequence1
// is reconstructible"). 4-5x fewer device-to-host bytes.
int vg_get_results_compact(vg_ctx* c, vg_pair_result* out, int64_t n_out, uint32_t* ops, int64_t ops_cap, int64_t* n_ops) {
  if (!c) return VG_ERR_INVALID;
  VG_CUDA(c, cudaSetDevice(c->device));
  if (n_ops) *n_ops = 0;
  const int64_t n = c->nResults;
  if (n_out < n) {
    c->required = n;
    return vg_fail(c, VG_ERR_CAPACITY, "record buffer too small (%lld results)", (long long)n);
  }
  if (n == 0) return VG_OK;
  if (!out) return vg_fail(c, VG_ERR_INVALID, "null output");
  const int64_t n2 = 2 * n;
  VG_CUDA(c, c->d_bamN.reserve((size_t)n2 * 4 + 64));
  VG_CUDA(c, c->d_bamOff.reserve((size_t)n2 * 8 + 64));
  VG_CUDA(c, c->d_bamRec.reserve((size_t)n * sizeof(vg_pair_result)));
  const vg_pair_result* d_res = c->d_results.as<vg_pair_result>();
  int64_t total = 0;
  const unsigned wblocks = (unsigned)((n2 * 32 + 255) / 256);
  if (c->resultsMsa) {  // no alignments in MSA mode: the records as they are (seq1_offset = -1)
    VG_CUDA(c, cudaMemsetAsync(c->d_bamN.p, 0, (size_t)n2 * 4, c->stream));
    VG_CUDA(c, cudaMemsetAsync(c->d_bamOff.p, 0, (size_t)n2 * 8, c->stream));
  } else {
    mh_ops_kernel<false><<<wblocks, 256, 0, c->stream>>>(d_res, n2, c->d_seq1pool.as<uint8_t>(), c->d_bamN.as<int32_t>(), nullptr, nullptr);
    VG_LAUNCH_CHECK(c);
    VG_TRY(mh_exclusive_scan(c, c->d_bamN.as<int32_t>(), c->d_bamOff.as<int64_t>(), n2, &total, true));
  }
  if (n_ops) *n_ops = total;
  if (ops_cap < total) {
    c->required = total;
    return vg_fail(c, VG_ERR_CAPACITY, "operation pool too small (%lld elements)", (long long)total);
  }
  if (total > 0 && !ops) return vg_fail(c, VG_ERR_INVALID, "null operation pool");
  VG_CUDA(c, c->d_bamOps.reserve((size_t)std::max<int64_t>(total, 1) * 4));
  if (total > 0) {
    mh_ops_kernel<true><<<wblocks, 256, 0, c->stream>>>(d_res, n2, c->d_seq1pool.as<uint8_t>(), nullptr, c->d_bamOff.as<int64_t>(),
                                                        c->d_bamOps.as<uint32_t>());
    VG_LAUNCH_CHECK(c);
  }
  mh_compact_records_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, c->stream>>>(d_res, n2, c->d_bamN.as<int32_t>(),
                                                                                c->d_bamOff.as<int64_t>(), c->d_bamRec.as<vg_pair_result>());
  VG_LAUNCH_CHECK(c);
  VG_CUDA(c, cudaMemcpyAsync(out, c->d_bamRec.p, (size_t)n * sizeof(vg_pair_result), cudaMemcpyDeviceToHost, c->stream));
  if (total > 0) VG_CUDA(c, cudaMemcpyAsync(ops, c->d_bamOps.p, (size_t)total * 4, cudaMemcpyDeviceToHost, c->stream));
  VG_CUDA(c, cudaStreamSynchronize(c->stream));
  return VG_OK;
}

int vg_set_results(vg_ctx* c, const vg_pair_result* rec, int64_t n, const int64_t* pair_idx, const uint8_t* seq1_pool,
                   int64_t pool_bytes) {
  if (!c) return VG_ERR_INVALID;
  VG_CUDA(c, cudaSetDevice(c->device));
  if (!c->haveRef) return vg_fail(c, VG_ERR_STATE, "vg_set_reference has not been called");
  if (!c->haveReads) return vg_fail(c, VG_ERR_STATE, "vg_upload_reads has not been called");
  if (n < 0 || pool_bytes < 0 || (n > 0 && (!rec || !pair_idx)) || (pool_bytes > 0 && !seq1_pool))
    return vg_fail(c, VG_ERR_INVALID, "bad arguments");
  for (int64_t i = 0; i < n; i++) {
    if (pair_idx[i] < 0 || pair_idx[i] >= c->nPairs) return vg_fail(c, VG_ERR_INVALID, "pair index %lld out of range", (long long)pair_idx[i]);
    if (rec[i].cls < VG_CLS_CONCORDANT_FR || rec[i].cls > VG_CLS_UNMAPPED) return vg_fail(c, VG_ERR_INVALID, "record %lld has class %d", (long long)i, rec[i].cls);
    for (int k = 0; k < 2; k++) {
      const vg_mate& m = rec[i].m[k];
      if (!m.present) continue;
      if (m.length < 0 || m.seq1_offset < 0 || m.seq1_offset + m.length > pool_bytes)
        return vg_fail(c, VG_ERR_INVALID, "record %lld mate %d: sequence1 lies outside the pool", (long long)i, k + 1);
      if (m.length > c->maxReadLen * 3 + 16)  // an alignment is at most |read| + |window| columns long
        return vg_fail(c, VG_ERR_INVALID, "record %lld mate %d: alignment of %d columns is longer than any read allows", (long long)i, k + 1, m.length);
    }
  }
  // the previous results are gone whatever happens below
  c->nResults = 0;
  c->seq1Bytes = 0;
  c->resultsMsa = false;
  VG_CUDA(c, c->d_results.reserve((size_t)std::max<int64_t>(n, 1) * sizeof(vg_pair_result)));
  VG_CUDA(c, c->d_pairIdx.reserve((size_t)std::max<int64_t>(n, 1) * 8));
  VG_CUDA(c, c->d_seq1pool.reserve((size_t)pool_bytes + 64));
  VG_CUDA(c, c->d_rowpool.reserve((size_t)pool_bytes + 64));
  VG_CUDA(c, c->d_rowLen.reserve((size_t)std::max<int64_t>(n, 1) * 2 * 4 + 64));
  VG_CUDA(c, c->d_counter.reserve(16));
  VG_CUDA(c, cudaMemsetAsync(c->d_counter.p, 0, 16, c->stream));
  if (n > 0) {
    VG_CUDA(c, cudaMemcpyAsync(c->d_results.p, rec, (size_t)n * sizeof(vg_pair_result), cudaMemcpyHostToDevice, c->stream));
    VG_CUDA(c, cudaMemcpyAsync(c->d_pairIdx.p, pair_idx, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
    if (pool_bytes > 0) VG_CUDA(c, cudaMemcpyAsync(c->d_seq1pool.p, seq1_pool, (size_t)pool_bytes, cudaMemcpyHostToDevice, c->stream));
    pu_rows_kernel<<<(unsigned)((n * 2 * 32 + 255) / 256), 256, 0, c->stream>>>(c->d_results.as<vg_pair_result>(), n * 2,
                                                                                c->d_seq1pool.as<uint8_t>(), c->d_rowpool.as<uint8_t>(),
                                                                                c->d_rowLen.as<int32_t>(), c->d_counter.as<int>());
    VG_LAUNCH_CHECK(c);
  }
  int flag = 0;
  VG_CUDA(c, cudaMemcpyAsync(&flag, c->d_counter.p, 4, cudaMemcpyDeviceToHost, c->stream));
  VG_CUDA(c, cudaStreamSynchronize(c->stream));
  if (flag) return vg_fail(c, VG_ERR_ALPHABET, "a sequence1 byte lies in ('T','a'): the Java pileup throws on it");
  c->nResults = n;
  c->seq1Bytes = pool_bytes;
  return VG_OK;
}

int vg_bam_fields(vg_ctx* c, vg_bam_record* out, int64_t n_records, uint32_t* cigar_pool, int64_t pool_cap, int64_t* n_ops) {
  if (!c) return VG_ERR_INVALID;
  VG_CUDA(c, cudaSetDevice(c->device));
  if (!c->haveRef) return vg_fail(c, VG_ERR_STATE, "vg_set_reference has not been called");
  if (c->resultsMsa) return vg_fail(c, VG_ERR_STATE, "the last mapping was MSA mode (no alignments, no BAM records)");
  const int64_t n = c->nResults;
  if (n_ops) *n_ops = 0;
  if (n_records < 2 * n) {
    c->required = 2 * n;
    return vg_fail(c, VG_ERR_CAPACITY, "record buffer too small (%lld records)", (long long)(2 * n));
  }
  if (n == 0) return VG_OK;
  if (!out) return vg_fail(c, VG_ERR_INVALID, "null output");
  BamArgs A;
  memset(&A, 0, sizeof A);
  A.res = c->d_results.as<vg_pair_result>();
  A.nPairs = n;
  A.seq1pool = c->d_seq1pool.as<uint8_t>();
  A.pairIdx = c->d_pairIdx.as<int64_t>();
  A.len1 = c->d_len1.as<int32_t>();
  A.len2 = c->d_len2.as<int32_t>();
  A.contigEnds = c->d_contigEnds.as<int32_t>();
  A.nContigs = (int)c->contigEnds.size();
  VG_CUDA(c, c->d_bamN.reserve((size_t)n * 2 * 4 + 64));
  VG_CUDA(c, c->d_bamOff.reserve((size_t)n * 2 * 8 + 64));
  VG_CUDA(c, c->d_bamRec.reserve((size_t)n * 2 * sizeof(vg_bam_record)));
  A.nOps = c->d_bamN.as<int32_t>();
  const unsigned blocks = (unsigned)((2 * n + 127) / 128);
  bam_fields_kernel<false><<<blocks, 128, 0, c->stream>>>(A);
  VG_LAUNCH_CHECK(c);
  int64_t total = 0;
  VG_TRY(mh_exclusive_scan(c, A.nOps, c->d_bamOff.as<int64_t>(), 2 * n, &total));
  if (n_ops) *n_ops = total;
  if (pool_cap < total) {
    c->required = total;
    return vg_fail(c, VG_ERR_CAPACITY, "CIGAR pool too small (%lld elements)", (long long)total);
  }
  if (total > 0 && !cigar_pool) return vg_fail(c, VG_ERR_INVALID, "null CIGAR pool");
  VG_CUDA(c, c->d_bamOps.reserve((size_t)std::max<int64_t>(total, 1) * 4));
  A.opOff = c->d_bamOff.as<int64_t>();
  A.out = c->d_bamRec.as<vg_bam_record>();
  A.ops = c->d_bamOps.as<uint32_t>();
  bam_fields_kernel<true><<<blocks, 128, 0, c->stream>>>(A);
  VG_LAUNCH_CHECK(c);
  VG_CUDA(c, cudaMemcpyAsync(out, A.out, (size_t)n * 2 * sizeof(vg_bam_record), cudaMemcpyDeviceToHost, c->stream));
  if (total > 0) VG_CUDA(c, cudaMemcpyAsync(cigar_pool, A.ops, (size_t)total * 4, cudaMemcpyDeviceToHost, c->stream));
  VG_CUDA(c, cudaStreamSynchronize(c->stream));
  return VG_OK;
}

// ProblemIntervalBuilder.countIdentity (ProblemIntervalBuilder.java:35-160) over the alignments of the last
// vg_map_pairs call and the reference they were mapped to; thread_num = the reference's ThreadNumber (it fixes the order
// of the binary32 sums). identity / coverage: G entries each.
int vg_identity_counting(vg_ctx* c, int32_t window_size, int32_t thread_num, float* identity, int32_t* coverage) {
  if (!c) return VG_ERR_INVALID;
  VG_CUDA(c, cudaSetDevice(c->device));
  if (!c->haveRef) return vg_fail(c, VG_ERR_STATE, "vg_set_reference has not been called");
  if (c->resultsMsa) return vg_fail(c, VG_ERR_STATE, "the last mapping was MSA mode (no alignments)");
  if (window_size < 1 || thread_num < 1 || thread_num > 65535) return vg_fail(c, VG_ERR_INVALID, "bad window size / thread number");
  const int G = c->G;
  const int64_t n = c->nResults;
  if (G == 0) return VG_OK;
  if (!identity || !coverage) return vg_fail(c, VG_ERR_INVALID, "null output");
  if (n == 0) {
    memset(identity, 0, (size_t)G * 4);
    memset(coverage, 0, (size_t)G * 4);
    return VG_OK;
  }
  const int64_t n2 = n * 2;
  IdcArgs A;
  memset(&A, 0, sizeof A);
  A.res = c->d_results.as<vg_pair_result>();
  A.nPairs = n;
  A.seq1pool = c->d_seq1pool.as<uint8_t>();
  A.pairIdx = c->d_pairIdx.as<int64_t>();
  A.len1 = c->d_len1.as<int32_t>();
  A.len2 = c->d_len2.as<int32_t>();
  A.genome = c->d_ref.as<uint8_t>();
  A.G = G;
  A.windowSize = window_size;
  VG_CUDA(c, c->d_bamN.reserve((size_t)n2 * 2 * 4 + 64));
  VG_CUDA(c, c->d_bamOff.reserve((size_t)n2 * 2 * 8 + 64));
  VG_CUDA(c, c->d_bamRec.reserve((size_t)n2 * sizeof(IdcRead)));
  VG_CUDA(c, c->d_counter.reserve(16));
  VG_CUDA(c, cudaMemsetAsync(c->d_counter.p, 0, 16, c->stream));
  A.isMapped = c->d_bamN.as<int32_t>();
  A.npos = A.isMapped + n2;
  A.reads = c->d_bamRec.as<IdcRead>();
  A.err = c->d_counter.as<int>();
  const unsigned blocks = (unsigned)((n2 + 127) / 128);
  idc_plan_kernel<<<blocks, 128, 0, c->stream>>>(A);
  VG_LAUNCH_CHECK(c);
  int64_t nMapped = 0, nVals = 0;
  int64_t* d_listIdx = c->d_bamOff.as<int64_t>();
  int64_t* d_valOff = d_listIdx + n2;
  VG_TRY(mh_exclusive_scan(c, A.isMapped, d_listIdx, n2, &nMapped, true));
  VG_TRY(mh_exclusive_scan(c, A.npos, d_valOff, n2, &nVals, true));
  A.listIdx = d_listIdx;
  A.valOff = d_valOff;
  VG_CUDA(c, c->d_bamOps.reserve((size_t)std::max<int64_t>(nVals, 1) * 4));
  A.vals = c->d_bamOps.as<float>();
  idc_walk_kernel<<<blocks, 128, 0, c->stream>>>(A);
  VG_LAUNCH_CHECK(c);
  const int nTiles = (G + IDC_TILE - 1) / IDC_TILE, T = thread_num;
  VG_CUDA(c, c->d_scratch.reserve((size_t)T * G * 8 + (size_t)G * 8 + 64));
  float* d_partial = c->d_scratch.as<float>();
  int32_t* d_partialCov = (int32_t*)(d_partial + (size_t)T * G);
  float* d_identity = (float*)(d_partialCov + (size_t)T * G);
  int32_t* d_coverage = (int32_t*)(d_identity + G);
  idc_sum_kernel<<<dim3(nTiles, T), IDC_TILE, 0, c->stream>>>(A.reads, d_listIdx, n2, nMapped, T, A.vals, G, d_partial, d_partialCov);
  VG_LAUNCH_CHECK(c);
  idc_final_kernel<<<(G + 255) / 256, 256, 0, c->stream>>>(d_partial, d_partialCov, G, T, d_identity, d_coverage);
  VG_LAUNCH_CHECK(c);
  int flag = 0;
  VG_CUDA(c, cudaMemcpyAsync(&flag, c->d_counter.p, 4, cudaMemcpyDeviceToHost, c->stream));
  VG_CUDA(c, cudaMemcpyAsync(identity, d_identity, (size_t)G * 4, cudaMemcpyDeviceToHost, c->stream));
  VG_CUDA(c, cudaMemcpyAsync(coverage, d_coverage, (size_t)G * 4, cudaMemcpyDeviceToHost, c->stream));
  VG_CUDA(c, cudaStreamSynchronize(c->stream));
  if (flag) return vg_fail(c, VG_ERR_INVALID, "an alignment walks outside its sequence or the reference (the Java would throw)");
  return VG_OK;
}

int vg_pileup(vg_ctx* c, int32_t accumulate) {
  if (!c) return VG_ERR_INVALID;
  VG_CUDA(c, cudaSetDevice(c->device));
  if (!c->haveRef) return vg_fail(c, VG_ERR_STATE, "vg_set_reference has not been called");
  if (c->resultsMsa) return vg_fail(c, VG_ERR_STATE, "the last mapping was MSA mode (no alignments to pile up)");
  const int G = c->G;
  const int64_t n = c->nResults;
  if (accumulate && !c->haveCounts)
    return vg_fail(c, VG_ERR_STATE, "accumulate requested but there are no counts to add to (vg_set_reference clears them)");
  VG_CUDA(c, c->d_counts.reserve((size_t)std::max(G, 1) * 5 * 4));
  if (!accumulate) VG_CUDA(c, cudaMemsetAsync(c->d_counts.p, 0, (size_t)std::max(G, 1) * 5 * 4, c->stream));
  c->haveCounts = true;
  if (G == 0 || n == 0) {
    VG_CUDA(c, cudaStreamSynchronize(c->stream));
    return VG_OK;
  }
  VG_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
  const int nTiles = (G + PU_TILE - 1) / PU_TILE;
  const int64_t n2 = n * 2;
  VG_CUDA(c, c->d_tasks.reserve((size_t)n2 * sizeof(PuRead) * 2 + 64));  // unsorted + sorted lists
  PuRead* d_reads = c->d_tasks.as<PuRead>();
  PuRead* d_sorted = d_reads + n2;
  VG_CUDA(c, c->d_fill.reserve((size_t)n2 * 4 + 64));
  int32_t* d_tileOf = c->d_fill.as<int32_t>();
  const size_t cntBytes = ((size_t)(nTiles + 2) * 4 + 15) & ~(size_t)15;
  VG_CUDA(c, c->d_misc.reserve(cntBytes + (size_t)(nTiles + 2) * 16 + 64));
  int32_t* d_tileCount = c->d_misc.as<int32_t>();
  int64_t* d_tileStart = (int64_t*)(c->d_misc.as<uint8_t>() + cntBytes);
  int64_t* d_cursor = d_tileStart + nTiles + 2;
  VG_CUDA(c, cudaMemsetAsync(d_tileCount, 0, (size_t)(nTiles + 2) * 4, c->stream));
  VG_CUDA(c, c->d_counter.reserve(16));
  VG_CUDA(c, cudaMemsetAsync(c->d_counter.p, 0, 16, c->stream));
  pu_list_kernel<<<(unsigned)((n2 + 255) / 256), 256, nTiles * 4, c->stream>>>(
      c->d_results.as<vg_pair_result>(), n, c->d_rowLen.as<int32_t>(), G, nTiles, d_reads, d_tileOf, d_tileCount,
      c->d_counter.as<int>());
  VG_LAUNCH_CHECK(c);
  pu_scan_kernel<<<1, 32, 0, c->stream>>>(d_tileCount, nTiles, d_tileStart, d_cursor);
  VG_LAUNCH_CHECK(c);
  pu_scatter_kernel<<<(unsigned)((n2 + 255) / 256), 256, (size_t)(2 * nTiles + 2) * 4 + (size_t)nTiles * 8, c->stream>>>(
      d_reads, d_tileOf, n2, nTiles, (unsigned long long*)d_cursor, d_sorted);
  VG_LAUNCH_CHECK(c);
  const int nSlices = std::max(1, std::min(64, (c->smCount * 8) / nTiles));  // 8 CTAs of 256 threads per SM
  VG_CUDA(c, c->d_scratch.reserve((size_t)nSlices * G * 5 * 4 + 64));
  dim3 grid(nTiles, nSlices);
  pu_count_kernel<<<grid, PU_TILE, 0, c->stream>>>(d_sorted, d_tileStart, nTiles, c->d_counter.as<int>() + 1, c->d_rowpool.as<uint8_t>(), G,
                                                   nSlices, c->d_scratch.as<int32_t>());
  VG_LAUNCH_CHECK(c);
  pu_reduce_kernel<<<(unsigned)(((int64_t)G * 5 + 255) / 256), 256, 0, c->stream>>>(c->d_scratch.as<int32_t>(), G, nSlices,
                                                                                    accumulate ? 1 : 0, c->d_counts.as<int32_t>());
  VG_LAUNCH_CHECK(c);
  VG_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
  int flag = 0;
  VG_CUDA(c, cudaMemcpyAsync(&flag, c->d_counter.p, 4, cudaMemcpyDeviceToHost, c->stream));
  VG_CUDA(c, cudaStreamSynchronize(c->stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
  c->timings[3] = ms;
  if (flag) return vg_fail(c, VG_ERR_ALPHABET, "an aligned read runs past the end of the reference (the Java pileup would throw)");
  return VG_OK;
}

int vg_pileup_counts_device(vg_ctx* c, void** dev_ptr, int64_t* n_int32) {
  if (!c || !dev_ptr) return VG_ERR_INVALID;
  if (!c->haveCounts) return vg_fail(c, VG_ERR_STATE, "vg_pileup has not been called");
  *dev_ptr = c->d_counts.p;
  if (n_int32) *n_int32 = (int64_t)c->G * 5;
  return VG_OK;
}

int vg_pileup_counts_host(vg_ctx* c, int32_t* out, int64_t n_int32) {
  if (!c || !out) return VG_ERR_INVALID;
  VG_CUDA(c, cudaSetDevice(c->device));
  if (!c->haveCounts) return vg_fail(c, VG_ERR_STATE, "vg_pileup has not been called");
  if (n_int32 < (int64_t)c->G * 5) {
    c->required = (int64_t)c->G * 5;
    return vg_fail(c, VG_ERR_CAPACITY, "counts buffer too small");
  }
  VG_CUDA(c, cudaMemcpyAsync(out, c->d_counts.p, (size_t)c->G * 5 * 4, cudaMemcpyDeviceToHost, c->stream));
  VG_CUDA(c, cudaStreamSynchronize(c->stream));
  return VG_OK;
}

int vg_consensus(vg_ctx* c, int32_t save_uncovered, int32_t coverage_threshold, const uint8_t* genome, int32_t G,
                 uint8_t* out) {
  if (!c || !out) return VG_ERR_INVALID;
  VG_CUDA(c, cudaSetDevice(c->device));
  if (!c->haveCounts) return vg_fail(c, VG_ERR_STATE, "vg_pileup has not been called");
  if (G != c->G) return vg_fail(c, VG_ERR_INVALID, "genome length %d differs from the reference length %d", G, c->G);
  if (G == 0) return VG_OK;
  if (save_uncovered && !genome) return vg_fail(c, VG_ERR_INVALID, "genome missing");
  VG_CUDA(c, c->d_s1pool.reserve((size_t)G * 2 + 64));
  uint8_t* d_gen = c->d_s1pool.as<uint8_t>();
  uint8_t* d_out = d_gen + G;
  if (genome) VG_CUDA(c, cudaMemcpyAsync(d_gen, genome, G, cudaMemcpyHostToDevice, c->stream));
  pu_consensus_kernel<<<(G + 255) / 256, 256, 0, c->stream>>>(c->d_counts.as<int32_t>(), G, save_uncovered, coverage_threshold,
                                                             d_gen, d_out);
  VG_LAUNCH_CHECK(c);
  VG_CUDA(c, cudaMemcpyAsync(out, d_out, G, cudaMemcpyDeviceToHost, c->stream));
  VG_CUDA(c, cudaStreamSynchronize(c->stream));
  return VG_OK;
}

static int mh_seed_batch(vg_ctx* c, int32_t msa, const uint8_t* pool, const int64_t* off, int64_t n, int32_t max_windows,
                         int32_t* windows, int32_t* n_windows, int32_t* dbgHits, int32_t dbgCap, int32_t* dbgN, bool bestOnly = false) {
  if (!c) return VG_ERR_INVALID;
  VG_CUDA(c, cudaSetDevice(c->device));
  if (msa ? !c->haveMsa : !c->haveRef) return vg_fail(c, VG_ERR_STATE, "reference / MSA index not set");
  if (n < 0 || max_windows < 1 || (n > 0 && (!pool || !off))) return vg_fail(c, VG_ERR_INVALID, "bad arguments");
  if (n == 0) return VG_OK;
  std::vector<int32_t> len(n);
  int maxLen = 0;
  for (int64_t i = 0; i < n; i++) {
    int64_t l = off[i + 1] - off[i];
    if (l < 0 || l > (1 << 20)) return vg_fail(c, VG_ERR_INVALID, "bad offsets");
    len[i] = (int32_t)l;
    maxLen = std::max(maxLen, len[i]);
  }
  for (int64_t i = 0; i < off[n]; i++)
    if (pool[i] >= 127) return vg_fail(c, VG_ERR_ALPHABET, "read byte >= 127");
  VG_CUDA(c, c->d_s1pool.reserve((size_t)off[n] + 64));
  VG_CUDA(c, c->d_s1off.reserve((size_t)(n + 1) * 8));
  VG_CUDA(c, c->d_s2off.reserve((size_t)n * 4 + 64));
  VG_CUDA(c, cudaMemcpyAsync(c->d_s1pool.p, pool, (size_t)off[n], cudaMemcpyHostToDevice, c->stream));
  VG_CUDA(c, cudaMemcpyAsync(c->d_s1off.p, off, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, c->stream));
  VG_CUDA(c, cudaMemcpyAsync(c->d_s2off.p, len.data(), (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
  int32_t* d_dbg = nullptr;
  int32_t* d_dbgN = nullptr;
  if (dbgHits) {
    VG_CUDA(c, c->d_rev.reserve((size_t)16384 * 3 * 4 + 64));
    d_dbgN = c->d_rev.as<int32_t>();
    d_dbg = d_dbgN + 4;
    VG_CUDA(c, cudaMemsetAsync(d_dbgN, 0, 16, c->stream));
  }
  VG_TRY(seed_launch(c, msa != 0, c->d_s1pool.as<uint8_t>(), c->d_s1off.as<int64_t>(), c->d_s2off.as<int32_t>(), nullptr,
                     nullptr, nullptr, n, true, maxLen, max_windows, d_dbg, d_dbgN, bestOnly));
  VG_CUDA(c, cudaMemcpyAsync(windows, c->d_cand.p, (size_t)n * max_windows * 3 * 4, cudaMemcpyDeviceToHost, c->stream));
  VG_CUDA(c, cudaMemcpyAsync(n_windows, c->d_ncand.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
  if (dbgHits) {
    int nh = 0;
    VG_CUDA(c, cudaMemcpyAsync(&nh, d_dbgN, 4, cudaMemcpyDeviceToHost, c->stream));
    VG_CUDA(c, cudaStreamSynchronize(c->stream));
    *dbgN = nh;
    VG_CUDA(c, cudaMemcpyAsync(dbgHits, d_dbg, (size_t)std::min(nh, dbgCap) * 3 * 4, cudaMemcpyDeviceToHost, c->stream));
  }
  VG_CUDA(c, cudaStreamSynchronize(c->stream));
  return VG_OK;
}

// ---- random model ------------------------------------------------------------------------------------
// MarkovModel.generate (MarkovModel.java:94-119): one thread per random read. The reference draws from an unseeded
// java.util.Random created per read (:95, Q10), so no stream is "the reference's"; the policy here is SplitMix64 with the state
// seed + (first_draw + r * draws_per_read) * golden for read r, draws_per_read = 2 + read_len - order (genome, position, one
// draw per further base) — the stream the oracle's seeded restatement consumes sequentially.
struct RmRng {
  uint64_t s;
  __device__ __forceinline__ uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
  }
  __device__ __forceinline__ int nextInt(int n) { return (int)(next() % (uint64_t)n); }
  __device__ __forceinline__ float nextFloat() { return (float)(next() >> 40) / 16777216.0f; }
};
// freqs: [4^order][4] cumulative frequencies in the order A, T, G, C (fr[3] < 0: the k-mer is not in the model);
// k-mer code: 2 bits per base in that order, first base most significant
__global__ void rm_generate_kernel(const uint8_t* __restrict__ genomes, const int64_t* __restrict__ genomeOff, int nGenomes,
                                   int order, const float4* __restrict__ freqs, uint64_t seed, int64_t firstDraw, int readLen,
                                   int64_t nReads, uint8_t* __restrict__ out) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nReads) return;
  RmRng rnd;
  rnd.s = seed + (uint64_t)(firstDraw + r * (int64_t)(2 + readLen - order)) * 0x9E3779B97F4A7C15ULL;
  const int gid = rnd.nextInt(nGenomes);
  const uint8_t* genome = genomes + genomeOff[gid];
  const int glen = (int)(genomeOff[gid + 1] - genomeOff[gid]);
  const int gpos = rnd.nextInt(glen - order + 1);
  uint8_t* dst = out + r * (int64_t)readLen;
  const uint32_t cmask = (order >= 16) ? 0xFFFFFFFFu : ((1u << (2 * order)) - 1u);
  uint32_t code = 0;
  int good = 0;  // trailing bases that are A/T/G/C
  auto push = [&](uint8_t b) {
    const int v = b == 'A' ? 0 : b == 'T' ? 1 : b == 'G' ? 2 : b == 'C' ? 3 : -1;
    code = ((code << 2) | (uint32_t)(v & 3)) & cmask;
    good = v >= 0 ? good + 1 : 0;
  };
  for (int j = 0; j < order && j < readLen; j++) {
    dst[j] = genome[gpos + j];
    push(genome[gpos + j]);
  }
  for (int j = order; j < readLen; j++) {
    float4 fr = make_float4(0.f, 0.f, 0.f, -1.f);
    if (good >= order) fr = freqs[code];
    uint8_t b;
    if (fr.w < 0.f) {  // freqs.get(...) == null
      const int n = rnd.nextInt(4);
      b = n == 0 ? 'A' : n == 1 ? 'T' : n == 2 ? 'G' : 'C';  // Constants.iToN
    } else {
      const float pr = rnd.nextFloat();
      b = pr <= fr.x ? 'A' : pr <= fr.y ? 'T' : pr <= fr.z ? 'G' : 'C';
    }
    dst[j] = b;
    push(b);
  }
}
__global__ void rm_offsets_kernel(int64_t n, int readLen, int64_t* __restrict__ off, int32_t* __restrict__ len) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r <= n) off[r] = r * readLen;
  if (r < n) len[r] = readLen;
}

// RandomModel.RandomCountsComputing.run (RandomModel.java:113-128): counts[r] = computeKMerCount(model.generate(len), index, 0, length)
int vg_random_model_counts(vg_ctx* c, int32_t msa, const uint8_t* genomes, const int64_t* genome_off, int32_t n_genomes,
                           int32_t order, uint64_t seed, int64_t first_draw, int32_t read_len, int64_t n_reads,
                           int32_t* counts, uint8_t* reads_out) {
  if (!c) return VG_ERR_INVALID;
  VG_CUDA(c, cudaSetDevice(c->device));
  if (msa ? !c->haveMsa : !c->haveRef) return vg_fail(c, VG_ERR_STATE, "reference / MSA index not set");
  if (!genomes || !genome_off || n_genomes < 1 || order < 1 || order > 8 || read_len < order || read_len > SD_MAXL || n_reads < 0 ||
      first_draw < 0 || (n_reads > 0 && !counts))
    return vg_fail(c, VG_ERR_INVALID, "bad arguments (order 1..8, order <= read_len <= %d)", SD_MAXL);
  if (n_reads == 0) return VG_OK;
  // MarkovModel(String[] genomes, int order) (MarkovModel.java:18-71): float32 arithmetic in the reference's order
  const int nCodes = 1 << (2 * order);
  std::vector<int32_t> cnt((size_t)nCodes * 4, 0);
  std::vector<uint8_t> seen(nCodes, 0);
  auto v4 = [](uint8_t b) { return b == 'A' ? 0 : b == 'T' ? 1 : b == 'G' ? 2 : b == 'C' ? 3 : -1; };
  for (int g = 0; g < n_genomes; g++) {
    const uint8_t* s = genomes + genome_off[g];
    const int64_t gl = genome_off[g + 1] - genome_off[g];
    if (gl < order) return vg_fail(c, VG_ERR_INVALID, "genome %d is shorter than the model order (Random.nextInt would throw)", g);
    for (int64_t j = 0; j < gl - order; j++) {
      uint32_t code = 0;
      bool ok = true;
      for (int t = 0; t < order; t++) {
        const int v = v4(s[j + t]);
        ok &= v >= 0;
        code = (code << 2) | (uint32_t)(v & 3);
      }
      if (!ok) continue;
      seen[code] = 1;
      const int v = v4(s[j + order]);
      if (v >= 0) cnt[(size_t)code * 4 + v]++;
    }
  }
  std::vector<float> fr((size_t)nCodes * 4);
  for (int k = 0; k < nCodes; k++) {
    float* f = &fr[(size_t)k * 4];
    if (!seen[k]) {
      f[0] = f[1] = f[2] = 0.f, f[3] = -1.f;
      continue;
    }
    const int* q = &cnt[(size_t)k * 4];
    const int total = q[0] + q[1] + q[2] + q[3];
    volatile float cf = 1.0f / (float)total;  // total == 0: +Inf, the products are NaN and every comparison fails -> 'C'
    volatile float a = (float)q[0] * cf;
    volatile float t1 = (float)q[1] * cf;
    volatile float b = a + t1;
    volatile float t2 = (float)q[2] * cf;
    volatile float d = b + t2;
    f[0] = a, f[1] = b, f[2] = d, f[3] = 1.0f;
  }
  const size_t gbytes = (size_t)genome_off[n_genomes];
  VG_CUDA(c, c->d_rand.reserve(gbytes + 64 + (size_t)(n_genomes + 1) * 8 + fr.size() * 4 + 64));
  uint8_t* d_gen = c->d_rand.as<uint8_t>();
  float4* d_fr = (float4*)(d_gen + ((gbytes + 63) & ~(size_t)63));
  int64_t* d_goff = (int64_t*)((uint8_t*)d_fr + fr.size() * 4);
  VG_CUDA(c, cudaMemcpyAsync(d_gen, genomes, gbytes, cudaMemcpyHostToDevice, c->stream));
  VG_CUDA(c, cudaMemcpyAsync(d_fr, fr.data(), fr.size() * 4, cudaMemcpyHostToDevice, c->stream));
  VG_CUDA(c, cudaMemcpyAsync(d_goff, genome_off, (size_t)(n_genomes + 1) * 8, cudaMemcpyHostToDevice, c->stream));
  VG_CUDA(c, c->d_s1pool.reserve((size_t)n_reads * read_len + 64));
  VG_CUDA(c, c->d_s1off.reserve((size_t)(n_reads + 1) * 8));
  VG_CUDA(c, c->d_s2off.reserve((size_t)n_reads * 4 + 64));
  const int blocks = (int)((n_reads + 1 + 127) / 128);
  rm_generate_kernel<<<blocks, 128, 0, c->stream>>>(d_gen, d_goff, n_genomes, order, d_fr, seed, first_draw, read_len, n_reads,
                                                    c->d_s1pool.as<uint8_t>());
  VG_LAUNCH_CHECK(c);
  rm_offsets_kernel<<<blocks, 128, 0, c->stream>>>(n_reads, read_len, c->d_s1off.as<int64_t>(), c->d_s2off.as<int32_t>());
  VG_LAUNCH_CHECK(c);
  if (reads_out)
    VG_CUDA(c, cudaMemcpyAsync(reads_out, c->d_s1pool.p, (size_t)n_reads * read_len, cudaMemcpyDeviceToHost, c->stream));
  VG_TRY(seed_launch(c, msa != 0, c->d_s1pool.as<uint8_t>(), c->d_s1off.as<int64_t>(), c->d_s2off.as<int32_t>(), nullptr,
                     nullptr, nullptr, n_reads, true, read_len, 1, nullptr, nullptr, true));
  std::vector<int32_t> win((size_t)n_reads * 3), found(n_reads);
  VG_CUDA(c, cudaMemcpyAsync(win.data(), c->d_cand.p, (size_t)n_reads * 3 * 4, cudaMemcpyDeviceToHost, c->stream));
  VG_CUDA(c, cudaMemcpyAsync(found.data(), c->d_ncand.p, (size_t)n_reads * 4, cudaMemcpyDeviceToHost, c->stream));
  VG_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int64_t r = 0; r < n_reads; r++) counts[r] = found[r] ? win[(size_t)r * 3 + 2] : 0;  // no hit: computeKMerCount returns 0
  return VG_OK;
}

int vg_seed_batch(vg_ctx* c, int32_t msa, const uint8_t* pool, const int64_t* off, int64_t n, int32_t max_windows,
                  int32_t* windows, int32_t* n_windows) {
  return mh_seed_batch(c, msa, pool, off, n, max_windows, windows, n_windows, nullptr, 0, nullptr);
}

int vg_best_window_batch(vg_ctx* c, int32_t msa, const uint8_t* pool, const int64_t* off, int64_t n, int32_t* windows,
                         int32_t* found) {
  return mh_seed_batch(c, msa, pool, off, n, 1, windows, found, nullptr, 0, nullptr, true);
}

// ContigBuilder.countReads / Graph.SWScore seeding step: KMerCounter.mapReadToRegion(read, index of the
// region, 0, |region|) for (read, region) pairs, every region with its own StringIndexer.buildIndex index (single-threaded:
// no duplicated positions). The per-region tables are built here and live in global memory.
int vg_map_to_regions(vg_ctx* c, const uint8_t* region_pool, const int64_t* region_off, int64_t n_regions,
                      const uint8_t* read_pool, const int64_t* read_off, int64_t n_reads, const int32_t* read_region,
                      int32_t* windows, int32_t* found) {
  if (!c) return VG_ERR_INVALID;
  VG_CUDA(c, cudaSetDevice(c->device));
  if (n_regions < 0 || n_reads < 0 || (n_regions > 0 && (!region_pool || !region_off)) ||
      (n_reads > 0 && (!read_pool || !read_off || !read_region || !windows || !found)))
    return vg_fail(c, VG_ERR_INVALID, "bad arguments");
  if (n_reads == 0) return VG_OK;
  if (n_regions == 0) return vg_fail(c, VG_ERR_INVALID, "%lld reads but no region", (long long)n_reads);
  const int K = c->seed.K;
  if (K < 1 || K > 7) return vg_fail(c, VG_ERR_INVALID, "region seeding supports K = 1..7 (got %d)", K);
  // layout: region bytes padded by 16, codes padded to a multiple of SD2_SCAN entries
  std::vector<int64_t> refOff(n_regions + 1, 0), codeOff(n_regions + 1, 0);
  std::vector<int32_t> rlen(std::max<int64_t>(n_regions, 1), 0);
  int maxLen = 0;
  for (int64_t r = 0; r < n_regions; r++) {
    const int64_t l = region_off[r + 1] - region_off[r];
    if (l < 0 || l >= 65000) return vg_fail(c, VG_ERR_INVALID, "region %lld has unsupported length %lld", (long long)r, (long long)l);
    rlen[r] = (int32_t)l;
    maxLen = std::max(maxLen, (int)l);
    refOff[r + 1] = refOff[r] + ((l + 15) & ~(int64_t)15) + 16;
    codeOff[r + 1] = codeOff[r] + std::max<int64_t>(SD2_SCAN, (l + SD2_SCAN - 1) & ~(int64_t)(SD2_SCAN - 1));
  }
  for (int64_t i = 0; i < region_off[n_regions]; i++)
    if (region_pool[i] >= 127) return vg_fail(c, VG_ERR_ALPHABET, "region byte >= 127");
  std::vector<int32_t> len(n_reads);
  int maxRead = 0;
  for (int64_t i = 0; i < n_reads; i++) {
    const int64_t l = read_off[i + 1] - read_off[i];
    if (l < 0 || l > (1 << 20)) return vg_fail(c, VG_ERR_INVALID, "bad read offsets");
    if (read_region[i] < 0 || read_region[i] >= n_regions) return vg_fail(c, VG_ERR_INVALID, "read %lld names region %d", (long long)i, read_region[i]);
    len[i] = (int32_t)l;
    maxRead = std::max(maxRead, (int)l);
  }
  for (int64_t i = 0; i < read_off[n_reads]; i++)
    if (read_pool[i] >= 127) return vg_fail(c, VG_ERR_ALPHABET, "read byte >= 127");
  // exotic k-mers (a byte outside ACGT) of all regions: one sorted list, ids 4^K + rank
  std::vector<uint64_t> exo;
  for (int64_t r = 0; r < n_regions; r++) {
    const uint8_t* p = region_pool + region_off[r];
    for (int g = 0; g + K <= rlen[r]; g++) {
      uint32_t code;
      uint64_t key;
      if (!sdh_pack(p + g, K, &code, &key)) exo.push_back(key);
    }
  }
  std::sort(exo.begin(), exo.end());
  exo.erase(std::unique(exo.begin(), exo.end()), exo.end());
  std::vector<uint8_t> hRef((size_t)refOff[n_regions] + 16, 0);
  std::vector<uint16_t> hCode((size_t)codeOff[n_regions] + SD2_SCAN, (uint16_t)((1u << (2 * K)) + exo.size()));  // padding: the id of no k-mer
  for (int64_t r = 0; r < n_regions; r++) {
    const uint8_t* p = region_pool + region_off[r];
    memcpy(hRef.data() + refOff[r], p, rlen[r]);
    for (int g = 0; g + K <= rlen[r]; g++) {
      uint32_t code;
      uint64_t key;
      hCode[codeOff[r] + g] = sdh_pack(p + g, K, &code, &key)
                                  ? (uint16_t)code
                                  : (uint16_t)((1u << (2 * K)) + (std::lower_bound(exo.begin(), exo.end(), key) - exo.begin()));
    }
  }
  // device copies (scratch buffers that no mapping result depends on)
  VG_CUDA(c, c->d_s1pool.reserve((size_t)read_off[n_reads] + 64));
  VG_CUDA(c, c->d_s1off.reserve((size_t)(n_reads + 1) * 8));
  VG_CUDA(c, c->d_s2off.reserve((size_t)n_reads * 4 + 64));
  VG_CUDA(c, c->d_s2pool.reserve(hRef.size() + hCode.size() * 2 + (size_t)(n_regions + 1) * 20 + (size_t)n_reads * 4 + exo.size() * 8 + 1024));
  uint8_t* base = c->d_s2pool.as<uint8_t>();
  size_t o = 0;
  auto put = [&](const void* src, size_t bytes) -> uint8_t* {
    uint8_t* d = base + o;
    if (bytes) cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, c->stream);
    o = (o + bytes + 255) & ~(size_t)255;
    return d;
  };
  SeedRegions R;
  R.d_ref = put(hRef.data(), hRef.size());
  R.d_code = (const uint16_t*)put(hCode.data(), hCode.size() * 2);
  R.d_refOff = (const int64_t*)put(refOff.data(), refOff.size() * 8);
  R.d_codeOff = (const int64_t*)put(codeOff.data(), codeOff.size() * 8);
  R.d_len = (const int32_t*)put(rlen.data(), rlen.size() * 4);
  R.d_itemRegion = (const int32_t*)put(read_region, (size_t)n_reads * 4);
  R.d_exo = (const uint64_t*)put(exo.data(), exo.size() * 8);
  R.nExo = (int)exo.size();
  R.maxLen = maxLen;
  VG_CUDA(c, cudaMemcpyAsync(c->d_s1pool.p, read_pool, (size_t)read_off[n_reads], cudaMemcpyHostToDevice, c->stream));
  VG_CUDA(c, cudaMemcpyAsync(c->d_s1off.p, read_off, (size_t)(n_reads + 1) * 8, cudaMemcpyHostToDevice, c->stream));
  VG_CUDA(c, cudaMemcpyAsync(c->d_s2off.p, len.data(), (size_t)n_reads * 4, cudaMemcpyHostToDevice, c->stream));
  VG_CUDA(c, cudaGetLastError());
  VG_TRY(seed_launch(c, false, c->d_s1pool.as<uint8_t>(), c->d_s1off.as<int64_t>(), c->d_s2off.as<int32_t>(), nullptr, nullptr,
                     nullptr, n_reads, true, maxRead, 1, nullptr, nullptr, true, &R));
  VG_CUDA(c, cudaMemcpyAsync(windows, c->d_cand.p, (size_t)n_reads * 3 * 4, cudaMemcpyDeviceToHost, c->stream));
  VG_CUDA(c, cudaMemcpyAsync(found, c->d_ncand.p, (size_t)n_reads * 4, cudaMemcpyDeviceToHost, c->stream));
  VG_CUDA(c, cudaStreamSynchronize(c->stream));
  return VG_OK;
}

int vg_pileup_alignments(vg_ctx* c, const uint8_t* seq1_pool, const int64_t* seq1_off, const int32_t* length, const int64_t* pos,
                         const int64_t* end, int64_t n, int64_t total_positions, int32_t* counts) {
  if (!c) return VG_ERR_INVALID;
  VG_CUDA(c, cudaSetDevice(c->device));
  if (n < 0 || total_positions < 0 || (n > 0 && (!seq1_pool || !seq1_off || !length || !pos)) || (total_positions > 0 && !counts))
    return vg_fail(c, VG_ERR_INVALID, "bad arguments");
  if (total_positions == 0) return VG_OK;
  int64_t bytes = 0;
  for (int64_t i = 0; i < n; i++) {
    if (length[i] < 0 || seq1_off[i] < 0) return vg_fail(c, VG_ERR_INVALID, "bad alignment %lld", (long long)i);
    bytes = std::max(bytes, seq1_off[i] + length[i]);
  }
  VG_CUDA(c, c->d_s1pool.reserve((size_t)bytes + 64));
  VG_CUDA(c, c->d_s1off.reserve((size_t)std::max<int64_t>(n, 1) * 8));
  VG_CUDA(c, c->d_s2off.reserve((size_t)std::max<int64_t>(n, 1) * 4));
  VG_CUDA(c, c->d_s2pool.reserve((size_t)std::max<int64_t>(n, 1) * 16 + (size_t)total_positions * 20 + 256));
  int64_t* d_pos = c->d_s2pool.as<int64_t>();
  int64_t* d_end = d_pos + std::max<int64_t>(n, 1);
  int32_t* d_counts = (int32_t*)(c->d_s2pool.as<uint8_t>() + (((size_t)std::max<int64_t>(n, 1) * 16 + 255) & ~(size_t)255));
  VG_CUDA(c, c->d_counter.reserve(512));
  int* d_err = c->d_counter.as<int>();
  VG_CUDA(c, cudaMemsetAsync(d_err, 0, 4, c->stream));
  VG_CUDA(c, cudaMemsetAsync(d_counts, 0, (size_t)total_positions * 20, c->stream));
  if (n > 0) {
    VG_CUDA(c, cudaMemcpyAsync(c->d_s1pool.p, seq1_pool, (size_t)bytes, cudaMemcpyHostToDevice, c->stream));
    VG_CUDA(c, cudaMemcpyAsync(c->d_s1off.p, seq1_off, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
    VG_CUDA(c, cudaMemcpyAsync(c->d_s2off.p, length, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
    VG_CUDA(c, cudaMemcpyAsync(d_pos, pos, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
    if (end) VG_CUDA(c, cudaMemcpyAsync(d_end, end, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
    pu_alignments_kernel<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(c->d_s1pool.as<uint8_t>(), c->d_s1off.as<int64_t>(),
                                                                           c->d_s2off.as<int32_t>(), d_pos, end ? d_end : nullptr, n,
                                                                           total_positions, d_counts, d_err);
    VG_LAUNCH_CHECK(c);
  }
  int err = 0;
  VG_CUDA(c, cudaMemcpyAsync(&err, d_err, 4, cudaMemcpyDeviceToHost, c->stream));
  VG_CUDA(c, cudaMemcpyAsync(counts, d_counts, (size_t)total_positions * 20, cudaMemcpyDeviceToHost, c->stream));
  VG_CUDA(c, cudaStreamSynchronize(c->stream));
  if (err) return vg_fail(c, VG_ERR_ALPHABET, "an aligned byte lies in ('T', 'a') or an alignment runs past the end of its region (the Java would throw)");
  return VG_OK;
}

// Test hook (not part of the reference-facing ABI): LongHits of one read after mergeHits + splitLongHits.
int vg_debug_seed_hits(vg_ctx* c, int32_t msa, const uint8_t* read, int32_t L, int32_t* hits, int32_t cap, int32_t* n_hits) {
  int64_t off[2] = {0, L};
  if (!c) return VG_ERR_INVALID;
  // room for every window the read can have: pruned windows do not overlap and are at least K wide
  const int K = std::max(1, msa ? c->seedMsa.K : c->seed.K);
  const int mw = (msa ? c->msaLength : c->G) / K + 2 * (int)c->contigEnds.size() + 8;
  std::vector<int32_t> win((size_t)mw * 3);
  int32_t nw = 0;
  return mh_seed_batch(c, msa, read, off, 1, mw, win.data(), &nw, hits, cap, n_hits);
}

}  // extern "C"

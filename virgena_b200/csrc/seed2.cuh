// Seeding fast path for sm_100a: the single-reference KMerCounter chain split into two kernels so that each
// runs at 32 warps per SM (the monolithic seed_kernel of seed.cuh is limited to 14 by its per-warp shared
// memory and is latency bound there). Same restatement of the Java as seed.cuh (see its header for the
// file:line map); the phases are unchanged, only where the data lives between them:
//
//   seed_gen_kernel   A (read k-mer table) + B (genome-ordered raw hits -> LongHit stream)
//                     KMerCounterBase.getHits + mergeHits phase 1 (KMerCounterBase.java:337-421,452-466)
//                     -> per item: S[item][0..nS) = LongHits sorted by (genomePos, readPos), in HBM
//   seed_win_kernel   C (mergeHits phase 2, :422-449) + D (splitLongHits, KMerCounter.java:142-185)
//                     + E (concordanceArray, :266-279) + F (window counts, KMerCounter.java:72-221)
//                     + G (pruneWindows, KMerCounterBase.java:317-335) -> windows
//
// Anything the fast path cannot hold (a k-mer occurring > 127 times in a read, more LongHits than the per-item
// slot, a sweep segment longer than the shared-memory window, a full re-insertion queue) is not an error: the
// item is appended to a fallback list and redone by seed_kernel, which has the general (hash / serial) paths.
#pragma once
#include <type_traits>

#include "seed.cuh"

#ifndef SD2_PN
#define SD2_PN 8               // genome positions per lane and trip of the scan (4 or 8)
#endif
#define SD2_SCAN (32 * SD2_PN)   // positions per trip; the position-code tables are padded to a multiple of it

struct Seed2Args {
  // reference
  const uint8_t* blob;         // [ref padded to 16][poscode u16 padded to a multiple of 128 entries]
  int blobRefBytes, blobBytes;
  const uint64_t* exoKeys;
  int nExo;
  const int32_t* contigEnds;
  int nContigs;
  int G, K;
  float lowCoef, highCoef;
  const int32_t* cutoffs;
  int nCutoffs;
  // reads: global work item gi = itemBase + local item
  const uint8_t* bases;
  const int64_t* off1;
  const int32_t* len1;
  const int64_t* off2;
  const int32_t* len2;
  const int64_t* pairIdx;
  int64_t itemBase;
  int nItems;                  // items of this launch
  int singleReads;
  // per-item scratch
  uint64_t* S;                 // [nItems][capS] stream elements: g << 32 | r << 16 | len, or (reference shorter than 16384, reads up
                               // to 512 bases) 32 bits g << 18 | r << 9 | len — the element of the sweep window, SdWin<P32>::E
  int capS;
  int32_t* nS;                 // [nItems]; -1 = handed to the fallback
  // per-warp-slot scratch of the window kernel
  uint32_t* candKey;           // [warpSlots][2][capC]
  uint32_t* candVal;
  int capC;
  uint16_t* scratchR;          // [warpSlots][2][capR]: hitG, P when they do not fit in shared memory
  int capR;
  // outputs
  int32_t* windows;            // [all items][maxWin][3]
  int32_t* nWindows;
  int maxWin;
  int* counters;               // [0] gen work counter, [1] window work counter, [2] error flags, [3] fallback count,
                               // [4..8] fallbacks by reason, [10..15] three u64: LongHits in, LongHits kept, windows out
  int32_t* fallback;           // global item indices to redo with seed_kernel
  int32_t* dbgHits;
  int32_t* dbgNHits;
  unsigned long long* prof;    // optional: cycles per phase of the window kernel (VG_SEED_DEBUG=2)
  // geometry (host computed)
  int emptyId;                 // 4^K + nExo: a table slot that is never filled
  int tabBytes, nxtBytes, qcap, rdBytes, genWarpBytes;
  int capRs, nBuckets, winWarpBytes;
  // MSA mode (seed_gen_msa_kernel): CSR index id -> sorted distinct alignment columns, per-warp sort scratch
  const int32_t* msaOff;
  const int32_t* msaCols;
  uint32_t* rawBuf;            // [warpSlots][4][capRaw]
  int msaPacked;               // MSA stream kernel: LongHits packed into one word for the second sort
  int capRaw;
  int msaWarpBytes, idsBytes;
  // region mode (vg_map_to_regions): every work item has its own small reference (cluster centroid / reference
  // sub-sequence); tables live in global memory, nothing is staged
  int regions;
  const uint8_t* regRef;       // region bytes, each region padded by >= 16
  const uint16_t* regCode;     // per-position k-mer ids, each region padded with SD_NONE16 to a multiple of 128 entries
  const int64_t* regRefOff;
  const int64_t* regCodeOff;
  const int32_t* regLen;
  const int32_t* itemRegion;   // [all items]
  int bestOnly;                // 1: mapReadToRegion / computeKMerCount over [0, G): the first window of maximal count, no cutoff,
                               //    no splitLongHits, no pruning (KMerCounter.java:119-140, KMerCounterBase.java:468-487)
  int key24;                   // G < 16384: candidates sort on a 24-bit key (start << 10 | 1023 - (end - start)), 3 passes
};

__device__ __forceinline__ void sd2_item(const Seed2Args& A, int local, int64_t& off, int& L, int& rev) {
  const int64_t gi = A.itemBase + local;
  if (A.singleReads) {
    off = A.off1[gi];
    L = A.len1[gi];
    rev = 0;
  } else {
    int64_t p = gi >> 2;
    if (A.pairIdx) p = A.pairIdx[p];
    const int o = (int)(gi & 3);  // 0 fwd1, 1 rev1, 2 fwd2, 3 rev2
    off = (o < 2) ? A.off1[p] : A.off2[p];
    L = (o < 2) ? A.len1[p] : A.len2[p];
    rev = o & 1;
  }
}

// reasons: 0 k-mer count > 127 / queue, 1 slot full, 2 long segment, 3 re-insertion queue, 4 split / LongHit capacity
__device__ __forceinline__ void sd2_fallback(const Seed2Args& A, int local, int reason) {
  atomicAdd(A.counters + 4 + reason, 1);
  const int slot = atomicAdd(A.counters + 3, 1);
  A.fallback[slot] = (int32_t)(A.itemBase + local);
  A.nS[local] = -1;
}

// ------------------------------------------------------------------------------------------------
// Kernel 1: LongHit stream. One warp per read orientation; blockDim = 32 * warps.
// Dynamic shared memory: [staged blob][per warp: tab u16 | nxt u16 | queue u32[qcap + 4] | rd]
//   tab[id] = head | count << 9 (count 0 = the k-mer does not occur in the read), nxt = chain of ascending
//   read positions, queue = flattened raw hits (g << 16 | i) of the current batch of genome positions.
// ------------------------------------------------------------------------------------------------
// REGIONS: every work item has its own reference in global memory (vg_map_to_regions); otherwise the one reference is
// staged in shared memory and addressed as such.
template <bool P32>
struct SdWin;  // element of the LongHit stream and of the sweep window (below)

// sd_kmer_id for the common case of a k-mer of pure A/C/G/T: the K <= 7 bytes come from two unaligned 4-byte loads
// (rd is 4-byte aligned and padded), the codes from (b >> 1 & 3) ^ (b >> 2 & 1): A 0, C 1, G 2, T 3. Anything else takes the
// general function (exotic k-mers of the reference, or none).
__device__ __forceinline__ uint32_t sd2_kmer_id(const uint8_t* rd, int i, int K, const uint64_t* exoKeys, int nExo) {
  const uint32_t w[2] = {sd_load4(rd, i), sd_load4(rd, i + 4)};
  uint32_t code = 0;
  bool pure = true;
#pragma unroll
  for (int t = 0; t < 7; t++) {
    if (t < K) {
      const uint32_t b = (w[t >> 2] >> (8 * (t & 3))) & 0xFFu;
      const uint32_t x = b - 0x41u;
      pure &= (x < 32u) & (((0x00080045u >> (x & 31u)) & 1u) != 0u);  // A C G T
      uint32_t c = (b >> 1) & 3u;
      c ^= c >> 1;
      code = (code << 2) | c;
    }
  }
  return pure ? code : sd_kmer_id(rd + i, K, exoKeys, nExo);
}

template <bool REGIONS, bool P32>
__global__ void __launch_bounds__(1024, 1) seed_gen_kernel(Seed2Args A) {
  extern __shared__ __align__(16) uint8_t sd_smem[];
  __shared__ uint64_t sd_mbar;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int K = A.K;
  int G = A.G;
  if (!REGIONS) sd_tma_stage(sd_smem, A.blob, A.blobBytes, &sd_mbar);
  const uint8_t* ref = sd_smem;
  const uint16_t* poscode = (const uint16_t*)(sd_smem + A.blobRefBytes);
  uint8_t* wbase = sd_smem + A.blobBytes + (size_t)warp * A.genWarpBytes;
  uint16_t* tab = (uint16_t*)wbase;
  uint16_t* nxt = (uint16_t*)(wbase + A.tabBytes);
  uint32_t* queue = (uint32_t*)(wbase + A.tabBytes + A.nxtBytes);
  uint8_t* rd = (uint8_t*)(queue + A.qcap + 4) + 4;  // rd[-1] = 0xFF: no base in front of read position 0
  if (lane == 0) rd[-1] = 0xFFu;
  const int QCAP = A.qcap;
  const unsigned lt = (1u << lane) - 1u;

#pragma unroll 1
  for (;;) {
    int it = 0;
    if (lane == 0) it = atomicAdd(A.counters, 1);
    it = __shfl_sync(VG_FULL, it, 0);
    if (it >= A.nItems) break;
    int64_t off;
    int L, rev;
    sd2_item(A, it, off, L, rev);
    if (REGIONS) {
      const int rg = A.itemRegion[A.itemBase + it];
      ref = A.regRef + A.regRefOff[rg];
      poscode = A.regCode + A.regCodeOff[rg];
      G = A.regLen[rg];
    }
    const int nk = L - K + 1;
    if (nk < 1 || L > SD_MAXL) {  // L < K: no k-mer (Q6); too long: reported by the window kernel
      if (lane == 0) A.nS[it] = 0;
      continue;
    }
    const uint8_t* src = A.bases + off;
    __syncwarp();
#pragma unroll 1
    for (int i = lane; i < L; i += 32) rd[i] = rev ? vg_complement(src[L - 1 - i]) : src[i];
#pragma unroll 1
    for (int t = lane; t < (A.tabBytes >> 2); t += 32) ((uint32_t*)tab)[t] = 0;
    __syncwarp();
    // ---- A: read k-mer table, chains built from the last chunk down so that they ascend ------------
    bool bad = false;
    int why = 0;
#pragma unroll 1
    for (int base = ((nk - 1) >> 5) << 5; base >= 0; base -= 32) {
      const int i = base + lane;
      uint32_t id = SD_NONE16;
      if (i < nk) id = sd2_kmer_id(rd, i, K, A.exoKeys, A.nExo);
      const bool valid = id != SD_NONE16;
      const unsigned grp = __match_any_sync(VG_FULL, valid ? id : (0x80000000u | lane));
      if (valid) {
        const int leader = __ffs(grp) - 1;
        const uint32_t old = tab[id];  // read by the whole group before the leader overwrites it
        const unsigned higher = grp & ~((2u << lane) - 1u);
        nxt[i] = higher ? (uint16_t)(base + __ffs(higher) - 1) : (uint16_t)(old & 511u);
        __syncwarp(grp);
        if (lane == leader) {
          const uint32_t cnt = (old >> 9) + (uint32_t)__popc(grp);
          if (cnt > 127u) bad = true;
          tab[id] = (uint16_t)((uint32_t)i | (cnt << 9));
        }
      }
      __syncwarp();
    }
    bad = __any_sync(VG_FULL, bad);
    // ---- B: scan the reference in genome order -> LongHits sorted by (g, i) -------------------------
    typename SdWin<P32>::E* S = (typename SdWin<P32>::E*)A.S + (size_t)it * A.capS;
    int nS = 0;
#pragma unroll 1
    for (int gb = 0; gb <= G - K && !bad;) {
      int nq = 0;
      bool qfull = false;
      // B1: SD2_SCAN genome positions per trip, SD2_PN consecutive positions per lane; every raw hit (g, i) is queued in
      // (g, i) order. Offsets come from ballots over the bits of the per-lane hit count.
#pragma unroll 1
      for (; gb <= G - K && !qfull; gb += SD2_SCAN) {
        const int g0 = gb + SD2_PN * lane;
        uint32_t hv[SD2_PN];
#if SD2_PN == 8
        const uint4 pc = *(const uint4*)(poscode + g0);  // beyond G-K the table holds emptyId
        const uint32_t pw[4] = {pc.x, pc.y, pc.z, pc.w};
#else
        const uint2 pc = *(const uint2*)(poscode + g0);
        const uint32_t pw[2] = {pc.x, pc.y};
#endif
        int tot = 0;
#pragma unroll
        for (int u = 0; u < SD2_PN; u++) {
          const uint32_t code = (u & 1) ? (pw[u >> 1] >> 16) : pw[u >> 1];
          hv[u] = tab[code & 0x7FFFu];  // beyond the last k-mer the table holds emptyId: a slot that is never filled
          tot += hv[u] >> 9;
        }
        const unsigned b0 = __ballot_sync(VG_FULL, tot & 1), b1 = __ballot_sync(VG_FULL, tot & 2),
                       b2 = __ballot_sync(VG_FULL, tot & 4);
#if SD2_PN == 8
        const unsigned b3 = __ballot_sync(VG_FULL, tot & 8), bh = __ballot_sync(VG_FULL, tot >> 4);
#else
        const unsigned b3 = 0, bh = __ballot_sync(VG_FULL, tot >> 3);
#endif
        int excl, total;
        if (bh == 0) {
          total = __popc(b0) + 2 * __popc(b1) + 4 * __popc(b2) + 8 * __popc(b3);
          if (total == 0) continue;
          excl = __popc(b0 & lt) + 2 * __popc(b1 & lt) + 4 * __popc(b2 & lt) + 8 * __popc(b3 & lt);
        } else {
          const int incl = sd_warp_incl_scan(tot, lane);
          total = __shfl_sync(VG_FULL, incl, 31);
          excl = incl - tot;
        }
        if (total > QCAP) {  // low-complexity read: the general kernel has a path for it
          bad = true;
          break;
        }
        if (nq + total > QCAP) {  // does not fit any more: flush first, then redo this block
          qfull = true;
          gb -= SD2_SCAN;
          continue;
        }
        int o = nq + excl;
#pragma unroll
        for (int u = 0; u < SD2_PN; u++) {  // chains of 1 or 2 read positions are the rule: straight-line code with
          const int cu = hv[u] >> 9;        // predicated stores (@P STS; a select of a spare slot cost one more instruction each)
          const uint32_t i0 = hv[u] & 511u;
          // bit 15: the position was indexed twice (Q1); the read position needs 9 bits
          const uint32_t gk = ((uint32_t)(g0 + u) << 16) | (((u & 1) ? (pw[u >> 1] >> 16) : pw[u >> 1]) & 0x8000u);
          if (cu) queue[o] = gk | i0;
          uint32_t i = nxt[i0];
          if (cu > 1) queue[o + 1] = gk | i;
          if (cu > 2) {
#pragma unroll 1
            for (int t = 2; t < cu; t++) {
              i = nxt[i];
              queue[o + t] = gk | i;
            }
          }
          o += cu;
        }
        nq += total;
      }
      __syncwarp();
      // B2: one lane per raw hit (two per trip): chain-start test, extension, ballot-compacted ordered writes
#pragma unroll 1
      for (int q0 = 0; q0 < nq && !bad; q0 += 64) {
        // straight-line for both hits of the lane: chain-start test, first 4 bytes of the extension (they decide almost
        // every chain), Q1 flag; only chains that go on take the loop below
        bool em[2], more[2];
        int gg[2], ii[2], ll[2], mx[2];
#pragma unroll
        for (int u = 0; u < 2; u++) {
          const int q = q0 + 32 * u + lane;
          const bool ok = q < nq;
          const uint32_t e = ok ? queue[q] : 0u;
          const int g = (int)(e >> 16), i = (int)(e & 0x1FFu);
          gg[u] = g, ii[u] = i;
          const uint8_t ca = rd[i - 1], cb = ref[max(g - 1, 0)];  // rd[-1] never equals a reference byte
          const bool start = (g == 0) | (ca != cb);
          const int maxLen = min(L - i - K, G - g - K);
          const uint32_t x0 = sd_load4(rd, i + K) ^ sd_load4(ref, g + K);
          const int len = x0 ? (__ffs(x0) - 1) >> 3 : 4;
          // Q1: a chain passing through a twice-indexed position adds a zero-length hit
          em[u] = ok & (start | ((e & 0x8000u) != 0));
          ll[u] = start ? min(len, maxLen) : 0;
          mx[u] = maxLen;
          more[u] = ok & start & (x0 == 0) & (maxLen > 4);
        }
        if (more[0] | more[1]) {
#pragma unroll
          for (int u = 0; u < 2; u++) {
            if (!more[u]) continue;
            int len = 4;
#pragma unroll 1
            while (len < mx[u]) {
              const uint32_t x = sd_load4(rd, ii[u] + K + len) ^ sd_load4(ref, gg[u] + K + len);
              if (x) {
                len += (__ffs(x) - 1) >> 3;
                break;
              }
              len += 4;
            }
            ll[u] = min(len, mx[u]);
          }
        }
        const unsigned m0 = __ballot_sync(VG_FULL, em[0]);
        const unsigned m1 = __ballot_sync(VG_FULL, em[1]);
        const int t0 = __popc(m0), total = t0 + __popc(m1);
        if (nS + total > A.capS) {
          bad = true;
          why = 1;
          break;
        }
        if (em[0]) S[nS + __popc(m0 & lt)] = SdWin<P32>::pack(gg[0], ii[0], ll[0]);
        if (em[1]) S[nS + t0 + __popc(m1 & lt)] = SdWin<P32>::pack(gg[1], ii[1], ll[1]);
        nS += total;
      }
      __syncwarp();
    }
    if (lane == 0) {
      if (bad) sd2_fallback(A, it, why);
      else A.nS[it] = nS;
    }
  }
}

#define SD2_PH(k)                                                            \
  if (A.prof) {                                                              \
    __syncwarp();                                                            \
    const long long t2 = clock64();                                          \
    if (lane == 0) atomicAdd(A.prof + (k), (unsigned long long)(t2 - tph)); \
    tph = t2;                                                                \
  }
// Stable LSD radix sort on the low 8 x passes bits of the keys, values optional (v0 == nullptr). The caller has already
// counted the digits of every key into binsAll[pass][256] (shared memory), e.g. while it produced the keys, so the buffers — which live in L2 / HBM — are read only by the two scatter passes, with the next four chunks
// of keys in flight while four are ranked. Ping-pongs between (k0, v0) and (k1, v1); returns which pair holds the result.
__device__ __forceinline__ void sd2_scan_bins(int* bins, int lane) {
  int loc[8], sum = 0;
#pragma unroll
  for (int t = 0; t < 8; t++) {
    loc[t] = bins[lane * 8 + t];
    sum += loc[t];
  }
  int run = sd_warp_incl_scan(sum, lane) - sum;
#pragma unroll
  for (int t = 0; t < 8; t++) {
    bins[lane * 8 + t] = run;
    run += loc[t];
  }
}

__device__ __noinline__ int sd2_radix_sort_counted(uint32_t* k0, uint32_t* k1, uint32_t* v0, uint32_t* v1, int n, int* binsAll,
                                                   int passes, int lane, int bits = 8) {
  const uint32_t dmask = (1u << bits) - 1u;  // digit width (the bins of a pass stay 256 apart)
#pragma unroll 1
  for (int pass = 0; pass < passes; pass++) {
    const uint32_t* ks = (pass & 1) ? k1 : k0;
    const uint32_t* vs = (pass & 1) ? v1 : v0;
    uint32_t* kd = (pass & 1) ? k0 : k1;
    uint32_t* vd = (pass & 1) ? v0 : v1;
    int* bins = binsAll + 256 * pass;
    const int shift = pass * bits;
    __syncwarp();
    sd2_scan_bins(bins, lane);
    __syncwarp();
    uint32_t kn[4], vn[4];  // the next four chunks are requested while the current four are ranked (serial part)
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int e = 32 * u + lane;
      kn[u] = (e < n) ? ks[e] : 0;
      vn[u] = (e < n && v0) ? vs[e] : 0;
    }
#pragma unroll 1
    for (int e0 = 0; e0 < n; e0 += 128) {
      uint32_t kk[4], vv[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        kk[u] = kn[u], vv[u] = vn[u];
        const int e = e0 + 128 + 32 * u + lane;
        kn[u] = (e < n) ? ks[e] : 0;
        vn[u] = (e < n && v0) ? vs[e] : 0;
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        if (e0 + 32 * u >= n) break;
        const int e = e0 + 32 * u + lane;
        const bool ok = e < n;
        const uint32_t d = ok ? ((kk[u] >> shift) & dmask) : (256u + lane);
        const unsigned grp = __match_any_sync(VG_FULL, d);
        const int rank = __popc(grp & ((1u << lane) - 1));
        const int leader = __ffs(grp) - 1;
        int base = 0;
        if (ok && lane == leader) {
          base = bins[d];
          bins[d] = base + __popc(grp);
        }
        base = __shfl_sync(VG_FULL, base, leader);
        if (ok) {
          kd[base + rank] = kk[u];
          if (v0) vd[base + rank] = vv[u];
        }
        __syncwarp();
      }
    }
    __threadfence_block();
    __syncwarp();
  }
  return passes & 1;  // 0: result in (k0, v0), 1: in (k1, v1)
}

// ------------------------------------------------------------------------------------------------
// Kernel 1, MSA flavour (KMerCounterMSA.getNBestRegions, KMerCounterMSA.java:71): the index maps a k-mer to the
// sorted distinct alignment columns where it starts in some reference (ReferenceAlignment.java:104-209), so a
// column holds many k-mers and the genome-ordered scan of seed_gen_kernel does not apply. Instead:
//   1. every raw hit (i, c) is listed read position by read position (lane = read position, its list is contiguous);
//   2. stable sort by diagonal c - i (descending, see 4.): a chain of mergeHits phase 1 (KMerCounterBase.java:341-421,
//      hits on one diagonal at consecutive read positions) is now a run of consecutive i inside one diagonal;
//   3. run starts -> LongHits (c, i, len = run length - 1);
//   4. stable sort by c: LongHits with equal c come in descending-diagonal = ascending-i order, the TreeSet order.
// Dynamic shared memory per warp: ids u16[idsBytes / 2] | binsLo int[256] | binsHi int[256] | rd.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024, 1) seed_gen_msa_kernel(Seed2Args A) {
  extern __shared__ __align__(16) uint8_t sd_smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int warpSlot = blockIdx.x * (blockDim.x >> 5) + warp;
  const int K = A.K, G = A.G;
  uint8_t* wbase = sd_smem + (size_t)warp * A.msaWarpBytes;
  uint16_t* ids = (uint16_t*)wbase;
  int* binsLo = (int*)(wbase + A.idsBytes);
  int* binsHi = binsLo + 256;
  uint8_t* rd = (uint8_t*)(binsHi + 256);
  const int capRaw = A.capRaw;
  uint32_t* K0 = A.rawBuf + (size_t)warpSlot * 4 * capRaw;
  uint32_t* K1 = K0 + capRaw;
  uint32_t* V0 = K1 + capRaw;
  uint32_t* V1 = V0 + capRaw;
  const unsigned lt = (1u << lane) - 1u;
  const uint32_t dBias = (uint32_t)(G - 1);  // key = dBias - (c - i) in [0, G + L): larger diagonal first

#pragma unroll 1
  for (;;) {
    int it = 0;
    if (lane == 0) it = atomicAdd(A.counters, 1);
    it = __shfl_sync(VG_FULL, it, 0);
    if (it >= A.nItems) break;
    int64_t off;
    int L, rev;
    sd2_item(A, it, off, L, rev);
    const int nk = L - K + 1;
    if (nk < 1 || L > SD_MAXL) {  // NULL_SEQ / shorter than K: no k-mer (MapperMSA.java:57)
      if (lane == 0) A.nS[it] = 0;
      continue;
    }
    long long tph = A.prof ? clock64() : 0;
    const uint8_t* src = A.bases + off;
    __syncwarp();
#pragma unroll 1
    for (int i = lane; i < L; i += 32) rd[i] = rev ? vg_complement(src[L - 1 - i]) : src[i];
    __syncwarp();
#pragma unroll 1
    for (int i = lane; i < nk; i += 32) ids[i] = (uint16_t)sd_kmer_id(rd + i, K, A.exoKeys, A.nExo);
    __syncwarp();
#pragma unroll 1
    for (int b = lane; b < 512; b += 32) binsLo[b] = 0;
    __syncwarp();
    // 1. raw hits, read position by read position (the two digits of every sort key are counted on the way)
    int nRaw = 0, why = 0;
    bool bad = false;
#pragma unroll 1
    for (int base = 0; base < nk; base += 32) {
      const int i = base + lane;
      int lo = 0, hi = 0;
      if (i < nk && ids[i] != SD_NONE16) {
        lo = A.msaOff[ids[i]];
        hi = A.msaOff[ids[i] + 1];
      }
      const int cnt = hi - lo;
      const int incl = sd_warp_incl_scan(cnt, lane);
      const int total = __shfl_sync(VG_FULL, incl, 31);
      if (nRaw + total > capRaw) {
        bad = true;
        break;
      }
      // the lists are of very different lengths: the entries of the 32 lists are dealt out evenly, entry t of the
      // round to lane t mod 32 (its read position = the last lane whose exclusive offset is <= t, by bisection over the
      // lanes), so that loads and stores are coalesced and no lane waits for the longest list
      const int excl = incl - cnt;
      const int src0 = lo - excl;  // msaCols index of round entry t when this lane owns it: src0 + t
#pragma unroll 1
      for (int t0 = 0; t0 < total; t0 += 64) {  // two trips at a time: both column loads are in flight together
        int own[2], col[2];
#pragma unroll
        for (int u = 0; u < 2; u++) {
          const int t = t0 + 32 * u + lane;
          int owner = 0;
#pragma unroll
          for (int d = 16; d >= 1; d >>= 1) {
            const int ex = __shfl_sync(VG_FULL, excl, owner + d);  // owner + d <= 31
            if (ex <= t) owner += d;
          }
          const int oSrc = __shfl_sync(VG_FULL, src0, owner);
          own[u] = owner;
          col[u] = (t < total) ? A.msaCols[oSrc + t] : 0;
        }
#pragma unroll
        for (int u = 0; u < 2; u++) {
          const int t = t0 + 32 * u + lane;
          if (t < total) {
            const uint32_t io = (uint32_t)(base + own[u]);
            const uint32_t key = (io << 16) + dBias + io - (uint32_t)col[u];  // low half: dBias - (c - i)
            K0[nRaw + t] = key;
            atomicAdd(&binsLo[key & 255u], 1);
            atomicAdd(&binsHi[(key >> 8) & 255u], 1);
          }
        }
      }
      nRaw += total;
    }
    __syncwarp();
    SD2_PH(8)
    if (A.prof && lane == 0) atomicAdd(A.prof + 13, (unsigned long long)nRaw);
    int nSt = 0;
    uint64_t* S = A.S + (size_t)it * A.capS;            // wide stream; packed stream (msaPacked): 32-bit elements
    uint32_t* S32 = (uint32_t*)A.S + (size_t)it * A.capS;
    if (!bad && nRaw > 0) {
      __threadfence_block();
      // 2. stable sort by diagonal (16-bit key in the low half, read position rides in the high half)
      sd2_radix_sort_counted(K0, K1, nullptr, nullptr, nRaw, binsLo, 2, lane);
      const uint32_t* ks = K0;
#pragma unroll 1
      for (int b = lane; b < 512; b += 32) binsLo[b] = 0;  // for the second sort
      __syncwarp();
      SD2_PH(9)
      // 3. chain starts (first hit of a diagonal, or read position not the successor of the previous hit's) -> LongHits
      //    in one pass: a start is closed by the next start (len = distance - 1); the last start of a chunk of 32 stays
      //    open (warp-uniform) until a later chunk, or the end of the list, closes it. (The digits of the second sort's
      //    keys are counted on the way.)
      // packed (alignment shorter than 16384 columns, reads up to 512): column | readPos << 14 | len << 23 in one word,
      // sorted by two 7-bit digits of the column, so that only K0 / K1 are touched (they stay in L2)
      const bool pk = A.msaPacked != 0;
      auto emit = [&](int idx, uint32_t k, int len) {
        const int i = (int)(k >> 16);
        const int c = (int)dBias - (int)(k & 0xFFFFu) + i;
        if (pk) {
          K1[idx] = (uint32_t)c | ((uint32_t)i << 14) | ((uint32_t)len << 23);
          atomicAdd(&binsLo[c & 127], 1);
          atomicAdd(&binsHi[(c >> 7) & 127], 1);
        } else {
          V0[idx] = ((uint32_t)i << 16) | (uint32_t)c;
          V1[idx] = (uint32_t)len;
          atomicAdd(&binsLo[c & 255], 1);
          atomicAdd(&binsHi[(c >> 8) & 255], 1);
        }
      };
      bool haveOpen = false;
      uint32_t openKey = 0;
      int openE = 0;
#pragma unroll 1
      for (int e0 = 0; e0 < nRaw; e0 += 128) {
        uint32_t kc[4], kp[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int e = e0 + 32 * u + lane;
          kc[u] = (e < nRaw) ? ks[e] : 0;
          kp[u] = (e < nRaw && e > 0) ? ks[e - 1] : ~kc[u];
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int eb = e0 + 32 * u, e = eb + lane;
          const bool st = e < nRaw && !(((kc[u] ^ kp[u]) & 0xFFFFu) == 0 && (kc[u] >> 16) == (kp[u] >> 16) + 1);
          const unsigned m = __ballot_sync(VG_FULL, st);
          if (m) {
            if (haveOpen && lane == 0) emit(nSt, openKey, eb + __ffs(m) - 1 - openE - 1);
            const int base = nSt + (haveOpen ? 1 : 0);
            const unsigned higher = m & ~((2u << lane) - 1u);
            if (st && higher) emit(base + __popc(m & lt), kc[u], __ffs(higher) - 1 - lane - 1);
            nSt = base + __popc(m) - 1;
            const int lastLane = 31 - __clz(m);
            openKey = __shfl_sync(VG_FULL, kc[u], lastLane);
            openE = eb + lastLane;
            haveOpen = true;
          }
        }
      }
      if (haveOpen) {
        if (lane == 0) emit(nSt, openKey, nRaw - openE - 1);
        nSt++;
      }
      __syncwarp();
      __threadfence_block();
      SD2_PH(10)
      if (nSt > A.capS) {
        bad = true;
        why = 1;
      } else {
        SD2_PH(11)
        // 4. stable sort by column; K0 / K1 are free again and serve as the other halves
        if (pk) {
          sd2_radix_sort_counted(K1, K0, nullptr, nullptr, nSt, binsLo, 2, lane, 7);
#pragma unroll 4
          for (int e = lane; e < nSt; e += 32) {
            const uint32_t k = K1[e];
            S32[e] = ((k & 0x3FFFu) << 18) | (((k >> 14) & 0x1FFu) << 9) | (k >> 23);  // SdWin<true>::pack
          }
        } else {
          sd2_radix_sort_counted(V0, K0, V1, K1, nSt, binsLo, 2, lane);
          const uint32_t* k2 = V0;
          const uint32_t* v2 = V1;
#pragma unroll 4
          for (int e = lane; e < nSt; e += 32) {
            const uint32_t k = k2[e];
            S[e] = ((uint64_t)(k & 0xFFFFu) << 32) | ((uint64_t)(k >> 16) << 16) | (uint64_t)v2[e];
          }
        }
      }
    }
    __syncwarp();
    SD2_PH(12)
    if (lane == 0) {
      if (bad) sd2_fallback(A, it, why);
      else A.nS[it] = nSt;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Kernel 2: overlap resolution, window counts, pruning. One warp per read orientation.
// Per-warp shared memory (winWarpBytes), two aliased uses:
//   C   : win (SdWin: 512 x u64 or 1024 x u32, + 32 spare slots; re-inserted LongHits go back into it)
//   E-G : hitG u16[capRs + 8] | P u16[capRs] | len u8[capRs] | bucket u16[nBuckets]   (bins int[256] of the sort
//         overlay hitG); len saturates at 255 (then the exact value is re-read from the stream)
// ------------------------------------------------------------------------------------------------
// The sweep window of phase C. Wide: the stream's own 64-bit entries, 512 of them (16 per lane). Packed (reference
// shorter than 16384, reads up to 512): genomePos << 18 | readPos << 9 | len in 32 bits, 1024 entries in the same shared
// memory (32 per lane): half as many windows per read, 32-bit shared-memory traffic. One spare slot per lane chunk keeps
// the lanes' chunks on different banks.
template <bool P32>
struct SdWin;
template <>
struct SdWin<false> {
  using E = uint64_t;
  static constexpr int N = SD_WIN, CH = SD_WIN / 32, SH = 4;
  static __device__ __forceinline__ E pack(int g, int r, int l) { return ((uint64_t)g << 32) | ((uint64_t)r << 16) | (uint64_t)l; }
  static __device__ __forceinline__ int g(E v) { return (int)(v >> 32); }
  static __device__ __forceinline__ int r(E v) { return (int)((v >> 16) & 0xFFFFu); }
  static __device__ __forceinline__ int l(E v) { return (int)(v & 0xFFFFu); }
  static __device__ __forceinline__ E withLen(E v, int l) { return (v & ~(uint64_t)0xFFFFu) | (uint64_t)l; }
  static __device__ __forceinline__ uint64_t key(E v) { return v >> 16; }  // (genomePos, readPos): the TreeSet order
  static __device__ __forceinline__ uint64_t mkkey(int g, int r) { return ((uint64_t)g << 16) | (uint64_t)r; }
};
template <>
struct SdWin<true> {
  using E = uint32_t;
  static constexpr int N = 2 * SD_WIN, CH = 2 * SD_WIN / 32, SH = 5;
  static __device__ __forceinline__ E pack(int g, int r, int l) { return ((uint32_t)g << 18) | ((uint32_t)r << 9) | (uint32_t)l; }
  static __device__ __forceinline__ int g(E v) { return (int)(v >> 18); }
  static __device__ __forceinline__ int r(E v) { return (int)((v >> 9) & 0x1FFu); }
  static __device__ __forceinline__ int l(E v) { return (int)(v & 0x1FFu); }
  static __device__ __forceinline__ E withLen(E v, int l) { return (v & ~0x1FFu) | (uint32_t)l; }
  static __device__ __forceinline__ uint32_t key(E v) { return v >> 9; }
  static __device__ __forceinline__ uint32_t mkkey(int g, int r) { return ((uint32_t)g << 9) | (uint32_t)r; }
};
template <bool P32>
__device__ __forceinline__ int sd2_wi(int e) { return e + (e >> SdWin<P32>::SH); }

// TreeSet.add of a clipped LongHit (KMerCounterBase.java:435-441) into the sorted window: it goes in front of the first
// unpolled element with a larger (genomePos, readPos); dropped when that key is already there (Q2). The slot of the
// element polled last is free (the in-place output is always at least two behind), so the elements in between move
// down by one and the stream index steps back. Its genomePos is the end of prev, which lies before the next
// synchronisation point, i.e. inside [si, s1]. Returns the new stream index.
template <bool P32>
__device__ __noinline__ int sd2_reinsert(typename SdWin<P32>::E* win, int si, int s1, int g, int r, int len) {
  using W = SdWin<P32>;
  const auto key = W::mkkey(g, r);
  int p = si;
#pragma unroll 1
  while (p < s1 && W::key(win[sd2_wi<P32>(p)]) < key) p++;
  if (p < s1 && W::key(win[sd2_wi<P32>(p)]) == key) return si;
#pragma unroll 1
  for (int j = si; j < p; j++) win[sd2_wi<P32>(j - 1)] = win[sd2_wi<P32>(j)];
  win[sd2_wi<P32>(p - 1)] = W::pack(g, r, len);
  return si - 1;
}

template <bool P32>
__global__ void __launch_bounds__(1024, 1) seed_win_kernel(Seed2Args A) {
  extern __shared__ __align__(16) uint8_t sd_smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int warpSlot = blockIdx.x * (blockDim.x >> 5) + warp;
  const int K = A.K;
  int G = A.G;
  uint8_t* wbase = sd_smem + (size_t)warp * A.winWarpBytes;
  using W = SdWin<P32>;
  typename W::E* win = (typename W::E*)wbase;
  const int capRs = A.capRs;
  uint16_t* bucket = (uint16_t*)(wbase + (size_t)capRs * 5 + 16);
  int* bins = (int*)wbase;
  const int capC = A.capC;
  uint32_t* ckey = A.candKey + (size_t)warpSlot * 2 * capC;
  uint32_t* cval = A.candVal + (size_t)warpSlot * 2 * capC;
  const unsigned lt = (1u << lane) - 1u;
  // contig ends: binary-searched per hit in phases D and F; a copy in shared memory when there are few of them
  __shared__ int32_t sd_contigs[64];  // the host takes references with more than 64 contigs to the general kernel
  const int32_t* cEnds = sd_contigs;
  if (A.nContigs > 1) {
    if (threadIdx.x < A.nContigs) sd_contigs[threadIdx.x] = A.contigEnds[threadIdx.x];
    __syncthreads();
  }

#pragma unroll 1
  for (;;) {
    int it = 0;
    if (lane == 0) it = atomicAdd(A.counters + 1, 1);
    it = __shfl_sync(VG_FULL, it, 0);
    if (it >= A.nItems) break;
    long long tph = A.prof ? clock64() : 0;
    const int64_t gi = A.itemBase + it;
    int64_t off;
    int L, rev;
    sd2_item(A, it, off, L, rev);
    const int nS = A.nS[it];
    if (nS < 0) continue;  // redone by the general kernel
    if (A.regions) G = A.regLen[A.itemRegion[gi]];
    SD2_PH(0)
    if (L > SD_MAXL && lane == 0) atomicOr(A.counters + 2, 16);
    using SE = typename SdWin<P32>::E;  // stream element = window element (32 bits packed or 64 bits)
    SE* S = (SE*)A.S + (size_t)it * A.capS;
    // the stream was written by the other kernel and comes from HBM: pull all of it into L2 now, so that only the
    // first window pays the DRAM latency
    constexpr int perLine = 128 / (int)sizeof(SE);
#pragma unroll 1
    for (int e = lane * perLine; e < nS; e += 32 * perLine) asm volatile("prefetch.global.L2 [%0];" ::"l"(S + e));
    int nWin = 0, nR = 0;
    bool giveUp = false;
    int why = 4;
    __syncwarp();
    // ---- C: mergeHits phase 2 (KMerCounterBase.java:422-449); see seed.cuh for the decomposition into
    // independent segments at synchronisation points. The LongHits that survive are written back to the front
    // of the item's slot (never ahead of what has been read into the window).
    if (nS > 0) {
      constexpr int CH = W::CH;  // entries per lane when flagging
      int tbase = 0, runMaxE = -1;
#pragma unroll 1
      while (tbase < nS) {
        const int cnt = min(W::N, nS - tbase);
#pragma unroll 4
        for (int e = lane; e < cnt; e += 32) win[sd2_wi<P32>(e)] = S[tbase + e];
        __syncwarp();
        const int lo = lane * CH, nOwn = max(0, min(CH, cnt - lo));
        const typename W::E* wl = win + lane * (CH + 1);  // = win + sd2_wi(lo): the lane's chunk is contiguous
        int locMax = -1;
#pragma unroll 4
        for (int t = 0; t < nOwn; t++) {
          const typename W::E v = wl[t];
          locMax = max(locMax, W::g(v) + W::l(v) + K);
        }
        int incl = locMax;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int t = __shfl_up_sync(VG_FULL, incl, d);
          if (lane >= d) incl = max(incl, t);
        }
        int excl = __shfl_up_sync(VG_FULL, incl, 1);
        excl = (lane == 0) ? runMaxE : max(excl, runMaxE);
        uint32_t fmask = 0;
        int mLast = excl;  // running maximum end in front of the lane's last flagged entry
        {
          int m = excl;
#pragma unroll 4
          for (int t = 0; t < nOwn; t++) {
            const typename W::E v = wl[t];
            const int g = W::g(v);
            if (g > m) {
              fmask |= 1u << t;
              mLast = m;
            }
            m = max(m, g + W::l(v) + K);
          }
        }
        // lane l sweeps the segments that START in its chunk: from its first flag to the first flag of the next
        // flagged lane; the last segment of a window that is not the last one is carried over (tailStart)
        const int nSeg = __reduce_add_sync(VG_FULL, __popc(fmask));
        const int firstFlag = fmask ? lo + __ffs(fmask) - 1 : 0x7fff;
        const int lastFlag = fmask ? lo + 31 - __clz(fmask) : -1;
        const bool more = tbase + cnt < nS;
        if ((more ? nSeg - 1 : nSeg) <= 0) {  // a single segment longer than the window
          giveUp = true;
          why = 2;
          break;
        }
        const int tailStart = more ? __reduce_max_sync(VG_FULL, lastFlag) : cnt;
        int nextFirst;
        {
          int v = firstFlag;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_down_sync(VG_FULL, v, d);
            if (lane + d < 32) v = min(v, o);
          }
          nextFirst = __shfl_down_sync(VG_FULL, v, 1);
          if (lane == 31) nextFirst = 0x7fff;
        }
        // maximum end over the entries consumed now = what the owner of the carried-over flag saw in front of it
        if (more) runMaxE = __shfl_sync(VG_FULL, mLast, tailStart / CH);
        const int s0 = min(firstFlag, tailStart), s1 = min(nextFirst, tailStart);  // s0 == s1: nothing to do
        __syncwarp();
        // All lanes step through their ranges in lock-step, one polled element per trip: the sweep of
        // KMerCounterBase.java:422-449, started afresh at a synchronisation point; results in place. The common
        // cases are branch-free: with prev = (pg, pr, pl) and the polled element (cg, cr, cl),
        //   no overlap                      -> emit prev, prev = curr              (:443-446)
        //   overlap, prev shorter           -> emit prev truncated to cg - pg - K if >= 0, prev = curr (:428-433)
        //   overlap, prev at least as long  -> curr clipped on the left and re-inserted if it keeps >= K (:435-441)
        int si = s0, w = s0;
        bool have = false;
        typename W::E pv = 0;  // prev, as it sits in the window
        // one polled element per lane; branch-free except for the (rare) re-insertion. A lane only touches its own
        // range of the window, so the trips need no barrier.
        auto step = [&]() {
          const bool act = si < s1;
          const typename W::E v = win[sd2_wi<P32>(act ? si : s0)];  // TreeSet.pollFirst
          const int cg = W::g(v), cl = W::l(v);
          const int pg = W::g(pv), pl = W::l(pv);
          const int pe = pg + pl + K;
          const bool ov = have & (pe > cg);  // prev.genomePos + prev.len + K - 1 >= curr.genomePos
          const bool keepPrev = ov & (pl >= cl);
          const int tl = cg - pg - K;
          if (act & have & !keepPrev & (!ov | (tl >= 0))) {
            win[sd2_wi<P32>(w)] = ov ? W::withLen(pv, tl) : pv;
            w++;
          }
          const int rest = cl - (pe - cg);
          if (act) {
            si++;
            have = true;
            if (!keepPrev) pv = v;
            else if (rest >= 0) si = sd2_reinsert<P32>(win, si, s1, pe, W::r(v) + (pe - cg), rest);
          }
        };
        {
          const int n0 = __reduce_max_sync(VG_FULL, s1 - s0);  // without re-insertions the longest range decides
#pragma unroll 1
          for (int t = 0; t < n0; t++) step();
#pragma unroll 1
          while (__any_sync(VG_FULL, si < s1)) step();
        }
        if (have) {
          win[sd2_wi<P32>(w)] = pv;
          w++;
        }
        __syncwarp();
        {  // ordered copy-out: every lane appends its run of survivors behind those of the lanes below
          const int c = w - s0;
          const int ic = sd_warp_incl_scan(c, lane);
          SE* dst = S + nR + ic - c;
#pragma unroll 1
          for (int t = 0; t < c; t++) dst[t] = win[sd2_wi<P32>(s0 + t)];
          nR += __shfl_sync(VG_FULL, ic, 31);
        }
        __syncwarp();
        tbase += tailStart;
      }
    }
    __syncwarp();
    SD2_PH(1)
    // ---- D: KMerCounter.splitLongHits (KMerCounter.java:142-185); hits are independent given their contig
    if (!giveUp && A.nContigs > 1 && nR > 0 && !A.bestOnly) {
      __threadfence_block();
      // the pieces go to the upper half of the item's slot (a hit yields at most 3), which then replaces the lower half
      const int half = A.capS >> 1;
      SE* S2 = S + half;
      int nOut = 0;
      bool bad = nR > half;
#pragma unroll 1
      for (int j0 = 0; j0 < nR && !bad; j0 += 32) {
        const int j = j0 + lane;
        SE q0 = 0, q1 = 0, q2 = 0;  // the pieces, packed (three named registers: an indexed array would live in local memory)
        int np2 = 0;
        auto put = [&](int g_, int r_, int l_) {
          const SE e = W::pack(g_, r_, l_);
          if (np2 == 0) q0 = e;
          else if (np2 == 1) q1 = e;
          else if (np2 == 2) q2 = e;
          np2++;
        };
        if (j < nR) {
          const SE v = S[j];
          int g = W::g(v), r = W::r(v), len = W::l(v);
          for (;;) {
            int lo2 = 0, hi2 = A.nContigs - 1;
#pragma unroll 1
            while (lo2 < hi2) {
              const int mid = (lo2 + hi2) >> 1;
              if (cEnds[mid] > g) hi2 = mid;
              else lo2 = mid + 1;
            }
            const int contigEnd = cEnds[lo2];
            if (g + K + len > contigEnd) {
              if (contigEnd - g >= K) put(g, r, contigEnd - g - K);
              if (g + K + len - contigEnd >= K) {
                r += contigEnd - g;
                len -= contigEnd - g;
                g = contigEnd;
                continue;
              }
              break;
            }
            put(g, r, len);
            break;
          }
          if (np2 > 3) bad = true, np2 = 3;  // a hit across > 2 contig boundaries (contigs shorter than K + 1)
        }
        const unsigned m1 = __ballot_sync(VG_FULL, np2 >= 1), m2 = __ballot_sync(VG_FULL, np2 >= 2), m3 = __ballot_sync(VG_FULL, np2 >= 3);
        int o = nOut + __popc(m1 & lt) + __popc(m2 & lt) + __popc(m3 & lt);
        nOut += __popc(m1) + __popc(m2) + __popc(m3);
        if (__any_sync(VG_FULL, bad) || nOut > half) {
          bad = true;
          break;
        }
        if (np2 >= 1) S2[o] = q0;
        if (np2 >= 2) S2[o + 1] = q1;
        if (np2 >= 3) S2[o + 2] = q2;
      }
      __syncwarp();
      __threadfence_block();
      if (__any_sync(VG_FULL, bad)) {
        giveUp = true;
      } else {
        S = S2;
        nR = nOut;
      }
    }
    if (giveUp || nR > A.capR) {
      if (lane == 0) sd2_fallback(A, it, why);
      continue;
    }
    __threadfence_block();
    if (A.dbgHits && gi == 0) {
#pragma unroll 1
      for (int j = lane; j < nR; j += 32) {
        const SE v = S[j];
        A.dbgHits[3 * j] = W::g(v);
        A.dbgHits[3 * j + 1] = W::r(v);
        A.dbgHits[3 * j + 2] = W::l(v);
      }
      if (lane == 0) *A.dbgNHits = nR;
    }
    if (nR > 0) {
      // ---- E: isAdjacent (IEEE binary32 division, Q3 coefficients verbatim) + prefix sums
      //         P[j] = sum_{t<=j}(len_t + isAdjacent[t-1]) (mod 2^16); genomePos table + buckets for the searches
      const int cutoff = A.bestOnly ? -1 : (L < A.nCutoffs) ? A.cutoffs[L] : 0x7fffffff;
      int nC = 0;
      int bestCnt = -1, bestJ = 0x7fffffff, bestS = 0, bestE = 0;  // bestOnly: per lane, first maximum in hit order
      // The arrays live in shared memory when the LongHits fit (the rule) and in L2-resident per-warp scratch otherwise;
      // the phases are instantiated once per address space so that the common case is compiled to LDS, not generic loads.
      auto phaseEF = [&](auto inShared, uint16_t* hitG, uint16_t* Pp, uint8_t* hitL) {
        (void)inShared;
        if (lane < 8) hitG[nR + lane] = 0xFFFFu;  // sentinels for the branch-free searches of phase F
        // bucket[b] = first hit index with g >= 32 b: the first hit of every bucket is marked on the way through the loop
        // below (it has the previous hit's genomePos at hand), the empty buckets are filled from the right afterwards
        const int nb = A.nBuckets;
  #pragma unroll 1
        for (int b = lane; b < nb; b += 32) bucket[b] = (uint16_t)nR;
        __syncwarp();
        int carry = 0, pgLast = 0, prLast = 0;
        SE vNext = (lane < nR) ? S[lane] : 0;  // the slot is in L2: the next chunk is requested one trip ahead
  #pragma unroll 1
        for (int j0 = 0; j0 < nR; j0 += 32) {
          const int j = j0 + lane;
          int val = 0, g = 0, r = 0;
          const SE v = vNext;
          if (j + 32 < nR) vNext = S[j + 32];
          if (j < nR) {
            g = W::g(v), r = W::r(v);
            val = W::l(v);
            hitG[j] = (uint16_t)g;
            hitL[j] = (uint8_t)min(val, 255);
          }
          int pg = __shfl_up_sync(VG_FULL, g, 1), pr = __shfl_up_sync(VG_FULL, r, 1);
          if (lane == 0) pg = pgLast, pr = prLast;
          if (j < nR && j > 0) {
            const float rate = __fdiv_rn((float)(g - pg), (float)(r - pr));
            if (A.lowCoef <= rate && rate <= A.highCoef) val++;
          }
          if (j < nR && (j == 0 || (pg >> 5) != (g >> 5))) bucket[g >> 5] = (uint16_t)j;
          const int incl = sd_warp_incl_scan(val, lane);
          if (j < nR) Pp[j] = (uint16_t)(carry + incl);
          carry += __shfl_sync(VG_FULL, incl, 31);
          pgLast = __shfl_sync(VG_FULL, g, 31);
          prLast = __shfl_sync(VG_FULL, r, 31);
        }
        __syncwarp();
        {  // suffix fill of empty buckets (serial over chunks from the top, parallel inside a chunk)
          int nextv = nR;
  #pragma unroll 1
          for (int b0 = ((nb - 1) >> 5) << 5; b0 >= 0; b0 -= 32) {
            const int b = b0 + lane;
            int v = (b < nb) ? bucket[b] : nR;
            const bool empty = (v == nR);
  #pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
              const int o = __shfl_down_sync(VG_FULL, v, d);
              if (lane + d < 32) v = min(v, o);
            }
            v = min(v, nextv);
            if (b < nb && empty) bucket[b] = (uint16_t)v;
            nextv = __shfl_sync(VG_FULL, v, 0);
          }
        }
        __syncwarp();
        SD2_PH(2)
        // ---- F: every window independently (updateCount equals a direct recount on sorted non-overlapping
        //         hits); candidates (count > cutoff, strict) compacted in hit order
        SE vNextF = (lane < nR) ? S[lane] : 0;
  #pragma unroll 1
        for (int j0 = 0; j0 < nR; j0 += 32) {
          const int j = j0 + lane;
          bool isCand = false;
          uint32_t key = 0, cnt = 0;
          const SE v = vNextF;
          if (j + 32 < nR) vNextF = S[j + 32];
          if (j < nR) {
            const int g = W::g(v), r = W::r(v), len = W::l(v);
            int cs = 0, ce = G;
            if (A.nContigs > 1 && !A.bestOnly) {
              int lo = 0, hi = A.nContigs - 1;
  #pragma unroll 1
              while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (cEnds[mid] > g) hi = mid;
                else lo = mid + 1;
              }
              ce = cEnds[lo];
              cs = lo ? cEnds[lo - 1] : 0;
            }
            const int start = max(cs, g - r * 3 / 2);
            const int end = min(g + len + K + (L - r - len - K) * 3 / 2, ce);
            // a = first hit with genomePos >= start, u = first with genomePos >= end: a 32-position bucket holds at most
            // 32 / K + 1 hits (they do not overlap), three branch-free halving steps cover 8; the loops are the safety net
            int a = bucket[start >> 5];
            a += (hitG[a + 3] < start) ? 4 : 0;
            a += (hitG[a + 1] < start) ? 2 : 0;
            a += (hitG[a] < start) ? 1 : 0;
  #pragma unroll 1
            while (hitG[a] < start) a++;
            int u = bucket[min(end, G) >> 5];
            u += (hitG[u + 3] < end) ? 4 : 0;
            u += (hitG[u + 1] < end) ? 2 : 0;
            u += (hitG[u] < end) ? 1 : 0;
  #pragma unroll 1
            while (hitG[u] < end) u++;  // stops at the sentinel at the latest
            int b = u - 1;
            int lenA = hitL[a], lenB = hitL[b];
            if (lenA == 255) lenA = W::l(S[a]);
            if (lenB == 255) lenB = W::l(S[b]);
            if ((int)hitG[b] + K + lenB > end) b--;
            cnt = (uint32_t)((uint16_t)(Pp[b] - Pp[a])) + (uint32_t)lenA;
            isCand = (int)cnt > cutoff;
            if (A.bestOnly) {
              if ((int)cnt > bestCnt) bestCnt = (int)cnt, bestJ = j, bestS = start, bestE = end;
              isCand = false;
            }
            // order (start asc, end desc); a window is at most 1.5 L <= 768 wide
            key = A.key24 ? ((uint32_t)start << 10) | (uint32_t)(1023 - (end - start)) : ((uint32_t)start << 16) | (uint32_t)(0xFFFF - end);
          }
          const unsigned m = __ballot_sync(VG_FULL, isCand);
          if (isCand) {
            const int o = nC + __popc(m & lt);
            ckey[o] = key;
            cval[o] = cnt;
          }
          nC += __popc(m);
        }
      };
      if (nR <= capRs) {
        uint16_t* hs = (uint16_t*)wbase;
        phaseEF(std::true_type{}, hs, hs + capRs + 8, (uint8_t*)(hs + 2 * capRs + 8));
      } else {
        uint16_t* hg = A.scratchR + (size_t)warpSlot * 3 * (A.capR + 8);
        phaseEF(std::false_type{}, hg, hg + A.capR + 8, (uint8_t*)(hg + 2 * (A.capR + 8)));
      }
      __syncwarp();
      if (A.bestOnly) {  // strict '>' in hit order: the first hit that reaches the maximum wins
        const int mx = __reduce_max_sync(VG_FULL, bestCnt);
        const int jw = __reduce_min_sync(VG_FULL, bestCnt == mx ? bestJ : 0x7fffffff);
        if (bestCnt == mx && bestJ == jw) {
          int32_t* W = A.windows + (size_t)gi * A.maxWin * 3;
          W[0] = bestS, W[1] = bestE, W[2] = mx;
        }
        nWin = 1;
      }
      SD2_PH(3)
      // ---- G: stable sort by (start asc, end desc) + pruneWindows (KMerCounterBase.java:317-335). The pivot only
      //         changes at the first later candidate that either does not overlap it (emit, new pivot) or overlaps
      //         with a larger count (new pivot); the warp looks at 32 candidates at a time.
      if (nC > 0) {
        const int passes = A.key24 ? 3 : 4;
        __threadfence_block();
#pragma unroll 1
        for (int b = lane; b < 256 * passes; b += 32) bins[b] = 0;  // hitG / P / len are dead now
        __syncwarp();
#pragma unroll 4
        for (int e = lane; e < nC; e += 32) {  // all digits counted in one pass over the keys
          const uint32_t k = ckey[e];
          atomicAdd(&bins[k & 255u], 1);
          atomicAdd(&bins[256 + ((k >> 8) & 255u)], 1);
          atomicAdd(&bins[512 + ((k >> 16) & 255u)], 1);
          if (passes == 4) atomicAdd(&bins[768 + (k >> 24)], 1);
        }
        __syncwarp();
        const int srcb = sd2_radix_sort_counted(ckey, ckey + capC, cval, cval + capC, nC, bins, passes, lane);
        SD2_PH(4)
        const uint32_t* ks = ckey + (size_t)srcb * capC;
        const uint32_t* vs = cval + (size_t)srcb * capC;
        int32_t* W = A.windows + (size_t)gi * A.maxWin * 3;
        const bool k24 = A.key24 != 0;
        const uint32_t k0 = ks[0];
        int ps = k24 ? k0 >> 10 : k0 >> 16, pe = k24 ? ps + 1023 - (int)(k0 & 1023u) : 0xFFFF - (int)(k0 & 0xFFFFu), pc = vs[0];
        uint32_t kNext = (1 + lane < nC) ? ks[1 + lane] : 0u, cNext = (1 + lane < nC) ? vs[1 + lane] : 0u;
#pragma unroll 1
        for (int t0 = 1; t0 < nC; t0 += 32) {
          const int idx = t0 + lane;
          const bool valid = idx < nC;
          const uint32_t k = kNext;
          const int cc = (int)cNext;
          if (idx + 32 < nC) kNext = ks[idx + 32], cNext = vs[idx + 32];  // the next batch is requested one trip ahead
          const int cs_ = k24 ? k >> 10 : k >> 16, ce_ = k24 ? cs_ + 1023 - (int)(k & 1023u) : 0xFFFF - (int)(k & 0xFFFFu);
          int from = 0;  // lanes below `from` are already behind the pivot
#pragma unroll 1
          for (;;) {
            const unsigned m = __ballot_sync(VG_FULL, valid && lane >= from && (pe <= cs_ || pc < cc));
            if (!m) break;
            const int f = __ffs(m) - 1;
            const int ns = __shfl_sync(VG_FULL, cs_, f), ne = __shfl_sync(VG_FULL, ce_, f), nc = __shfl_sync(VG_FULL, cc, f);
            if (pe <= ns) {
              if (lane == 0 && nWin < A.maxWin) W[3 * nWin] = ps, W[3 * nWin + 1] = pe, W[3 * nWin + 2] = pc;
              nWin++;
            }
            ps = ns, pe = ne, pc = nc;
            from = f + 1;
          }
        }
        if (lane == 0 && nWin < A.maxWin) W[3 * nWin] = ps, W[3 * nWin + 1] = pe, W[3 * nWin + 2] = pc;
        nWin++;
        if (nWin > A.maxWin) {
          if (lane == 0) atomicOr(A.counters + 2, 8);
          nWin = 0;
        }
      }
    }
    if (lane == 0) {
      A.nWindows[gi] = nWin;
      unsigned long long* sums = (unsigned long long*)(A.counters + 10);
      atomicAdd(sums, (unsigned long long)nS);
      atomicAdd(sums + 1, (unsigned long long)nR);
      atomicAdd(sums + 2, (unsigned long long)nWin);
    }
    __syncwarp();
    SD2_PH(5)
  }
}

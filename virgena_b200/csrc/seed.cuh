// Seeding + candidate-window voting for sm_100a.
//
// Replaces, per read orientation, the chain
//   KMerCounterBase.getHits -> mergeHits (KMerCounterBase.java:337-466)
//   KMerCounter.splitLongHits            (KMerCounter.java:142-185)
//   KMerCounterBase.concordanceArray     (KMerCounterBase.java:266-279)
//   getInitState / computeCount / updateCount / moveWindow (KMerCounter.java:72-221, KMerCounterBase.java:13-191)
//   pruneWindows                         (KMerCounterBase.java:317-335)
// i.e. KMerCounter.getNBestRegions (KMerCounter.java:223) and KMerCounterMSA.getNBestRegions
// (KMerCounterMSA.java:71), and then the pair selection of Mapper.mapRead / MapperMSA.mapRead.
//
// How the sequential Java is restated for a warp (one warp = one read orientation):
//  A. the read's k-mers go into a per-warp shared-memory hash (k-mer id -> ascending read positions);
//  B. single reference: the reference's per-position k-mer ids (staged in shared memory by a TMA bulk
//     copy) are scanned in genome order; a raw hit (read pos i, genome pos g) is the start of a LongHit
//     iff the hit (i-1, g-1) does not exist, i.e. iff i==0 || g==0 || read[i-1] != ref[g-1], and its
//     `len` is the number of further matching bases. This is exactly phase 1 of mergeHits (diagonal
//     chaining over consecutive read positions) and yields the LongHits already ordered by
//     (genomePos, readPos), the TreeSet order of phase 2. A position that StringIndexer indexed twice
//     (Q1) contributes one extra zero-length LongHit whenever a chain passes through it. Hits are
//     compacted with warp ballot-free prefix scans (counts -> exclusive scan -> ordered writes).
//     MSA: raw hits come from the CSR index (k-mer id -> sorted columns); chain starts/lengths are found
//     by membership tests in the neighbouring lists, then the LongHits are radix-sorted by (column, i).
//  C. phase 2 of mergeHits (overlap resolution with re-insertion into the TreeSet) is inherently serial
//     and is executed verbatim by one lane over the sorted stream plus a small pending queue;
//     splitLongHits is applied as each LongHit is emitted.
//  E/F. after C the LongHits are sorted and non-overlapping, and then updateCount's incremental window
//     count provably equals a direct recount, so every window is evaluated independently:
//     count_i = P[b] - P[a] + len_a with P the prefix sums of (len_j + isAdjacent[j-1]).
//  G. candidates (count > cutoff) are stably radix-sorted by (start asc, end desc) and pruned by the
//     serial sweep of pruneWindows.
#pragma once
#include "common.cuh"

#define SD_NONE16 0x7FFFu       // k-mer id "absent" in the 16-bit position-code table
#define SD_EMPTY 0xFFFFFFFFu    // empty hash slot
#define SD_NOPOS 0xFFFFu        // end of a read-position chain
#define SD_HS 1024              // hash slots (>= 2 x k-mers of a 512-base read)
#define SD_MAXL 512             // longest supported read
#define SD_PENDING 32           // capacity of the re-insertion queue of phase C
#define SD_WIN 512              // entries of the shared-memory window over the sorted LongHit stream
#define SD_LPEND 11             // per-lane re-insertion queue of the segment-parallel sweep
// padded window index (one spare slot per 16 entries: no bank conflicts when every lane walks its own 16-entry chunk)
__device__ __forceinline__ int SD_WI(int e) { return e + (e >> 4); }

struct SeedGeom {       // sizes of the per-warp shared-memory regions (bytes), computed on the host
  int regA;             // bytes of the phase-aliased region
  int capR;             // max LongHits after phase C (per orientation)
  int nBuckets;         // (coordinate range >> 5) + 2
  int warpBytes;        // total per warp
  int refBytes;         // staged reference blob (0 = not staged)
};

struct SeedArgs {
  // reference / index
  const uint8_t* ref;          // G bytes
  const uint16_t* poscode;     // per position: id | dup << 15 (SD_NONE16 beyond G-K)
  const uint8_t* blob;         // [ref padded to 16][poscode padded to 16] for the TMA stage
  int blobRefBytes, blobBytes;
  const uint64_t* exoKeys;     // sorted packed exotic k-mers
  int nExo;
  const int32_t* contigEnds;
  int nContigs;
  int G;                       // coordinate range: genome length or MSA length
  int K;
  float lowCoef, highCoef;
  const int32_t* cutoffs;
  int nCutoffs;
  // MSA index (CSR over ids)
  int lockstep;                // 1: static work assignment with a CTA barrier per item
  int direct;                  // 1: ids < 2048 -> direct-addressed read table instead of the hash
  int msa;
  const int32_t* msaOff;
  const int32_t* msaCols;
  // reads: work item w -> (pool offset, length, reverse-complement flag)
  const uint8_t* bases;
  const int64_t* off1;
  const int32_t* len1;
  const int64_t* off2;
  const int32_t* len2;
  const int64_t* pairIdx;      // optional indirection (remap subset)
  int64_t nItems;              // 4 x pairs (pair mode) or number of single reads
  const int32_t* itemList;     // optional: nItems indices of the items to do (fallback of the fast path)
  int singleReads;             // 1: items are single forward reads (vg_seed_batch)
  int bestOnly;                // 1: mapReadToRegion / computeKMerCount over [0, G): first window of maximal count, no cutoff,
                               //    no splitLongHits, no pruning (see seed2.cuh)
  // region mode (vg_map_to_regions): item w is seeded against its own small reference, tables in global memory
  const uint8_t* regRef;
  const uint16_t* regCode;
  const int64_t* regRefOff;
  const int64_t* regCodeOff;
  const int32_t* regLen;
  const int32_t* itemRegion;
  // scratch + outputs
  uint64_t* scratchS;          // [warpSlots][capS]
  uint16_t* scratchR;          // [warpSlots][4][capR]: hitG, hitR, hitL, P
  int capS;
  uint32_t* candKey;           // [warpSlots][2][capC]
  uint32_t* candVal;           // [warpSlots][2][capC]
  int capC;                    // >= capR; larger in MSA mode (holds the unsorted raw LongHits)
  int32_t* windows;            // [nItems][maxWin][3]
  int32_t* nWindows;           // [nItems]
  int maxWin;
  int* counter;                // work counter
  int* errFlag;                // 1: capS overflow, 2: capR overflow, 4: pending overflow, 8: window overflow
  int32_t* dbgHits;            // optional: [capR*3] LongHits of item 0 after phase C (tests)
  int32_t* dbgNHits;
  unsigned long long* prof;    // optional: cycles per phase [A,B,C,D,E,F,G,total] (VG_SEED_PROF=1)
  SeedGeom geom;
};

__device__ __forceinline__ uint32_t sd_hash(uint32_t id) { return (id * 2654435761u) >> 22; }  // 10 bits

__device__ __forceinline__ int sd_code2(uint8_t b) {  // A,C,G,T -> 0..3, else -1
  return b == 'A' ? 0 : b == 'C' ? 1 : b == 'G' ? 2 : b == 'T' ? 3 : -1;
}

// id of the K-mer at p: 2-bit code when purely ACGT, else 4^K + rank among the reference's exotic
// k-mers, or SD_NONE16 when the reference has no such k-mer.
__device__ __noinline__ uint32_t sd_kmer_id(const uint8_t* p, int K, const uint64_t* exoKeys, int nExo) {
  uint32_t code = 0;
  uint64_t key = 0;
  bool pure = true;
  for (int t = 0; t < K; t++) {
    uint8_t b = p[t];
    int c = sd_code2(b);
    pure &= c >= 0;
    code = (code << 2) | (uint32_t)(c & 3);
    key = (key << 8) | b;
  }
  if (pure) return code;
  int lo = 0, hi = nExo;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (exoKeys[mid] < key) lo = mid + 1;
    else hi = mid;
  }
  if (lo < nExo && exoKeys[lo] == key) return (1u << (2 * K)) + (uint32_t)lo;
  return SD_NONE16;
}

// 4 bytes starting at base + pos (base 4-byte aligned; may read up to 7 bytes past pos)
__device__ __forceinline__ uint32_t sd_load4(const uint8_t* base, int pos) {
  const uint32_t* p = (const uint32_t*)(base + (pos & ~3));
  return __funnelshift_r(p[0], p[1], (pos & 3) * 8);
}
// number of further matching bases of the chain that starts with the K-mer hit (i, g): 4 bytes per trip
__device__ __forceinline__ int sd_extend(const uint8_t* rd, const uint8_t* ref, int i, int g, int K, int L, int G) {
  const int maxLen = min(L - i - K, G - g - K);
  // the first 4 bytes decide almost every chain: straight-line; longer matches take the loop
  const uint32_t x0 = sd_load4(rd, i + K) ^ sd_load4(ref, g + K);
  int len = x0 ? (__ffs(x0) - 1) >> 3 : 4;
  if (x0 == 0 && maxLen > 4) {
#pragma unroll 1
    while (len < maxLen) {
      const uint32_t x = sd_load4(rd, i + K + len) ^ sd_load4(ref, g + K + len);
      if (x) {
        len += (__ffs(x) - 1) >> 3;
        break;
      }
      len += 4;
    }
  }
  return min(len, maxLen);
}

__device__ __forceinline__ int sd_warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int n = __shfl_up_sync(VG_FULL, v, d);
    if (lane >= d) v += n;
  }
  return v;
}

// Stable LSD radix sort of n (key,val) pairs by 32-bit key, 8 bits per pass; buffers k[2][cap], v[2][cap].
// Returns the index (0/1) of the buffer holding the result. `bins` = 256 ints of shared memory.
__device__ __noinline__ int sd_radix_sort(uint32_t* kbuf, uint32_t* vbuf, int cap, int n, int* bins, int lane, uint32_t keyBits) {
  int src = 0;
  for (int shift = 0; shift < (int)keyBits; shift += 8) {
    uint32_t* ks = kbuf + (size_t)src * cap;
    uint32_t* vs = vbuf + (size_t)src * cap;
    uint32_t* kd = kbuf + (size_t)(src ^ 1) * cap;
    uint32_t* vd = vbuf + (size_t)(src ^ 1) * cap;
    for (int b = lane; b < 256; b += 32) bins[b] = 0;
    __syncwarp();
    for (int e = lane; e < n; e += 32) atomicAdd(&bins[(ks[e] >> shift) & 255], 1);
    __syncwarp();
    {  // exclusive scan of the 256 bins: 8 per lane
      int loc[8], s = 0;
#pragma unroll
      for (int t = 0; t < 8; t++) {
        loc[t] = bins[lane * 8 + t];
        s += loc[t];
      }
      int incl = sd_warp_incl_scan(s, lane);
      int run = incl - s;
#pragma unroll
      for (int t = 0; t < 8; t++) {
        bins[lane * 8 + t] = run;
        run += loc[t];
      }
    }
    __syncwarp();
    for (int e0 = 0; e0 < n; e0 += 32) {
      int e = e0 + lane;
      bool ok = e < n;
      uint32_t k = ok ? ks[e] : 0, v = ok ? vs[e] : 0;
      uint32_t d = ok ? ((k >> shift) & 255u) : (256u + lane);
      unsigned grp = __match_any_sync(VG_FULL, d);
      int rank = __popc(grp & ((1u << lane) - 1));
      int leader = __ffs(grp) - 1;
      int base = 0;
      if (ok && lane == leader) {
        base = bins[d];
        bins[d] = base + __popc(grp);
      }
      base = __shfl_sync(VG_FULL, base, leader);
      if (ok) {
        kd[base + rank] = k;
        vd[base + rank] = v;
      }
      __syncwarp();
    }
    src ^= 1;
    __syncwarp();
  }
  return src;
}

// TMA bulk copy global -> shared with an mbarrier (cp.async.bulk; UBLKCP in SASS).
__device__ __forceinline__ void sd_tma_stage(uint8_t* smemDst, const uint8_t* gsrc, int bytes, uint64_t* mbar) {
  const uint32_t bar = (uint32_t)__cvta_generic_to_shared(mbar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes));
    int done = 0;
    while (done < bytes) {
      int chunk = min(bytes - done, 32768);
      uint32_t dst = (uint32_t)__cvta_generic_to_shared(smemDst + done);
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
          "l"(gsrc + done), "r"(chunk), "r"(bar)
          : "memory");
      done += chunk;
    }
  }
  // everyone waits for phase 0 of the barrier
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(0));
  }
}

// ------------------------------------------------------------------------------------------------
// The seeding kernel. blockDim = 32 * warpsPerBlock. Dynamic shared memory:
//   [staged reference blob][per-warp regions ...]
// Per-warp region (geom.warpBytes):
//   R-region   : hitG u16[capR], hitR u16[capR], hitL u16[capR], P u16[capR]
//                (during phases A/B it is aliased by hkey u32[SD_HS], hval u32[SD_HS], nxt u16[SD_MAXL])
//   aux        : pending u64[SD_PENDING] | bins int[256] | bucket u16[nBuckets] | win u64[SD_WIN]
//   rd         : u8[SD_MAXL + 16], ids u16[SD_MAXL] (MSA)
// ------------------------------------------------------------------------------------------------
template <bool REGIONS>
__global__ void __launch_bounds__(512) seed_kernel(SeedArgs A) {
  extern __shared__ __align__(16) uint8_t sd_smem[];
  __shared__ uint64_t sd_mbar;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int warpsPerBlock = blockDim.x >> 5;
  const int warpSlot = blockIdx.x * warpsPerBlock + warp;
  const int K = A.K;
  int G = A.G;

  const uint8_t* ref = A.ref;
  const uint16_t* poscode = A.poscode;
  if (!REGIONS && A.geom.refBytes > 0) {
    sd_tma_stage(sd_smem, A.blob, A.blobBytes, &sd_mbar);
    ref = sd_smem;
    poscode = (const uint16_t*)(sd_smem + A.blobRefBytes);
  }
  // Per-warp shared memory (geom.warpBytes):
  //   regA  (SD_REGA bytes): phases A/B: tab u32[2048] (or hash hkey/hval u32[1024] each) + nxt u16[SD_MAXL];
  //                          phases C..G (aliased): pending | bins | bucket | pendL | segBuf
  //   win   (SD_WIN x 8)   : B: raw-hit queue; C: window over the sorted stream; MSA B': ids
  //   rd    (SD_MAXL + 16) : the read in its orientation
  // The LongHit arrays (hitG/hitR/hitL/P, capR each) live in an L2-resident per-warp global scratch.
  uint8_t* wbase = sd_smem + A.geom.refBytes + (size_t)warp * A.geom.warpBytes;
  const int capR = A.geom.capR;
  uint16_t* hitG = A.scratchR + (size_t)warpSlot * 4 * capR;
  uint16_t* hitR = hitG + capR;
  uint16_t* hitL = hitR + capR;
  uint16_t* Pp = hitL + capR;
  uint32_t* hkey = (uint32_t*)wbase;
  uint32_t* hval = hkey + SD_HS;
  uint16_t* nxt = (uint16_t*)(hval + SD_HS);
  uint8_t* aux = wbase;  // aliases regA after phase B
  // (the first SD_PENDING x 8 bytes of aux were the serial sweep's re-insertion queue; it now lives in ckey)
  int* bins = (int*)(aux + SD_PENDING * 8);                            // 256 x 4
  uint16_t* bucket = (uint16_t*)(aux + SD_PENDING * 8 + 1024);         // nBuckets x 2
  uint64_t* pendL = (uint64_t*)(aux + SD_PENDING * 8 + 1024 + ((A.geom.nBuckets * 2 + 15) & ~15));  // 32 x SD_LPEND x 8
  uint64_t* win = (uint64_t*)(wbase + A.geom.regA);                    // (SD_WIN + 32) x 8 (padded, SD_WI)
  uint8_t* rd = (uint8_t*)(win + SD_WIN + 32);
  uint16_t* ids = (uint16_t*)win;                                      // MSA only (win is unused until phase C)

  uint64_t* S = A.scratchS + (size_t)warpSlot * A.capS;
  const int capC = A.capC;
  uint32_t* ckey = A.candKey + (size_t)warpSlot * 2 * capC;
  uint32_t* cval = A.candVal + (size_t)warpSlot * 2 * capC;

  // Work distribution. lockstep: static round-robin with a CTA barrier per item, so that all warps of the
  // SM run the same phase at the same time (the kernel is instruction-cache bound otherwise); else a
  // global atomic work counter.
  const int totalWarps = gridDim.x * warpsPerBlock;
#pragma unroll 1
  for (int64_t itBase = 0;; itBase += totalWarps) {
    int wi = 0;
    if (A.lockstep) {
      __syncthreads();
      if (itBase >= A.nItems) break;
      if (itBase + warpSlot >= A.nItems) continue;
      wi = (int)(itBase + warpSlot);
    } else {
      if (lane == 0) wi = atomicAdd(A.counter, 1);
      wi = __shfl_sync(VG_FULL, wi, 0);
      if (wi >= A.nItems) break;
    }
    if (A.itemList) wi = A.itemList[wi];
    if (REGIONS) {
      const int rg = A.itemRegion[wi];
      ref = A.regRef + A.regRefOff[rg];
      poscode = A.regCode + A.regCodeOff[rg];
      G = A.regLen[rg];
    }
    hitG = A.scratchR + (size_t)warpSlot * 4 * capR;  // (phase E may have redirected it to shared memory)
    // ---- which read, which orientation ---------------------------------------------------------
    int64_t off;
    int L, rev;
    if (A.singleReads) {
      off = A.off1[wi];
      L = A.len1[wi];
      rev = 0;
    } else {
      int64_t p = wi >> 2;
      if (A.pairIdx) p = A.pairIdx[p];
      const int o = wi & 3;  // 0 fwd1, 1 rev1, 2 fwd2, 3 rev2
      off = (o < 2) ? A.off1[p] : A.off2[p];
      L = (o < 2) ? A.len1[p] : A.len2[p];
      rev = o & 1;
    }
    int nWin = 0;
    const int nk = L - K + 1;
    const long long tph0 = A.prof ? clock64() : 0;
    // MapperMSA.mapReadFast short-circuits NULL_SEQ (Q6); the single-reference path relies on L < K.
    if (nk >= 1 && L <= SD_MAXL) {
      const uint8_t* src = A.bases + off;
      __syncwarp();
#pragma unroll 1
      for (int i = lane; i < L; i += 32) rd[i] = rev ? vg_complement(src[L - 1 - i]) : src[i];
#pragma unroll 1
      for (int s = lane; s < SD_HS; s += 32) hkey[s] = SD_EMPTY;
      __syncwarp();
      int nS = 0;
      bool overflow = false;
      long long tph = A.prof ? clock64() : 0;
#define SD_PH(k)                                                        \
  if (A.prof) {                                                         \
    __syncwarp();                                                       \
    long long t2 = clock64();                                           \
    if (lane == 0) atomicAdd(A.prof + (k), (unsigned long long)(t2 - tph)); \
    tph = t2;                                                           \
  }
      if (!A.msa) {
        // ---- A: read k-mer table (id -> ascending chain of read positions) ----------------------
        // direct mode: tab[id] = head | count << 16 (ids < 2048, i.e. K <= 5); else an open-addressing hash.
        uint32_t* tab = hkey;  // 2048 x u32 (aliases hkey + hval)
        bool direct = A.direct != 0;
      retry_hash:
        if (direct) {
#pragma unroll 1
          for (int t = lane; t < 2 * SD_HS; t += 32) tab[t] = SD_NOPOS;
          __syncwarp();
#pragma unroll 1
          for (int base = ((nk - 1) >> 5) << 5; base >= 0; base -= 32) {
            const int i = base + lane;
            uint32_t id = SD_NONE16;
            if (i < nk) id = sd_kmer_id(rd + i, K, A.exoKeys, A.nExo);
            const bool valid = id != SD_NONE16;
            const unsigned grp = __match_any_sync(VG_FULL, valid ? id : (0x80000000u | lane));
            if (valid) {
              const int leader = __ffs(grp) - 1;
              const uint32_t old = tab[id];  // read by the whole group before the leader overwrites it
              const unsigned higher = grp & ~((2u << lane) - 1u);
              nxt[i] = higher ? (uint16_t)(base + __ffs(higher) - 1) : (uint16_t)(old & 0xFFFFu);
              __syncwarp(grp);
              if (lane == leader) tab[id] = (uint32_t)i | ((((old >> 16) + (uint32_t)__popc(grp)) & 0xFFFFu) << 16);
            }
            __syncwarp();
          }
        } else {
#pragma unroll 1
        for (int s2 = lane; s2 < SD_HS; s2 += 32) hkey[s2] = SD_EMPTY;
        __syncwarp();
#pragma unroll 1
        for (int base = ((nk - 1) >> 5) << 5; base >= 0; base -= 32) {
          const int i = base + lane;
          uint32_t id = SD_NONE16;
          if (i < nk) id = sd_kmer_id(rd + i, K, A.exoKeys, A.nExo);
          const bool valid = id != SD_NONE16;
          const unsigned grp = __match_any_sync(VG_FULL, valid ? id : (0x80000000u | lane));
          const int leader = __ffs(grp) - 1;
          uint32_t slot = 0, old = SD_NOPOS;  // old = head | count << 16
          if (valid && lane == leader) {
            slot = sd_hash(id);
            for (;;) {
              uint32_t k = atomicCAS(&hkey[slot], SD_EMPTY, id);
              if (k == SD_EMPTY) break;
              if (k == id) {
                old = hval[slot];
                break;
              }
              slot = (slot + 1) & (SD_HS - 1);
            }
          }
          slot = __shfl_sync(VG_FULL, slot, leader);
          old = __shfl_sync(VG_FULL, old, leader);
          if (valid) {
            const unsigned higher = grp & ~((2u << lane) - 1u);
            nxt[i] = higher ? (uint16_t)(base + __ffs(higher) - 1) : (uint16_t)(old & 0xFFFFu);
            if (lane == leader) hval[slot] = (uint32_t)i | ((((old >> 16) + (uint32_t)__popc(grp)) & 0xFFFFu) << 16);
          }
          __syncwarp();
        }
        }
        SD_PH(0)
        // ---- B: scan the reference in genome order -> LongHits sorted by (g, i) ------------------
        // B1: genome positions whose k-mer occurs in the read are queued (ballot compaction, genome order);
        // B2: one lane per queued position walks its chain of read positions: count the LongHit starts,
        //     exclusive scan, ordered writes (two passes over a chain of ~1.3 entries).
        if (direct) {
          // B1 (direct): every raw hit (g, i) is queued, flattened, in (g, i) order: lanes look their id up
          //     in the table (head, count), offsets come from three ballots (chains longer than 3 take a scan);
          // B2: one lane per raw hit: chain-start test, extension, ballot-compacted ordered writes. One pass.
          uint32_t* queue = (uint32_t*)win;  // g << 16 | i ; SD_WIN * 2 entries
          const int QCAP = SD_WIN * 2;
#pragma unroll 1
          for (int gb = 0; gb <= G - K && !overflow;) {
            int nq = 0;
            bool qfull = false;
            long long tb1 = A.prof ? clock64() : 0;
            // B1: 128 genome positions per iteration, 4 consecutive positions per lane (4 independent lookups)
#pragma unroll 1
            for (; gb <= G - K && !qfull; gb += 128) {
              const int g0 = gb + 4 * lane;
              uint32_t hv[4];
              int cnt[4], tot = 0;
#pragma unroll
              for (int u = 0; u < 4; u++) {
                const int g = g0 + u;
                const uint32_t code = (g <= G - K) ? poscode[g] : SD_NONE16;
                const uint32_t id = code & 0x7FFFu;
                hv[u] = (id != SD_NONE16) ? tab[id] : SD_NOPOS;
              }
#pragma unroll
              for (int u = 0; u < 4; u++) {
                cnt[u] = (int)(hv[u] >> 16);  // SD_NOPOS carries count 0
                tot += cnt[u];
              }
              if (!__any_sync(VG_FULL, tot > 0)) continue;
              const int incl = sd_warp_incl_scan(tot, lane);
              const int total = __shfl_sync(VG_FULL, incl, 31);
              if (total > QCAP) {  // > QCAP raw hits in 128 positions (low-complexity read): hash path
                direct = false;
                nS = 0;
                goto retry_hash;
              }
              if (nq + total > QCAP) {  // does not fit any more: flush first, then redo this block
                qfull = true;
                gb -= 128;
                continue;
              }
              int o = nq + incl - tot;
#pragma unroll
              for (int u = 0; u < 4; u++) {  // chains of 1 or 2 read positions are the rule: straight-line code
                if (cnt[u]) {
                  uint32_t i = hv[u] & 0xFFFFu;
                  const uint32_t gk = (uint32_t)(g0 + u) << 16;
                  queue[o] = gk | i;
                  if (cnt[u] > 1) {
                    i = nxt[i];
                    queue[o + 1] = gk | i;
#pragma unroll 1
                    for (int t = 2; t < cnt[u]; t++) {
                      i = nxt[i];
                      queue[o + t] = gk | i;
                    }
                  }
                  o += cnt[u];
                }
              }
              nq += total;
            }
            __syncwarp();
            long long tb2 = A.prof ? clock64() : 0;
            if (A.prof && lane == 0) atomicAdd(A.prof + 8, (unsigned long long)(tb2 - tb1));
            // B2: two raw hits per lane per iteration (independent loads in flight)
#pragma unroll 1
            for (int q0 = 0; q0 < nq; q0 += 64) {
              bool em[2] = {false, false};
              int gg[2] = {0, 0}, ii[2] = {0, 0}, ll[2] = {0, 0};
              uint32_t e[2];
              uint8_t ca[2], cb[2];
#pragma unroll
              for (int u = 0; u < 2; u++) {
                const int q = q0 + 32 * u + lane;
                e[u] = (q < nq) ? queue[q] : 0xFFFFFFFFu;
              }
#pragma unroll
              for (int u = 0; u < 2; u++) {
                gg[u] = (int)(e[u] >> 16);
                ii[u] = (int)(e[u] & 0xFFFFu);
                const bool ok = e[u] != 0xFFFFFFFFu;
                ca[u] = (ok && ii[u] > 0) ? rd[ii[u] - 1] : 0;
                cb[u] = (ok && gg[u] > 0) ? ref[gg[u] - 1] : 1;
              }
#pragma unroll
              for (int u = 0; u < 2; u++) {
                if (e[u] == 0xFFFFFFFFu) continue;
                const int g = gg[u], i = ii[u];
                const bool start = (i == 0) || (g == 0) || (ca[u] != cb[u]);
                if (start) {
                  em[u] = true;
                  ll[u] = sd_extend(rd, ref, i, g, K, L, G);
                } else if (poscode[g] & 0x8000u) {
                  em[u] = true;  // Q1: a chain passing through a twice-indexed position adds a zero-length hit
                }
              }
              const unsigned m0 = __ballot_sync(VG_FULL, em[0]);
              const unsigned m1 = __ballot_sync(VG_FULL, em[1]);
              const int t0 = __popc(m0), total = t0 + __popc(m1);
              if (nS + total > A.capS) {
                overflow = true;
                break;
              }
              const unsigned lt = (1u << lane) - 1u;
              if (em[0]) S[nS + __popc(m0 & lt)] = ((uint64_t)gg[0] << 32) | ((uint64_t)ii[0] << 16) | (uint64_t)ll[0];
              if (em[1]) S[nS + t0 + __popc(m1 & lt)] = ((uint64_t)gg[1] << 32) | ((uint64_t)ii[1] << 16) | (uint64_t)ll[1];
              nS += total;
            }
            __syncwarp();
            if (A.prof && lane == 0) atomicAdd(A.prof + 9, (unsigned long long)(clock64() - tb2));
          }
        } else {
        uint32_t* queue = (uint32_t*)win;                 // g << 16 | chain head ; SD_WIN * 2 entries
        const int QCAP = SD_WIN * 2;
#pragma unroll 1
        for (int gb = 0; gb <= G - K && !overflow;) {
          int nq = 0;
#pragma unroll 1
          for (; gb <= G - K && nq + 32 <= QCAP; gb += 32) {
            const int g = gb + lane;
            const uint32_t code = (g <= G - K) ? poscode[g] : SD_NONE16;
            const uint32_t id = code & 0x7FFFu;
            uint32_t head = SD_NOPOS;
            if (id != SD_NONE16) {
              uint32_t slot = sd_hash(id);
              for (;;) {
                const uint32_t k = hkey[slot];
                if (k == SD_EMPTY) break;
                if (k == id) {
                  head = hval[slot] & 0xFFFFu;
                  break;
                }
                slot = (slot + 1) & (SD_HS - 1);
              }
            }
            const unsigned m = __ballot_sync(VG_FULL, head != SD_NOPOS);
            if (head != SD_NOPOS) queue[nq + __popc(m & ((1u << lane) - 1))] = ((uint32_t)g << 16) | head;
            nq += __popc(m);
          }
          __syncwarp();
#pragma unroll 1
          for (int q0 = 0; q0 < nq; q0 += 32) {
            const int q = q0 + lane;
            int g = 0;
            uint32_t head = SD_NOPOS;
            bool dup = false;
            if (q < nq) {
              g = (int)(queue[q] >> 16);
              head = queue[q] & 0xFFFFu;
              dup = (poscode[g] & 0x8000u) != 0;
            }
            int cnt = 0;
#pragma unroll 1
            for (uint32_t i = head; i != SD_NOPOS; i = nxt[i]) {
              const bool start = (i == 0) || (g == 0) || (rd[i - 1] != ref[g - 1]);
              cnt += (start || dup) ? 1 : 0;
            }
            const int incl = sd_warp_incl_scan(cnt, lane);
            const int total = __shfl_sync(VG_FULL, incl, 31);
            if (nS + total > A.capS) {
              overflow = true;
              break;
            }
            int o = nS + incl - cnt;
#pragma unroll 1
            for (uint32_t i = head; i != SD_NOPOS; i = nxt[i]) {
              const bool start = (i == 0) || (g == 0) || (rd[i - 1] != ref[g - 1]);
              if (start) {
                const int len = sd_extend(rd, ref, (int)i, g, K, L, G);
                S[o++] = ((uint64_t)g << 32) | ((uint64_t)i << 16) | (uint64_t)len;
              } else if (dup) {
                S[o++] = ((uint64_t)g << 32) | ((uint64_t)i << 16);
              }
            }
            nS += total;
          }
          __syncwarp();
        }
        }
      } else {
        // ---- MSA: raw hits from the CSR index; chain starts by membership in the neighbour list --
#pragma unroll 1
        for (int i = lane; i < nk; i += 32) ids[i] = (uint16_t)sd_kmer_id(rd + i, K, A.exoKeys, A.nExo);
        __syncwarp();
#pragma unroll 1
        for (int base = 0; base < nk && !overflow; base += 32) {
          const int i = base + lane;
          int lo = 0, hi = 0, plo = 0, phi = 0;
          if (i < nk && ids[i] != SD_NONE16) {
            lo = A.msaOff[ids[i]];
            hi = A.msaOff[ids[i] + 1];
            if (i > 0 && ids[i - 1] != SD_NONE16) {
              plo = A.msaOff[ids[i - 1]];
              phi = A.msaOff[ids[i - 1] + 1];
            }
          }
          int cnt = 0;
#pragma unroll 1
          for (int q = lo; q < hi; q++) {
            const int c = A.msaCols[q];
            bool cont = false;  // does hit (i-1, c-1) exist?
#pragma unroll 1
            for (int z = plo; z < phi; z++) cont |= (A.msaCols[z] == c - 1);
            cnt += cont ? 0 : 1;
          }
          const int incl = sd_warp_incl_scan(cnt, lane);
          const int total = __shfl_sync(VG_FULL, incl, 31);
          if (nS + total > A.capS || nS + total > capC) {
            overflow = true;
            break;
          }
          int o = nS + incl - cnt;
#pragma unroll 1
          for (int q = lo; q < hi; q++) {
            const int c = A.msaCols[q];
            bool cont = false;
#pragma unroll 1
            for (int z = plo; z < phi; z++) cont |= (A.msaCols[z] == c - 1);
            if (cont) continue;
            int len = 0;
            for (;;) {  // extend while hit (i+len+1, c+len+1) exists
              const int ii = i + len + 1;
              if (ii >= nk || ids[ii] == SD_NONE16) break;
              bool found = false;
#pragma unroll 1
              for (int z = A.msaOff[ids[ii]], ze = A.msaOff[ids[ii] + 1]; z < ze; z++) found |= (A.msaCols[z] == c + len + 1);
              if (!found) break;
              len++;
            }
            ckey[o] = ((uint32_t)c << 16) | (uint32_t)i;
            cval[o] = (uint32_t)len;
            o++;
          }
          nS += total;
        }
        __syncwarp();
        if (!overflow && nS > 0) {
          const int srcb = sd_radix_sort(ckey, cval, capC, nS, bins, lane, 32);
          const uint32_t* ks = ckey + (size_t)srcb * capC;
          const uint32_t* vs = cval + (size_t)srcb * capC;
#pragma unroll 1
          for (int e = lane; e < nS; e += 32)
            S[e] = ((uint64_t)(ks[e] >> 16) << 32) | ((uint64_t)(ks[e] & 0xFFFFu) << 16) | (uint64_t)vs[e];
        }
      }
      __syncwarp();
      __threadfence_block();
      if (overflow) {
        if (lane == 0) atomicOr(A.errFlag, 1);
        nS = 0;
      }
      SD_PH(1)
      // ---- C: mergeHits phase 2 (KMerCounterBase.java:422-449) ------------------------------------
      // The sweep keeps one `prev` and re-inserts clipped hits into the TreeSet. An element x whose
      // genomePos exceeds the maximum end (g + len + K) of ALL earlier elements is a synchronisation
      // point: when it is polled nothing earlier is left in the set (re-inserted keys are ends of earlier
      // elements, hence < g_x) and prev cannot overlap it, so the state becomes (prev = x) whatever the
      // history was. The stream therefore splits into independent segments; each lane sweeps whole
      // segments verbatim (in place in the shared-memory window), results are compacted in order.
      int nR = 0;
      const uint64_t* Sall = S;
      const int nSall = nS;
      int serialFrom = -1;
      int err = 0;
      if (nS > 0) {
        constexpr int CH = SD_WIN / 32;  // entries per lane when flagging
        uint64_t* pend = pendL + lane * SD_LPEND;
        int tbase = 0, runMaxE = -1;
#pragma unroll 1
        while (tbase < nS) {
          const int cnt = min(SD_WIN, nS - tbase);
#pragma unroll 8
          for (int e = lane; e < cnt; e += 32) win[SD_WI(e)] = S[tbase + e];
          __syncwarp();
          const int lo = lane * CH, hi = min(lo + CH, cnt);
          int locMax = -1;
#pragma unroll 1
          for (int j = lo; j < hi; j++) {
            const uint64_t v = win[SD_WI(j)];
            locMax = max(locMax, (int)(v >> 32) + (int)(v & 0xFFFFu) + K);
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
          {
            int m = excl;
#pragma unroll 1
            for (int j = lo; j < hi; j++) {
              const uint64_t v = win[SD_WI(j)];
              const int g = (int)(v >> 32);
              if (g > m) fmask |= 1u << (j - lo);
              m = max(m, g + (int)(v & 0xFFFFu) + K);
            }
          }
          // Lane ranges: lane l sweeps all segments that START in its chunk, i.e. the entries from its first
          // flag up to the first flag of the next flagged lane. The last segment of a window that is not the
          // last one may continue in the next window and is carried over (tailStart).
          const int nSeg = __reduce_add_sync(VG_FULL, __popc(fmask));
          const int firstFlag = fmask ? lo + __ffs(fmask) - 1 : 0x7fff;
          const int lastFlag = fmask ? lo + 31 - __clz(fmask) : -1;
          const bool more = tbase + cnt < nS;
          const int nDo = more ? nSeg - 1 : nSeg;
          if (nDo <= 0) {  // a single segment longer than the window
            serialFrom = tbase;
            break;
          }
          const int tailStart = more ? __reduce_max_sync(VG_FULL, lastFlag) : cnt;
          int nextFirst = 0x7fff;  // first flag of the lanes above
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
          {  // running maximum end over the entries that are consumed now (original ends)
            int m2 = -1;
#pragma unroll 1
            for (int j = lo; j < hi && j < tailStart; j++) {
              const uint64_t v = win[SD_WI(j)];
              m2 = max(m2, (int)(v >> 32) + (int)(v & 0xFFFFu) + K);
            }
            runMaxE = max(runMaxE, __reduce_max_sync(VG_FULL, m2));
          }
          const int s0 = min(firstFlag, tailStart), s1 = min(nextFirst, tailStart);  // s0 == s1: nothing to do
          __syncwarp();
          // All lanes step through their ranges in lock-step (one polled element per trip), so the warp stays
          // converged: the sweep of KMerCounterBase.java:422-449 verbatim, started afresh at a synchronisation
          // point. Results are written in place (w never overtakes si).
          bool ovf = false;
          int si = s0, w = s0, np = 0;
          bool have = false;
          int pg = 0, pr = 0, pl = 0;
#pragma unroll 1
          for (;;) {
            const bool act = si < s1 || np > 0;
            if (!__any_sync(VG_FULL, act)) break;
            if (act) {
              uint64_t v;
              if (np > 0 && (si >= s1 || (pend[0] >> 16) < (win[SD_WI(si)] >> 16))) {  // TreeSet.pollFirst
                v = pend[0];
#pragma unroll 1
                for (int t = 1; t < np; t++) pend[t - 1] = pend[t];
                np--;
              } else {
                v = win[SD_WI(si++)];
              }
              int cg = (int)(v >> 32), cr = (int)((v >> 16) & 0xFFFFu), cl = (int)(v & 0xFFFFu);
              if (!have) {
                have = true;
                pg = cg, pr = cr, pl = cl;
              } else if (pg + pl + K - 1 >= cg) {
                if (pl < cl) {
                  pl = cg - pg - K;
                  if (pl >= 0) win[SD_WI(w++)] = ((uint64_t)pg << 32) | ((uint64_t)pr << 16) | (uint64_t)pl;
                  pg = cg, pr = cr, pl = cl;
                } else {
                  const int lenDiff = pg + pl + K - cg;
                  cl -= lenDiff;
                  if (cl >= 0) {
                    cg += lenDiff;
                    cr += lenDiff;
                    const uint64_t key = ((uint64_t)cg << 16) | (uint64_t)cr;
                    bool present = false;
                    int ins = np;
#pragma unroll 1
                    for (int t = 0; t < np; t++) {
                      const uint64_t pk = pend[t] >> 16;
                      if (pk == key) present = true;
                      if (pk > key && ins == np) ins = t;
                    }
                    if (!present)
#pragma unroll 1
                      for (int j = si; j < s1; j++) {  // the key lies before the next synchronisation point
                        const uint64_t k = win[SD_WI(j)] >> 16;
                        if (k >= key) {
                          present = (k == key);
                          break;
                        }
                      }
                    if (!present) {
                      if (np >= SD_LPEND) {
                        ovf = true;
                        np = 0;
                        si = s1;
                      } else {
#pragma unroll 1
                        for (int t = np; t > ins; t--) pend[t] = pend[t - 1];
                        pend[ins] = (key << 16) | (uint64_t)cl;
                        np++;
                      }
                    }
                  }
                }
              } else {
                win[SD_WI(w++)] = ((uint64_t)pg << 32) | ((uint64_t)pr << 16) | (uint64_t)pl;
                pg = cg, pr = cr, pl = cl;
              }
            }
            __syncwarp();
          }
          if (have) win[SD_WI(w++)] = ((uint64_t)pg << 32) | ((uint64_t)pr << 16) | (uint64_t)pl;
#pragma unroll 1
          for (int t = w; t < s1; t++) win[SD_WI(t)] = ~0ull;  // consumed, not emitted
          __syncwarp();
          if (__any_sync(VG_FULL, ovf)) {
            serialFrom = tbase;
            break;
          }
#pragma unroll 1
          for (int e0 = 0; e0 < tailStart; e0 += 32) {  // ordered compaction into the LongHit arrays
            const int e = e0 + lane;
            const uint64_t v = (e < tailStart) ? win[SD_WI(e)] : ~0ull;
            const bool ok = v != ~0ull;
            const unsigned m = __ballot_sync(VG_FULL, ok);
            const int o = nR + __popc(m & ((1u << lane) - 1u));
            if (ok && o < capR) hitG[o] = (uint16_t)(v >> 32), hitR[o] = (uint16_t)((v >> 16) & 0xFFFFu), hitL[o] = (uint16_t)(v & 0xFFFFu);
            nR += __popc(m);
          }
          __syncwarp();
          tbase += tailStart;
        }
      }
      if (serialFrom >= 0) {
        // Fallback (a segment longer than the window, or a full re-insertion queue): the verbatim serial
        // sweep by lane 0 from the synchronisation point `serialFrom`, through a sliding window.
        const uint64_t* S = Sall + serialFrom;
        const int nS = nSall - serialFrom;
        int wbase = 0;
#pragma unroll 1
        for (int e = lane; e < SD_WIN && e < nS; e += 32) win[e] = S[e];
        __syncwarp();
        // lane-0 state of the sweep. The re-insertion queue of this path lives in the (idle) candidate scratch: low-
        // complexity reads against low-complexity stretches re-insert hundreds of clipped hits. Sorted, popped at `ph`.
        uint64_t* pq = (uint64_t*)ckey;
        const int PCAP = capC;
        int ph = 0;
        int si = 0, np = 0;
        int pg = 0, pr = 0, pl = 0;
        bool started = false, done = false;
        for (;;) {
          if (lane == 0) {
            auto emit = [&](int g, int r, int len) {
              for (;;) {
                if (nR < capR) hitG[nR] = (uint16_t)g, hitR[nR] = (uint16_t)r, hitL[nR] = (uint16_t)len;
                nR++;
                break;
              }
            };
#pragma unroll 1
            while (!done) {
              if (si >= wbase + SD_WIN / 2 && wbase + SD_WIN < nS) break;  // slide the window first
              if (si >= nS && np == 0) {
                if (started) emit(pg, pr, pl);
                done = true;
                break;
              }
              // TreeSet.pollFirst: the smaller of the stream head and the re-insertion queue head
              uint64_t v;
              if (np > 0 && (si >= nS || (pq[ph] >> 16) < (win[si - wbase] >> 16))) {
                v = pq[ph];
                ph++;
                np--;
              } else {
                v = win[si - wbase];
                si++;
              }
              int cg = (int)(v >> 32), cr = (int)((v >> 16) & 0xFFFFu), cl = (int)(v & 0xFFFFu);
              if (!started) {
                pg = cg, pr = cr, pl = cl;
                started = true;
                continue;
              }
              if (pg + pl + K - 1 >= cg) {
                if (pl < cl) {
                  pl = cg - pg - K;
                  if (pl >= 0) emit(pg, pr, pl);
                  pg = cg, pr = cr, pl = cl;
                } else {
                  const int lenDiff = pg + pl + K - cg;
                  cl -= lenDiff;
                  if (cl >= 0) {
                    cg += lenDiff;
                    cr += lenDiff;
                    // TreeSet.add: dropped when an element with the same (genomePos, readPos) is present
                    const uint64_t key = ((uint64_t)cg << 16) | (uint64_t)cr;
                    int ins = ph;
                    {
                      int hi2 = ph + np;
#pragma unroll 1
                      while (ins < hi2) {
                        const int mid = (ins + hi2) >> 1;
                        if ((pq[mid] >> 16) < key) ins = mid + 1;
                        else hi2 = mid;
                      }
                    }
                    bool present = ins < ph + np && (pq[ins] >> 16) == key;
                    if (!present) {
                      bool resolved = false;
                      const int wend = min(nS, wbase + SD_WIN);
#pragma unroll 1
                      for (int j = si; j < wend; j++) {
                        const uint64_t k = win[j - wbase] >> 16;
                        if (k >= key) {
                          present = (k == key);
                          resolved = true;
                          break;
                        }
                      }
                      if (!resolved && wend < nS) {  // beyond the window (dense repeats): search the global stream
                        int lo = wend, hi = nS;
#pragma unroll 1
                        while (lo < hi) {
                          int mid = (lo + hi) >> 1;
                          if ((S[mid] >> 16) < key) lo = mid + 1;
                          else hi = mid;
                        }
                        if (lo < nS && (S[lo] >> 16) == key) present = true;
                      }
                    }
                    if (!present) {
                      if (ph + np >= PCAP && ph > 0) {  // out of room at the top: move the queue down
#pragma unroll 1
                        for (int t = 0; t < np; t++) pq[t] = pq[ph + t];
                        ins -= ph;
                        ph = 0;
                      }
                      if (ph + np >= PCAP) {
                        err |= 4;
                      } else {
#pragma unroll 1
                        for (int t = ph + np; t > ins; t--) pq[t] = pq[t - 1];
                        pq[ins] = (key << 16) | (uint64_t)cl;
                        np++;
                      }
                    }
                  }
                }
              } else {
                emit(pg, pr, pl);
                pg = cg, pr = cr, pl = cl;
              }
            }
          }
          done = __shfl_sync(VG_FULL, (int)done, 0) != 0;
          if (done) break;
          // slide by half a window
#pragma unroll 1
          for (int e = lane; e < SD_WIN / 2; e += 32) win[e] = win[e + SD_WIN / 2];
          __syncwarp();
          wbase += SD_WIN / 2;
#pragma unroll 1
          for (int e = lane; e < SD_WIN / 2; e += 32) {
            const int idx = wbase + SD_WIN / 2 + e;
            if (idx < nS) win[SD_WIN / 2 + e] = S[idx];
          }
          __syncwarp();
        }
        nR = __shfl_sync(VG_FULL, nR, 0);
        err = __shfl_sync(VG_FULL, err, 0);
      }
      SD_PH(2)
      if (nS > 0) {
        __syncwarp();
        // ---- D: KMerCounter.splitLongHits (KMerCounter.java:142-185); hits are independent given their contig
        if (!err && nR <= capR && A.nContigs > 1 && nR > 0 && !A.bestOnly) {
          uint32_t* tg = ckey;   // pieces: g | (r << 16) in ckey, len in cval
          uint32_t* tl = cval;
          int nOut = 0;
          // the pieces of one hit, walking from contig to contig (any number of boundaries); counted first, then written
          auto pieces = [&](int j, bool write, int o) -> int {
            int g = hitG[j], r = hitR[j], len = hitL[j], n = 0;
            for (;;) {
              int lo2 = 0, hi2 = A.nContigs - 1;
#pragma unroll 1
              while (lo2 < hi2) {
                const int mid = (lo2 + hi2) >> 1;
                if (A.contigEnds[mid] > g) hi2 = mid;
                else lo2 = mid + 1;
              }
              const int contigEnd = A.contigEnds[lo2];
              if (g + K + len > contigEnd) {
                if (contigEnd - g >= K) {
                  if (write && o + n < capC) tg[o + n] = (uint32_t)g | ((uint32_t)r << 16), tl[o + n] = (uint32_t)(contigEnd - g - K);
                  n++;
                }
                if (g + K + len - contigEnd >= K) {
                  r += contigEnd - g;
                  len -= contigEnd - g;
                  g = contigEnd;
                  continue;
                }
                break;
              }
              if (write && o + n < capC) tg[o + n] = (uint32_t)g | ((uint32_t)r << 16), tl[o + n] = (uint32_t)len;
              n++;
              break;
            }
            return n;
          };
#pragma unroll 1
          for (int j0 = 0; j0 < nR; j0 += 32) {
            const int j = j0 + lane;
            const int np2 = (j < nR) ? pieces(j, false, 0) : 0;
            const int ic = sd_warp_incl_scan(np2, lane);
            if (j < nR) pieces(j, true, nOut + ic - np2);
            nOut += __shfl_sync(VG_FULL, ic, 31);
          }
          __syncwarp();
          if (nOut > capR) {
            err |= 2;
          } else {
#pragma unroll 1
            for (int j = lane; j < nOut; j += 32) hitG[j] = (uint16_t)(tg[j] & 0xFFFFu), hitR[j] = (uint16_t)(tg[j] >> 16), hitL[j] = (uint16_t)tl[j];
            nR = nOut;
          }
          __syncwarp();
        }
        if (nR > capR) err |= 2;
        if (err) {
          if (lane == 0) atomicOr(A.errFlag, err);
          nR = 0;
        }
      }
      __syncwarp();
      if (A.dbgHits && wi == 0) {
#pragma unroll 1
        for (int j = lane; j < nR; j += 32) {
          A.dbgHits[3 * j] = hitG[j];
          A.dbgHits[3 * j + 1] = hitR[j];
          A.dbgHits[3 * j + 2] = hitL[j];
        }
        if (lane == 0) *A.dbgNHits = nR;
      }
      SD_PH(3)
      if (nR > 0) {
        // ---- E: isAdjacent + prefix sums P[j] = sum_{t<=j}(len_t + isAdjacent[t-1]) (mod 2^16) ------
        // genomePos of the LongHits is searched at random by E/F: keep a shared-memory copy when it fits
        // (it overlays the phase-C buffers pendL/segBuf, which are dead now)
        const uint16_t* hitGg = hitG;
        {
          uint16_t* cp = (uint16_t*)pendL;
          const int room = (int)((wbase + A.geom.regA - (uint8_t*)cp) / 2);
          if (nR <= room) {
#pragma unroll 4
            for (int j = lane; j < nR; j += 32) cp[j] = hitGg[j];
            __syncwarp();
            hitG = cp;
          }
        }
        int carry = 0;
#pragma unroll 2
        for (int j0 = 0; j0 < nR; j0 += 32) {
          const int j = j0 + lane;
          int val = 0;
          if (j < nR) {
            val = hitL[j];
            if (j > 0) {
              const float rate = __fdiv_rn((float)((int)hitG[j] - (int)hitG[j - 1]), (float)((int)hitR[j] - (int)hitR[j - 1]));
              if (A.lowCoef <= rate && rate <= A.highCoef) val++;
            }
          }
          const int incl = sd_warp_incl_scan(val, lane);
          if (j < nR) Pp[j] = (uint16_t)(carry + incl);
          carry += __shfl_sync(VG_FULL, incl, 31);
        }
        // bucket[b] = first hit index with g >= 32 b
        const int nb = A.geom.nBuckets;
#pragma unroll 1
        for (int b = lane; b < nb; b += 32) bucket[b] = (uint16_t)nR;
        __syncwarp();
#pragma unroll 1
        for (int j = lane; j < nR; j += 32) {
          const int b = hitG[j] >> 5;
          if (j == 0 || (hitG[j - 1] >> 5) != b) bucket[b] = (uint16_t)j;
        }
        __syncwarp();
        {  // suffix fill of empty buckets (serial over chunks from the top, parallel inside a chunk)
          int nextv = nR;
#pragma unroll 1
          for (int b0 = ((nb - 1) >> 5) << 5; b0 >= 0; b0 -= 32) {
            const int b = b0 + lane;
            int v = (b < nb) ? bucket[b] : nR;
            const bool empty = (v == nR);
            // suffix-min within the warp over lanes >= me
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
              int o = __shfl_down_sync(VG_FULL, v, d);
              if (lane + d < 32) v = min(v, o);
            }
            v = min(v, nextv);
            if (b < nb && empty) bucket[b] = (uint16_t)v;
            nextv = __shfl_sync(VG_FULL, v, 0);
          }
        }
        __syncwarp();
        SD_PH(4)
        // ---- F: every window independently; candidates (count > cutoff) compacted in hit order ------
        const int cutoff = A.bestOnly ? -1 : (L < A.nCutoffs) ? A.cutoffs[L] : 0x7fffffff;
        int nC = 0;
        int bestCnt = -1, bestJ = 0x7fffffff, bestS = 0, bestE = 0;  // bestOnly: per lane, first maximum in hit order
#pragma unroll 2
        for (int j0 = 0; j0 < nR; j0 += 32) {
          const int j = j0 + lane;
          bool isCand = false;
          uint32_t key = 0, cnt = 0;
          if (j < nR) {
            const int g = hitG[j], r = hitR[j], len = hitL[j];
            int cs = 0, ce = G;
            if (A.nContigs > 1 && !A.bestOnly) {
              int lo = 0, hi = A.nContigs - 1;
#pragma unroll 1
              while (lo < hi) {
                int mid = (lo + hi) >> 1;
                if (A.contigEnds[mid] > g) hi = mid;
                else lo = mid + 1;
              }
              ce = A.contigEnds[lo];
              cs = lo ? A.contigEnds[lo - 1] : 0;
            }
            const int start = max(cs, g - r * 3 / 2);
            const int end = min(g + len + K + (L - r - len - K) * 3 / 2, ce);
            int a = bucket[start >> 5];
#pragma unroll 1
            while (hitG[a] < start) a++;
            int u = bucket[min(end, G) >> 5];
#pragma unroll 1
            while (u < nR && hitG[u] < end) u++;
            int b = u - 1;
            if ((int)hitG[b] + K + (int)hitL[b] > end) b--;
            cnt = (uint32_t)((uint16_t)(Pp[b] - Pp[a])) + hitL[a];
            isCand = (int)cnt > cutoff;
            key = ((uint32_t)start << 16) | (uint32_t)(0xFFFF - end);
            if (A.bestOnly) {
              if ((int)cnt > bestCnt) bestCnt = (int)cnt, bestJ = j, bestS = start, bestE = end;
              isCand = false;
            }
          }
          const unsigned m = __ballot_sync(VG_FULL, isCand);
          if (isCand) {
            const int o = nC + __popc(m & ((1u << lane) - 1));
            ckey[o] = key;
            cval[o] = cnt;
          }
          nC += __popc(m);
        }
        __syncwarp();
        if (A.bestOnly) {  // strict '>' in hit order: the first hit that reaches the maximum wins
          const int mx = __reduce_max_sync(VG_FULL, bestCnt);
          const int jw = __reduce_min_sync(VG_FULL, bestCnt == mx ? bestJ : 0x7fffffff);
          if (bestCnt == mx && bestJ == jw) {
            int32_t* W = A.windows + (size_t)wi * A.maxWin * 3;
            W[0] = bestS, W[1] = bestE, W[2] = mx;
          }
          nWin = 1;
        }
        SD_PH(5)
        // ---- G: stable sort by (start asc, end desc) + pruneWindows (serial, lane 0) ----------------
        if (nC > 0) {
          const int srcb = sd_radix_sort(ckey, cval, capC, nC, bins, lane, 32);
          SD_PH(6)
          __threadfence_block();
          {
            // pruneWindows (KMerCounterBase.java:317-335). The pivot only changes at the first later candidate
            // that either does not overlap it (emit, new pivot) or overlaps with a larger count (new pivot);
            // everything in between is dropped. The warp looks at 32 candidates at a time.
            const uint32_t* ks = ckey + (size_t)srcb * capC;
            const uint32_t* vs = cval + (size_t)srcb * capC;
            int32_t* W = A.windows + (size_t)wi * A.maxWin * 3;
            int ps = ks[0] >> 16, pe = 0xFFFF - (ks[0] & 0xFFFF), pc = vs[0];
#pragma unroll 1
            for (int t0 = 1; t0 < nC; t0 += 32) {
              const int idx = t0 + lane;
              const bool valid = idx < nC;
              const uint32_t k = valid ? ks[idx] : 0u;
              const int cs_ = k >> 16, ce_ = 0xFFFF - (k & 0xFFFF), cc = valid ? (int)vs[idx] : 0;
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
              if (lane == 0) atomicOr(A.errFlag, 8);
              nWin = 0;
            }
          }
        }
      }
    } else if (L > SD_MAXL) {
      if (lane == 0) atomicOr(A.errFlag, 16);
    }
    if (lane == 0) A.nWindows[wi] = nWin;
    __syncwarp();
    if (A.prof && nk >= 1 && L <= SD_MAXL) {
      long long t2 = clock64();
      if (lane == 0) atomicAdd(A.prof + 7, (unsigned long long)(t2 - tph0));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Pair selection: Mapper.ReadsMapping.mapRead (Mapper.java:92-390) and MapperMSA.ReadMapping.mapRead
// (MapperMSA.java:72-244) after the four mapReadFast calls. One thread per pair.
// ------------------------------------------------------------------------------------------------
struct PairArgs {
  const int32_t* windows;   // [4*nPairs][maxWin][3]
  const int32_t* nWindows;
  int maxWin;
  int64_t nPairs;
  const int64_t* pairIdx;
  const int64_t* off1;
  const int32_t* len1;
  const int64_t* off2;
  const int32_t* len2;
  const int32_t* fragmentEnds;
  int nFragments;
  int isFragmented;
  int insertLen;
  int msa;
  vg_pair_result* results;
  SwTask* tasks;
  int* nTasks;              // [0] task count, [1] max window length, [2..3] sum of |read| x |window| (u64)
  int64_t outBase;          // output slot of (pair 0, mate 0) of this chunk
  int32_t* fragScratch;     // [nPairs][6][nFragments] when fragmented
};

__device__ __forceinline__ int pr_fragment(int start, int end, const int32_t* fe, int nf) {
  // MappedRead.chooseFragment (MappedRead.java:39-50); Q7: stays 0 when the window fits no fragment
  int s = 0;
  for (int i = 0; i < nf; i++) {
    const int e = fe[i];
    if (start >= s && end <= e) return i;
    s = e;
  }
  return 0;
}

__global__ void pair_select_kernel(PairArgs A) {
  const int64_t pi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pi >= A.nPairs) return;
  const int64_t p = A.pairIdx ? A.pairIdx[pi] : pi;
  const int32_t* W[4];
  int n[4];
  for (int o = 0; o < 4; o++) {
    W[o] = A.windows + (size_t)(pi * 4 + o) * A.maxWin * 3;
    n[o] = A.nWindows[pi * 4 + o];
  }
  // orientation index: 0 forward1, 1 reverse1, 2 forward2, 3 reverse2
  const bool frag = A.isFragmented && !A.msa;
  const int nf = A.nFragments;
  int32_t* fs = frag ? A.fragScratch + (size_t)pi * 6 * nf : nullptr;
  int32_t *countFrag1 = fs, *sideFrag1 = fs + nf, *maxIndFrag1 = fs + 2 * nf, *countFrag2 = fs + 3 * nf,
          *sideFrag2 = fs + 4 * nf, *maxIndFrag2 = fs + 5 * nf;
  if (frag)
    for (int i = 0; i < 6 * nf; i++) fs[i] = 0;
  int count1 = 0, side1 = -1, maxInd1 = -1, count2 = 0, side2 = -1, maxInd2 = -1;
  auto scan = [&](int o, int sd, int& cnt, int& side, int& maxInd, int32_t* cf, int32_t* sf, int32_t* mf) {
    for (int j = 0; j < n[o]; j++) {
      const int c = W[o][3 * j + 2];
      if (frag) {
        const int f = pr_fragment(W[o][3 * j], W[o][3 * j + 1], A.fragmentEnds, nf);
        if (c > cf[f]) cf[f] = c, sf[f] = sd, mf[f] = j;
      }
      if (c > cnt) cnt = c, side = sd, maxInd = j;
    }
  };
  scan(0, 0, count1, side1, maxInd1, countFrag1, sideFrag1, maxIndFrag1);  // Mapper.java:124-157
  scan(1, 1, count1, side1, maxInd1, countFrag1, sideFrag1, maxIndFrag1);
  scan(3, 1, count2, side2, maxInd2, countFrag2, sideFrag2, maxIndFrag2);  // mate 2: reverse first (:158-191)
  scan(2, 0, count2, side2, maxInd2, countFrag2, sideFrag2, maxIndFrag2);
  int maxCount = 0, max1 = -1, max2 = -1, side = -1;
  for (int j = 0; j < n[0]; j++) {  // forward1 x reverse2 (:193-208)
    const int s1 = W[0][3 * j], e1 = W[0][3 * j + 1], c1 = W[0][3 * j + 2];
    const int f1 = frag ? pr_fragment(s1, e1, A.fragmentEnds, nf) : 0;
    for (int k = 0; k < n[3]; k++) {
      const int s2 = W[3][3 * k], e2 = W[3][3 * k + 1], c2 = W[3][3 * k + 2];
      if (s1 < e2 && e2 - s1 <= A.insertLen) {
        const int f2 = frag ? pr_fragment(s2, e2, A.fragmentEnds, nf) : 0;
        if (f1 == f2 && c1 + c2 > maxCount) maxCount = c1 + c2, max1 = j, max2 = k, side = 0;
      }
    }
  }
  for (int j = 0; j < n[1]; j++) {  // reverse1 x forward2 (:209-224)
    const int s1 = W[1][3 * j], e1 = W[1][3 * j + 1], c1 = W[1][3 * j + 2];
    const int f1 = frag ? pr_fragment(s1, e1, A.fragmentEnds, nf) : 0;
    for (int k = 0; k < n[2]; k++) {
      const int s2 = W[2][3 * k], e2 = W[2][3 * k + 1], c2 = W[2][3 * k + 2];
      if (s2 < e1 && e1 - s2 <= A.insertLen) {
        const int f2 = frag ? pr_fragment(s2, e2, A.fragmentEnds, nf) : 0;
        if (f1 == f2 && c1 + c2 > maxCount) maxCount = c1 + c2, max1 = j, max2 = k, side = 1;
      }
    }
  }
  int cls = VG_CLS_UNMAPPED, o1 = -1, i1 = -1, o2 = -1, i2 = -1;
  if (side == 0) {
    cls = VG_CLS_CONCORDANT_FR, o1 = 0, i1 = max1, o2 = 3, i2 = max2;
  } else if (side == 1) {
    cls = VG_CLS_CONCORDANT_RF, o1 = 1, i1 = max1, o2 = 2, i2 = max2;
  } else if (side1 == -1) {
    if (side2 != -1) cls = VG_CLS_RIGHT, o2 = side2 ? 3 : 2, i2 = maxInd2;
  } else if (side2 == -1) {
    cls = VG_CLS_LEFT, o1 = side1 ? 1 : 0, i1 = maxInd1;
  } else if (frag) {  // Mapper.java:309-353
    int maxSum = 0, maxIndex = -1;
    for (int i = 0; i < nf; i++)
      if (countFrag1[i] > 0 && countFrag2[i] > 0 && countFrag1[i] + countFrag2[i] > maxSum)
        maxSum = countFrag1[i] + countFrag2[i], maxIndex = i;
    if (maxIndex != -1) {
      cls = VG_CLS_DISCORDANT;
      o1 = sideFrag1[maxIndex] ? 1 : 0, i1 = maxIndFrag1[maxIndex];
      o2 = sideFrag2[maxIndex] ? 3 : 2, i2 = maxIndFrag2[maxIndex];
    }
  } else {
    cls = VG_CLS_DISCORDANT, o1 = side1 ? 1 : 0, i1 = maxInd1, o2 = side2 ? 3 : 2, i2 = maxInd2;
  }
  vg_pair_result res;
  memset(&res, 0, sizeof res);
  res.cls = cls;
  res.m[0].seq1_offset = res.m[1].seq1_offset = -1;
  int t = 0;
  if (!A.msa) t = atomicAdd(A.nTasks, (o1 >= 0) + (o2 >= 0));  // mates of a pair stay adjacent
  for (int mate = 0; mate < 2; mate++) {
    const int o = mate ? o2 : o1, ix = mate ? i2 : i1;
    if (o < 0) continue;
    vg_mate& m = res.m[mate];
    m.present = 1;
    m.start = W[o][3 * ix];
    m.end = W[o][3 * ix + 1];
    m.count = W[o][3 * ix + 2];
    m.reverse = o & 1;
    m.fragment_id = frag ? pr_fragment(m.start, m.end, A.fragmentEnds, nf) : 0;
    if (!A.msa) {
      SwTask tk;
      tk.s1_off = mate ? A.off2[p] : A.off1[p];
      tk.m = mate ? A.len2[p] : A.len1[p];
      tk.rc = m.reverse;
      tk.s2_off = m.start;
      tk.n = m.end - m.start;
      tk.out = (int32_t)(A.outBase + pi * 2 + mate);
      A.tasks[t++] = tk;
      atomicMax(A.nTasks + 1, tk.n);
      atomicAdd((unsigned long long*)(A.nTasks + 2), (unsigned long long)tk.m * (unsigned long long)tk.n);
    }
  }
  A.results[pi] = res;
}

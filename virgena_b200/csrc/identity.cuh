// Sliding-window identity of the mapped reads.
//
// Replaces ProblemIntervalBuilder.IdentityCounting.run (ProblemIntervalBuilder.java:35-133) and the sum of the per-thread
// arrays in countIdentity (:135-160). The reference walks every mapped read serially: the identity of the first window
// of `windowSize` read bases, then one step per window start, and adds identity/len (binary32) to sumIdentity[window
// start]. Those sums are floating point, so their order matters: reference thread t adds its reads
// [t*size, (t+1)*size) in list order into its own array, and the arrays are added in thread order.
//
// Here: one thread per mapped read does the walk and stores its values (idc_walk_kernel); one CTA per (256-position tile,
// reference thread t) then adds, per position, the values of the reads of chunk t in list order: the reads are scanned
// in batches of 256, the ones reaching the tile are compacted in order into shared memory, and every position-thread
// adds the ones that cover it (idc_sum_kernel); a last kernel adds the T partial arrays in order (idc_final_kernel).
#pragma once
#include "ctx.cuh"

#define IDC_TILE 256

struct IdcRead {   // one mapped read: its window starts [start, start + npos) and where its values live
  int32_t start;
  int32_t npos;
  int64_t off;
};

struct IdcArgs {
  const vg_pair_result* res;
  int64_t nPairs;
  const uint8_t* seq1pool;
  const int64_t* pairIdx;
  const int32_t* len1;
  const int32_t* len2;
  const uint8_t* genome;
  int G, windowSize;
  int32_t* isMapped;  // [2 nPairs] 1 where the slot (pair * 2 + mate) is an element of mappedReads
  int32_t* npos;      // [2 nPairs] window starts the read contributes to
  const int64_t* listIdx;  // exclusive scans of the two
  const int64_t* valOff;
  IdcRead* reads;
  float* vals;
  int* err;
};

__device__ __forceinline__ int idc_seq_len(const IdcArgs& A, int64_t q) {
  const int64_t p = A.pairIdx[q >> 1];
  return (q & 1) ? A.len2[p] : A.len1[p];
}

// pass 1: membership in mappedReads and the number of values per read
__global__ void idc_plan_kernel(IdcArgs A) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= A.nPairs * 2) return;
  const vg_mate& m = A.res[q >> 1].m[q & 1];
  int mapped = 0, np = 0;
  if (m.present && m.seq1_offset >= 0) {
    mapped = 1;
    const int seqLen = idc_seq_len(A, q);
    if (seqLen >= A.windowSize) {
      const int startNoClip = m.start + m.start2;
      const int start = max(0, startNoClip - m.start1);
      const int end = min(A.G, m.start + m.end2 + seqLen - m.end1);
      np = 1 + max(0, end - A.windowSize - start);
    }
  }
  A.isMapped[q] = mapped;
  A.npos[q] = np;
}

// pass 2: the walk of IdentityCounting.run, one thread per read; values to vals[off .. off + npos)
__global__ void idc_walk_kernel(IdcArgs A) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= A.nPairs * 2) return;
  IdcRead r;
  r.start = 0, r.npos = A.npos[q], r.off = A.valOff[q];
  if (r.npos > 0) {
    const vg_mate& m = A.res[q >> 1].m[q & 1];
    const int G = A.G, W = A.windowSize;
    const uint8_t* __restrict__ seq1 = A.seq1pool + m.seq1_offset;
    const uint8_t* __restrict__ gen = A.genome;
    const int alnLen = m.length;
    const int seqLen = idc_seq_len(A, q);
    const int startNoClip = m.start + m.start2;
    const int start = max(0, startNoClip - m.start1);
    const int endNoClip = m.start + m.end2;
    const int end = min(G, endNoClip + seqLen - m.end1);
    bool oob = start >= G;
    auto S = [&](int p) -> uint8_t {  // seq1[p]; out of range = the Java throws
      if ((unsigned)p >= (unsigned)alnLen) {
        oob = true;
        return 0;
      }
      return seq1[p];
    };
    auto Gn = [&](int p) -> uint8_t {
      if ((unsigned)p >= (unsigned)G) {
        oob = true;
        return 0;
      }
      return gen[p];
    };
    int identity = 0, alnEndPos = 0, alnStartPos = 0, len;
    if (m.start1 >= W) {
      len = W;
    } else {
      len = m.start1;
      if (alnLen > W - m.start1) {
#pragma unroll 1
        for (int i = startNoClip; i < startNoClip + W - m.start1 && alnEndPos < alnLen && !oob;) {
          const uint8_t c = seq1[alnEndPos];
          if (c != '-' && c == Gn(i)) {
            identity++, alnEndPos++, len++, i++;
          } else if (c < 'a') {
            alnEndPos++, len++, i++;
          } else {
            do {
              alnEndPos++, len++;
            } while (S(alnEndPos) >= 'a');
          }
        }
        while (alnEndPos < alnLen && seq1[alnEndPos] >= 'a') alnEndPos++, len++;
      } else {
        identity += m.identity;
        len += W - m.start1;
        alnEndPos = alnLen;
      }
    }
    float* __restrict__ out = A.vals + r.off;
    if (!oob) out[0] = __fdiv_rn((float)identity, (float)len);
#pragma unroll 1
    for (int i = start + 1; i <= end - W && !oob; i++) {
      if (i > startNoClip && i <= endNoClip) {
        const uint8_t c = S(alnStartPos);
        if (c != '-' && c == Gn(i - 1)) identity--;
        alnStartPos++;
        if (alnStartPos < alnLen)
          while (S(alnStartPos) >= 'a') alnStartPos++, len--;
      }
      if (i > startNoClip - W && i <= endNoClip - W) {
        const uint8_t c = S(alnEndPos);
        if (c != '-' && c == Gn(i + W - 1)) identity++;
        alnEndPos++;
        if (alnEndPos < alnLen)
          while (S(alnEndPos) >= 'a') alnEndPos++, len++;
      }
      out[i - start] = __fdiv_rn((float)identity, (float)len);
    }
    if (oob) atomicOr(A.err, 1);
    r.start = start;
  }
  A.reads[q] = r;
}

// first slot whose list index is >= want (listIdx is the exclusive scan of isMapped over n slots)
__device__ __forceinline__ int64_t idc_first_slot(const int64_t* __restrict__ listIdx, int64_t n, int64_t want) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (listIdx[mid] >= want) hi = mid;
    else lo = mid + 1;
  }
  return lo;
}

// grid (nTiles, T): partial[t][x] = sum, in list order, of the values of chunk t's reads at window start x
__global__ void __launch_bounds__(IDC_TILE) idc_sum_kernel(const IdcRead* __restrict__ reads, const int64_t* __restrict__ listIdx,
                                                         int64_t nSlots, int64_t nMapped, int T, const float* __restrict__ vals,
                                                         int G, float* __restrict__ partial, int32_t* __restrict__ partialCov) {
  const int tile = blockIdx.x, t = blockIdx.y;
  const int x = tile * IDC_TILE + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t size = nMapped / T;
  __shared__ int64_t range[2];
  __shared__ IdcRead list[IDC_TILE];
  __shared__ int warpCount[IDC_TILE / 32];
  if (threadIdx.x == 0) {
    range[0] = idc_first_slot(listIdx, nSlots, (int64_t)t * size);
    range[1] = (t == T - 1) ? nSlots : idc_first_slot(listIdx, nSlots, (int64_t)(t + 1) * size);
  }
  __syncthreads();
  const int64_t lo = range[0], hi = range[1];
  const int tileLo = tile * IDC_TILE, tileHi = tileLo + IDC_TILE;
  float sum = 0.0f;
  int cov = 0;
  for (int64_t q0 = lo; q0 < hi; q0 += IDC_TILE) {
    const int64_t q = q0 + threadIdx.x;
    IdcRead r;
    r.start = 0, r.npos = 0, r.off = 0;
    if (q < hi) r = reads[q];
    const bool hit = r.npos > 0 && r.start < tileHi && r.start + r.npos > tileLo;
    const unsigned bal = __ballot_sync(VG_FULL, hit);
    if (lane == 0) warpCount[warp] = __popc(bal);
    __syncthreads();
    int base = 0, total = 0;
#pragma unroll
    for (int w = 0; w < IDC_TILE / 32; w++) {
      base += (w < warp) ? warpCount[w] : 0;
      total += warpCount[w];
    }
    if (hit) list[base + __popc(bal & ((1u << lane) - 1))] = r;
    __syncthreads();
    for (int k = 0; k < total; k++) {
      const IdcRead e = list[k];
      const unsigned d = (unsigned)(x - e.start);
      if (d < (unsigned)e.npos) {
        sum = __fadd_rn(sum, vals[e.off + d]);
        cov++;
      }
    }
    __syncthreads();
  }
  if (x < G) {
    partial[(size_t)t * G + x] = sum;
    partialCov[(size_t)t * G + x] = cov;
  }
}

// identity[i] = ((0 + p_0[i]) + p_1[i]) + ...   (countIdentity, ProblemIntervalBuilder.java:154-159)
__global__ void idc_final_kernel(const float* __restrict__ partial, const int32_t* __restrict__ partialCov, int G, int T,
                                 float* __restrict__ identity, int32_t* __restrict__ coverage) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= G) return;
  float s = 0.0f;
  int c = 0;
  for (int t = 0; t < T; t++) {
    s = __fadd_rn(s, partial[(size_t)t * G + x]);
    c += partialCov[(size_t)t * G + x];
  }
  identity[x] = s;
  coverage[x] = c;
}

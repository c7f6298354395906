// Consensus pileup for sm_100a.
//
// Replaces ConsensusBuilderSimple.ReadCounting.run (ConsensusBuilderSimple.java:56-75), the serial sum
// of the per-thread matrices (:138-145) and ConsensusBuildingSimple.run (:98-120).
//
// The reference walks every aligned read, and for each column c of sequence1 with c < 'a' does
//   counts[j][nToI[c]]++; j++   starting at j = read.start + aln.start2          (lower case = insertion, skipped)
// Here every alignment additionally gets a "reference row": one code byte (nToI[c], Q11: every unlisted
// byte counts as A) per reference position it covers, written next to sequence1 by the emit kernel.
// The count itself is a segmented, atomic-free gather: reads are binned by the 256-position tile of their
// first reference position; one CTA = one tile x one slice of the reads that can reach the tile; one
// thread = one reference position accumulating its 5 counters in registers; the slices are then summed
// in a fixed order. Integer sums, so the result is exact and independent of the binning order.
#pragma once
#include "ctx.cuh"

#define PU_TILE 256

struct PuRead {     // one mapped read for the pileup
  int32_t s;        // first reference position = MappedRead.start + Alignment.start2
  int32_t len;      // reference positions covered (row length)
  int64_t off;      // offset of its row in the row pool
};

// nToI (Constants.java:51-60): A,T,G,C,'-' -> 0..4, every other byte <= 'T' -> 0
__device__ __forceinline__ uint8_t pu_ntoi(uint8_t c) {
  return c == 'T' ? 1 : c == 'G' ? 2 : c == 'C' ? 3 : c == '-' ? 4 : 0;
}

// sequence1 = reverse(reversed1) into the pool, plus the reference row. One warp per alignment.
__global__ void pu_emit_kernel(const SwAln* __restrict__ alns, int64_t n, const uint8_t* __restrict__ rev,
                               int64_t revStride, const int64_t* __restrict__ seqOff, uint8_t* __restrict__ pool,
                               uint8_t* __restrict__ rowpool, int32_t* __restrict__ rowLen) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= n) return;
  const int len = alns[w].length;
  const int64_t off = seqOff[w];
  if (off < 0) {
    if (lane == 0 && rowLen) rowLen[w] = 0;
    return;
  }
  const uint8_t* src = rev + w * revStride;
  int j = 0;
  for (int t0 = 0; t0 < len; t0 += 32) {
    const int t = t0 + lane;
    uint8_t ch = 0xFF;
    if (t < len) {
      ch = src[len - 1 - t];
      pool[off + t] = ch;
    }
    const bool col = (t < len) && ch < 'a';
    const unsigned m = __ballot_sync(VG_FULL, col);
    if (rowpool && col) rowpool[off + j + __popc(m & ((1u << lane) - 1))] = pu_ntoi(ch);
    j += __popc(m);
  }
  if (lane == 0 && rowLen) rowLen[w] = j;
}

// Reference rows of alignments whose sequence1 is already in the pool (vg_set_results: a MappedData edited by the host,
// e.g. after markRemapReads / adjustCoord). One warp per (pair, mate). err: a byte in ('T', 'a') (the Java pileup throws).
__global__ void pu_rows_kernel(const vg_pair_result* __restrict__ res, int64_t n2, const uint8_t* __restrict__ pool,
                               uint8_t* __restrict__ rowpool, int32_t* __restrict__ rowLen, int* __restrict__ err) {
  const int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (q >= n2) return;
  const vg_mate& m = res[q >> 1].m[q & 1];
  if (!m.present || m.seq1_offset < 0) {
    if (lane == 0) rowLen[q] = 0;
    return;
  }
  const int len = m.length;
  const int64_t off = m.seq1_offset;
  int j = 0;
  for (int t0 = 0; t0 < len; t0 += 32) {
    const int t = t0 + lane;
    const uint8_t ch = (t < len) ? pool[off + t] : 0xFF;
    const bool col = (t < len) && ch < 'a';
    if (col && ch > 'T') atomicOr(err, 1);
    const unsigned mk = __ballot_sync(VG_FULL, col);
    if (col) rowpool[off + j + __popc(mk & ((1u << lane) - 1))] = pu_ntoi(ch);
    j += __popc(mk);
  }
  if (lane == 0) rowLen[q] = j;
}

// PuRead list of the mapped reads of the last mapping + per-tile histogram (block-aggregated).
__global__ void pu_list_kernel(const vg_pair_result* __restrict__ res, int64_t nPairs, const int32_t* __restrict__ rowLen,
                               int G, int nTiles, PuRead* __restrict__ reads, int32_t* __restrict__ tileOf,
                               int32_t* __restrict__ tileCount, int* __restrict__ err) {
  extern __shared__ int pu_hist[];
  for (int t = threadIdx.x; t < nTiles; t += blockDim.x) pu_hist[t] = 0;
  __syncthreads();
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // q = pair*2 + mate
  if (q < nPairs * 2) {
    const vg_mate& m = res[q >> 1].m[q & 1];
    int tile = -1;
    PuRead r;
    r.s = 0, r.len = 0, r.off = 0;
    if (m.present && m.seq1_offset >= 0) {
      r.s = m.start + m.start2;
      r.len = rowLen[q];
      r.off = m.seq1_offset;
      if (r.len > 0) {
        if (r.s < 0 || r.s + r.len > G) {
          atomicOr(err, 1);  // the Java would throw ArrayIndexOutOfBounds
          r.len = 0;
        } else {
          tile = r.s / PU_TILE;
          atomicAdd(&pu_hist[tile], 1);
          if (r.len > err[1]) atomicMax(err + 1, r.len);  // longest row: how many tiles back a read can start (pu_count_kernel)
        }
      }
    }
    reads[q] = r;
    tileOf[q] = tile;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < nTiles; t += blockDim.x)
    if (pu_hist[t]) atomicAdd(&tileCount[t], pu_hist[t]);
}

// exclusive scan of the (few) tile counts -> tileStart[nTiles+1]; also resets the fill cursors
__global__ void pu_scan_kernel(const int32_t* __restrict__ tileCount, int nTiles, int64_t* __restrict__ tileStart,
                               int64_t* __restrict__ cursor) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int64_t run = 0;
    for (int t = 0; t < nTiles; t++) {
      tileStart[t] = run;
      cursor[t] = run;
      run += tileCount[t];
    }
    tileStart[nTiles] = run;
  }
}

// scatter the reads into tile order (order inside a tile is irrelevant: integer sums)
__global__ void pu_scatter_kernel(const PuRead* __restrict__ reads, const int32_t* __restrict__ tileOf, int64_t n,
                                  int nTiles, unsigned long long* __restrict__ cursor, PuRead* __restrict__ sorted) {
  extern __shared__ __align__(16) int pu_sm[];
  int* hist = pu_sm;                // per-block count per tile
  unsigned long long* base64 = (unsigned long long*)(pu_sm + 2 * nTiles + 2);  // 8-byte aligned whatever nTiles is
  for (int t = threadIdx.x; t < nTiles; t += blockDim.x) hist[t] = 0;
  __syncthreads();
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int tile = -1, rank = 0;
  if (q < n) {
    tile = tileOf[q];
    if (tile >= 0) rank = atomicAdd(&hist[tile], 1);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < nTiles; t += blockDim.x)
    if (hist[t]) base64[t] = atomicAdd(&cursor[t], (unsigned long long)hist[t]);
  __syncthreads();
  if (tile >= 0) sorted[base64[tile] + rank] = reads[q];
}

// The count kernel: grid (nTiles, nSlices), PU_TILE threads. Atomic-free.
__global__ void __launch_bounds__(PU_TILE) pu_count_kernel(const PuRead* __restrict__ sorted,
                                                          const int64_t* __restrict__ tileStart, int nTiles,
                                                          const int* __restrict__ maxRowLen, const uint8_t* __restrict__ rowpool, int G,
                                                          int nSlices, int32_t* __restrict__ partial) {
  const int T = blockIdx.x, sl = blockIdx.y;
  const int x = T * PU_TILE + threadIdx.x;
  // a row of at most *maxRowLen positions that covers this tile starts at most ceil(maxRowLen / tile) tiles before it
  const int t0 = max(0, T - (*maxRowLen + PU_TILE - 1) / PU_TILE);
  const int64_t lo = tileStart[t0], hi = tileStart[T + 1];
  const int64_t per = (hi - lo + nSlices - 1) / nSlices;
  const int64_t a = lo + per * sl, b = (a + per < hi) ? (a + per) : hi;
  int c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0;
  __shared__ PuRead buf[PU_TILE];
  for (int64_t q0 = a; q0 < b; q0 += PU_TILE) {
    const int nb = (int)((b - q0) < (int64_t)PU_TILE ? (b - q0) : (int64_t)PU_TILE);
    __syncthreads();
    if (threadIdx.x < nb) buf[threadIdx.x] = sorted[q0 + threadIdx.x];
    __syncthreads();
#pragma unroll 8
    for (int i = 0; i < nb; i++) {
      const PuRead r = buf[i];
      const unsigned d = (unsigned)(x - r.s);
      if (d < (unsigned)r.len) {
        const uint8_t c = rowpool[r.off + d];
        c0 += (c == 0);
        c1 += (c == 1);
        c2 += (c == 2);
        c3 += (c == 3);
        c4 += (c == 4);
      }
    }
  }
  if (x < G) {
    int32_t* dst = partial + ((size_t)sl * G + x) * 5;
    dst[0] = c0, dst[1] = c1, dst[2] = c2, dst[3] = c3, dst[4] = c4;
  }
}

// counts[x][c] (+)= sum over slices, in slice order
__global__ void pu_reduce_kernel(const int32_t* __restrict__ partial, int G, int nSlices, int accumulate,
                                 int32_t* __restrict__ counts) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)G * 5) return;
  int s = accumulate ? counts[e] : 0;
  for (int sl = 0; sl < nSlices; sl++) s += partial[(size_t)sl * G * 5 + e];
  counts[e] = s;
}

// ConsensusBuildingSimple.run (ConsensusBuilderSimple.java:98-120)
__global__ void pu_consensus_kernel(const int32_t* __restrict__ counts, int G, int saveUncovered, int coverageThreshold,
                                    const uint8_t* __restrict__ genome, uint8_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= G) return;
  const int32_t* c = counts + (size_t)i * 5;
  int maxCount = c[0], maxIdx = 0;
#pragma unroll
  for (int j = 1; j < 4; j++)
    if (c[j] > maxCount) maxCount = c[j], maxIdx = j;
  uint8_t r;
  if (maxCount <= coverageThreshold)
    r = saveUncovered ? genome[i] : (uint8_t)'-';
  else
    r = maxIdx == 0 ? 'A' : maxIdx == 1 ? 'T' : maxIdx == 2 ? 'G' : 'C';  // Constants.iToN
  out[i] = r;
}

// ContigBuilder.countReads pileup (ContigBuilder.java:93-99; the same loop as ConsensusBuilderSimple.java:56-75) for a batch
// of alignments against many small regions laid out back to back: alignment i starts at position pos[i] of the
// concatenated coordinate system. One thread per alignment; integer atomics (the result does not depend on their order).
__global__ void pu_alignments_kernel(const uint8_t* __restrict__ seq1, const int64_t* __restrict__ off,
                                     const int32_t* __restrict__ len, const int64_t* __restrict__ pos,
                                     const int64_t* __restrict__ end, int64_t n, int64_t total,
                                     int32_t* __restrict__ counts, int* __restrict__ err) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t k = pos[i];
  const int64_t kEnd = end ? min(end[i], total) : total;  // end of the alignment's own region (its counts array in the Java)
  const uint8_t* s = seq1 + off[i];
  for (int j = 0; j < len[i]; j++) {
    const uint8_t ch = s[j];
    if (ch >= 'a') continue;               // insertion in the read: no reference position
    if (ch > 'T' || k < 0 || k >= kEnd) {  // nToI has 'T' + 1 entries / counts[k] out of range: the Java throws
      *err = 1;
      return;
    }
    atomicAdd(&counts[k * 5 + pu_ntoi(ch)], 1);
    k++;
  }
}

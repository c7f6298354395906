// BAM record fields on the device: what BamPrinter derives per read from the mapping results
// before htsjdk serialises it — alignment start inside its contig, flags, mate fields (BamPrinter.java:16-115)
// and the CIGAR + XN attribute of setCigar (BamPrinter.java:121-171). The wire format (BGZF, index) stays with
// htsjdk in the Java host.
//
// One thread per record (pair result x mate): sequence1 is walked serially exactly like the Java loop. Two passes:
// count the CIGAR elements, exclusive scan, then emit (len << 4 | op, BAM op codes M=0 I=1 D=2 S=4).
// Quirk kept: the loop starts with operator M and length 0, so an alignment that begins with an insertion or a
// deletion gets a leading zero-length M element, and an empty alignment is the single element 0M.
#pragma once
#include "common.cuh"

struct BamArgs {
  const vg_pair_result* res;
  int64_t nPairs;
  const uint8_t* seq1pool;
  const int64_t* pairIdx;
  const int32_t* len1;
  const int32_t* len2;
  const int32_t* contigEnds;
  int nContigs;
  int32_t* nOps;            // [2 * nPairs]
  const int64_t* opOff;     // [2 * nPairs] (emit pass)
  vg_bam_record* out;
  uint32_t* ops;
};

// the element walk of setCigar; EMIT writes the elements, otherwise they are only counted
template <bool EMIT>
__device__ __forceinline__ int bam_cigar(const uint8_t* seq1, int length, int start1, int end1, int readLen, uint32_t* ops,
                                         int* mNumOut) {
  int n = 0;
  if (start1 > 0) {
    if (EMIT) ops[n] = ((uint32_t)start1 << 4) | 4u;
    n++;
  }
  int op = 0, len = 0, mNum = 0;
  for (int i = 0; i < length; i++) {
    const uint8_t ch = seq1[i];
    const int cls = (ch == '-') ? 2 : (ch < 'a') ? 0 : 1;
    mNum += (cls == 0);
    if (cls == op) {
      len++;
    } else {
      if (EMIT) ops[n] = ((uint32_t)len << 4) | (uint32_t)op;
      n++;
      op = cls;
      len = 1;
    }
  }
  if (EMIT) ops[n] = ((uint32_t)len << 4) | (uint32_t)op;
  n++;
  if (end1 < readLen) {
    if (EMIT) ops[n] = ((uint32_t)(readLen - end1) << 4) | 4u;
    n++;
  }
  *mNumOut = mNum;
  return n;
}

// BamPrinter.java:21-30: the contig that holds position s and the offset inside it (refIndex = nContigs and s unchanged
// when s lies behind the last contig, like the Java loop that runs off the end)
__device__ __forceinline__ void bam_locate(int s, const int32_t* contigEnds, int nContigs, int* refIndex, int* pos) {
  int prev = 0, r = 0;
  for (; r < nContigs; r++) {
    const int end = contigEnds[r];
    if (s < end) {
      s -= prev;
      break;
    }
    prev = end;
  }
  *refIndex = r;
  *pos = s + 1;
}

template <bool EMIT>
__global__ void bam_fields_kernel(BamArgs A) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 2 * A.nPairs) return;
  const int64_t pi = idx >> 1;
  const int mate = (int)(idx & 1);
  const vg_pair_result& r = A.res[pi];
  const vg_mate& me = r.m[mate];
  const vg_mate& other = r.m[mate ^ 1];
  const bool mapped = me.present != 0 && me.seq1_offset >= 0;
  int nOps = 0, mNum = 0;
  const uint8_t* seq1 = A.seq1pool + (mapped ? me.seq1_offset : 0);
  int readLen = 0;
  if (mapped) {
    const int64_t p = A.pairIdx[pi];
    readLen = mate ? A.len2[p] : A.len1[p];
  }
  if (!EMIT) {
    if (mapped) nOps = bam_cigar<false>(seq1, me.length, me.start1, me.end1, readLen, nullptr, &mNum);
    A.nOps[idx] = nOps;
    return;
  }
  vg_bam_record o;
  o.flag = 0x1 | (mate ? 0x80 : 0x40);
  if (r.cls == VG_CLS_CONCORDANT_FR || r.cls == VG_CLS_CONCORDANT_RF) o.flag |= 0x2;
  o.ref_index = -1, o.pos = 0, o.mapq = 0, o.mate_ref_index = -1, o.mate_pos = 0, o.xn = 0, o.n_cigar = 0;
  o.cigar_offset = A.opOff[idx];
  if (mapped) {
    bam_locate(me.start + me.start2, A.contigEnds, A.nContigs, &o.ref_index, &o.pos);
    o.mapq = 255;
    if (me.reverse) o.flag |= 0x10;
    o.n_cigar = bam_cigar<true>(seq1, me.length, me.start1, me.end1, readLen, A.ops + o.cigar_offset, &mNum);
    o.xn = mNum - me.identity;
  } else {
    o.flag |= 0x4;
  }
  if (other.present != 0 && other.seq1_offset >= 0) {
    bam_locate(other.start + other.start2, A.contigEnds, A.nContigs, &o.mate_ref_index, &o.mate_pos);
    if (other.reverse) o.flag |= 0x20;
  } else {
    o.flag |= 0x8;
  }
  A.out[idx] = o;
}

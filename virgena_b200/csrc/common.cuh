// Shared declarations for libvirgena_gpu.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/virgena_gpu.h"

#define VG_WARP 32
#define VG_FULL 0xffffffffu

// Packed s16x2 helpers ---------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t vg_pk(int lo, int hi) {
  return (uint32_t)(lo & 0xffff) | ((uint32_t)(hi & 0xffff) << 16);
}
__host__ __device__ __forceinline__ uint32_t vg_pk2(int v) { return vg_pk(v, v); }
__device__ __forceinline__ int vg_lo(uint32_t x) { return (int)(int16_t)(x & 0xffff); }
__device__ __forceinline__ int vg_hi(uint32_t x) { return (int)(int16_t)(x >> 16); }

// Constants.complement (Constants.java:11-20): A<->T, G<->C, '-'->'-', N->N, any other byte < 127 -> 0.
__device__ __forceinline__ uint8_t vg_complement(uint8_t c) {
  switch (c) {
    case 'A': return 'T';
    case 'T': return 'A';
    case 'G': return 'C';
    case 'C': return 'G';
    case '-': return '-';
    case 'N': return 'N';
    default: return 0;
  }
}
// Constants.lower (Constants.java:22-30): A,T,G,C,N -> lower case, anything else -> 0 (Q9).
__device__ __forceinline__ uint8_t vg_lower(uint8_t c) {
  switch (c) {
    case 'A': return 'a';
    case 'T': return 't';
    case 'G': return 'g';
    case 'C': return 'c';
    case 'N': return 'n';
    default: return 0;
  }
}

// One Smith-Waterman task: s1 = read (optionally reverse-complemented on load), s2 = window.
struct SwTask {
  int64_t s1_off;  // into the s1 byte pool
  int64_t s2_off;  // into the s2 byte pool
  int32_t m;       // |s1|
  int32_t n;       // |s2|
  int32_t rc;      // 1: s1 is the reverse complement of the stored bytes
  int32_t out;     // index of the output record
};

// Output of the fill kernel for one task: the exact maximum and where to look for its cell.
struct SwFillOut {
  int32_t score;   // S = max V
  int32_t strip;   // smallest strip index whose maximum equals S
  int32_t col;     // last column of the group of 4 wavefront steps in which that strip first reached S
  int32_t tie;     // 1: the strip reached S again in a later group (end cell needs a scan of the strip)
};

// Alignment record written by the traceback kernel (Alignment.java:4-35).
struct SwAln {
  int32_t score, start1, end1, start2, end2, identity, length;
};

struct SwParams {
  int32_t match, mismatch, gop, gep;
};

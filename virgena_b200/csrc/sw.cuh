// Smith-Waterman-Gotoh, batched, for sm_100a.
//
// Replaces SmithWatermanGotoh.align / construct / traceback
// (SmithWatermanGotoh.java:51-222) and findMaxScore (:240-300).
//
// Two kernels:
//   sw_fill_kernel<R>      inter-read batched, score-only fill of the full |s1| x |s2| rectangle.
//                          One warp = one *pair* of alignments packed in the two s16 lanes of every
//                          register (DPX: VIADDMNMX.S16x2, VIMNMX3-family). The 32 lanes own R
//                          consecutive rows each and sweep the columns as a systolic wavefront
//                          (lane k is one column behind lane k-1; the boundary row travels by SHFL).
//                          It stores no per-cell direction: only (V, E) along strip boundaries and
//                          (V, F) along every SW_TC-th column ("checkpoints"), plus the exact maximum
//                          and the strip/column where it was first reached.
//   sw_traceback_kernel<R> one thread = one alignment. Recomputes, from the checkpoints, only the
//                          R x SW_TC tiles that the traceback path crosses (about |s1|/R + |s2|/SW_TC
//                          of them), derives the reference's pointer byte and gap-extension flags for
//                          those cells and walks them exactly like SmithWatermanGotoh.traceback.
//
// Exactness notes (SURVEY.md §7 "hard parts"):
//   * E and F are kept shifted by +gop (Es = E + gop), so that
//       Es' = max(Es_up - gep, V_up)            "extend only if strictly greater" (:115) == (Es_up - gep > V_up)
//       V   = max(max(Es, Fs) - gop, f, 0)
//   * pointer priority STOP > DIAGONAL > UP > LEFT (:135-143) and the row-major "first strictly
//     greater" maximum (:146-150) are reproduced in the traceback kernel, which sees exact cell values.
//   * scores are exact small integers (the Java keeps them in float); they must fit int16.
#pragma once
#include "common.cuh"

#define SW_TC 32          // checkpoint column-block width
#define SW_NEG (-16384)   // stands in for Float.NEGATIVE_INFINITY (consumed only in row 1 / column 1)
#define SW_PAD_S1 0x7F00  // symbol code of a padded row (never equals a real code or SW_PAD_S2)
#define SW_PAD_S2 0x00FF  // symbol code of a padded column

__device__ __forceinline__ uint8_t sw_s1_byte(const uint8_t* s1, int m, int rc, int idx) {
  return rc ? vg_complement(s1[m - 1 - idx]) : s1[idx];
}

template <int R, bool STORE>
__global__ void __launch_bounds__(256) sw_fill_kernel(const SwTask* __restrict__ tasks, int ntasks,
                                                      const uint8_t* __restrict__ s1pool,
                                                      const uint8_t* __restrict__ s2pool, SwParams prm,
                                                      uint32_t* __restrict__ rowck, uint32_t* __restrict__ colck,
                                                      int nW, int ncb, SwFillOut* __restrict__ out,
                                                      int* __restrict__ counter, int smemWordsPerWarp) {
  extern __shared__ uint32_t sw_smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  uint32_t* wsym = sw_smem + (size_t)warp * smemWordsPerWarp;
  const int mW = 32 * R + 1;
  const uint32_t NEGp = vg_pk2(SW_NEG);
  const uint32_t NGEPp = vg_pk2(-prm.gep);
  const uint32_t NGOPp = vg_pk2(-prm.gop);
  const uint32_t MATCH1p = vg_pk2(prm.match + 1);
  const uint32_t MISp = vg_pk2(prm.mismatch);
  const int npairs = (ntasks + 1) >> 1;

  for (;;) {
    int pi = 0;
    if (lane == 0) pi = atomicAdd(counter, 1);
    pi = __shfl_sync(VG_FULL, pi, 0);
    if (pi >= npairs) break;
    const int ia = 2 * pi, ib = 2 * pi + 1;
    const SwTask ta = tasks[ia];
    SwTask tb;
    if (ib < ntasks) {
      tb = tasks[ib];
    } else {
      tb.s1_off = tb.s2_off = 0;
      tb.m = tb.n = 0;
      tb.rc = 0;
      tb.out = -1;
    }
    const uint8_t* a1 = s1pool + ta.s1_off;
    const uint8_t* b1 = s1pool + tb.s1_off;
    const uint8_t* a2 = s2pool + ta.s2_off;
    const uint8_t* b2 = s2pool + tb.s2_off;
    const int nmax = max(ta.n, tb.n);
    const int mmax = max(ta.m, tb.m);

    __syncwarp();
    for (int j = lane; j < nmax; j += 32) {
      uint32_t ca = (j < ta.n) ? ((uint32_t)a2[j] << 8) : SW_PAD_S2;
      uint32_t cb = (j < tb.n) ? ((uint32_t)b2[j] << 8) : SW_PAD_S2;
      wsym[j + 1] = ca | (cb << 16);
    }
    uint32_t a[R], V[R], F[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
      int idx = lane * R + r;
      uint32_t ca = (idx < ta.m) ? ((uint32_t)sw_s1_byte(a1, ta.m, ta.rc, idx) << 8) : SW_PAD_S1;
      uint32_t cb = (idx < tb.m) ? ((uint32_t)sw_s1_byte(b1, tb.m, tb.rc, idx) << 8) : SW_PAD_S1;
      a[r] = ca | (cb << 16);
      V[r] = 0;
      F[r] = NEGp;
    }
    __syncwarp();

    uint32_t vup_prev = 0, vbot = 0, ebot = NEGp, lanemax = 0;
    int colA = 0, colB = 0;
    bool tieA = false, tieB = false;
    // lanes whose strip starts beyond both reads never influence a result
    const int lastLane = min(31, (mmax - 1) / R);
    const int64_t rowBaseA = ((int64_t)ia * 32 + lane) * nW;
    const int64_t rowBaseB = ((int64_t)ib * 32 + lane) * nW;
    const bool rowStoreA = STORE && ((lane + 1) * R < ta.m);
    const bool rowStoreB = STORE && ((lane + 1) * R < tb.m);
    const int tEnd = nmax + lastLane;

    for (int t = 1; t <= tEnd; t++) {
      uint32_t vin = __shfl_up_sync(VG_FULL, vbot, 1);
      uint32_t ein = __shfl_up_sync(VG_FULL, ebot, 1);
      if (lane == 0) {
        vin = 0;
        ein = NEGp;
      }
      const int j = t - lane;
      if (j >= 1 && j <= nmax && lane <= lastLane) {
        const uint32_t b = wsym[j];
        uint32_t vd = vup_prev, vu = vin, e = ein, cm = 0;
#pragma unroll
        for (int r = 0; r < R; r++) {
          uint32_t nx = ~(a[r] ^ b);
          uint32_t sim = __viaddmax_s16x2(nx, MATCH1p, MISp);  // equal: match, else mismatch
          uint32_t f = __vadd2(vd, sim);
          e = __viaddmax_s16x2(e, NGEPp, vu);
          F[r] = __viaddmax_s16x2(F[r], NGEPp, V[r]);
          uint32_t tm = __vmaxs2(e, F[r]);
          vd = V[r];
          uint32_t v = __viaddmax_s16x2_relu(tm, NGOPp, f);
          V[r] = v;
          vu = v;
          if (r & 1)
            cm = __vimax3_s16x2(cm, V[r - 1], v);
          else if (r == R - 1)
            cm = __vmaxs2(cm, v);
        }
        vbot = vu;
        ebot = e;
        vup_prev = vin;
        {  // exact maximum: value per strip, first column reaching it, and whether it was reached again
          bool ph, pl, qh, ql;
          uint32_t nm = __vibmax_s16x2(lanemax, cm, &ph, &pl);  // p = lanemax >= cm
          (void)__vibmax_s16x2(cm, lanemax, &qh, &ql);          // q = cm >= lanemax
          tieA = pl ? (tieA | ql) : false;
          colA = pl ? colA : j;
          tieB = ph ? (tieB | qh) : false;
          colB = ph ? colB : j;
          lanemax = nm;
        }
        if (STORE) {
          if (rowStoreA && j <= ta.n) rowck[rowBaseA + j] = __byte_perm(vbot, ebot, 0x5410);
          if (rowStoreB && j <= tb.n) rowck[rowBaseB + j] = __byte_perm(vbot, ebot, 0x7632);
          if ((j & (SW_TC - 1)) == 0) {
            const int c = j / SW_TC - 1;
            if (j < ta.n) {
              uint32_t* dst = colck + ((int64_t)ia * ncb + c) * mW + lane * R + 1;
#pragma unroll
              for (int r = 0; r < R; r++)
                if (lane * R + r < ta.m) dst[r] = __byte_perm(V[r], F[r], 0x5410);
            }
            if (j < tb.n) {
              uint32_t* dst = colck + ((int64_t)ib * ncb + c) * mW + lane * R + 1;
#pragma unroll
              for (int r = 0; r < R; r++)
                if (lane * R + r < tb.m) dst[r] = __byte_perm(V[r], F[r], 0x7632);
            }
          }
        }
      }
    }
    // warp reduction: S, the smallest strip holding it, its first column and tie flag
    int sA = __reduce_max_sync(VG_FULL, vg_lo(lanemax));
    int sB = __reduce_max_sync(VG_FULL, vg_hi(lanemax));
    unsigned candA = __ballot_sync(VG_FULL, vg_lo(lanemax) == sA);
    unsigned candB = __ballot_sync(VG_FULL, vg_hi(lanemax) == sB);
    int kA = __ffs(candA) - 1, kB = __ffs(candB) - 1;
    int cA = __shfl_sync(VG_FULL, colA, kA), cB = __shfl_sync(VG_FULL, colB, kB);
    int tA = __shfl_sync(VG_FULL, (int)tieA, kA), tB = __shfl_sync(VG_FULL, (int)tieB, kB);
    if (lane == 0) {
      SwFillOut o;
      o.score = sA, o.strip = kA, o.col = cA, o.tie = tA;
      out[ia] = o;
      if (ib < ntasks) {
        o.score = sB, o.strip = kB, o.col = cB, o.tie = tB;
        out[ib] = o;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Traceback: recompute one R x SW_TC tile from the checkpoints. Cell flags (4 bits):
//   bits 0-1 pointer (STOP=0, LEFT=1, DIAGONAL=2, UP=3 — SmithWatermanGotoh.Directions :31-48),
//   bit 2 vertical-gap extended here (sizesOfVerticalGaps[l] = [l-n] + 1, :117),
//   bit 3 horizontal-gap extended here (:126).
// Optionally probes for cells equal to S (end-cell search, row-major minimum).
// ---------------------------------------------------------------------------------------------
template <int R>
struct SwTile {
  int k, c;  // strip / column block held in flags
};

template <int R>
__device__ __forceinline__ void sw_tile_recompute(const SwTask& tk, const uint8_t* __restrict__ s1,
                                                  const uint8_t* __restrict__ s2, const SwParams& prm,
                                                  const uint32_t* __restrict__ rowckT,  // [32][nW] of this task
                                                  const uint32_t* __restrict__ colckT,  // [ncb][mW] of this task
                                                  int nW, int k, int c, uint32_t* __restrict__ flags, int fstride,
                                                  int probeS, int probeCol, int& bestI, int& bestJ) {
  const int mW = 32 * R + 1;
  const int r0 = k * R, c0 = c * SW_TC;
  int v[SW_TC + 1], g[SW_TC + 1];
  uint8_t bsym[SW_TC + 1];
#pragma unroll
  for (int jj = 0; jj <= SW_TC; jj++) {
    int col = c0 + jj;
    int vv = 0, ee = SW_NEG;
    if (k > 0 && col >= 1 && col <= tk.n) {
      uint32_t w = rowckT[(int64_t)(k - 1) * nW + col];
      vv = vg_lo(w);
      ee = vg_hi(w);
    }
    if (k > 0 && col == 0) vv = 0;
    v[jj] = vv;
    g[jj] = ee;
    bsym[jj] = (col >= 1 && col <= tk.n) ? s2[col - 1] : 0;
  }
  int vdiagRow = v[0];  // V(r0 + rr, c0) of the previous row
  for (int rr = 0; rr < R; rr++) {
    const int i = r0 + rr + 1;
    if (i > tk.m) break;
    int vleft = 0, hs = SW_NEG;
    if (c > 0) {
      uint32_t w = colckT[(int64_t)(c - 1) * mW + i];
      vleft = vg_lo(w);
      hs = vg_hi(w);
    }
    const uint8_t asym = sw_s1_byte(s1, tk.m, tk.rc, i - 1);
    int vd = vdiagRow;
    vdiagRow = vleft;
    uint32_t words[SW_TC / 8];
#pragma unroll
    for (int w = 0; w < SW_TC / 8; w++) words[w] = 0;
#pragma unroll
    for (int jj = 1; jj <= SW_TC; jj++) {
      const int j = c0 + jj;
      if (j <= tk.n) {
        int sim = (asym == bsym[jj]) ? prm.match : prm.mismatch;
        int f = vd + sim;
        int g1 = g[jj] - prm.gep, g2 = v[jj];
        bool vext = g1 > g2;
        int es = vext ? g1 : g2;
        int h1 = hs - prm.gep, h2 = vleft;
        bool hext = h1 > h2;
        hs = hext ? h1 : h2;
        vd = v[jj];
        int E = es - prm.gop, Fv = hs - prm.gop;
        int V = max(max(f, E), max(Fv, 0));
        uint32_t dir = (V == 0) ? 0u : (V == f) ? 2u : (V == E) ? 3u : 1u;
        uint32_t nib = dir | (vext ? 4u : 0u) | (hext ? 8u : 0u);
        words[(jj - 1) >> 3] |= nib << (((jj - 1) & 7) * 4);
        g[jj] = es;
        v[jj] = V;
        vleft = V;
        if (V == probeS && (probeCol < 0 || probeCol == j)) {
          if (i < bestI || (i == bestI && j < bestJ)) {
            bestI = i;
            bestJ = j;
          }
        }
      }
    }
#pragma unroll
    for (int w = 0; w < SW_TC / 8; w++) flags[(rr * (SW_TC / 8) + w) * fstride] = words[w];
  }
}

template <int R>
__global__ void __launch_bounds__(128) sw_traceback_kernel(const SwTask* __restrict__ tasks, int ntasks,
                                                           const uint8_t* __restrict__ s1pool,
                                                           const uint8_t* __restrict__ s2pool, SwParams prm,
                                                           const uint32_t* __restrict__ rowck,
                                                           const uint32_t* __restrict__ colck, int nW, int ncb,
                                                           const SwFillOut* __restrict__ fill,
                                                           SwAln* __restrict__ alns, uint8_t* __restrict__ rev,
                                                           int64_t revStride) {
  __shared__ uint32_t flagS[R * (SW_TC / 8) * 128];
  const int tid = threadIdx.x;
  const int ti = blockIdx.x * blockDim.x + tid;
  if (ti >= ntasks) return;
  const SwTask tk = tasks[ti];
  const uint8_t* s1 = s1pool + tk.s1_off;
  const uint8_t* s2 = s2pool + tk.s2_off;
  const int mW = 32 * R + 1;
  const uint32_t* rowckT = rowck + (int64_t)ti * 32 * nW;
  const uint32_t* colckT = colck + (int64_t)ti * ncb * mW;
  uint32_t* flags = flagS + tid;
  const SwFillOut fo = fill[ti];
  SwAln al;
  al.score = fo.score;
  uint8_t* out = rev + (int64_t)tk.out * revStride;
  if (fo.score <= 0) {  // Q8: no cell > 0 -> empty alignment at (0,0)
    al.score = 0, al.start1 = al.end1 = al.start2 = al.end2 = al.identity = al.length = 0;
    alns[tk.out] = al;
    return;
  }
  int no1 = 1 << 30, no2 = 1 << 30;
  int bi = 1 << 30, bj = 1 << 30;
  int curK = -1, curC = -1;
  if (!fo.tie) {
    curK = fo.strip, curC = (fo.col - 1) / SW_TC;
    sw_tile_recompute<R>(tk, s1, s2, prm, rowckT, colckT, nW, curK, curC, flags, 128, fo.score, fo.col, bi, bj);
  } else {
    const int nc = (tk.n + SW_TC - 1) / SW_TC;
    for (int c = 0; c < nc; c++)
      sw_tile_recompute<R>(tk, s1, s2, prm, rowckT, colckT, nW, fo.strip, c, flags, 128, fo.score, -1, bi, bj);
    curK = -1;
  }
  int i = bi, j = bj;
  al.end1 = i;
  al.end2 = j;
  int len = 0, identity = 0;
  int mode = 0;  // 0 normal, 1 inside a vertical gap run, 2 inside a horizontal gap run
  for (;;) {
    uint32_t nib = 0;
    if (i >= 1 && j >= 1) {
      int k = (i - 1) / R, c = (j - 1) / SW_TC;
      if (k != curK || c != curC) {
        curK = k, curC = c;
        sw_tile_recompute<R>(tk, s1, s2, prm, rowckT, colckT, nW, k, c, flags, 128, -1, -1, no1, no2);
      }
      int rr = (i - 1) - k * R, jj = (j - 1) - c * SW_TC;
      nib = (flags[(rr * (SW_TC / 8) + (jj >> 3)) * 128] >> ((jj & 7) * 4)) & 15u;
    }
    if (mode == 0) {
      uint32_t dir = nib & 3u;
      if (dir == 0u) break;  // STOP (also the row-0 / column-0 boundary)
      if (dir == 2u) {       // DIAGONAL
        uint8_t c1 = sw_s1_byte(s1, tk.m, tk.rc, i - 1), c2 = s2[j - 1];
        out[len++] = c1;
        if (c1 == c2) identity++;
        i--, j--;
        continue;
      }
      mode = (dir == 3u) ? 1 : 2;
    }
    if (mode == 1) {  // UP: sizesOfVerticalGaps[k+j] lower-case read bases
      out[len++] = vg_lower(sw_s1_byte(s1, tk.m, tk.rc, i - 1));
      bool ext = (nib & 4u) != 0;
      i--;
      if (!ext) mode = 0;
    } else {  // LEFT: sizesOfHorizontalGaps[k+j] gap characters
      out[len++] = '-';
      bool ext = (nib & 8u) != 0;
      j--;
      if (!ext) mode = 0;
    }
  }
  al.start1 = i;
  al.start2 = j;
  al.identity = identity;
  al.length = len;
  alns[tk.out] = al;
}

// sequence1 = reverse(reversed1) into the caller's pool (SmithWatermanGotoh.reverse :309-315)
__global__ void sw_emit_seq1_kernel(const SwAln* __restrict__ alns, int n, const uint8_t* __restrict__ rev,
                                    int64_t revStride, const int64_t* __restrict__ seqOff,
                                    uint8_t* __restrict__ pool) {
  int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (w >= n) return;
  int len = alns[w].length;
  int64_t off = seqOff[w];
  if (off < 0) return;
  const uint8_t* src = rev + (int64_t)w * revStride;
  for (int t = lane; t < len; t += 32) pool[off + t] = src[len - 1 - t];
}

// DPX issue-rate microbenchmark: eight independent chains of VIADDMNMX.S16x2 / VIMNMX3.S16x2 per thread.
__global__ void dpx_peak_kernel(uint32_t* out, int iters) {
  uint32_t a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const uint32_t c = vg_pk2(-2), d = blockIdx.x;
  for (int i = 0; i < iters; i++) {
    a0 = __viaddmax_s16x2(a0, c, d);
    a1 = __viaddmax_s16x2(a1, c, d);
    a2 = __viaddmax_s16x2(a2, c, d);
    a3 = __viaddmax_s16x2(a3, c, d);
    a4 = __vimax3_s16x2(a4, c, d);
    a5 = __vimax3_s16x2(a5, c, d);
    a6 = __viaddmax_s16x2_relu(a6, c, d);
    a7 = __viaddmax_s16x2_relu(a7, c, d);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

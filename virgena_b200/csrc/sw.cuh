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

#define SW_TC 16          // checkpoint column-block width (wavefront steps between column checkpoints)
#define SW_NEG (-16384)   // stands in for Float.NEGATIVE_INFINITY (consumed only in row 1 / column 1)
#define SW_PAD_S1 0x7F00  // symbol code of a padded row (never equals a real code or SW_PAD_S2)
#define SW_PAD_S2 0x00FF  // symbol code of a padded column

__device__ __forceinline__ uint8_t sw_s1_byte(const uint8_t* s1, int m, int rc, int idx) {
  return rc ? vg_complement(s1[m - 1 - idx]) : s1[idx];
}
// the same with Constants.complement as a 128-byte table in shared memory: one load, no divergence on the base (the traceback
// fetches a read byte per row of every tile and per walk step, and the lanes of its warps hold different reads)
__device__ __forceinline__ uint8_t sw_s1_byte_lut(const uint8_t* s1, int m, int rc, int idx, const uint8_t* lut) {
  const uint8_t b = s1[rc ? m - 1 - idx : idx];
  return rc ? lut[b & 127] : b;
}

// Checkpoint layout (written by the fill kernel, read by the traceback kernel), per *pair* p of tasks
// (2p, 2p+1) whose values sit in the lo / hi s16 lanes:
//   rowck[p][lane][t-1] = uint2{V, Es} of lane's bottom row at wavefront step t (column j = t - lane); nT entries per lane
//   colck[p][tb][row]  = {V, Fs} of row (0-based) at step t = SW_TC*tb             (column j = SW_TC*tb - row/R)
// Indexing by the wavefront step makes every store warp-uniform; the row checkpoints go out as whole 32-byte sectors
// (4 steps per lane) and come back to the traceback as contiguous runs.
// Checkpoint element. Wide: uint2{V, Xs} (Xs = Es or Fs, the gap state shifted by +gop), both s16x2 for the two
// alignments of the pair. Compact (scores < 4096 and 0 <= gop - gep <= 15): one 32-bit word, per 16-bit lane V in the low
// 12 bits and d = max(Xs - gep - V, 0) in the top 4. That is all the traceback needs of Xs: the cell below / to the right
// extends the gap iff Xs - gep > V, i.e. iff d > 0, and then continues with Xs - gep = V + d; otherwise it restarts from V.
// (Xs <= V + gop because E, F <= V, hence d <= gop - gep.) Xs is handed back as V + d + gep.
template <bool COMPACT>
struct SwCk;
template <>
struct SwCk<false> {
  using T = uint2;
  static __device__ __forceinline__ T make(uint32_t v, uint32_t xs, uint32_t) { return make_uint2(v, xs); }
  static __device__ __forceinline__ void get(T w, int hi, int, int& v, int& xs) {
    v = hi ? vg_hi(w.x) : vg_lo(w.x);
    xs = hi ? vg_hi(w.y) : vg_lo(w.y);
  }
};
template <>
struct SwCk<true> {
  using T = uint32_t;
  static __device__ __forceinline__ T make(uint32_t v, uint32_t xs, uint32_t ngepP) {
    // max(Xs - gep, V) - V per lane, 0 .. 15: no half is negative, so the 32-bit subtraction is exact per half (FMA pipe)
    // t = max(Xs - gep, V) >= V in both halves, so (t - v) * 4096 + v is exact per half in 32-bit arithmetic; written as
    // t * 4096 + v * (1 - 4096) it is two multiply-adds on the FMA pipe and nothing on the ALU / DPX pipe that bounds the kernel
    const uint32_t t = __viaddmax_s16x2(xs, ngepP, v);
    return t * 4096u + v * 0xFFFFF001u;  // = v | d << 12, d = t - v (the fields do not overlap)
  }
  static __device__ __forceinline__ void get(T w, int hi, int gep, int& v, int& xs) {
    const uint32_t h = hi ? (w >> 16) : (w & 0xFFFFu);
    v = (int)(h & 0xFFFu);
    xs = v + (int)(h >> 12) + gep;
  }
};

// colA/colB = last column of the group of 4 wavefront steps in which the strip first reached its maximum (the group
// (4q, 4q+4] lies inside one checkpoint tile because the column checkpoints are taken at steps that are multiples of 32)
#define SW_MAX_BOOKKEEPING(JEND)                                              \
  {                                                                           \
    bool ph, pl, qh, ql;                                                      \
    const uint32_t nm = __vibmax_s16x2(lanemax, cm4, &ph, &pl); /* p = lanemax >= cm4 */ \
    (void)__vibmax_s16x2(cm4, lanemax, &qh, &ql);              /* q = cm4 >= lanemax */ \
    tieA = pl ? (tieA | ql) : false;                                          \
    colA = pl ? colA : (JEND);                                                \
    tieB = ph ? (tieB | qh) : false;                                          \
    colB = ph ? colB : (JEND);                                                \
    lanemax = nm;                                                             \
    cm4 = 0;                                                                  \
  }

// FASTF (gop + mismatch >= 0): the diagonal term is kept shifted by +gop like E and F, g_s = V_diag + (equal ? match + gop :
// mismatch + gop). Both summands are non-negative in both 16-bit halves and the sum stays below 2^15, so a plain 32-bit
// add is exact per half and issues on the FMA pipe (IMAD.IADD) instead of the ALU / DPX pipe that bounds the kernel;
// V = relu(max(Es, Fs, g_s) - gop) is VIMNMX3 + VIADDMNMX.RELU. 6.5 instead of 7.5 ALU-pipe instructions per cell-pair.
// W (32, 16 or 8) = lanes per pair of alignments: reads of at most W * R bases run 32 / W pairs per warp, every group of W
// lanes its own wavefront (shuffles of width W, a skew of W - 1 steps instead of 31, no idle lanes for short reads). The
// checkpoint layouts are the same with W in the place of 32.
template <int R, bool STORE, bool COMPACT, bool FASTF, int W = 32>
__global__ void __launch_bounds__(256) sw_fill_kernel(const SwTask* __restrict__ tasks, int ntasks,
                                                      const uint8_t* __restrict__ s1pool,
                                                      const uint8_t* __restrict__ s2pool, SwParams prm,
                                                      typename SwCk<COMPACT>::T* __restrict__ rowck,
                                                      typename SwCk<COMPACT>::T* __restrict__ colck,
                                                      int nT, int nTb, SwFillOut* __restrict__ out,
                                                      int* __restrict__ counter, int smemWordsPerWarp) {
  extern __shared__ uint32_t sw_smem[];
  constexpr int NG = 32 / W;  // pairs per warp
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int grp = lane / W, gl = lane % W;  // group of the lane, lane inside the group (= strip)
  const unsigned gmask = (W == 32) ? VG_FULL : (((1u << (W & 31)) - 1u) << (grp * W));
  uint32_t* wsym = sw_smem + ((size_t)warp * NG + grp) * smemWordsPerWarp;
  const uint32_t NEGp = vg_pk2(SW_NEG);
  const uint32_t NGEPp = vg_pk2(-prm.gep);
  const uint32_t NGOPp = vg_pk2(-prm.gop);
  const uint32_t MATCH1p = vg_pk2(prm.match + 1);
  const uint32_t MISp = vg_pk2(prm.mismatch);
  const uint32_t MATCHG1p = vg_pk2(prm.match + prm.gop + 1);  // FASTF: equal -> nx + this = match + gop
  const uint32_t MISGp = vg_pk2(prm.mismatch + prm.gop);      //        else mismatch + gop (>= 0)
  const int npairs = (ntasks + 1) >> 1;

  for (;;) {
    int pi = 0;
    if (lane == 0) pi = atomicAdd(counter, NG);
    pi = __shfl_sync(VG_FULL, pi, 0);
    if (pi >= npairs) break;
    pi += grp;
    const bool have = pi < npairs;  // a group beyond the last pair runs along with empty sequences
    const int ia = 2 * pi, ib = 2 * pi + 1;
    SwTask ta;
    if (have) {
      ta = tasks[ia];
    } else {
      ta.s1_off = ta.s2_off = 0;
      ta.m = ta.n = 0;
      ta.rc = 0;
      ta.out = -1;
    }
    SwTask tb;
    if (have && ib < ntasks) {
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
    for (int j = gl; j < nmax; j += W) {
      uint32_t ca = (j < ta.n) ? ((uint32_t)a2[j] << 8) : SW_PAD_S2;
      uint32_t cb = (j < tb.n) ? ((uint32_t)b2[j] << 8) : SW_PAD_S2;
      wsym[j + 1] = ca | (cb << 16);
    }
    uint32_t a[R], V[R], F[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
      int idx = gl * R + r;
      uint32_t ca = (idx < ta.m) ? ((uint32_t)sw_s1_byte(a1, ta.m, ta.rc, idx) << 8) : SW_PAD_S1;
      uint32_t cb = (idx < tb.m) ? ((uint32_t)sw_s1_byte(b1, tb.m, tb.rc, idx) << 8) : SW_PAD_S1;
      a[r] = ~(ca | (cb << 16));  // complemented: a[r] ^ b is the XNOR of the codes in one LOP3
      V[r] = 0;
      F[r] = NEGp;
    }
    __syncwarp();

    uint32_t vup_prev = 0, vbot = 0, ebot = NEGp, lanemax = 0, cm4 = 0;
    int colA = 0, colB = 0;
    bool tieA = false, tieB = false;
    // lanes whose strip starts beyond both reads never influence a result
    const int lastLane = mmax > 0 ? min(W - 1, (mmax - 1) / R) : 0;
    const int tEnd = (W == 32) ? nmax + lastLane : __reduce_max_sync(VG_FULL, nmax > 0 ? nmax + lastLane : 0);
    // checkpoints of the lane's bottom row: nT entries per lane, entry t - 1 = wavefront step t, written 4 steps
    // (one 32-byte sector) at a time so that the traceback reads the boundary of a tile as one contiguous run
    using CK = SwCk<COMPACT>;
    typename CK::T* rowp = rowck + ((int64_t)pi * W + gl) * nT;
    typename CK::T* colp = colck + ((int64_t)pi * nTb) * (W * R) + gl * R;
    const int tEnd4 = (tEnd + 3) & ~3;  // steps beyond tEnd find every lane inactive
    const unsigned nAct = gl <= lastLane ? (unsigned)nmax : 0u;  // columns this lane computes

#pragma unroll 1
    for (int tg = 0; tg < tEnd4; tg += 4) {
      typename CK::T rb[4];
#pragma unroll
      for (int s4 = 1; s4 <= 4; s4++) {
        const int t = tg + s4;
        uint32_t vin = __shfl_up_sync(VG_FULL, vbot, 1, W);
        uint32_t ein = __shfl_up_sync(VG_FULL, ebot, 1, W);
        if (gl == 0) {
          vin = 0;
          ein = NEGp;
        }
        const int j = t - gl;
        const bool active = (unsigned)(j - 1) < nAct;  // 1 <= j <= nmax on a lane that holds rows: one unsigned compare
        if (active) {
          const uint32_t b = wsym[j];
          uint32_t vd = vup_prev, vu = vin, e = ein, cm = cm4;  // running maximum of the current group of 4 steps
#pragma unroll
          for (int r = 0; r < R; r++) {
            uint32_t nx = a[r] ^ b;  // 0xFFFF in a lane whose symbols are equal, negative and < -256 otherwise
            uint32_t v;
            if constexpr (FASTF) {
              const uint32_t simg = __viaddmax_s16x2(nx, MATCHG1p, MISGp);  // equal: match + gop, else mismatch + gop
              const uint32_t gs = vd + simg;                               // exact per half (see above): FMA pipe
              e = __viaddmax_s16x2(e, NGEPp, vu);
              F[r] = __viaddmax_s16x2(F[r], NGEPp, V[r]);
              const uint32_t t2 = __vimax3_s16x2(e, F[r], gs);
              vd = V[r];
              v = __viaddmax_s16x2_relu(t2, NGOPp, NGOPp);  // relu(t2 - gop): the third operand (negative) never wins
            } else {
              uint32_t sim = __viaddmax_s16x2(nx, MATCH1p, MISp);  // equal: match, else mismatch
              uint32_t f = __vadd2(vd, sim);
              e = __viaddmax_s16x2(e, NGEPp, vu);
              F[r] = __viaddmax_s16x2(F[r], NGEPp, V[r]);
              uint32_t tm = __vmaxs2(e, F[r]);
              vd = V[r];
              v = __viaddmax_s16x2_relu(tm, NGOPp, f);
            }
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
          cm4 = cm;
        }
        if (STORE) rb[s4 - 1] = CK::make(vbot, ebot, NGEPp);
      }
      // once per group of 4 steps (warp-uniform): exact maximum of the strip, the group of 4 columns that reached it
      // first, and whether a later group reached it again
      SW_MAX_BOOKKEEPING(tg + 4 - gl)
      if (STORE) {
        const int j1 = tg + 1 - gl, j4 = tg + 4 - gl;
        if (gl <= lastLane && j4 >= 1 && j1 <= nmax) {
          uint4* dst = (uint4*)(rowp + tg);
          if constexpr (COMPACT) {
            dst[0] = make_uint4(rb[0], rb[1], rb[2], rb[3]);
          } else {
            dst[0] = make_uint4(rb[0].x, rb[0].y, rb[1].x, rb[1].y);
            dst[1] = make_uint4(rb[2].x, rb[2].y, rb[3].x, rb[3].y);
          }
        }
        if (((tg + 4) & (SW_TC - 1)) == 0 && gl <= lastLane && j4 >= 1 && j4 <= nmax) {  // column checkpoint of the rows
          typename CK::T* dst = colp + (int64_t)((tg + 4) / SW_TC) * (W * R);
          if constexpr (COMPACT && (R % 4 == 0)) {  // a lane's R words are contiguous and 16-byte aligned: whole-sector stores
#pragma unroll
            for (int r = 0; r < R; r += 4)
              *(uint4*)(dst + r) = make_uint4(CK::make(V[r], F[r], NGEPp), CK::make(V[r + 1], F[r + 1], NGEPp),
                                              CK::make(V[r + 2], F[r + 2], NGEPp), CK::make(V[r + 3], F[r + 3], NGEPp));
          } else {
#pragma unroll
            for (int r = 0; r < R; r++) dst[r] = CK::make(V[r], F[r], NGEPp);
          }
        }
      }
    }
    // warp reduction: S, the smallest strip holding it, its first column group and tie flag
    int sA = __reduce_max_sync(gmask, vg_lo(lanemax));
    int sB = __reduce_max_sync(gmask, vg_hi(lanemax));
    unsigned candA = __ballot_sync(VG_FULL, vg_lo(lanemax) == sA) & gmask;
    unsigned candB = __ballot_sync(VG_FULL, vg_hi(lanemax) == sB) & gmask;
    int kA = __ffs(candA) - 1, kB = __ffs(candB) - 1;  // lanes of the warp; the strip is the lane inside the group
    int cA = __shfl_sync(VG_FULL, colA, kA), cB = __shfl_sync(VG_FULL, colB, kB);
    int tA = __shfl_sync(VG_FULL, (int)tieA, kA), tB = __shfl_sync(VG_FULL, (int)tieB, kB);
    if (gl == 0 && have) {
      SwFillOut o;
      o.score = sA, o.strip = kA - grp * W, o.col = cA, o.tie = tA;
      out[ia] = o;
      if (ib < ntasks) {
        o.score = sB, o.strip = kB - grp * W, o.col = cB, o.tie = tB;
        out[ib] = o;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Traceback. One thread = one alignment. A "tile" is the strip k (rows k*R+1 .. k*R+R) between two
// column checkpoints of that strip: columns cl+1 .. cl+32 where cl = 32*tb - k (or 0). The tile is
// recomputed column by column from its top boundary (rowck of strip k-1) and left boundary (colck),
// producing for every cell the 4 flag bits
//   bits 0-1 pointer (STOP=0, LEFT=1, DIAGONAL=2, UP=3 — SmithWatermanGotoh.Directions :31-48),
//   bit 2 vertical-gap extended here (sizesOfVerticalGaps[l] = [l-n] + 1, :117),
//   bit 3 horizontal-gap extended here (:126),
// one 32-bit word per column (R <= 8) or two (R <= 16), kept in shared memory.
// All lanes of a warp recompute together (one tile per outer iteration), then walk inside their tile.
// ---------------------------------------------------------------------------------------------
template <int R>
struct SwTb {
  static constexpr int WPC = (R + 7) / 8;  // flag words per column
};

// Flag word of 8 rows of one column: bits 2r..2r+1 pointer of row r (STOP=0, LEFT=1, DIAGONAL=2, UP=3), bit 16+r vertical
// gap extended at row r, bit 24+r horizontal gap extended.
// PROBE: also look for the end cell (smallest (row, column) with V == probeS); only the first tile(s) of an alignment need it.
template <int R, bool COMPACT, bool PROBE, int W = 32>
__device__ __forceinline__ void sw_tile_recompute(const SwTask& tk, const uint8_t* __restrict__ s1,
                                                  const uint8_t* __restrict__ s2, const SwParams& prm,
                                                  const typename SwCk<COMPACT>::T* __restrict__ rowckP,
                                                  const typename SwCk<COMPACT>::T* __restrict__ colckP,
                                                  int nT, int hi, int k, int cl, int jmax, uint32_t* __restrict__ flags,
                                                  int fstride, int probeS, int probeCol, int& bestI, int& bestJ,
                                                  const uint8_t* __restrict__ lut) {
  constexpr int WPC = SwTb<R>::WPC;
  const int r0 = k * R;
  int v[R], hs[R];
  uint8_t asym[R];
  // left boundary: column cl
  const int tb = (cl + k) / SW_TC;
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int i = r0 + r + 1;
    v[r] = 0;
    hs[r] = SW_NEG;
    asym[r] = 0;
    if (i <= tk.m) {
      asym[r] = sw_s1_byte_lut(s1, tk.m, tk.rc, i - 1, lut);
      if (cl > 0) {
        SwCk<COMPACT>::get(colckP[(int64_t)tb * (W * R) + (i - 1)], hi, prm.gep, v[r], hs[r]);
      }
    }
  }
  // V(r0, cl): corner for the first diagonal
  // top boundary = bottom row of strip k-1 (lane k-1): column j was computed at step j + k - 1, entry j + k - 2
  using CK = SwCk<COMPACT>;
  const typename CK::T* topRow = rowckP + (int64_t)(k - 1) * nT + (k - 2);  // [j] (unused when k == 0)
  int vtopPrev = 0;
  if (k > 0 && cl > 0) {
    int unusedXs;
    CK::get(topRow[cl], hi, prm.gep, vtopPrev, unusedXs);
  }
  // the boundary of the tile is one contiguous run of 8-byte checkpoints in HBM: both of its lines are requested up
  // front and the load of the next column is issued before this column is computed
  typename CK::T wNext = typename CK::T();
  if (k > 0 && cl + 1 <= jmax) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(topRow + min(jmax, cl + SW_TC)));
    wNext = topRow[cl + 1];
  }
  const int gep = prm.gep, gop = prm.gop, match = prm.match, mismatch = prm.mismatch;
  for (int j = cl + 1; j <= jmax; j++) {
    int vtop = 0, es = SW_NEG;
    if (k > 0) {
      const typename CK::T w = wNext;
      if (j + 1 <= jmax) wNext = topRow[j + 1];
      CK::get(w, hi, gep, vtop, es);
    }
    const uint8_t bsym = s2[j - 1];
    int vd = vtopPrev, vu = vtop;
    vtopPrev = vtop;
    uint32_t word[WPC];
#pragma unroll
    for (int w = 0; w < WPC; w++) word[w] = 0;
#pragma unroll
    for (int r = 0; r < R; r++) {
      const int f = vd + ((asym[r] == bsym) ? match : mismatch);
      // "extend only if strictly greater" (:115, :124): the predicate of the maximum is (kept >= extended)
      bool keepV, keepH, fGeE, mGeF;
      es = __vibmax_s32(vu, es - gep, &keepV);         // Es' = max(V_up, Es_up - gep)
      hs[r] = __vibmax_s32(v[r], hs[r] - gep, &keepH);  // Fs' = max(V_left, Fs_left - gep)
      vd = v[r];
      const int m1 = __vibmax_s32(f, es - gop, &fGeE);
      const int m2 = __vibmax_s32(m1, hs[r] - gop, &mGeF);
      bool nonPos;
      const int V = __vibmax_s32(0, m2, &nonPos);  // nonPos = (0 >= m2): the maximum and the STOP test in one instruction
      // STOP (V == 0) > DIAGONAL (V == f) > UP (V == E) > LEFT (:135-143)
      const uint32_t dir = nonPos ? 0u : (mGeF ? (fGeE ? 2u : 3u) : 1u);
      uint32_t wv = word[r >> 3] + (dir << (2 * (r & 7)));
      if (!keepV) wv |= 1u << (16 + (r & 7));
      if (!keepH) wv |= 1u << (24 + (r & 7));
      word[r >> 3] = wv;
      v[r] = V;
      vu = V;
      if constexpr (PROBE) {
        const int i = r0 + r + 1;
        if (V == probeS && i <= tk.m && (probeCol < 0 || (j <= probeCol && j + 3 >= probeCol))) {
          if (i < bestI || (i == bestI && j < bestJ)) {
            bestI = i;
            bestJ = j;
          }
        }
      }
    }
#pragma unroll
    for (int w = 0; w < WPC; w++) flags[((j - cl - 1) * WPC + w) * fstride] = word[w];
  }
}

// left checkpoint column of the tile of strip k that contains column j (0 = the matrix edge)
__device__ __forceinline__ int sw_tile_left(int k, int j) {
  int tb = (j - 1 + k) / SW_TC;
  int cl = SW_TC * tb - k;
  return cl < 1 ? 0 : cl;
}

template <int R, bool COMPACT, int W = 32>
__global__ void __launch_bounds__(128, 7) sw_traceback_kernel(const SwTask* __restrict__ tasks, int ntasks,
                                                           const uint8_t* __restrict__ s1pool,
                                                           const uint8_t* __restrict__ s2pool, SwParams prm,
                                                           const typename SwCk<COMPACT>::T* __restrict__ rowck,
                                                           const typename SwCk<COMPACT>::T* __restrict__ colck, int nT, int nTb,
                                                           const SwFillOut* __restrict__ fill,
                                                           SwAln* __restrict__ alns, uint8_t* __restrict__ rev,
                                                           int64_t revStride) {
  constexpr int WPC = SwTb<R>::WPC;
  __shared__ uint32_t flagS[SW_TC * WPC * 128];
  __shared__ uint8_t compLut[128];
  const int tid = threadIdx.x;
  compLut[tid & 127] = vg_complement((uint8_t)(tid & 127));  // blockDim.x == 128
  __syncthreads();
  const int ti = blockIdx.x * blockDim.x + tid;
  const bool valid = ti < ntasks;
  SwTask tk;
  tk.m = tk.n = 0;
  tk.s1_off = tk.s2_off = 0;
  tk.rc = 0;
  tk.out = 0;
  SwFillOut fo;
  fo.score = 0, fo.strip = 0, fo.col = 0, fo.tie = 0;
  if (valid) {
    tk = tasks[ti];
    fo = fill[ti];
  }
  const uint8_t* s1 = s1pool + tk.s1_off;
  const uint8_t* s2 = s2pool + tk.s2_off;
  const int hi = ti & 1;
  const typename SwCk<COMPACT>::T* rowckP = rowck + ((int64_t)(ti >> 1) * W) * nT;
  const typename SwCk<COMPACT>::T* colckP = colck + ((int64_t)(ti >> 1) * nTb) * (W * R);
  uint32_t* flags = flagS + tid;
  uint8_t* out = rev + (int64_t)tk.out * revStride;
  int no1 = 1 << 30, no2 = 1 << 30;
  int i = 0, j = 0, len = 0, identity = 0, end1 = 0, end2 = 0;
  bool done = !valid || fo.score <= 0;  // Q8: no cell > 0 -> empty alignment at (0,0)

  // ---- locate the end cell: smallest (row, column) with V == S ------------------------------------------
  int haveK = -1, haveCl = -1;  // tile whose flags are in shared memory
  if (!done) {
    int bi = 1 << 30, bj = 1 << 30;
    if (!fo.tie) {
      const int jEnd = min(fo.col, tk.n);  // fo.col = last column of the group of 4 that reached the maximum
      int cl = sw_tile_left(fo.strip, jEnd);
      sw_tile_recompute<R, COMPACT, true, W>(tk, s1, s2, prm, rowckP, colckP, nT, hi, fo.strip, cl, jEnd, flags, 128, fo.score, fo.col, bi, bj, compLut);
      haveK = fo.strip, haveCl = cl;
    } else {
      for (int jj = tk.n; jj >= 1;) {  // rare: the strip reached S at several columns
        int cl = sw_tile_left(fo.strip, jj);
        sw_tile_recompute<R, COMPACT, true, W>(tk, s1, s2, prm, rowckP, colckP, nT, hi, fo.strip, cl, jj, flags, 128, fo.score, -1, bi, bj, compLut);
        jj = cl;
      }
    }
    i = bi, j = bj;
    end1 = i, end2 = j;
  }
  // ---- walk: one tile recompute per outer iteration, all lanes of the warp together ---------------------
  int mode = 0;  // 0 normal, 1 inside a vertical gap run, 2 inside a horizontal gap run
  while (__any_sync(VG_FULL, !done)) {
    int k = 0, cl = 0;
    if (!done) {
      k = (i - 1) / R;
      cl = sw_tile_left(k, j);
      {  // the path leaves this tile to the left (same strip, previous checkpoint column) or upwards (strip k - 1, this or the
         // previous checkpoint column): ask for the left boundaries of the three candidates now, they come from HBM
        const int tbc = (cl + k) / SW_TC;
        const typename SwCk<COMPACT>::T* cb = colckP + (int64_t)tbc * (W * R) + k * R;
        if (tbc > 1) asm volatile("prefetch.global.L2 [%0];" ::"l"(cb - (W * R)));
        if (k > 0) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(cb - R));
          if (tbc > 1) asm volatile("prefetch.global.L2 [%0];" ::"l"(cb - (W * R) - R));
          if (k > 1) asm volatile("prefetch.global.L2 [%0];" ::"l"(rowckP + (int64_t)(k - 2) * nT + (k - 3) + max(cl, 1)));
        }
      }
      if (k != haveK || cl != haveCl)
        sw_tile_recompute<R, COMPACT, false, W>(tk, s1, s2, prm, rowckP, colckP, nT, hi, k, cl, j, flags, 128, -1, -1, no1, no2, compLut);
      haveK = -1;
    }
    while (!done) {
      if (i < 1 || j < 1) {  // row 0 / column 0 pointers are STOP (:58-64)
        done = true;
        break;
      }
      if ((i - 1) / R != k || j <= cl) break;  // left the tile
      const int rr = (i - 1) - k * R;
      const uint32_t fw = flags[((j - cl - 1) * WPC + (rr >> 3)) * 128];
      if (mode == 0) {
        const uint32_t dir = (fw >> (2 * (rr & 7))) & 3u;
        if (dir == 0u) {
          done = true;
          break;
        }
        if (dir == 2u) {  // DIAGONAL
          uint8_t c1 = sw_s1_byte_lut(s1, tk.m, tk.rc, i - 1, compLut), c2 = s2[j - 1];
          out[len++] = c1;
          if (c1 == c2) identity++;
          i--, j--;
          continue;
        }
        mode = (dir == 3u) ? 1 : 2;
      }
      if (mode == 1) {  // UP: sizesOfVerticalGaps[k+j] lower-case read bases
        out[len++] = vg_lower(sw_s1_byte_lut(s1, tk.m, tk.rc, i - 1, compLut));
        const bool ext = ((fw >> (16 + (rr & 7))) & 1u) != 0;
        i--;
        if (!ext) mode = 0;
      } else {  // LEFT: sizesOfHorizontalGaps[k+j] gap characters
        out[len++] = '-';
        const bool ext = ((fw >> (24 + (rr & 7))) & 1u) != 0;
        j--;
        if (!ext) mode = 0;
      }
    }
  }
  if (valid) {
    SwAln al;
    al.score = fo.score > 0 ? fo.score : 0;
    al.start1 = i, al.end1 = end1, al.start2 = j, al.end2 = end2, al.identity = identity, al.length = len;
    alns[tk.out] = al;
  }
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

// Host side of seeding: index tables (vg_set_reference / vg_set_msa_index) and kernel launches.
#pragma once
#include <stdlib.h>

#include "ctx.cuh"
#include "seed.cuh"
#include "seed2.cuh"

struct SeedHostState {  // lives in vg_ctx via the DevBufs; plain helpers here
};

// blob = [ref padded to 16][poscode u16, padded with the id of no k-mer (4^K + exotic k-mers) to a multiple of 256 entries (the scan's stride, SD2_SCAN)]
static inline int sdh_ref_pad(int G) { return (G + 15) & ~15; }
static inline int sdh_code_pad(int G) { return std::max(256, (G + 255) & ~255) * 2; }

static inline int sdh_code2(uint8_t b) { return b == 'A' ? 0 : b == 'C' ? 1 : b == 'G' ? 2 : b == 'T' ? 3 : -1; }

// k-mer of K bytes -> (pure, code, packed key)
static inline bool sdh_pack(const uint8_t* p, int K, uint32_t* code, uint64_t* key) {
  bool pure = true;
  uint32_t c = 0;
  uint64_t k = 0;
  for (int t = 0; t < K; t++) {
    int v = sdh_code2(p[t]);
    pure &= v >= 0;
    c = (c << 2) | (uint32_t)(v & 3);
    k = (k << 8) | p[t];
  }
  *code = c;
  *key = k;
  return pure;
}

// StringIndexer.buildIndexPara position multiplicities (Q1): thread i indexes [i*size, (i+1)*size]
// inclusive (clipped to G-K), the last thread [(T-1)*size, G-K].
static void sdh_multiplicity_from_threads(int G, int K, int threadNum, std::vector<uint8_t>& mult) {
  mult.assign(std::max(G, 1), 0);
  if (threadNum < 1) threadNum = 1;
  const int size = G / threadNum;
  const int last = G - K;
  for (int i = 0; i < threadNum; i++) {
    int from = i * size;
    int to = (i == threadNum - 1) ? last : std::min((i + 1) * size, last);
    for (int p = from; p <= to; p++)
      if (mult[p] < 255) mult[p]++;
  }
}

static int seed_build_reference(vg_ctx* c, const uint8_t* keys, const int64_t* offsets, const int32_t* positions,
                                int n_keys, int index_thread_num) {
  const int G = c->G, K = c->seed.K;
  const uint8_t* ref = c->h_ref.data();
  std::vector<uint8_t> mult;
  const int nPos = G - K + 1;  // indexed positions 0..G-K (may be <= 0)
  if (keys) {
    mult.assign(std::max(G, 1), 0);
    for (int k = 0; k < n_keys; k++) {
      for (int64_t q = offsets[k]; q < offsets[k + 1]; q++) {
        const int p = positions[q];
        if (p < 0 || p > G - K || memcmp(ref + p, keys + (size_t)k * K, K) != 0)
          return vg_fail(c, VG_ERR_INDEX, "index entry (key %d, position %d) does not match the reference", k, p);
        if (q > offsets[k] && positions[q - 1] > p)
          return vg_fail(c, VG_ERR_INDEX, "index list of key %d is not ascending", k);
        if (mult[p] < 255) mult[p]++;
      }
    }
  } else {
    sdh_multiplicity_from_threads(G, K, index_thread_num, mult);
  }
  for (int p = 0; p < nPos; p++) {
    if (mult[p] == 0) return vg_fail(c, VG_ERR_INDEX, "reference position %d is missing from the k-mer index", p);
    if (mult[p] > 2)
      return vg_fail(c, VG_ERR_INDEX, "reference position %d is indexed %d times (supported: <= 2; the reference is "
                                      "shorter than the indexing thread count?)", p, (int)mult[p]);
    if (mult[p] == 2 && p + 1 < nPos && mult[p + 1] == 2)
      return vg_fail(c, VG_ERR_INDEX, "adjacent duplicated index positions %d,%d are not supported", p, p + 1);
  }
  // exotic k-mers (containing a byte outside ACGT): sorted distinct packed keys
  std::vector<uint64_t> exo;
  for (int p = 0; p < nPos; p++) {
    uint32_t code;
    uint64_t key;
    if (!sdh_pack(ref + p, K, &code, &key)) exo.push_back(key);
  }
  std::sort(exo.begin(), exo.end());
  exo.erase(std::unique(exo.begin(), exo.end()), exo.end());
  if ((1u << (2 * K)) + exo.size() >= SD_NONE16)
    return vg_fail(c, VG_ERR_INVALID, "too many distinct k-mers for the 15-bit id table (K=%d, %zu exotic)", K, exo.size());
  // blob = [ref padded to 16][poscode u16 padded to 16]
  const int refPad = sdh_ref_pad(G);
  const int codePad = sdh_code_pad(G);
  std::vector<uint8_t> blob((size_t)refPad + codePad + 16, 0);
  memcpy(blob.data(), ref, G);
  uint16_t* pc = (uint16_t*)(blob.data() + refPad);
  for (int p = 0; p < codePad / 2; p++) {
    uint16_t v = (uint16_t)((1u << (2 * K)) + exo.size());  // padding: the id of no k-mer (Seed2Args::emptyId)
    if (p < nPos) {
      uint32_t code;
      uint64_t key;
      if (sdh_pack(ref + p, K, &code, &key))
        v = (uint16_t)code;
      else
        v = (uint16_t)((1u << (2 * K)) + (std::lower_bound(exo.begin(), exo.end(), key) - exo.begin()));
      if (mult[p] == 2) v |= 0x8000u;
    }
    pc[p] = v;
  }
  {  // expected raw hits per read k-mer when the read is drawn from this reference: sum_id count(id)^2 / positions
    std::unordered_map<uint32_t, int> cnt;
    for (int p = 0; p < nPos; p++) cnt[pc[p] & 0x7FFFu]++;
    double ss = 0;
    for (auto& kv : cnt) ss += (double)kv.second * kv.second;
    c->seedHitsPerKmer = nPos > 0 ? ss / nPos : 0.0;
  }
  VG_CUDA(c, c->d_poscode.reserve(blob.size()));
  VG_CUDA(c, cudaMemcpy(c->d_poscode.p, blob.data(), blob.size(), cudaMemcpyHostToDevice));
  VG_CUDA(c, c->d_exoKeys.reserve(exo.size() * 8 + 8));
  if (!exo.empty()) VG_CUDA(c, cudaMemcpy(c->d_exoKeys.p, exo.data(), exo.size() * 8, cudaMemcpyHostToDevice));
  c->nExo = (int)exo.size();
  return VG_OK;
}

static int seed_build_msa(vg_ctx* c, const uint8_t* keys, const int64_t* offsets, const int32_t* columns, int n_keys) {
  const int K = c->seedMsa.K;
  std::vector<std::pair<uint64_t, int>> exo;  // packed key -> key index
  const uint32_t nPure = 1u << (2 * K);
  for (int k = 0; k < n_keys; k++) {
    uint32_t code;
    uint64_t key;
    if (!sdh_pack(keys + (size_t)k * K, K, &code, &key)) exo.push_back({key, k});
  }
  std::sort(exo.begin(), exo.end());
  if (nPure + exo.size() >= SD_NONE16)
    return vg_fail(c, VG_ERR_INVALID, "too many distinct k-mers for the 15-bit id table (K=%d)", K);
  const size_t nIds = nPure + exo.size();
  std::vector<int32_t> off(nIds + 1, 0);
  std::vector<int> keyOfId(nIds, -1);
  for (int k = 0; k < n_keys; k++) {
    uint32_t code;
    uint64_t key;
    if (sdh_pack(keys + (size_t)k * K, K, &code, &key)) keyOfId[code] = k;
  }
  for (size_t e = 0; e < exo.size(); e++) keyOfId[nPure + e] = exo[e].second;
  std::vector<int32_t> cols;
  cols.reserve((size_t)offsets[n_keys]);
  for (size_t id = 0; id < nIds; id++) {
    off[id] = (int32_t)cols.size();
    const int k = keyOfId[id];
    if (k >= 0)
      for (int64_t q = offsets[k]; q < offsets[k + 1]; q++) {
        if (columns[q] < 0 || columns[q] >= c->msaLength || (q > offsets[k] && columns[q - 1] >= columns[q]))
          return vg_fail(c, VG_ERR_INDEX, "MSA index list of key %d is not strictly ascending within [0, length)", k);
        cols.push_back(columns[q]);
      }
  }
  off[nIds] = (int32_t)cols.size();
  {  // expected raw hits of one read k-mer when the read is drawn from the aligned references: sum |list|^2 / sum |list|
    double ss = 0;
    for (size_t id = 0; id < nIds; id++) ss += (double)(off[id + 1] - off[id]) * (off[id + 1] - off[id]);
    c->msaHitsPerKmer = cols.empty() ? 0.0 : ss / (double)cols.size();
  }
  std::vector<uint64_t> exoKeys(exo.size());
  for (size_t e = 0; e < exo.size(); e++) exoKeys[e] = exo[e].first;
  VG_CUDA(c, c->d_msaOff.reserve(off.size() * 4));
  VG_CUDA(c, cudaMemcpy(c->d_msaOff.p, off.data(), off.size() * 4, cudaMemcpyHostToDevice));
  VG_CUDA(c, c->d_msaCols.reserve(cols.size() * 4 + 4));
  if (!cols.empty()) VG_CUDA(c, cudaMemcpy(c->d_msaCols.p, cols.data(), cols.size() * 4, cudaMemcpyHostToDevice));
  VG_CUDA(c, c->d_msaExoKeys.reserve(exoKeys.size() * 8 + 8));
  if (!exoKeys.empty()) VG_CUDA(c, cudaMemcpy(c->d_msaExoKeys.p, exoKeys.data(), exoKeys.size() * 8, cudaMemcpyHostToDevice));
  c->nMsaExo = (int)exoKeys.size();
  return VG_OK;
}

// Launches seed_kernel over nItems work items. Windows land in c->d_cand ([nItems][maxWin][3]) and
// c->d_ncand ([nItems]).
// Region mode (vg_map_to_regions): per-item references prepared by the caller on the device.
struct SeedRegions {
  const uint8_t* d_ref;
  const uint16_t* d_code;
  const int64_t* d_refOff;
  const int64_t* d_codeOff;
  const int32_t* d_len;
  const int32_t* d_itemRegion;
  const uint64_t* d_exo;
  int nExo;
  int maxLen;
};

static int seed_launch_v1(vg_ctx* c, bool msa, const uint8_t* d_bases, const int64_t* d_off1, const int32_t* d_len1,
                          const int64_t* d_off2, const int32_t* d_len2, const int64_t* d_pairIdx, int64_t nItems,
                          bool singleReads, int maxReadLen, int maxWin, int32_t* dbgHits, int32_t* dbgNHits,
                          const int32_t* d_itemList = nullptr, bool bestOnly = false, const SeedRegions* regs = nullptr) {
  const SeedConfig& sc = msa ? c->seedMsa : c->seed;
  const int K = sc.K;
  const int range = regs ? regs->maxLen : msa ? c->msaLength : c->G;
  if (K < 1 || K > 7) return vg_fail(c, VG_ERR_INVALID, "K=%d outside the supported range 1..7", K);
  if (range >= 65535) return vg_fail(c, VG_ERR_INVALID, "reference longer than 65534 is not supported by the 16-bit seeding tables");
  if (maxReadLen > SD_MAXL) return vg_fail(c, VG_ERR_INVALID, "read length %d exceeds the supported %d", maxReadLen, SD_MAXL);
  if (!bestOnly && maxReadLen >= K && maxReadLen >= (int)sc.cutoffs.size())
    return vg_fail(c, VG_ERR_INVALID, "cutoffs has %zu entries but a read has length %d (the Java would throw)",
                   sc.cutoffs.size(), maxReadLen);
  if (nItems > 0x7fffffff) return vg_fail(c, VG_ERR_INVALID, "too many reads in one call");
  SeedArgs A;
  memset(&A, 0, sizeof A);
  A.G = range;
  A.K = K;
  A.lowCoef = sc.lowCoef;
  A.highCoef = sc.highCoef;
  A.cutoffs = sc.d_cutoffs.as<int32_t>();
  A.nCutoffs = (int)sc.cutoffs.size();
  A.msa = msa ? 1 : 0;
  SeedGeom g;
  const int nContigs = (msa || regs) ? 1 : (int)c->contigEnds.size();
  g.capR = std::max(256, std::min(16384, range / K + nContigs + 16));
  g.nBuckets = (range >> 5) + 2;
  const int regAB = 2 * SD_HS * 4 + SD_MAXL * 2;  // table/hash + nxt
  const int regCG = SD_PENDING * 8 + 1024 + ((g.nBuckets * 2 + 15) & ~15) + 32 * SD_LPEND * 8 + 16;
  g.regA = (std::max(regAB, regCG) + 15) & ~15;
  g.warpBytes = (g.regA + (SD_WIN + 32) * 8 + SD_MAXL + 16 + 15) & ~15;
  g.refBytes = 0;
  if (regs) {
    A.regRef = regs->d_ref, A.regCode = regs->d_code, A.regRefOff = regs->d_refOff, A.regCodeOff = regs->d_codeOff;
    A.regLen = regs->d_len, A.itemRegion = regs->d_itemRegion;
    A.exoKeys = regs->d_exo;
    A.nExo = regs->nExo;
    A.contigEnds = nullptr;  // one contig = the whole region: never looked up
    A.nContigs = 1;
    A.direct = ((1 << (2 * K)) + regs->nExo <= 2 * SD_HS && !getenv("VG_SEED_HASH")) ? 1 : 0;
  } else if (!msa) {
    const int refPad = sdh_ref_pad(c->G), codePad = sdh_code_pad(c->G);
    A.blob = c->d_poscode.as<uint8_t>();
    A.blobRefBytes = refPad;
    A.blobBytes = refPad + codePad;
    A.ref = A.blob;
    A.poscode = (const uint16_t*)(A.blob + refPad);
    A.exoKeys = c->d_exoKeys.as<uint64_t>();
    A.nExo = c->nExo;
    A.contigEnds = c->d_contigEnds.as<int32_t>();
    A.nContigs = nContigs;
    A.direct = ((1 << (2 * K)) + c->nExo <= 2 * SD_HS && !getenv("VG_SEED_HASH")) ? 1 : 0;
    if (A.blobBytes + 4 * g.warpBytes <= 200 * 1024) g.refBytes = A.blobBytes;  // stage in shared memory via TMA
  } else {
    A.exoKeys = c->d_msaExoKeys.as<uint64_t>();
    A.nExo = c->nMsaExo;
    A.msaOff = c->d_msaOff.as<int32_t>();
    A.msaCols = c->d_msaCols.as<int32_t>();
    // single "contig" = the whole alignment: kept in d_misc[0]
    VG_CUDA(c, c->d_misc.reserve(64));
    int32_t len = c->msaLength;
    VG_CUDA(c, cudaMemcpyAsync(c->d_misc.p, &len, 4, cudaMemcpyHostToDevice, c->stream));
    A.contigEnds = c->d_misc.as<int32_t>();
    A.nContigs = 1;
  }
  const int smemMax = 220 * 1024;
  // 8 warps/SM measured best: the kernel is front-end (instruction fetch / branch) bound, more warps only
  // add instruction-cache pressure (profiles/seed_r1.md)
  int warps = std::min(16, (smemMax - g.refBytes) / g.warpBytes);
  if (const char* e = getenv("VG_SEED_WARPS")) warps = std::max(1, std::min(warps, atoi(e)));
  if (warps < 1) return vg_fail(c, VG_ERR_INVALID, "reference too large for the seeding kernel's shared memory");
  const size_t smem = (size_t)g.refBytes + (size_t)warps * g.warpBytes;
  auto kern = regs ? seed_kernel<true> : seed_kernel<false>;
  VG_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int perSm = 1;
  VG_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kern, warps * 32, smem));
  if (perSm < 1) return vg_fail(c, VG_ERR_INVALID, "seeding kernel does not fit on the device");
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((nItems + warps - 1) / warps, (int64_t)c->smCount * perSm));
  const int warpSlots = blocks * warps;
  // raw LongHits per read: 64k entries per resident warp, more when few warps run (the fallback list of the fast path
  // is where the low-complexity reads end up), within 2 GB of scratch
  A.capS = (int)std::max<int64_t>(1 << 16, std::min<int64_t>(1 << 22, ((int64_t)1 << 28) / warpSlots));
  if (const char* e = getenv("VG_SEED_CAPS")) A.capS = atoi(e);
  VG_CUDA(c, c->d_scratch.reserve((size_t)warpSlots * A.capS * 8));
  // candidate sort buffers; also the serial sweep's re-insertion queue (u64 x capC) and, in MSA mode, the chain starts
  // before the sweep (short k-mers on a long alignment give tens of thousands per read): within 2 GB
  A.capC = msa ? (int)std::max<int64_t>(std::max(g.capR, 16384), std::min<int64_t>(A.capS, ((int64_t)1 << 27) / warpSlots))
               : std::max(g.capR, 4096);
  VG_CUDA(c, c->d_seqOff.reserve((size_t)warpSlots * 2 * A.capC * 4 * 2));
  VG_CUDA(c, c->d_scratchR.reserve((size_t)warpSlots * 4 * g.capR * 2 + 64));
  A.scratchR = c->d_scratchR.as<uint16_t>();
  A.scratchS = c->d_scratch.as<uint64_t>();
  A.candKey = c->d_seqOff.as<uint32_t>();
  A.candVal = A.candKey + (size_t)warpSlots * 2 * A.capC;
  VG_CUDA(c, c->d_cand.reserve((size_t)std::max<int64_t>(nItems, 1) * maxWin * 3 * 4));
  VG_CUDA(c, c->d_ncand.reserve((size_t)std::max<int64_t>(nItems, 1) * 4));
  VG_CUDA(c, c->d_counter.reserve(16));
  VG_CUDA(c, cudaMemsetAsync(c->d_counter.p, 0, 16, c->stream));
  A.windows = c->d_cand.as<int32_t>();
  A.nWindows = c->d_ncand.as<int32_t>();
  A.maxWin = maxWin;
  A.counter = c->d_counter.as<int>();
  A.errFlag = c->d_counter.as<int>() + 1;
  A.bases = d_bases;
  A.off1 = d_off1;
  A.len1 = d_len1;
  A.off2 = d_off2;
  A.len2 = d_len2;
  A.pairIdx = d_pairIdx;
  A.nItems = nItems;
  A.itemList = d_itemList;  // when set: nItems entries, each the index of an item to (re)do
  A.singleReads = singleReads ? 1 : 0;
  A.bestOnly = bestOnly ? 1 : 0;
  A.dbgHits = dbgHits;
  A.dbgNHits = dbgNHits;
  A.geom = g;
  A.lockstep = getenv("VG_SEED_LOCKSTEP") ? 1 : 0;  // measured: no gain over the atomic work counter
  const bool prof = getenv("VG_SEED_PROF") != nullptr;
  if (prof) {
    VG_CUDA(c, c->d_rev.reserve(4096 * 3 * 4 + 512));
    A.prof = (unsigned long long*)(c->d_rev.as<uint8_t>() + 4096 * 3 * 4 + 64);
    VG_CUDA(c, cudaMemsetAsync(A.prof, 0, 128, c->stream));
  }
  if (nItems > 0) {
    kern<<<blocks, warps * 32, smem, c->stream>>>(A);
    VG_LAUNCH_CHECK(c);
  }
  int flags[2] = {0, 0};
  VG_CUDA(c, cudaMemcpyAsync(flags, c->d_counter.p, 8, cudaMemcpyDeviceToHost, c->stream));
  VG_CUDA(c, cudaStreamSynchronize(c->stream));
  if (prof) {
    unsigned long long ph[16];
    cudaMemcpy(ph, A.prof, 128, cudaMemcpyDeviceToHost);
    fprintf(stderr, "[seed prof] B1=%.0f B2=%.0f\n", ph[8] / (double)nItems, ph[9] / (double)nItems);
    fprintf(stderr, "[seed prof] items=%lld cycles/item: A=%.0f B=%.0f C=%.0f D=%.0f E=%.0f F=%.0f Gsort=%.0f (prune+rest in total) total=%.0f\n", (long long)nItems,
            ph[0] / (double)nItems, ph[1] / (double)nItems, ph[2] / (double)nItems, ph[3] / (double)nItems, ph[4] / (double)nItems,
            ph[5] / (double)nItems, ph[6] / (double)nItems, ph[7] / (double)nItems);
  }
  if (flags[1]) {
    const int f = flags[1];
    return vg_fail(c, VG_ERR_OVERFLOW,
                   "seeding work buffer overflow (flags 0x%x:%s%s%s%s%s); raise VG_SEED_CAPS or shorten reads", f,
                   (f & 1) ? " raw-hit buffer" : "", (f & 2) ? " LongHit buffer" : "", (f & 4) ? " re-insertion queue" : "",
                   (f & 8) ? " window list" : "", (f & 16) ? " read too long" : "");
  }
  return VG_OK;
}

// Seeding entry point. Single-reference mode with a direct-addressed k-mer table (K <= 5, few exotic k-mers) and a
// reference that fits in shared memory runs the two-kernel fast path (seed2.cuh) in sub-chunks of items; the items
// it hands back, and every other mode (MSA, K > 5), run the general kernel. Windows land in c->d_cand
// ([nItems][maxWin][3]) and c->d_ncand ([nItems]).
static int seed_launch(vg_ctx* c, bool msa, const uint8_t* d_bases, const int64_t* d_off1, const int32_t* d_len1,
                       const int64_t* d_off2, const int32_t* d_len2, const int64_t* d_pairIdx, int64_t nItems,
                       bool singleReads, int maxReadLen, int maxWin, int32_t* dbgHits, int32_t* dbgNHits, bool bestOnly = false,
                       const SeedRegions* regs = nullptr) {
  c->seedFallbacks = 0;
  const SeedConfig& sc = msa ? c->seedMsa : c->seed;
  const int K = sc.K, G = regs ? regs->maxLen : msa ? c->msaLength : c->G;
  const int nExoRef = regs ? regs->nExo : c->nExo;
  const int smemMax = 227 * 1024;
  bool fast = !getenv("VG_SEED_V1") && !getenv("VG_SEED_HASH") && !getenv("VG_SEED_PROF") && K >= 1 && K <= 7 && G >= K &&
              G + maxReadLen < 65535 && maxReadLen <= SD_MAXL && maxReadLen >= K && nItems > 0 && nItems <= 0x7fffffff &&
              (bestOnly || maxReadLen < (int)sc.cutoffs.size());
  Seed2Args A;
  memset(&A, 0, sizeof A);
  int genWarps = 0, winWarps = 0;
  if (fast) {
    A.rdBytes = (maxReadLen + 32 + 15) & ~15;  // 4 bytes in front (sentinel), >= 12 behind the read for the 4-byte compares
    if (!msa) {
      A.emptyId = (1 << (2 * K)) + nExoRef;
      A.tabBytes = ((A.emptyId + 1) * 2 + 15) & ~15;
      A.nxtBytes = (maxReadLen * 2 + 15) & ~15;
      A.blobRefBytes = regs ? 0 : sdh_ref_pad(G);
      A.blobBytes = regs ? 0 : A.blobRefBytes + sdh_code_pad(G);
      // raw-hit queue: 512 entries, up to 1024 when 32 warps per SM still fit (a trip of the scan must fit in it)
      const int perWarp = ((smemMax - 64 - A.blobBytes) / 32) & ~15;
      A.qcap = std::max(512, std::min(1024, ((perWarp - A.tabBytes - A.nxtBytes - A.rdBytes) / 4 - 4) & ~63));
      A.genWarpBytes = A.tabBytes + A.nxtBytes + (A.qcap + 4) * 4 + A.rdBytes;
      genWarps = std::min(32, (smemMax - 64 - A.blobBytes) / A.genWarpBytes);
      if (A.emptyId >= SD_NONE16 || A.emptyId > 8191) fast = false;
    } else {
      A.idsBytes = (maxReadLen * 2 + 15) & ~15;
      A.msaWarpBytes = A.idsBytes + 2048 + A.rdBytes;
      genWarps = std::min(32, smemMax / A.msaWarpBytes);
      if (const char* e = getenv("VG_SEED_MSAWARPS")) genWarps = std::max(8, std::min(std::min(32, smemMax / A.msaWarpBytes), atoi(e)));
    }
    const int nContigs = (msa || regs) ? 1 : (int)c->contigEnds.size();
    A.capR = std::max(256, std::min(16384, G / K + nContigs + 16));
    A.nBuckets = (G >> 5) + 2;
    const int bucketBytes = (A.nBuckets * 2 + 15) & ~15;
    const int cBytes = (SD_WIN + 32) * 8;
    const int winMax = smemMax - 512;  // the window kernel keeps the contig ends in 256 bytes of static shared memory
    int budgetWarps = 32;
    if (const char* e = getenv("VG_SEED_WINWARPS")) budgetWarps = std::max(8, std::min(32, atoi(e)));
    const int budget = (winMax / budgetWarps) & ~15;  // per-warp bytes at 32 warps per SM
    A.capRs = std::max(0, std::min(A.capR, (budget - bucketBytes - 16) / 5)) & ~7;
    A.winWarpBytes = (std::max(std::max(cBytes, 1024), A.capRs * 5 + 16 + bucketBytes) + 15) & ~15;
    winWarps = std::min(32, winMax / A.winWarpBytes);
    if (genWarps < 8 || winWarps < 8 || nContigs > 64) fast = false;
  }
  if (!fast)
    return seed_launch_v1(c, msa, d_bases, d_off1, d_len1, d_off2, d_len2, d_pairIdx, nItems, singleReads, maxReadLen, maxWin,
                          dbgHits, dbgNHits, nullptr, bestOnly, regs);
  if (const char* e = getenv("VG_SEED_WARPS")) {
    genWarps = std::max(1, std::min(genWarps, atoi(e)));
    winWarps = std::max(1, std::min(winWarps, atoi(e)));
  }
  double hitsPerKmer = c->seedHitsPerKmer;
  if (regs) {
    A.regions = 1;
    A.regRef = regs->d_ref, A.regCode = regs->d_code, A.regRefOff = regs->d_refOff, A.regCodeOff = regs->d_codeOff;
    A.regLen = regs->d_len, A.itemRegion = regs->d_itemRegion;
    A.exoKeys = regs->d_exo;
    A.nExo = regs->nExo;
    A.contigEnds = nullptr;
    A.nContigs = 1;
    hitsPerKmer = 4.0 * (double)G / (double)(1u << (2 * K)) + 1.0;  // no spectrum at hand: 4 x the uniform expectation
  } else if (!msa) {
    A.blob = c->d_poscode.as<uint8_t>();
    A.exoKeys = c->d_exoKeys.as<uint64_t>();
    A.nExo = c->nExo;
    A.contigEnds = c->d_contigEnds.as<int32_t>();
    A.nContigs = (int)c->contigEnds.size();
  } else {
    A.exoKeys = c->d_msaExoKeys.as<uint64_t>();
    A.nExo = c->nMsaExo;
    A.msaOff = c->d_msaOff.as<int32_t>();
    A.msaCols = c->d_msaCols.as<int32_t>();
    A.contigEnds = nullptr;  // one "contig" = the whole alignment: never looked up
    A.nContigs = 1;
    hitsPerKmer = c->msaHitsPerKmer;
  }
  A.G = G;
  A.K = K;
  A.lowCoef = sc.lowCoef;
  A.highCoef = sc.highCoef;
  A.cutoffs = sc.d_cutoffs.as<int32_t>();
  A.nCutoffs = (int)sc.cutoffs.size();
  A.bases = d_bases;
  A.off1 = d_off1, A.len1 = d_len1, A.off2 = d_off2, A.len2 = d_len2;
  A.pairIdx = d_pairIdx;
  A.singleReads = singleReads ? 1 : 0;
  A.bestOnly = bestOnly ? 1 : 0;
  A.maxWin = maxWin;
  A.dbgHits = dbgHits;
  A.dbgNHits = dbgNHits;
  // per-item slot: 1.25 x the expected number of raw LongHits of a read drawn from the reference (0.75 of the raw k-mer
  // hits start a chain) + slack; a read that needs more is redone by the general kernel
  const double nkMax = std::max(1, maxReadLen - K + 1);
  {
    const double expect = (msa ? 1.0 : 0.75) * nkMax * hitsPerKmer;
    A.capS = (int)std::min<double>(65536.0, std::max(1024.0, 1.25 * expect + 1024.0));
    A.capS = (A.capS + 255) & ~255;
    if (const char* e = getenv("VG_SEED_CAPS2")) A.capS = std::max(64, atoi(e));
  }
  size_t sBudget = 0;
  if (const char* e = getenv("VG_SEED_SUBMB")) sBudget = (size_t)std::max(1, atoi(e)) << 20;
  else {
    if (!c->budgetSeed) c->budgetSeed = vg_mem_budget(8ull << 30, 0.3, c->d_seedS.cap);
    sBudget = c->budgetSeed;
  }
  // packed 32-bit stream / sweep-window elements (genomePos 14 bits, readPos 9, len 9) when the coordinates fit
  const bool p32 = G < 16384 && maxReadLen <= 512 && !getenv("VG_SEED_WIDE");
  const size_t elemBytes = p32 ? 4 : 8;
  A.capC = A.capR;
  A.key24 = (G < 16384 && !getenv("VG_SEED_KEY32")) ? 1 : 0;
  const int64_t subItems = std::max<int64_t>(4096, std::min<int64_t>(nItems, (int64_t)(sBudget / ((size_t)A.capS * elemBytes))));
  const size_t genSmem = msa ? (size_t)genWarps * A.msaWarpBytes : (size_t)A.blobBytes + (size_t)genWarps * A.genWarpBytes;
  const size_t winSmem = (size_t)winWarps * A.winWarpBytes;
  if (msa) VG_CUDA(c, cudaFuncSetAttribute(seed_gen_msa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)genSmem));
  else if (regs && p32) VG_CUDA(c, cudaFuncSetAttribute(seed_gen_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)genSmem));
  else if (regs) VG_CUDA(c, cudaFuncSetAttribute(seed_gen_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)genSmem));
  else if (p32) VG_CUDA(c, cudaFuncSetAttribute(seed_gen_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)genSmem));
  else VG_CUDA(c, cudaFuncSetAttribute(seed_gen_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)genSmem));
  if (p32) VG_CUDA(c, cudaFuncSetAttribute(seed_win_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)winSmem));
  else VG_CUDA(c, cudaFuncSetAttribute(seed_win_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)winSmem));
  const int64_t sub0 = std::min(subItems, nItems);
  const int genBlocks = (int)std::max<int64_t>(1, std::min<int64_t>((sub0 + genWarps - 1) / genWarps, c->smCount));
  const int winBlocks = (int)std::max<int64_t>(1, std::min<int64_t>((sub0 + winWarps - 1) / winWarps, c->smCount));
  const int winSlots = winBlocks * winWarps;
  VG_CUDA(c, c->d_seedS.reserve((size_t)sub0 * A.capS * elemBytes));
  VG_CUDA(c, c->d_seedNS.reserve((size_t)sub0 * 4));
  VG_CUDA(c, c->d_seedFb.reserve((size_t)nItems * 4));
  VG_CUDA(c, c->d_seqOff.reserve((size_t)winSlots * 2 * A.capC * 4 * 2));
  VG_CUDA(c, c->d_scratchR.reserve((size_t)winSlots * 3 * (A.capR + 8) * 2 + 64));
  VG_CUDA(c, c->d_cand.reserve((size_t)nItems * maxWin * 3 * 4));
  VG_CUDA(c, c->d_ncand.reserve((size_t)nItems * 4));
  VG_CUDA(c, c->d_counter.reserve(512));
  if (msa) {  // raw-hit sort scratch of the MSA stream kernel: 4 arrays per resident warp, within 4 GB
    const int genSlots = genBlocks * genWarps;
    const double want = 2.0 * nkMax * hitsPerKmer + 4096.0;
    const int64_t fit = (int64_t)(((size_t)4 << 30) / ((size_t)genSlots * 16));
    A.capRaw = (int)std::max<int64_t>(A.capS, std::min<int64_t>(std::min<int64_t>((int64_t)want, 1 << 18), fit)) & ~255;
    if (const char* e = getenv("VG_SEED_CAPRAW")) A.capRaw = std::max(256, atoi(e)) & ~255;
    VG_CUDA(c, c->d_scratch.reserve((size_t)genSlots * 4 * A.capRaw * 4));
    A.rawBuf = c->d_scratch.as<uint32_t>();
    A.msaPacked = p32 ? 1 : 0;  // the stream's element format: must be the window kernel's
  }
  A.S = c->d_seedS.as<uint64_t>();
  A.nS = c->d_seedNS.as<int32_t>();
  A.fallback = c->d_seedFb.as<int32_t>();
  A.candKey = c->d_seqOff.as<uint32_t>();
  A.candVal = A.candKey + (size_t)winSlots * 2 * A.capC;
  A.scratchR = c->d_scratchR.as<uint16_t>();
  A.windows = c->d_cand.as<int32_t>();
  A.nWindows = c->d_ncand.as<int32_t>();
  A.counters = c->d_counter.as<int>() + 4;
  VG_CUDA(c, cudaMemsetAsync(A.counters, 0, 64, c->stream));
  const bool dbg = getenv("VG_SEED_DEBUG") != nullptr;
  if (dbg && atoi(getenv("VG_SEED_DEBUG")) >= 2) {
    A.prof = (unsigned long long*)(c->d_counter.as<int>() + 32);
    VG_CUDA(c, cudaMemsetAsync(A.prof, 0, 128, c->stream));
  }
  const int nSub = (int)((nItems + subItems - 1) / subItems);
  while ((int)c->seedEv.size() < 3 * nSub + 2) {
    cudaEvent_t e;
    VG_CUDA(c, cudaEventCreate(&e));
    c->seedEv.push_back(e);
  }
  int sub = 0;
  for (int64_t s = 0; s < nItems; s += subItems, sub++) {
    A.itemBase = s;
    A.nItems = (int)std::min(subItems, nItems - s);
    VG_CUDA(c, cudaMemsetAsync(A.counters, 0, 8, c->stream));
    VG_CUDA(c, cudaEventRecord(c->seedEv[3 * sub], c->stream));
    if (msa) seed_gen_msa_kernel<<<genBlocks, genWarps * 32, genSmem, c->stream>>>(A);
    else if (regs && p32) seed_gen_kernel<true, true><<<genBlocks, genWarps * 32, genSmem, c->stream>>>(A);
    else if (regs) seed_gen_kernel<true, false><<<genBlocks, genWarps * 32, genSmem, c->stream>>>(A);
    else if (p32) seed_gen_kernel<false, true><<<genBlocks, genWarps * 32, genSmem, c->stream>>>(A);
    else seed_gen_kernel<false, false><<<genBlocks, genWarps * 32, genSmem, c->stream>>>(A);
    VG_LAUNCH_CHECK(c);
    VG_CUDA(c, cudaEventRecord(c->seedEv[3 * sub + 1], c->stream));
    if (p32) seed_win_kernel<true><<<winBlocks, winWarps * 32, winSmem, c->stream>>>(A);
    else seed_win_kernel<false><<<winBlocks, winWarps * 32, winSmem, c->stream>>>(A);
    VG_LAUNCH_CHECK(c);
    VG_CUDA(c, cudaEventRecord(c->seedEv[3 * sub + 2], c->stream));
  }
  int flags[16] = {0};
  VG_CUDA(c, cudaMemcpyAsync(flags, A.counters, 64, cudaMemcpyDeviceToHost, c->stream));
  VG_CUDA(c, cudaStreamSynchronize(c->stream));
  float genMs = 0, winMs = 0;
  for (int i = 0; i < nSub; i++) {
    float a = 0, b = 0;
    cudaEventElapsedTime(&a, c->seedEv[3 * i], c->seedEv[3 * i + 1]);
    cudaEventElapsedTime(&b, c->seedEv[3 * i + 1], c->seedEv[3 * i + 2]);
    genMs += a, winMs += b;
  }
  unsigned long long sums[3];
  memcpy(sums, flags + 10, 24);
  c->seedStats[0] += genMs, c->seedStats[1] += winMs;
  c->seedStats[3] += (double)sums[0], c->seedStats[4] += (double)sums[1];
  c->seedStats[5] += flags[3], c->seedStats[6] += (double)sums[2], c->seedStats[7] += (double)nItems;
  if (A.prof) {
    unsigned long long ph[16];
    cudaMemcpy(ph, A.prof, 128, cudaMemcpyDeviceToHost);
    if (msa)
      fprintf(stderr, "[seed] MSA stream kernel cycles/item: list %.0f, sort by diagonal %.0f, chain starts %.0f, LongHits %.0f, sort by column + "
              "write %.0f; raw hits/item %.0f\n", ph[8] / (double)nItems, ph[9] / (double)nItems, ph[10] / (double)nItems,
              ph[11] / (double)nItems, ph[12] / (double)nItems, ph[13] / (double)nItems);
    fprintf(stderr, "[seed] window kernel cycles/item: item+prefetch %.0f, C %.0f, D+E %.0f, F %.0f, sort %.0f, prune %.0f\n", ph[0] / (double)nItems,
            ph[1] / (double)nItems, ph[2] / (double)nItems, ph[3] / (double)nItems, ph[4] / (double)nItems, ph[5] / (double)nItems);
  }
  if (dbg)
    fprintf(stderr, "[seed] items=%lld in %d sub-chunks of %lld: gen %.2f ms, win %.2f ms; capS=%d genWarps=%d winWarps=%d capRs=%d; LongHits "
            "%.0f -> %.0f per item; fallbacks=%d (count/queue %d, slot %d, long segment %d, re-insertion %d, capacity %d)\n",
            (long long)nItems, nSub, (long long)subItems, genMs, winMs, A.capS, genWarps, winWarps, A.capRs, sums[0] / (double)nItems,
            sums[1] / (double)nItems, flags[3], flags[4], flags[5], flags[6], flags[7], flags[8]);
  if (flags[2]) {
    const int f = flags[2];
    return vg_fail(c, VG_ERR_OVERFLOW, "seeding work buffer overflow (flags 0x%x:%s%s); shorten reads", f,
                   (f & 8) ? " window list" : "", (f & 16) ? " read too long" : "");
  }
  c->seedFallbacks = flags[3];
  if (flags[3] > 0) {
    VG_CUDA(c, cudaEventRecord(c->seedEv[0], c->stream));
    VG_TRY(seed_launch_v1(c, msa, d_bases, d_off1, d_len1, d_off2, d_len2, d_pairIdx, flags[3], singleReads, maxReadLen, maxWin,
                          dbgHits, dbgNHits, c->d_seedFb.as<int32_t>(), bestOnly, regs));
    VG_CUDA(c, cudaEventRecord(c->seedEv[1], c->stream));
    VG_CUDA(c, cudaEventSynchronize(c->seedEv[1]));
    float f = 0;
    cudaEventElapsedTime(&f, c->seedEv[0], c->seedEv[1]);
    c->seedStats[2] += f;
  }
  return VG_OK;
}

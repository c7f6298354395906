// Multi-GPU behind the C ABI: read pairs are sharded over the devices in contiguous blocks (the
// reference's task order is kept on gather), reference / index / cutoffs are replicated, and the only exchange is one
// ncclAllReduce(int32, sum) of the pileup counts — the multi-GPU replacement of the serial sum of the per-thread
// matrices at ConsensusBuilderSimple.java:138-145 (exact, because integer).
//
// Two ways in, both without any host-language collective:
//   vg_comm_unique_id + vg_comm_init_rank   one process per GPU (torchrun-style hosts): the 128-byte id travels by
//                                           whatever channel the host has; vg_pileup_allreduce then runs on the
//                                           context's own stream.
//   vg_multi_*                              one process, N devices (one JVM): N contexts, one host thread per device
//                                           per call, ncclCommInitAll. This is the `vg_ctx_create(params, device_ids[],
//                                           n_dev)` shape of SURVEY §8b.
// NCCL is loaded with dlopen at first use, so that the single-GPU library has no link-time dependency on it.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <thread>

#include "ctx.cuh"

struct VgNccl {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  char err[256] = {0};
};

static VgNccl* vg_nccl() {
  static VgNccl N;
  static bool tried = false;
  if (tried) return N.h ? &N : nullptr;
  tried = true;
  // a libnccl.so.2 that the process already holds (e.g. the one bundled with torch) is reused
  const char* names[] = {getenv("VG_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    if (!nm) continue;
    N.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (N.h) break;
  }
  if (!N.h) {
    snprintf(N.err, sizeof N.err, "libnccl.so.2 not found (%s)", dlerror());
    return nullptr;
  }
#define VG_NCCL_SYM(field, name)                                       \
  *(void**)(&N.field) = dlsym(N.h, name);                              \
  if (!N.field) {                                                      \
    snprintf(N.err, sizeof N.err, "NCCL symbol %s is missing", name);  \
    N.h = nullptr;                                                     \
    return nullptr;                                                    \
  }
  VG_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
  VG_NCCL_SYM(CommInitRank, "ncclCommInitRank")
  VG_NCCL_SYM(CommInitAll, "ncclCommInitAll")
  VG_NCCL_SYM(CommDestroy, "ncclCommDestroy")
  VG_NCCL_SYM(AllReduce, "ncclAllReduce")
  VG_NCCL_SYM(GroupStart, "ncclGroupStart")
  VG_NCCL_SYM(GroupEnd, "ncclGroupEnd")
  VG_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef VG_NCCL_SYM
  return &N;
}

#define VG_NCCL(c, N, call)                                                                                        \
  do {                                                                                                             \
    ncclResult_t r__ = (call);                                                                                     \
    if (r__ != ncclSuccess)                                                                                        \
      return vg_fail((c), VG_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, (N)->GetErrorString(r__)); \
  } while (0)

static void vg_comm_release(vg_ctx* c) {
  if (c && c->comm) {
    VgNccl* N = vg_nccl();
    if (N) {
      cudaSetDevice(c->device);
      N->CommDestroy((ncclComm_t)c->comm);
    }
    c->comm = nullptr;
    c->commSize = 0;
    c->commRank = -1;
  }
}

// One vg_multi = N contexts on N devices of one process.
struct vg_multi {
  std::vector<vg_ctx*> ctx;
  std::vector<int64_t> lo;            // resident pairs of device d: [lo[d], lo[d + 1])
  std::vector<std::vector<int64_t>> outPos;  // last mapping: position of device d's k-th result in the caller's order
  bool contiguous = true;             // last mapping covered all pairs in order (results of device d sit at lo[d] ...)
  int64_t nPairs = 0, nResults = 0, seq1Bytes = 0;
  std::vector<int64_t> nRes, poolBytes;  // per device, last mapping
  char err[1024] = {0};
  int64_t required = 0;
};

static int vgm_fail(vg_multi* m, int code, const char* fmt, ...) {
  if (m) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(m->err, sizeof m->err, fmt, ap);
    va_end(ap);
  }
  return code;
}

// runs fn(d) on one host thread per device; the first failure (lowest device) is reported
template <class F>
static int vgm_each(vg_multi* m, F fn) {
  const int n = (int)m->ctx.size();
  std::vector<int> rc(n, VG_OK);
  if (n == 1) {
    rc[0] = fn(0);
  } else {
    std::vector<std::thread> th;
    for (int d = 0; d < n; d++) th.emplace_back([&, d] { rc[d] = fn(d); });
    for (auto& t : th) t.join();
  }
  for (int d = 0; d < n; d++)
    if (rc[d] != VG_OK) return vgm_fail(m, rc[d], "device %d: %s", m->ctx[d]->device, m->ctx[d]->err);
  return VG_OK;
}

extern "C" {

int vg_comm_unique_id(uint8_t* id128) {
  if (!id128) return VG_ERR_INVALID;
  VgNccl* N = vg_nccl();
  if (!N) return VG_ERR_CUDA;
  static_assert(sizeof(ncclUniqueId) == VG_COMM_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId u;
  if (N->GetUniqueId(&u) != ncclSuccess) return VG_ERR_CUDA;
  memcpy(id128, &u, sizeof u);
  return VG_OK;
}

int vg_comm_init_rank(vg_ctx* c, const uint8_t* id128, int rank, int n_ranks) {
  if (!c || !id128 || n_ranks < 1 || rank < 0 || rank >= n_ranks) return vg_fail(c, VG_ERR_INVALID, "bad communicator arguments");
  VgNccl* N = vg_nccl();
  if (!N) return vg_fail(c, VG_ERR_CUDA, "NCCL is not available: %s", vg_nccl() ? "" : "libnccl.so.2 could not be loaded");
  VG_CUDA(c, cudaSetDevice(c->device));
  vg_comm_release(c);
  ncclUniqueId u;
  memcpy(&u, id128, sizeof u);
  ncclComm_t comm;
  VG_NCCL(c, N, N->CommInitRank(&comm, n_ranks, u, rank));
  c->comm = comm;
  c->commRank = rank;
  c->commSize = n_ranks;
  return VG_OK;
}

int vg_comm_init_all(vg_ctx** ctxs, int n) {
  if (!ctxs || n < 1) return VG_ERR_INVALID;
  for (int i = 0; i < n; i++)
    if (!ctxs[i]) return VG_ERR_INVALID;
  vg_ctx* c0 = ctxs[0];
  VgNccl* N = vg_nccl();
  if (!N) return vg_fail(c0, VG_ERR_CUDA, "NCCL is not available (libnccl.so.2 could not be loaded)");
  std::vector<int> devs(n);
  for (int i = 0; i < n; i++) {
    devs[i] = ctxs[i]->device;
    for (int j = 0; j < i; j++)
      if (devs[j] == devs[i]) return vg_fail(c0, VG_ERR_INVALID, "device %d is named twice", devs[i]);
    vg_comm_release(ctxs[i]);
  }
  std::vector<ncclComm_t> comms(n);
  VG_NCCL(c0, N, N->CommInitAll(comms.data(), n, devs.data()));
  for (int i = 0; i < n; i++) {
    ctxs[i]->comm = comms[i];
    ctxs[i]->commRank = i;
    ctxs[i]->commSize = n;
  }
  return VG_OK;
}

void vg_comm_destroy(vg_ctx* c) { vg_comm_release(c); }

int vg_comm_size(const vg_ctx* c) { return c && c->comm ? c->commSize : 0; }

// The exchange step of getConsensusSimple across GPUs (ConsensusBuilderSimple.java:138-145): in-place integer sum of
// the counts of all ranks, on the context's stream. Every rank of the communicator must call it (from its own thread or
// process). Without a communicator it is a no-op.
int vg_pileup_allreduce(vg_ctx* c) {
  if (!c) return VG_ERR_INVALID;
  if (!c->haveCounts) return vg_fail(c, VG_ERR_STATE, "vg_pileup has not been called");
  if (!c->comm || c->commSize <= 1) return VG_OK;
  VgNccl* N = vg_nccl();
  if (!N) return vg_fail(c, VG_ERR_CUDA, "NCCL is not available");
  VG_CUDA(c, cudaSetDevice(c->device));
  const size_t cnt = (size_t)c->G * 5;
  if (cnt == 0) return VG_OK;
  VG_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
  VG_NCCL(c, N, N->AllReduce(c->d_counts.p, c->d_counts.p, cnt, ncclInt32, ncclSum, (ncclComm_t)c->comm, c->stream));
  c->launches++;
  VG_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
  VG_CUDA(c, cudaStreamSynchronize(c->stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
  c->allreduceMs = ms;
  return VG_OK;
}

int vg_last_allreduce_ms(vg_ctx* c, double* ms) {
  if (!c || !ms) return VG_ERR_INVALID;
  *ms = c->allreduceMs;
  return VG_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// vg_multi: N devices in one process
// ---------------------------------------------------------------------------------------------------------------------
int vg_multi_create(const vg_params* p, const int32_t* device_ids, int32_t n_dev, vg_multi** out) {
  if (!out) return VG_ERR_INVALID;
  *out = nullptr;
  if (!p || !device_ids || n_dev < 1 || n_dev > 64) return VG_ERR_INVALID;
  vg_multi* m = new vg_multi();
  *out = m;  // returned even on failure so that the caller can read vg_multi_last_error
  for (int d = 0; d < n_dev; d++) {
    vg_ctx* c = nullptr;
    const int rc = vg_ctx_create(p, device_ids[d], &c);
    if (c) m->ctx.push_back(c);
    if (rc != VG_OK) return vgm_fail(m, rc, "device %d: %s", device_ids[d], c ? c->err : "allocation failed");
  }
  if (n_dev > 1) {
    const int rc = vg_comm_init_all(m->ctx.data(), n_dev);
    if (rc != VG_OK) return vgm_fail(m, rc, "%s", m->ctx[0]->err);
  }
  m->lo.assign(n_dev + 1, 0);
  m->nRes.assign(n_dev, 0);
  m->poolBytes.assign(n_dev, 0);
  m->outPos.resize(n_dev);
  return VG_OK;
}

void vg_multi_destroy(vg_multi* m) {
  if (!m) return;
  for (vg_ctx* c : m->ctx) vg_ctx_destroy(c);
  delete m;
}

const char* vg_multi_last_error(const vg_multi* m) { return m ? m->err : "null multi-context"; }
int64_t vg_multi_last_required(const vg_multi* m) { return m ? m->required : 0; }
int32_t vg_multi_size(const vg_multi* m) { return m ? (int32_t)m->ctx.size() : 0; }
vg_ctx* vg_multi_ctx(vg_multi* m, int32_t i) { return (m && i >= 0 && i < (int)m->ctx.size()) ? m->ctx[i] : nullptr; }

int vg_multi_set_reference(vg_multi* m, const uint8_t* seq, int32_t G, const int32_t* contig_ends, int32_t n_contigs,
                           const int32_t* fragment_ends, int32_t n_fragments, int32_t is_fragmented, const uint8_t* keys,
                           const int64_t* offsets, const int32_t* positions, int32_t n_keys, int32_t index_thread_num) {
  if (!m) return VG_ERR_INVALID;
  return vgm_each(m, [&](int d) {
    return vg_set_reference(m->ctx[d], seq, G, contig_ends, n_contigs, fragment_ends, n_fragments, is_fragmented, keys, offsets,
                            positions, n_keys, index_thread_num);
  });
}

int vg_multi_set_msa_index(vg_multi* m, int32_t K, float low_coef, float high_coef, const int32_t* cutoffs, int32_t n_cutoffs,
                           const uint8_t* keys, const int64_t* offsets, const int32_t* columns, int32_t n_keys,
                           int32_t msa_length) {
  if (!m) return VG_ERR_INVALID;
  return vgm_each(m, [&](int d) {
    return vg_set_msa_index(m->ctx[d], K, low_coef, high_coef, cutoffs, n_cutoffs, keys, offsets, columns, n_keys, msa_length);
  });
}

// Pairs are dealt out in contiguous blocks that differ by at most one pair; every device receives only the span of the
// byte pool that its own pairs touch.
int vg_multi_upload_reads(vg_multi* m, const uint8_t* bases, int64_t n_bases, const int64_t* off1, const int32_t* len1,
                          const int64_t* off2, const int32_t* len2, int64_t n_pairs) {
  if (!m) return VG_ERR_INVALID;
  if (n_pairs < 0 || n_bases < 0 || (n_pairs > 0 && (!bases || !off1 || !len1 || !off2 || !len2)))
    return vgm_fail(m, VG_ERR_INVALID, "bad read arrays");
  const int n = (int)m->ctx.size();
  const int64_t base = n_pairs / n, rem = n_pairs % n;
  for (int d = 0; d <= n; d++) m->lo[d] = d * base + std::min<int64_t>(d, rem);
  m->nPairs = n_pairs;
  m->nResults = 0;
  return vgm_each(m, [&](int d) {
    const int64_t a = m->lo[d], b = m->lo[d + 1], np = b - a;
    vg_ctx* c = m->ctx[d];
    int64_t lo = n_bases, hi = 0;
    for (int64_t i = a; i < b; i++) {
      if (len1[i] < 0 || len2[i] < 0 || off1[i] < 0 || off2[i] < 0 || off1[i] + len1[i] > n_bases || off2[i] + len2[i] > n_bases)
        return vg_fail(c, VG_ERR_INVALID, "read %lld lies outside the byte pool", (long long)i);
      lo = std::min(lo, std::min(off1[i], off2[i]));
      hi = std::max(hi, std::max(off1[i] + len1[i], off2[i] + len2[i]));
    }
    if (np == 0) lo = hi = 0;
    std::vector<int64_t> o1(np), o2(np);
    for (int64_t i = 0; i < np; i++) o1[i] = off1[a + i] - lo, o2[i] = off2[a + i] - lo;
    return vg_upload_reads(c, bases + lo, hi - lo, o1.data(), len1 + a, o2.data(), len2 + a, np);
  });
}

// The same with the packed transfer form (vg_upload_reads_packed): every device receives the bytes of the packed pool
// that hold its span (from the 4-base boundary below its first base) and the exceptions that fall into it.
int vg_multi_upload_reads_packed(vg_multi* m, const uint8_t* packed, int64_t n_bases, const int64_t* exc_pos,
                                 const uint8_t* exc_byte, int64_t n_exc, const int64_t* off1, const int32_t* len1,
                                 const int64_t* off2, const int32_t* len2, int64_t n_pairs) {
  if (!m) return VG_ERR_INVALID;
  if (n_pairs < 0 || n_bases < 0 || n_exc < 0 || (n_pairs > 0 && (!packed || !off1 || !len1 || !off2 || !len2)) ||
      (n_exc > 0 && (!exc_pos || !exc_byte)))
    return vgm_fail(m, VG_ERR_INVALID, "bad read arrays");
  const int n = (int)m->ctx.size();
  const int64_t base = n_pairs / n, rem = n_pairs % n;
  for (int d = 0; d <= n; d++) m->lo[d] = d * base + std::min<int64_t>(d, rem);
  m->nPairs = n_pairs;
  m->nResults = 0;
  return vgm_each(m, [&](int d) {
    const int64_t a = m->lo[d], b = m->lo[d + 1], np = b - a;
    vg_ctx* c = m->ctx[d];
    int64_t lo = n_bases, hi = 0;
    for (int64_t i = a; i < b; i++) {
      if (len1[i] < 0 || len2[i] < 0 || off1[i] < 0 || off2[i] < 0 || off1[i] + len1[i] > n_bases || off2[i] + len2[i] > n_bases)
        return vg_fail(c, VG_ERR_INVALID, "read %lld lies outside the byte pool", (long long)i);
      lo = std::min(lo, std::min(off1[i], off2[i]));
      hi = std::max(hi, std::max(off1[i] + len1[i], off2[i] + len2[i]));
    }
    if (np == 0) lo = hi = 0;
    lo &= ~(int64_t)3;  // a whole packed byte
    std::vector<int64_t> o1(np), o2(np), ep;
    std::vector<uint8_t> eb;
    for (int64_t i = 0; i < np; i++) o1[i] = off1[a + i] - lo, o2[i] = off2[a + i] - lo;
    for (int64_t k = 0; k < n_exc; k++)
      if (exc_pos[k] >= lo && exc_pos[k] < hi) ep.push_back(exc_pos[k] - lo), eb.push_back(exc_byte[k]);
    return vg_upload_reads_packed(c, packed + lo / 4, hi - lo, ep.data(), eb.data(), (int64_t)ep.size(), o1.data(), len1 + a,
                                  o2.data(), len2 + a, np);
  });
}

static void vgm_add_stats(vg_map_stats* acc, const vg_map_stats& s) {
  acc->total_pairs += s.total_pairs, acc->both_exist += s.both_exist, acc->reads_total += s.reads_total;
  acc->total_mapped += s.total_mapped, acc->mapped_forward += s.mapped_forward, acc->mapped_reverse += s.mapped_reverse;
  acc->both_mapped += s.both_mapped, acc->total_score += s.total_score, acc->total_concordant += s.total_concordant;
}

static int vgm_map(vg_multi* m, bool msa, const int64_t* pair_idx, int64_t n_idx, vg_map_stats* stats) {
  if (!m) return VG_ERR_INVALID;
  const int n = (int)m->ctx.size();
  if (stats) memset(stats, 0, sizeof *stats);
  std::vector<std::vector<int64_t>> local(n);
  m->contiguous = pair_idx == nullptr;
  if (pair_idx) {
    if (n_idx < 0) return vgm_fail(m, VG_ERR_INVALID, "negative pair count");
    for (int d = 0; d < n; d++) m->outPos[d].clear();
    for (int64_t k = 0; k < n_idx; k++) {
      const int64_t p = pair_idx[k];
      if (p < 0 || p >= m->nPairs) return vgm_fail(m, VG_ERR_INVALID, "pair index %lld out of range", (long long)p);
      const int d = (int)(std::upper_bound(m->lo.begin(), m->lo.end(), p) - m->lo.begin()) - 1;
      local[d].push_back(p - m->lo[d]);
      m->outPos[d].push_back(k);
    }
  }
  std::vector<vg_map_stats> st(n);
  const int rc = vgm_each(m, [&](int d) {
    if (msa) return vg_map_pairs_msa(m->ctx[d], &st[d]);
    if (pair_idx) return vg_map_pairs(m->ctx[d], local[d].data(), (int64_t)local[d].size(), &st[d]);
    return vg_map_pairs(m->ctx[d], nullptr, 0, &st[d]);
  });
  if (rc != VG_OK) return rc;
  m->nResults = 0, m->seq1Bytes = 0;
  for (int d = 0; d < n; d++) {
    m->nRes[d] = m->ctx[d]->nResults;
    m->poolBytes[d] = m->ctx[d]->seq1Bytes;
    m->nResults += m->nRes[d];
    m->seq1Bytes += m->poolBytes[d];
    if (stats) vgm_add_stats(stats, st[d]);
  }
  return VG_OK;
}

int vg_multi_map_pairs(vg_multi* m, const int64_t* pair_idx, int64_t n_idx, vg_map_stats* stats) {
  return vgm_map(m, false, pair_idx, n_idx, stats);
}
int vg_multi_map_pairs_msa(vg_multi* m, vg_map_stats* stats) { return vgm_map(m, true, nullptr, 0, stats); }

int vg_multi_result_sizes(vg_multi* m, int64_t* n_results, int64_t* seq1_bytes) {
  if (!m) return VG_ERR_INVALID;
  if (n_results) *n_results = m->nResults;
  if (seq1_bytes) *seq1_bytes = m->seq1Bytes;
  return VG_OK;
}

// Results in the caller's order (input order, or the order of pair_idx), with the sequence1 pool laid out in that same
// order: the output is byte-identical to what one device would have returned for the same call.
int vg_multi_get_results(vg_multi* m, vg_pair_result* out, int64_t n_out, uint8_t* seq1_pool, int64_t pool_cap) {
  if (!m) return VG_ERR_INVALID;
  if (n_out < m->nResults || pool_cap < m->seq1Bytes) {
    m->required = n_out < m->nResults ? m->nResults : m->seq1Bytes;
    return vgm_fail(m, VG_ERR_CAPACITY, "output buffers too small (%lld results, %lld sequence1 bytes)", (long long)m->nResults,
                    (long long)m->seq1Bytes);
  }
  if (m->nResults > 0 && !out) return vgm_fail(m, VG_ERR_INVALID, "null output");
  if (m->seq1Bytes > 0 && !seq1_pool) return vgm_fail(m, VG_ERR_INVALID, "null sequence1 pool");
  const int n = (int)m->ctx.size();
  std::vector<int64_t> resBase(n + 1, 0), poolBase(n + 1, 0);
  for (int d = 0; d < n; d++) resBase[d + 1] = resBase[d] + m->nRes[d], poolBase[d + 1] = poolBase[d] + m->poolBytes[d];
  if (m->contiguous) {
    // device d's results are the block [lo[d], lo[d + 1]) of the output and its pool the next block of the pool: the
    // offsets (prefix sums in (pair, mate) order on every device) only need the base of the block added
    return vgm_each(m, [&](int d) {
      vg_ctx* c = m->ctx[d];
      const int64_t nr = m->nRes[d];
      if (nr == 0) return (int)VG_OK;
      vg_pair_result* dst = out + resBase[d];
      const int rc = vg_get_results(c, dst, nr, seq1_pool + poolBase[d], m->poolBytes[d]);
      if (rc != VG_OK) return rc;
      if (poolBase[d])
        for (int64_t i = 0; i < nr; i++)
          for (int k = 0; k < 2; k++)
            if (dst[i].m[k].present && dst[i].m[k].seq1_offset >= 0) dst[i].m[k].seq1_offset += poolBase[d];
      return (int)VG_OK;
    });
  }
  // a subset in the caller's own order: fetch per device, then lay records and sequence1 bytes out in that order
  std::vector<std::vector<vg_pair_result>> tmp(n);
  std::vector<std::vector<uint8_t>> tpool(n);
  const int rc = vgm_each(m, [&](int d) {
    tmp[d].resize(m->nRes[d]);
    tpool[d].resize(m->poolBytes[d]);
    if (m->nRes[d] == 0) return (int)VG_OK;
    return vg_get_results(m->ctx[d], tmp[d].data(), m->nRes[d], tpool[d].data(), m->poolBytes[d]);
  });
  if (rc != VG_OK) return rc;
  for (int d = 0; d < n; d++)
    for (int64_t i = 0; i < m->nRes[d]; i++) out[m->outPos[d][i]] = tmp[d][i];
  std::vector<int> owner(m->nResults);
  for (int d = 0; d < n; d++)
    for (int64_t i = 0; i < m->nRes[d]; i++) owner[m->outPos[d][i]] = d;
  int64_t at = 0;
  for (int64_t i = 0; i < m->nResults; i++)
    for (int k = 0; k < 2; k++) {
      vg_mate& mt = out[i].m[k];
      if (!mt.present || mt.seq1_offset < 0) continue;
      if (mt.length > 0) memcpy(seq1_pool + at, tpool[owner[i]].data() + mt.seq1_offset, (size_t)mt.length);
      mt.seq1_offset = at;
      at += mt.length;
    }
  return VG_OK;
}

// vg_get_results_compact over all devices, in the caller's order: the records of device d with their operation indices moved
// behind those of the devices before it (input order), or re-laid in the order of pair_idx (a subset).
int vg_multi_get_results_compact(vg_multi* m, vg_pair_result* out, int64_t n_out, uint32_t* ops, int64_t ops_cap, int64_t* n_ops) {
  if (!m) return VG_ERR_INVALID;
  if (n_ops) *n_ops = 0;
  if (n_out < m->nResults) {
    m->required = m->nResults;
    return vgm_fail(m, VG_ERR_CAPACITY, "record buffer too small (%lld results)", (long long)m->nResults);
  }
  if (m->nResults > 0 && !out) return vgm_fail(m, VG_ERR_INVALID, "null output");
  const int n = (int)m->ctx.size();
  std::vector<std::vector<vg_pair_result>> tmp(n);
  std::vector<std::vector<uint32_t>> tops(n);
  const int rc = vgm_each(m, [&](int d) {
    vg_ctx* c = m->ctx[d];
    tmp[d].resize(m->nRes[d]);
    if (m->nRes[d] == 0) return (int)VG_OK;
    int64_t cnt = 0;
    int r = vg_get_results_compact(c, tmp[d].data(), m->nRes[d], nullptr, 0, &cnt);  // sizes first
    if (r == VG_ERR_CAPACITY) {
      tops[d].resize(cnt);
      r = vg_get_results_compact(c, tmp[d].data(), m->nRes[d], tops[d].data(), cnt, &cnt);
    }
    return r;
  });
  if (rc != VG_OK) return rc;
  int64_t total = 0;
  for (int d = 0; d < n; d++) total += (int64_t)tops[d].size();
  if (n_ops) *n_ops = total;
  if (ops_cap < total) {
    m->required = total;
    return vgm_fail(m, VG_ERR_CAPACITY, "operation pool too small (%lld elements)", (long long)total);
  }
  if (total > 0 && !ops) return vgm_fail(m, VG_ERR_INVALID, "null operation pool");
  std::vector<int64_t> resBase(n + 1, 0);
  for (int d = 0; d < n; d++) resBase[d + 1] = resBase[d] + m->nRes[d];
  std::vector<int> owner(m->nResults);
  for (int d = 0; d < n; d++)
    for (int64_t i = 0; i < m->nRes[d]; i++) {
      const int64_t at = m->contiguous ? resBase[d] + i : m->outPos[d][i];
      out[at] = tmp[d][i];
      owner[at] = d;
    }
  int64_t at = 0;
  for (int64_t i = 0; i < m->nResults; i++)
    for (int k = 0; k < 2; k++) {
      vg_mate& mt = out[i].m[k];
      if (!mt.present || mt.seq1_offset < 0) continue;
      if (mt.pad_ > 0) memcpy(ops + at, tops[owner[i]].data() + mt.seq1_offset, (size_t)mt.pad_ * 4);
      mt.seq1_offset = at;
      at += mt.pad_;
    }
  return VG_OK;
}

// getConsensusSimple's counting step on every device + the all-reduce: afterwards every device holds the global counts.
int vg_multi_pileup(vg_multi* m, int32_t accumulate) {
  if (!m) return VG_ERR_INVALID;
  return vgm_each(m, [&](int d) {
    const int rc = vg_pileup(m->ctx[d], accumulate);
    if (rc != VG_OK) return rc;
    return vg_pileup_allreduce(m->ctx[d]);
  });
}

int vg_multi_pileup_counts_host(vg_multi* m, int32_t* out, int64_t n_int32) {
  if (!m) return VG_ERR_INVALID;
  const int rc = vg_pileup_counts_host(m->ctx[0], out, n_int32);
  if (rc != VG_OK) return vgm_fail(m, rc, "%s", m->ctx[0]->err);
  return VG_OK;
}

int vg_multi_consensus(vg_multi* m, int32_t save_uncovered, int32_t coverage_threshold, const uint8_t* genome, int32_t G,
                       uint8_t* out) {
  if (!m) return VG_ERR_INVALID;
  const int rc = vg_consensus(m->ctx[0], save_uncovered, coverage_threshold, genome, G, out);
  if (rc != VG_OK) return vgm_fail(m, rc, "%s", m->ctx[0]->err);
  return VG_OK;
}

}  // extern "C"

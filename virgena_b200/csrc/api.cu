// C ABI of libvirgena_gpu.so (include/virgena_gpu.h). sm_100a only; no CPU fallback.
#include "ctx.cuh"
#include "sw_host.cuh"
#include "map_host.cuh"
#include "multi.cuh"

extern "C" {

int vg_abi_version(void) { return VG_ABI_VERSION; }

int vg_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

static int vg_set_seed_config(vg_ctx* c, SeedConfig& s, int K, float lo, float hi, const int32_t* cutoffs, int n) {
  if (K < 1 || K > 12) return vg_fail(c, VG_ERR_INVALID, "K=%d outside the supported range 1..12", K);
  if (!cutoffs || n < 1) return vg_fail(c, VG_ERR_INVALID, "cutoffs missing");
  s.K = K;
  s.lowCoef = lo;
  s.highCoef = hi;
  s.cutoffs.assign(cutoffs, cutoffs + n);
  VG_CUDA(c, s.d_cutoffs.reserve((size_t)n * 4));
  VG_CUDA(c, cudaMemcpy(s.d_cutoffs.p, cutoffs, (size_t)n * 4, cudaMemcpyHostToDevice));
  return VG_OK;
}

int vg_ctx_create(const vg_params* p, int device, vg_ctx** out) {
  if (!out) return VG_ERR_INVALID;
  *out = nullptr;
  if (!p) return VG_ERR_INVALID;
  int n = vg_device_count();
  vg_ctx* c = new vg_ctx();
  *out = c;  // returned even on failure so that the caller can read vg_last_error
  if (n <= 0) return vg_fail(c, VG_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  if (device < 0 || device >= n) return vg_fail(c, VG_ERR_INVALID, "device %d out of range (0..%d)", device, n - 1);
  c->device = device;
  VG_CUDA(c, cudaSetDevice(device));
  cudaDeviceProp prop;
  VG_CUDA(c, cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return vg_fail(c, VG_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                   prop.minor);
  c->smCount = prop.multiProcessorCount;
  VG_CUDA(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  for (auto& e : c->ev) VG_CUDA(c, cudaEventCreate(&e));
  c->insertLen = p->insert_len;
  c->sw.match = p->match, c->sw.mismatch = p->mismatch, c->sw.gop = p->gop, c->sw.gep = p->gep;
  VG_TRY(vg_set_seed_config(c, c->seed, p->K, p->low_coef, p->high_coef, p->cutoffs, p->n_cutoffs));
  return VG_OK;
}

void vg_ctx_destroy(vg_ctx* c) {
  if (!c) return;
  if (c->stream) {
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
  }
  vg_comm_release(c);
  DevBuf* bufs[] = {&c->seed.d_cutoffs, &c->seedMsa.d_cutoffs, &c->d_ref, &c->d_contigEnds, &c->d_fragmentEnds,
                    &c->d_poscode, &c->d_exoKeys, &c->d_msaOff, &c->d_msaCols, &c->d_msaExoKeys, &c->d_bases,
                    &c->d_off1, &c->d_off2, &c->d_len1, &c->d_len2, &c->d_results, &c->d_seq1pool, &c->d_pairIdx,
                    &c->d_tasks, &c->d_fill, &c->d_alns, &c->d_rev, &c->d_rowck, &c->d_colck, &c->d_counter,
                    &c->d_scratch, &c->d_seqOff, &c->d_cand, &c->d_ncand, &c->d_misc, &c->d_s1pool, &c->d_s2pool,
                    &c->d_s1off, &c->d_s2off, &c->d_counts, &c->d_swSeq1, &c->d_packed, &c->d_swSorted, &c->d_swBins, &c->d_rowpool, &c->d_rowLen, &c->d_scratchR, &c->d_seedS, &c->d_seedNS, &c->d_seedFb, &c->d_bamN,
                    &c->d_bamOff, &c->d_bamRec, &c->d_bamOps, &c->d_rand};
  for (DevBuf* b : bufs) b->release();
  for (auto& e : c->ev)
    if (e) cudaEventDestroy(e);
  for (auto& e : c->seedEv) cudaEventDestroy(e);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

const char* vg_last_error(const vg_ctx* c) { return c ? c->err : "null context"; }
int64_t vg_last_required(const vg_ctx* c) { return c ? c->required : 0; }
int64_t vg_kernel_launches(const vg_ctx* c) { return c ? c->launches : 0; }

int vg_ctx_stream(vg_ctx* c, void** stream) {
  if (!c || !stream) return VG_ERR_INVALID;
  *stream = (void*)c->stream;
  return VG_OK;
}

int vg_seed_timings(vg_ctx* c, double* out8) {
  if (!c || !out8) return VG_ERR_INVALID;
  for (int i = 0; i < 8; i++) out8[i] = c->seedStats[i];
  return VG_OK;
}

int vg_last_timings(vg_ctx* c, double* out6) {
  if (!c || !out6) return VG_ERR_INVALID;
  for (int i = 0; i < 6; i++) out6[i] = c->timings[i];
  return VG_OK;
}

// ---------------------------------------------------------------------------------------------
// Aligner.align / Aligner.findMaxScore over a batch
// ---------------------------------------------------------------------------------------------
static int sw_batch_common(vg_ctx* c, const uint8_t* s1_pool, const int64_t* s1_off, const uint8_t* s2_pool,
                           const int64_t* s2_off, int64_t n, int* mmaxOut, int* nmaxOut) {
  if (!c) return VG_ERR_INVALID;
  if (n < 0 || (n > 0 && (!s1_pool || !s1_off || !s2_pool || !s2_off))) return vg_fail(c, VG_ERR_INVALID, "null input");
  VG_CUDA(c, cudaSetDevice(c->device));
  std::vector<SwTask> tasks(n);
  int mmax = 0, nmax = 0;
  for (int64_t i = 0; i < n; i++) {
    int64_t m = s1_off[i + 1] - s1_off[i], nn = s2_off[i + 1] - s2_off[i];
    if (m < 0 || nn < 0 || m > (1 << 20) || nn > (1 << 20)) return vg_fail(c, VG_ERR_INVALID, "bad offsets at task %lld", (long long)i);
    tasks[i] = SwTask{s1_off[i], s2_off[i], (int32_t)m, (int32_t)nn, 0, (int32_t)i};
    mmax = std::max(mmax, (int)m);
    nmax = std::max(nmax, (int)nn);
  }
  *mmaxOut = mmax;
  *nmaxOut = nmax;
  if (n == 0) return VG_OK;
  size_t b1 = (size_t)s1_off[n], b2 = (size_t)s2_off[n];
  // the packed kernel keeps a symbol in the high byte of an s16 lane and pads with 0x7F / 0x00FF: bytes >= 127 would
  // alias the padding or wrap the comparison (the Java's Constants tables have 127 entries: such bytes throw there)
  for (size_t i = 0; i < b1; i++)
    if (s1_pool[i] >= 127) return vg_fail(c, VG_ERR_ALPHABET, "sequence byte %d >= 127 in the first pool", (int)s1_pool[i]);
  for (size_t i = 0; i < b2; i++)
    if (s2_pool[i] >= 127) return vg_fail(c, VG_ERR_ALPHABET, "sequence byte %d >= 127 in the second pool", (int)s2_pool[i]);
  VG_CUDA(c, c->d_s1pool.reserve(b1 + 16));
  VG_CUDA(c, c->d_s2pool.reserve(b2 + 16));
  VG_CUDA(c, c->d_tasks.reserve(n * sizeof(SwTask)));
  VG_CUDA(c, c->d_fill.reserve(n * sizeof(SwFillOut)));
  VG_CUDA(c, cudaMemcpyAsync(c->d_s1pool.p, s1_pool, b1, cudaMemcpyHostToDevice, c->stream));
  VG_CUDA(c, cudaMemcpyAsync(c->d_s2pool.p, s2_pool, b2, cudaMemcpyHostToDevice, c->stream));
  VG_CUDA(c, cudaMemcpyAsync(c->d_tasks.p, tasks.data(), n * sizeof(SwTask), cudaMemcpyHostToDevice, c->stream));
  VG_CUDA(c, cudaStreamSynchronize(c->stream));
  return VG_OK;
}

int vg_sw_align_batch(vg_ctx* c, const uint8_t* s1_pool, const int64_t* s1_off, const uint8_t* s2_pool,
                      const int64_t* s2_off, int64_t n, int32_t* rec, uint8_t* seq1_pool, const int64_t* seq1_off) {
  int mmax = 0, nmax = 0;
  VG_TRY(sw_batch_common(c, s1_pool, s1_off, s2_pool, s2_off, n, &mmax, &nmax));
  if (n == 0) return VG_OK;
  if (!rec || !seq1_pool || !seq1_off) return vg_fail(c, VG_ERR_INVALID, "null output");
  const int64_t revStride = mmax + nmax + 1;
  int64_t poolBytes = 0;
  for (int64_t i = 0; i < n; i++) poolBytes = std::max(poolBytes, seq1_off[i] + (s1_off[i + 1] - s1_off[i]) + (s2_off[i + 1] - s2_off[i]));
  VG_CUDA(c, c->d_alns.reserve(n * sizeof(SwAln)));
  VG_CUDA(c, c->d_rev.reserve((size_t)n * revStride));
  VG_CUDA(c, c->d_seqOff.reserve(n * sizeof(int64_t)));
  VG_CUDA(c, c->d_swSeq1.reserve((size_t)poolBytes + 16));  // never the sequence1 pool of the last mapping
  VG_CUDA(c, cudaMemcpyAsync(c->d_seqOff.p, seq1_off, n * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
  c->timings[1] = c->timings[2] = 0;
  VG_TRY(sw_run(c, c->d_tasks.as<SwTask>(), n, c->d_s1pool.as<uint8_t>(), c->d_s2pool.as<uint8_t>(), mmax, nmax, true,
                c->d_fill.as<SwFillOut>(), c->d_alns.as<SwAln>(), c->d_rev.as<uint8_t>(), revStride, &c->timings[1],
                &c->timings[2]));
  sw_emit_seq1_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, c->stream>>>(
      c->d_alns.as<SwAln>(), (int)n, c->d_rev.as<uint8_t>(), revStride, c->d_seqOff.as<int64_t>(), c->d_swSeq1.as<uint8_t>());
  VG_LAUNCH_CHECK(c);
  static_assert(sizeof(SwAln) == 7 * sizeof(int32_t), "SwAln layout");
  VG_CUDA(c, cudaMemcpyAsync(rec, c->d_alns.p, n * sizeof(SwAln), cudaMemcpyDeviceToHost, c->stream));
  VG_CUDA(c, cudaMemcpyAsync(seq1_pool, c->d_swSeq1.p, (size_t)poolBytes, cudaMemcpyDeviceToHost, c->stream));
  VG_CUDA(c, cudaStreamSynchronize(c->stream));
  double cells = 0;
  for (int64_t i = 0; i < n; i++) cells += (double)(s1_off[i + 1] - s1_off[i]) * (double)(s2_off[i + 1] - s2_off[i]);
  c->timings[5] = cells;
  return VG_OK;
}

int vg_sw_max_score_batch(vg_ctx* c, const uint8_t* s1_pool, const int64_t* s1_off, const uint8_t* s2_pool,
                          const int64_t* s2_off, int64_t n, int32_t* score) {
  int mmax = 0, nmax = 0;
  VG_TRY(sw_batch_common(c, s1_pool, s1_off, s2_pool, s2_off, n, &mmax, &nmax));
  if (n == 0) return VG_OK;
  if (!score) return vg_fail(c, VG_ERR_INVALID, "null output");
  c->timings[1] = c->timings[2] = 0;
  VG_TRY(sw_run(c, c->d_tasks.as<SwTask>(), n, c->d_s1pool.as<uint8_t>(), c->d_s2pool.as<uint8_t>(), mmax, nmax, false,
                c->d_fill.as<SwFillOut>(), nullptr, nullptr, 0, &c->timings[1], &c->timings[2]));
  std::vector<SwFillOut> fo(n);
  VG_CUDA(c, cudaMemcpyAsync(fo.data(), c->d_fill.p, n * sizeof(SwFillOut), cudaMemcpyDeviceToHost, c->stream));
  VG_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int64_t i = 0; i < n; i++) score[i] = fo[i].score;
  return VG_OK;
}

int vg_dpx_peak(vg_ctx* c, double* ops_per_s) {
  if (!c || !ops_per_s) return VG_ERR_INVALID;
  VG_CUDA(c, cudaSetDevice(c->device));
  const int blocks = c->smCount * 8, threads = 256, iters = 20000;
  VG_CUDA(c, c->d_scratch.reserve((size_t)blocks * threads * 4));
  double best = 0;
  for (int rep = 0; rep < 4; rep++) {
    VG_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
    dpx_peak_kernel<<<blocks, threads, 0, c->stream>>>(c->d_scratch.as<uint32_t>(), iters);
    VG_LAUNCH_CHECK(c);
    VG_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
    VG_CUDA(c, cudaEventSynchronize(c->ev[1]));
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
    double ops = (double)blocks * threads * (double)iters * 8.0 / (ms * 1e-3);
    if (rep > 0) best = std::max(best, ops);
  }
  *ops_per_s = best;
  return VG_OK;
}

}  // extern "C"

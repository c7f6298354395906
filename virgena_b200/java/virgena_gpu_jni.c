/* JNI layer between VirgenaGpu.java and libvirgena_gpu.so (include/virgena_gpu.h).
 *
 * NEVER RUN: there is no JDK (and no jni.h) in the build environment. tests/test_host_cpu.py compiles this file with
 * `gcc -fsyntax-only` against tests/jni_stub/jni.h, a hand-written subset of the JNI declarations, which checks every
 * call into the library against the header; nothing more is claimed. Build on a host with a JDK:
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude \
 *       virgena_b200/java/virgena_gpu_jni.c -Lvirgena_b200 -lvirgena_gpu -o libvirgena_gpu_jni.so
 *
 * A handle is a vg_ctx* (one device) or a vg_multi* (several devices of this JVM); `multi` says which. Errors become
 * java.lang.RuntimeException(vg_last_error): there is no CPU fallback to fall back to. */
#include <jni.h>
#include <stdint.h>
#include <string.h>

#include "virgena_gpu.h"

static void throw_msg(JNIEnv* e, const char* msg) {
  jclass c = (*e)->FindClass(e, "java/lang/RuntimeException");
  if (c) (*e)->ThrowNew(e, c, msg ? msg : "virgena_gpu: unknown error");
}
static const char* err_of(jlong h, jboolean multi) {
  return multi ? vg_multi_last_error((const vg_multi*)(intptr_t)h) : vg_last_error((const vg_ctx*)(intptr_t)h);
}
#define CHECK(rc, h, multi, ret)          \
  if ((rc) != VG_OK) {                    \
    throw_msg(e, err_of((h), (multi)));   \
    return ret;                           \
  }
#define CTX(h) ((vg_ctx*)(intptr_t)(h))
#define MULTI(h) ((vg_multi*)(intptr_t)(h))

/* pinned views of primitive arrays (NULL array -> NULL pointer) */
#define GET_I(a) ((a) ? (*e)->GetIntArrayElements(e, (a), 0) : 0)
#define GET_L(a) ((a) ? (*e)->GetLongArrayElements(e, (a), 0) : 0)
#define GET_B(a) ((a) ? (*e)->GetByteArrayElements(e, (a), 0) : 0)
#define GET_F(a) ((a) ? (*e)->GetFloatArrayElements(e, (a), 0) : 0)
#define REL_I(a, p, mode) if (a) (*e)->ReleaseIntArrayElements(e, (a), (p), (mode))
#define REL_L(a, p, mode) if (a) (*e)->ReleaseLongArrayElements(e, (a), (p), (mode))
#define REL_B(a, p, mode) if (a) (*e)->ReleaseByteArrayElements(e, (a), (p), (mode))
#define REL_F(a, p, mode) if (a) (*e)->ReleaseFloatArrayElements(e, (a), (p), (mode))
#define LEN(a) ((a) ? (*e)->GetArrayLength(e, (a)) : 0)
#define DIRECT(b) ((b) ? (*e)->GetDirectBufferAddress(e, (b)) : 0)

/* vg_ctx_create / vg_multi_create  <-  new Mapper(document), Mapper.java:27 */
JNIEXPORT jlong JNICALL Java_VirgenaGpu_create(JNIEnv* e, jclass cls, jint K, jfloat lowCoef, jfloat highCoef,
                                               jintArray cutoffs, jint insertLen, jint match, jint mismatch, jint gop,
                                               jint gep, jintArray devices) {
  (void)cls;
  vg_params p;
  memset(&p, 0, sizeof p);
  jint* cut = GET_I(cutoffs);
  jint* dev = GET_I(devices);
  p.K = K, p.low_coef = lowCoef, p.high_coef = highCoef;
  p.cutoffs = (const int32_t*)cut, p.n_cutoffs = LEN(cutoffs);
  p.insert_len = insertLen, p.match = match, p.mismatch = mismatch, p.gop = gop, p.gep = gep;
  const int nDev = LEN(devices);
  jlong h = 0;
  int rc;
  if (nDev > 1) {
    vg_multi* m = 0;
    rc = vg_multi_create(&p, (const int32_t*)dev, nDev, &m);
    h = (jlong)(intptr_t)m;
  } else {
    vg_ctx* c = 0;
    rc = vg_ctx_create(&p, nDev == 1 ? dev[0] : 0, &c);
    h = (jlong)(intptr_t)c;
  }
  REL_I(cutoffs, cut, JNI_ABORT);
  REL_I(devices, dev, JNI_ABORT);
  if (rc != VG_OK) {
    throw_msg(e, rc == VG_ERR_CUDA ? "virgena_gpu: no usable CUDA device (there is no CPU fallback)"
                                   : "virgena_gpu: context creation failed (bad parameters)");
    return 0;
  }
  return h;
}

JNIEXPORT void JNICALL Java_VirgenaGpu_destroy(JNIEnv* e, jclass cls, jlong h, jboolean multi) {
  (void)e, (void)cls;
  if (multi) vg_multi_destroy(MULTI(h));
  else vg_ctx_destroy(CTX(h));
}

/* vg_set_reference  <-  new Reference(...) + StringIndexer.buildIndexPara, Reference.java:98-205 */
JNIEXPORT void JNICALL Java_VirgenaGpu_setReference(JNIEnv* e, jclass cls, jlong h, jboolean multi, jbyteArray seqB,
                                                    jintArray contigEnds, jintArray fragmentEnds, jboolean isFragmented,
                                                    jbyteArray keys, jlongArray offsets, jintArray positions, jint nKeys) {
  (void)cls;
  jbyte* s = GET_B(seqB);
  jint* ce = GET_I(contigEnds);
  jint* fe = GET_I(fragmentEnds);
  jbyte* k = GET_B(keys);
  jlong* o = GET_L(offsets);
  jint* p = GET_I(positions);
  const int32_t G = LEN(seqB);
  int rc = multi ? vg_multi_set_reference(MULTI(h), (const uint8_t*)s, G, (const int32_t*)ce, LEN(contigEnds), (const int32_t*)fe,
                                          LEN(fragmentEnds), isFragmented ? 1 : 0, (const uint8_t*)k, (const int64_t*)o,
                                          (const int32_t*)p, nKeys, 1)
                 : vg_set_reference(CTX(h), (const uint8_t*)s, G, (const int32_t*)ce, LEN(contigEnds), (const int32_t*)fe,
                                    LEN(fragmentEnds), isFragmented ? 1 : 0, (const uint8_t*)k, (const int64_t*)o,
                                    (const int32_t*)p, nKeys, 1);
  REL_B(seqB, s, JNI_ABORT);
  REL_I(contigEnds, ce, JNI_ABORT);
  REL_I(fragmentEnds, fe, JNI_ABORT);
  REL_B(keys, k, JNI_ABORT);
  REL_L(offsets, o, JNI_ABORT);
  REL_I(positions, p, JNI_ABORT);
  CHECK(rc, h, multi, )
}

/* vg_set_msa_index  <-  ReferenceAlignment.buildAlnIndex, ReferenceAlignment.java:138-209 */
JNIEXPORT void JNICALL Java_VirgenaGpu_setMsaIndex(JNIEnv* e, jclass cls, jlong h, jboolean multi, jint K, jfloat lowCoef,
                                                   jfloat highCoef, jintArray cutoffs, jbyteArray keys, jlongArray offsets,
                                                   jintArray columns, jint nKeys, jint msaLength) {
  (void)cls;
  jint* cut = GET_I(cutoffs);
  jbyte* k = GET_B(keys);
  jlong* o = GET_L(offsets);
  jint* c = GET_I(columns);
  int rc = multi ? vg_multi_set_msa_index(MULTI(h), K, lowCoef, highCoef, (const int32_t*)cut, LEN(cutoffs), (const uint8_t*)k,
                                          (const int64_t*)o, (const int32_t*)c, nKeys, msaLength)
                 : vg_set_msa_index(CTX(h), K, lowCoef, highCoef, (const int32_t*)cut, LEN(cutoffs), (const uint8_t*)k,
                                    (const int64_t*)o, (const int32_t*)c, nKeys, msaLength);
  REL_I(cutoffs, cut, JNI_ABORT);
  REL_B(keys, k, JNI_ABORT);
  REL_L(offsets, o, JNI_ABORT);
  REL_I(columns, c, JNI_ABORT);
  CHECK(rc, h, multi, )
}

/* vg_upload_reads  <-  DataReader.pairedReads */
JNIEXPORT void JNICALL Java_VirgenaGpu_uploadReads(JNIEnv* e, jclass cls, jlong h, jboolean multi, jobject bases, jlong nBases,
                                                   jlongArray off1, jintArray len1, jlongArray off2, jintArray len2) {
  (void)cls;
  jlong* o1 = GET_L(off1);
  jlong* o2 = GET_L(off2);
  jint* l1 = GET_I(len1);
  jint* l2 = GET_I(len2);
  const int64_t n = LEN(off1);
  int rc = multi ? vg_multi_upload_reads(MULTI(h), (const uint8_t*)DIRECT(bases), nBases, (const int64_t*)o1, (const int32_t*)l1,
                                         (const int64_t*)o2, (const int32_t*)l2, n)
                 : vg_upload_reads(CTX(h), (const uint8_t*)DIRECT(bases), nBases, (const int64_t*)o1, (const int32_t*)l1,
                                   (const int64_t*)o2, (const int32_t*)l2, n);
  REL_L(off1, o1, JNI_ABORT);
  REL_L(off2, o2, JNI_ABORT);
  REL_I(len1, l1, JNI_ABORT);
  REL_I(len2, l2, JNI_ABORT);
  CHECK(rc, h, multi, )
}

/* vg_map_pairs / vg_map_pairs_msa + vg_result_sizes  <-  Mapper.mapReads (Mapper.java:420,477), MapperMSA.mapAllReads (:274) */
JNIEXPORT jlongArray JNICALL Java_VirgenaGpu_mapPairs(JNIEnv* e, jclass cls, jlong h, jboolean multi, jlongArray idx,
                                                      jboolean msa) {
  (void)cls;
  vg_map_stats st;
  memset(&st, 0, sizeof st);
  jlong* p = GET_L(idx);
  const int64_t n = LEN(idx);
  int rc;
  if (multi) rc = msa ? vg_multi_map_pairs_msa(MULTI(h), &st) : vg_multi_map_pairs(MULTI(h), (const int64_t*)p, n, &st);
  else rc = msa ? vg_map_pairs_msa(CTX(h), &st) : vg_map_pairs(CTX(h), (const int64_t*)p, n, &st);
  REL_L(idx, p, JNI_ABORT);
  CHECK(rc, h, multi, 0)
  int64_t nRes = 0, nBytes = 0;
  rc = multi ? vg_multi_result_sizes(MULTI(h), &nRes, &nBytes) : vg_result_sizes(CTX(h), &nRes, &nBytes);
  CHECK(rc, h, multi, 0)
  jlong out[11];
  memcpy(out, &st, 9 * sizeof(jlong));
  out[9] = nRes, out[10] = nBytes;
  jlongArray res = (*e)->NewLongArray(e, 11);
  if (res) (*e)->SetLongArrayRegion(e, res, 0, 11, out);
  return res;
}

/* vg_get_results */
JNIEXPORT void JNICALL Java_VirgenaGpu_getResults(JNIEnv* e, jclass cls, jlong h, jboolean multi, jobject records,
                                                  jlong nRecords, jobject pool) {
  (void)cls;
  const int64_t cap = pool ? (*e)->GetDirectBufferCapacity(e, pool) : 0;
  int rc = multi ? vg_multi_get_results(MULTI(h), (vg_pair_result*)DIRECT(records), nRecords, (uint8_t*)DIRECT(pool), cap)
                 : vg_get_results(CTX(h), (vg_pair_result*)DIRECT(records), nRecords, (uint8_t*)DIRECT(pool), cap);
  CHECK(rc, h, multi, )
}

/* vg_set_results  <-  the MappedData edits of ConsensusBuilderWithReassembling + Mapper.java:477-523 */
JNIEXPORT void JNICALL Java_VirgenaGpu_setResults(JNIEnv* e, jclass cls, jlong h, jboolean multi, jobject records, jlong n,
                                                  jlongArray pairIdx, jobject pool, jlong poolBytes) {
  (void)cls;
  jlong* p = GET_L(pairIdx);
  vg_ctx* c = multi ? vg_multi_ctx(MULTI(h), 0) : CTX(h);
  int rc = vg_set_results(c, (const vg_pair_result*)DIRECT(records), n, (const int64_t*)p, (const uint8_t*)DIRECT(pool), poolBytes);
  REL_L(pairIdx, p, JNI_ABORT);
  if (rc != VG_OK) throw_msg(e, vg_last_error(c));
}

/* vg_pileup + vg_consensus  <-  getConsensusSimple, ConsensusBuilderSimple.java:123-167 */
JNIEXPORT void JNICALL Java_VirgenaGpu_consensus(JNIEnv* e, jclass cls, jlong h, jboolean multi, jboolean saveUncovered,
                                                 jint coverageThreshold, jbyteArray genome, jbyteArray out) {
  (void)cls;
  int rc = multi ? vg_multi_pileup(MULTI(h), 0) : vg_pileup(CTX(h), 0);
  CHECK(rc, h, multi, )
  if (!multi) {
    rc = vg_pileup_allreduce(CTX(h)); /* no-op unless the host set up a communicator (one JVM per GPU) */
    CHECK(rc, h, multi, )
  }
  jbyte* g = GET_B(genome);
  jbyte* o = GET_B(out);
  const int32_t G = LEN(genome);
  rc = multi ? vg_multi_consensus(MULTI(h), saveUncovered ? 1 : 0, coverageThreshold, (const uint8_t*)g, G, (uint8_t*)o)
             : vg_consensus(CTX(h), saveUncovered ? 1 : 0, coverageThreshold, (const uint8_t*)g, G, (uint8_t*)o);
  REL_B(genome, g, JNI_ABORT);
  REL_B(out, o, 0);
  CHECK(rc, h, multi, )
}

/* vg_bam_fields  <-  BamPrinter.setRecordFieldsMapped / Unmapped / setCigar, BamPrinter.java:16-171 */
JNIEXPORT jlong JNICALL Java_VirgenaGpu_bamFields(JNIEnv* e, jclass cls, jlong h, jobject records, jlong nRecords,
                                                  jobject cigarPool) {
  (void)cls;
  int64_t nOps = 0;
  const int64_t cap = cigarPool ? (*e)->GetDirectBufferCapacity(e, cigarPool) / 4 : 0;
  int rc = vg_bam_fields(CTX(h), (vg_bam_record*)DIRECT(records), nRecords, (uint32_t*)DIRECT(cigarPool), cap, &nOps);
  if (rc == VG_ERR_CAPACITY) return -vg_last_required(CTX(h)); /* caller grows the pool and calls again */
  CHECK(rc, h, JNI_FALSE, 0)
  return nOps;
}

/* vg_best_window_batch  <-  KMerCounter.mapReadToRegion (KMerCounter.java:119-140), computeKMerCount (KMerCounterBase.java:468-487) */
JNIEXPORT void JNICALL Java_VirgenaGpu_bestWindows(JNIEnv* e, jclass cls, jlong h, jboolean msa, jobject reads,
                                                   jlongArray readOff, jintArray windows, jintArray found) {
  (void)cls;
  jlong* ro = GET_L(readOff);
  jint* w = GET_I(windows);
  jint* f = GET_I(found);
  int rc = vg_best_window_batch(CTX(h), msa ? 1 : 0, (const uint8_t*)DIRECT(reads), (const int64_t*)ro, LEN(readOff) - 1,
                                (int32_t*)w, (int32_t*)f);
  REL_L(readOff, ro, JNI_ABORT);
  REL_I(windows, w, 0);
  REL_I(found, f, 0);
  CHECK(rc, h, JNI_FALSE, )
}

/* vg_random_model_counts  <-  RandomModel.RandomCountsComputing.run (RandomModel.java:113-128), MarkovModel.generate (:94-119) */
JNIEXPORT void JNICALL Java_VirgenaGpu_randomModelCounts(JNIEnv* e, jclass cls, jlong h, jboolean msa, jobject genomes,
                                                         jlongArray genomeOff, jint order, jlong seed, jlong firstDraw,
                                                         jint readLen, jintArray counts) {
  (void)cls;
  jlong* go = GET_L(genomeOff);
  jint* c = GET_I(counts);
  int rc = vg_random_model_counts(CTX(h), msa ? 1 : 0, (const uint8_t*)DIRECT(genomes), (const int64_t*)go,
                                  (int32_t)(LEN(genomeOff) - 1), order, (uint64_t)seed, firstDraw, readLen, LEN(counts),
                                  (int32_t*)c, NULL);
  REL_L(genomeOff, go, JNI_ABORT);
  REL_I(counts, c, 0);
  CHECK(rc, h, JNI_FALSE, )
}

/* vg_map_to_regions  <-  ContigBuilder.countReads (ContigBuilder.java:72-104), Graph.SWScore (Graph.java:493-550) */
JNIEXPORT void JNICALL Java_VirgenaGpu_mapToRegions(JNIEnv* e, jclass cls, jlong h, jobject regions, jlongArray regionOff,
                                                    jobject reads, jlongArray readOff, jintArray readRegion,
                                                    jintArray windows, jintArray found) {
  (void)cls;
  jlong* go = GET_L(regionOff);
  jlong* ro = GET_L(readOff);
  jint* rr = GET_I(readRegion);
  jint* w = GET_I(windows);
  jint* f = GET_I(found);
  int rc = vg_map_to_regions(CTX(h), (const uint8_t*)DIRECT(regions), (const int64_t*)go, LEN(regionOff) - 1,
                             (const uint8_t*)DIRECT(reads), (const int64_t*)ro, LEN(readOff) - 1, (const int32_t*)rr,
                             (int32_t*)w, (int32_t*)f);
  REL_L(regionOff, go, JNI_ABORT);
  REL_L(readOff, ro, JNI_ABORT);
  REL_I(readRegion, rr, JNI_ABORT);
  REL_I(windows, w, 0);
  REL_I(found, f, 0);
  CHECK(rc, h, JNI_FALSE, )
}

/* vg_sw_align_batch  <-  Aligner.align, SmithWatermanGotoh.java:51 */
JNIEXPORT void JNICALL Java_VirgenaGpu_alignBatch(JNIEnv* e, jclass cls, jlong h, jobject s1, jlongArray s1Off, jobject s2,
                                                  jlongArray s2Off, jintArray rec, jobject seq1, jlongArray seq1Off) {
  (void)cls;
  jlong* o1 = GET_L(s1Off);
  jlong* o2 = GET_L(s2Off);
  jlong* os = GET_L(seq1Off);
  jint* r = GET_I(rec);
  int rc = vg_sw_align_batch(CTX(h), (const uint8_t*)DIRECT(s1), (const int64_t*)o1, (const uint8_t*)DIRECT(s2),
                             (const int64_t*)o2, LEN(s1Off) - 1, (int32_t*)r, (uint8_t*)DIRECT(seq1), (const int64_t*)os);
  REL_L(s1Off, o1, JNI_ABORT);
  REL_L(s2Off, o2, JNI_ABORT);
  REL_L(seq1Off, os, JNI_ABORT);
  REL_I(rec, r, 0);
  CHECK(rc, h, JNI_FALSE, )
}

/* vg_pileup_alignments  <-  the counts loop of ContigBuilder.countReads, ContigBuilder.java:93-99 */
JNIEXPORT void JNICALL Java_VirgenaGpu_pileupAlignments(JNIEnv* e, jclass cls, jlong h, jobject seq1, jlongArray seq1Off,
                                                        jintArray length, jlongArray pos, jlongArray end,
                                                        jlong totalPositions, jintArray counts) {
  (void)cls;
  jlong* os = GET_L(seq1Off);
  jint* ln = GET_I(length);
  jlong* ps = GET_L(pos);
  jlong* en = GET_L(end);
  jint* ct = GET_I(counts);
  int rc = vg_pileup_alignments(CTX(h), (const uint8_t*)DIRECT(seq1), (const int64_t*)os, (const int32_t*)ln,
                                (const int64_t*)ps, (const int64_t*)en, LEN(length), totalPositions, (int32_t*)ct);
  REL_L(seq1Off, os, JNI_ABORT);
  REL_I(length, ln, JNI_ABORT);
  REL_L(pos, ps, JNI_ABORT);
  REL_L(end, en, JNI_ABORT);
  REL_I(counts, ct, 0);
  CHECK(rc, h, JNI_FALSE, )
}

/* vg_identity_counting  <-  ProblemIntervalBuilder.countIdentity, ProblemIntervalBuilder.java:135-160 */
JNIEXPORT void JNICALL Java_VirgenaGpu_identityCounting(JNIEnv* e, jclass cls, jlong h, jint windowSize, jint threadNum,
                                                        jfloatArray identity, jintArray coverage) {
  (void)cls;
  jfloat* id = GET_F(identity);
  jint* cv = GET_I(coverage);
  int rc = vg_identity_counting(CTX(h), windowSize, threadNum, (float*)id, (int32_t*)cv);
  REL_F(identity, id, 0);
  REL_I(coverage, cv, 0);
  CHECK(rc, h, JNI_FALSE, )
}

// UNCOMPILED SOURCE: no JDK exists in the build environment (SURVEY.md §0), so this file has never been through
// javac. It is the JNI binding a VirGenA maintainer adds for libvirgena_gpu.so (include/virgena_gpu.h); the C half is
// virgena_gpu_jni.c next to it (that one is syntax-checked against a stub jni.h by tests/test_host_cpu.py).
//
// Lives in VirGenA's default package next to Mapper.java. One instance = one vg_ctx (one GPU) or one vg_multi
// (several GPUs of this JVM, SURVEY.md §8b/§8e); the native layer tells them apart by the `multi` flag.
import java.nio.ByteBuffer;
import java.nio.ByteOrder;
import java.util.ArrayList;
import java.util.HashMap;
import java.util.Map;

final class VirgenaGpu implements AutoCloseable {
  static { System.loadLibrary("virgena_gpu_jni"); }  // thin JNI layer over libvirgena_gpu.so

  static final int PAIR_BYTES = 136, MATE_BYTES = 64, BAM_BYTES = 40;  // sizeof(vg_pair_result), vg_mate, vg_bam_record

  private long handle;          // vg_ctx* or vg_multi*
  private final boolean multi;

  // vg_ctx_create / vg_multi_create: parameters come from the LIVE Java objects at call time (KMerCounter is a
  // process-wide singleton that ignores later config documents, KMerCounter.java:57-62), never from the XML.
  VirgenaGpu(KMerCounterBase counter, float lowCoef, float highCoef, Aligner aligner, int insertLen, int[] devices) {
    multi = devices.length > 1;
    handle = create(counter.K, lowCoef, highCoef, counter.cutoffs, insertLen, aligner.match, aligner.mismatch, aligner.gop,
                    aligner.gep, devices);                      // Aligner.java:9-12
  }

  // ---- natives: one per C-ABI entry point (virgena_gpu_jni.c); every one throws RuntimeException(vg_last_error) ----
  private static native long create(int K, float lowCoef, float highCoef, int[] cutoffs, int insertLen, int match,
                                    int mismatch, int gop, int gep, int[] devices);
  private static native void destroy(long h, boolean multi);
  // vg_set_reference: the index of StringIndexer.buildIndexPara as CSR (keys = nKeys x K bytes), so that the duplicated
  // thread-boundary positions (Q1) are exactly the Java's whatever threadNum it was built with
  private static native void setReference(long h, boolean multi, byte[] seqB, int[] contigEnds, int[] fragmentEnds,
                                          boolean isFragmented, byte[] keys, long[] offsets, int[] positions, int nKeys);
  // vg_set_msa_index: CSR of ReferenceAlignment.refAlnIndex
  private static native void setMsaIndex(long h, boolean multi, int K, float lowCoef, float highCoef, int[] cutoffs,
                                         byte[] keys, long[] offsets, int[] columns, int nKeys, int msaLength);
  // vg_upload_reads
  private static native void uploadReads(long h, boolean multi, ByteBuffer bases, long nBases, long[] off1, int[] len1,
                                         long[] off2, int[] len2);
  // vg_map_pairs / vg_map_pairs_msa + vg_result_sizes: {9 stats, nResults, seq1Bytes}
  private static native long[] mapPairs(long h, boolean multi, long[] pairIdxOrNull, boolean msa);
  // vg_get_results: records = direct buffer of vg_pair_result
  private static native void getResults(long h, boolean multi, ByteBuffer records, long nRecords, ByteBuffer seq1Pool);
  // vg_set_results (single context only: the merged MappedData of a remap pass lives on device 0 of a multi handle)
  private static native void setResults(long h, boolean multi, ByteBuffer records, long n, long[] pairIdx, ByteBuffer seq1Pool,
                                        long poolBytes);
  // vg_pileup (+ all-reduce inside for a multi handle) + vg_consensus
  private static native void consensus(long h, boolean multi, boolean saveUncovered, int coverageThreshold, byte[] genome,
                                       byte[] out);
  // vg_bam_fields: records = direct buffer of vg_bam_record, cigar = len << 4 | op elements; returns the element count
  private static native long bamFields(long h, ByteBuffer records, long nRecords, ByteBuffer cigarPool);
  // vg_best_window_batch, vg_map_to_regions, vg_sw_align_batch, vg_pileup_alignments, vg_identity_counting (SURVEY §8f)
  static native void bestWindows(long h, boolean msa, ByteBuffer reads, long[] readOff, int[] windows, int[] found);
  static native void mapToRegions(long h, ByteBuffer regions, long[] regionOff, ByteBuffer reads, long[] readOff,
                                  int[] readRegion, int[] windows, int[] found);
  static native void alignBatch(long h, ByteBuffer s1, long[] s1Off, ByteBuffer s2, long[] s2Off, int[] rec, ByteBuffer seq1,
                                long[] seq1Off);
  static native void pileupAlignments(long h, ByteBuffer seq1, long[] seq1Off, int[] length, long[] pos, long[] endOrNull,
                                      long totalPositions, int[] counts);
  static native void identityCounting(long h, int windowSize, int threadNum, float[] identity, int[] coverage);
  // vg_random_model_counts: RandomModel.RandomCountsComputing.run (RandomModel.java:113-128) for readNum reads of one length
  static native void randomModelCounts(long h, boolean msa, ByteBuffer genomes, long[] genomeOff, int order, long seed,
                                       long firstDraw, int readLen, int[] counts);

  long handle() { return handle; }
  boolean isMulti() { return multi; }

  @Override public void close() { if (handle != 0) { destroy(handle, multi); handle = 0; } }

  // ---- host-side packing --------------------------------------------------------------------------------------------
  /** HashMap<String,int[]> index -> CSR (keys sorted so that the call is deterministic; the order is irrelevant to the library). */
  static final class Csr { byte[] keys; long[] offsets; int[] values; int nKeys; }
  static Csr toCsr(HashMap<String, int[]> index, int K) {
    ArrayList<String> ks = new ArrayList<>(index.keySet());
    java.util.Collections.sort(ks);
    Csr c = new Csr();
    c.nKeys = ks.size();
    c.keys = new byte[c.nKeys * K];
    c.offsets = new long[c.nKeys + 1];
    int total = 0;
    for (String k : ks) total += index.get(k).length;
    c.values = new int[total];
    int at = 0;
    for (int i = 0; i < c.nKeys; i++) {
      byte[] kb = ks.get(i).getBytes();
      System.arraycopy(kb, 0, c.keys, i * K, K);
      int[] v = index.get(ks.get(i));
      c.offsets[i] = at;
      System.arraycopy(v, 0, c.values, at, v.length);
      at += v.length;
    }
    c.offsets[c.nKeys] = at;
    return c;
  }

  void setReference(Reference genome, int K) {
    Csr c = toCsr(genome.index, K);
    int[] frag = genome.fragmentEnds != null ? genome.fragmentEnds : genome.contigEnds;
    setReference(handle, multi, genome.seqB, genome.contigEnds, frag, genome.isFragmented, c.keys, c.offsets, c.values, c.nKeys);
  }

  void setMsa(ReferenceAlignment aln, KMerCounterMSA counter, float lowCoef, float highCoef) {
    Csr c = toCsr(aln.refAlnIndex, counter.K);
    setMsaIndex(handle, multi, counter.K, lowCoef, highCoef, counter.cutoffs, c.keys, c.offsets, c.values, c.nKeys, aln.length);
  }

  /** DataReader's pairs as one byte pool; an absent mate stays the single byte '*' (Constants.NULL_SEQ). */
  void uploadReads(ArrayList<PairedRead> reads) {
    long total = 0;
    for (PairedRead r : reads) total += r.seq1.length + r.seq2.length;
    if (total > Integer.MAX_VALUE) throw new RuntimeException("read pool above 2 GB: upload in batches");
    ByteBuffer pool = ByteBuffer.allocateDirect((int) Math.max(total, 1));
    int n = reads.size();
    long[] off1 = new long[n], off2 = new long[n];
    int[] len1 = new int[n], len2 = new int[n];
    for (int i = 0; i < n; i++) {
      PairedRead r = reads.get(i);
      off1[i] = pool.position(); len1[i] = r.seq1.length; pool.put(r.seq1);
      off2[i] = pool.position(); len2[i] = r.seq2.length; pool.put(r.seq2);
    }
    uploadReads(handle, multi, pool, total, off1, len1, off2, len2);
  }

  /** Result of one mapping call, still in the library's record layout. */
  static final class Mapped { long[] stats; ByteBuffer records; ByteBuffer pool; int n; }
  Mapped map(long[] pairIdxOrNull, boolean msa) {
    long[] r = mapPairs(handle, multi, pairIdxOrNull, msa);
    Mapped m = new Mapped();
    m.stats = java.util.Arrays.copyOf(r, 9);
    m.n = (int) r[9];
    m.records = ByteBuffer.allocateDirect(Math.max(1, m.n * PAIR_BYTES)).order(ByteOrder.LITTLE_ENDIAN);
    m.pool = ByteBuffer.allocateDirect((int) Math.max(1, r[10]));
    getResults(handle, multi, m.records, m.n, m.pool);
    return m;
  }

  byte[] consensus(boolean saveUncovered, int coverageThreshold, byte[] genome) {
    byte[] out = new byte[genome.length];
    consensus(handle, multi, saveUncovered, coverageThreshold, genome, out);
    return out;
  }

  /** Pushes a host-edited MappedData back as the context's last mapping (vg_set_results). */
  void setResults(ArrayList<PairedRead> pairs, Map<PairedRead, Integer> residentIndex, Map<PairedRead, Integer> pairClass) {
    int n = pairs.size();
    long poolBytes = 0;
    for (PairedRead p : pairs) {
      if (p.r1 != null && p.r1.aln != null) poolBytes += p.r1.aln.length;
      if (p.r2 != null && p.r2.aln != null) poolBytes += p.r2.aln.length;
    }
    ByteBuffer rec = ByteBuffer.allocateDirect(Math.max(1, n * PAIR_BYTES)).order(ByteOrder.LITTLE_ENDIAN);
    ByteBuffer pool = ByteBuffer.allocateDirect((int) Math.max(1, poolBytes));
    long[] idx = new long[n];
    for (int i = 0; i < n; i++) {
      PairedRead p = pairs.get(i);
      idx[i] = residentIndex.get(p);
      Integer cls = pairClass.get(p);                              // concordant side / discordant: as the mapping call said
      rec.putInt(i * PAIR_BYTES, cls != null ? cls : classOf(p));
      encodeMate(rec, i, 0, p.r1, pool);
      encodeMate(rec, i, 1, p.r2, pool);
    }
    setResults(handle, multi, rec, n, idx, pool, poolBytes);
  }

  // fallback for a pair the shim has not seen mapped (cannot happen through MapperGpu): by which mates exist
  static int classOf(PairedRead p) {
    if (p.r1 != null && p.r2 != null) return 2;
    return p.r1 != null ? 3 : p.r2 != null ? 4 : 5;
  }

  private static void encodeMate(ByteBuffer rec, int i, int k, MappedRead m, ByteBuffer pool) {
    int base = i * PAIR_BYTES + 8 + k * MATE_BYTES;
    if (m == null) { rec.putInt(base, 0); rec.putLong(base + 56, -1L); return; }
    rec.putInt(base, 1); rec.putInt(base + 4, m.start); rec.putInt(base + 8, m.end); rec.putInt(base + 12, m.count);
    rec.putInt(base + 16, m.reverse); rec.putInt(base + 20, m.fragmentID);
    Alignment a = m.aln;
    rec.putInt(base + 24, a.score); rec.putInt(base + 28, a.start1); rec.putInt(base + 32, a.end1);
    rec.putInt(base + 36, a.start2); rec.putInt(base + 40, a.end2); rec.putInt(base + 44, a.identity);
    rec.putInt(base + 48, a.length);
    rec.putLong(base + 56, (long) pool.position());
    pool.put(a.sequence1, 0, a.length);
  }

  /** Decodes record i / mate k of a vg_pair_result buffer into the reference's MappedRead (+ Alignment). */
  static MappedRead decodeMate(ByteBuffer rec, int i, int k, PairedRead pair, ByteBuffer seq1Pool, boolean msa) {
    int base = i * PAIR_BYTES + 8 + k * MATE_BYTES;
    if (rec.getInt(base) == 0) return null;  // present == 0  <=>  r1/r2 == null
    MappedRead m = new MappedRead(rec.getInt(base + 4), rec.getInt(base + 8), rec.getInt(base + 12));
    m.reverse = (byte) rec.getInt(base + 16);
    m.fragmentID = rec.getInt(base + 20);
    byte[] fwd = k == 0 ? pair.seq1 : pair.seq2;
    m.seq = m.reverse == 1 ? Constants.getComplement(fwd) : fwd;   // Mapper.java:96-99
    m.q = k == 0 ? pair.q1 : pair.q2;
    m.n = (byte) (k + 1);
    m.name = pair.name;
    if (!msa) {
      Alignment a = new Alignment();
      a.score = rec.getInt(base + 24); a.start1 = rec.getInt(base + 28); a.end1 = rec.getInt(base + 32);
      a.start2 = rec.getInt(base + 36); a.end2 = rec.getInt(base + 40); a.identity = rec.getInt(base + 44);
      a.length = rec.getInt(base + 48);
      int off = (int) rec.getLong(base + 56);
      a.sequence1 = new byte[a.length];
      for (int t = 0; t < a.length; t++) a.sequence1[t] = seq1Pool.get(off + t);
      m.aln = a;
    }
    return m;
  }
}

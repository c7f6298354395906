// UNCOMPILED SOURCE (no JDK in the build environment). Drop-in for MapperMSA (MapperMSA.java:19,274): mapAllReads on the
// device (seeding + pairing in alignment-column coordinates, no Smith-Waterman; totalScore sums k-mer counts, :161).
// To switch: `new MapperMSAGpu(document)` where ReferenceFinder builds its MapperMSA (ReferenceFinder.java:522 uses it).
import java.util.ArrayList;
import org.jdom2.Document;

class MapperMSAGpu extends MapperMSA {
  private final VirgenaGpu gpu;
  private final KMerCounterMSA counterMSA;
  private ArrayList<PairedRead> resident;
  private ReferenceAlignment boundAln;

  MapperMSAGpu(Document document) {
    super(document);
    counterMSA = KMerCounterMSA.getInstance(document);           // the same singleton the super constructor took
    int insertLen = Integer.parseInt(document.getRootElement().getChild("Data").getChildText("InsertionLength"));
    batch = Integer.parseInt(document.getRootElement().getChildText("BatchSize"));
    // the single-reference parameters of the handle are unused in MSA mode; the aligner is never called (MapperMSA has none)
    gpu = new VirgenaGpu(counterMSA, counterMSA.lowCoef, counterMSA.highCoef, new SmithWatermanGotoh(document), insertLen,
                         new int[]{0});
  }

  @Override
  MappedData mapAllReads(ArrayList<PairedRead> reads, ReferenceAlignment refAlignment) throws InterruptedException {
    long time = System.currentTimeMillis();
    if (reads != resident) { gpu.uploadReads(reads); resident = reads; }
    if (refAlignment != boundAln) {
      gpu.setMsa(refAlignment, counterMSA, counterMSA.lowCoef, counterMSA.highCoef);   // lowCoef = 1/highCoef, KMerCounterMSA.java:22-23
      boundAln = refAlignment;
    }
    VirgenaGpu.Mapped m = gpu.map(null, true);
    MappedData res = new MappedData();
    // MapperMSA.java:300-329 builds mappedReads list by list (concordant, left, right, discordant) with `pairs` /
    // `isConcordant` alongside. The Java does it per task of BatchSize pairs; tasks are contiguous blocks in input order,
    // so the same loop over blocks of batchSize reproduces the order exactly.
    int batch = batchSizeOf();
    ArrayList<Integer> pairsL = new ArrayList<>();
    ArrayList<Byte> concL = new ArrayList<>();
    for (int from = 0; from < reads.size(); from += batch) {
      int to = Math.min(from + batch, reads.size());
      ArrayList<PairedRead> conc = new ArrayList<>(), disc = new ArrayList<>(), left = new ArrayList<>(),
          right = new ArrayList<>(), unm = new ArrayList<>();
      for (int i = from; i < to; i++) {
        PairedRead p = reads.get(i);
        p.r1 = VirgenaGpu.decodeMate(m.records, i, 0, p, m.pool, true);
        p.r2 = VirgenaGpu.decodeMate(m.records, i, 1, p, m.pool, true);
        switch (m.records.getInt(i * VirgenaGpu.PAIR_BYTES)) {
          case 0: case 1: conc.add(p); break;
          case 2: disc.add(p); break;
          case 3: left.add(p); break;
          case 4: right.add(p); break;
          default: unm.add(p);
        }
      }
      for (PairedRead p : conc) {
        int c = res.mappedReads.size();
        res.mappedReads.add(p.r1); pairsL.add(c + 1); concL.add((byte) 1);
        res.mappedReads.add(p.r2); pairsL.add(c); concL.add((byte) 1);
      }
      for (PairedRead p : left) { res.mappedReads.add(p.r1); pairsL.add(-1); concL.add((byte) 0); }
      for (PairedRead p : right) { res.mappedReads.add(p.r2); pairsL.add(-1); concL.add((byte) 0); }
      for (PairedRead p : disc) {
        int c = res.mappedReads.size();
        res.mappedReads.add(p.r1); pairsL.add(c + 1); concL.add((byte) 0);
        res.mappedReads.add(p.r2); pairsL.add(c); concL.add((byte) 0);
      }
      res.concordant.addAll(conc); res.discordant.addAll(disc); res.leftMateMapped.addAll(left);
      res.rightMateMapped.addAll(right); res.unmapped.addAll(unm);
    }
    res.pairs = new int[pairsL.size()];
    res.isConcordant = new byte[pairsL.size()];
    for (int i = 0; i < res.pairs.length; i++) { res.pairs[i] = pairsL.get(i); res.isConcordant[i] = concL.get(i); }
    long[] s = m.stats;                                           // MapperMSA.java:335-357
    logger.println("Mapping to MSA stats:");
    logger.printf("1) Total read pairs: %d\n", s[0]);
    logger.printf("2) Total pairs with both reads exists: %d\n", s[1]);
    logger.printf("3) Total reads: %d\n", s[2]);
    logger.printf("4) Total reads mapped from (3): %1.2f, forward: %1.2f, reverse: %1.2f\n",
        (float) s[3] / s[2], (float) s[4] / s[3], (float) s[5] / s[3]);
    logger.printf("5) Total pairs with both reads mapped from (2): %d, %1.2f\n", s[6], (float) s[6] / s[1]);
    logger.printf("6) Concordant pairs from (5): %1.2f\n", (float) s[8] / s[6]);
    logger.printf("7) Total score: %d, average score: %1.2f\n", s[7], (float) s[7] / s[3]);
    logger.printf("Total time, s: %d\n", (System.currentTimeMillis() - time) / 1000);
    return res;
  }

  // MapperMSA.batchSize is private upstream; the shim re-reads the same element of the same document
  private int batch;
  private int batchSizeOf() { return batch; }
}

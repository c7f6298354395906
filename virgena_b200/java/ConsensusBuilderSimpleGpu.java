// UNCOMPILED SOURCE (no JDK in the build environment). Drop-in for ConsensusBuilderSimple: getConsensusSimple
// (ConsensusBuilderSimple.java:123) = ReadCounting over mappedData.mappedReads (:56-75), the sum of the per-thread
// matrices (:138-145; across several GPUs the NCCL all-reduce inside vg_multi_pileup) and ConsensusBuildingSimple
// (:98-120), all on the device over the mapper's last mapping (vg_map_pairs, or vg_set_results after a remap pass).
import java.util.ArrayList;
import org.jdom2.Document;

class ConsensusBuilderSimpleGpu extends ConsensusBuilderSimple {
  ConsensusBuilderSimpleGpu(Document document) {
    super(document);
    mapper = new MapperGpu(document);                             // replaces the Mapper of ConsensusBuilderSimple.java:23
  }

  @Override
  public String buildConsensus(Reference genome, ArrayList<PairedRead> reads) {
    try {
      mappedData = mapper.mapReads(reads, genome);
      return new String(getConsensusSimple(false, coverageThreshold, genome.seqB));
    } catch (Exception e) {                                       // like upstream: printed, null returned (:36-39)
      e.printStackTrace();
    }
    return null;
  }

  @Override
  byte[] getConsensusSimple(boolean saveUncovered, int coverageThreshold, byte[] genome) throws InterruptedException {
    if (genome.length == 0) return new byte[0];
    // the device holds exactly mappedData.mappedReads: MapperGpu installs every mapping it returns (and the merged lists
    // after mapReads(MappedData, Reference)) as the context's last mapping
    return ((MapperGpu) mapper).gpu.consensus(saveUncovered, coverageThreshold, genome);
  }
}

/*
 * virgena_gpu.h — C ABI of libvirgena_gpu.so, the B200 (sm_100a) implementation
 * of VirGenA's read-mapping hot path.
 *
 * The reference (gFedonin/VirGenA, pure Java) has no FFI/plugin interface; the
 * seam this library replaces is the batch API used by RefBasedAssembler and
 * MassRefBasedAssembler:
 *
 *   Mapper.mapReads(ArrayList<PairedRead>, Reference)          Mapper.java:420
 *   Mapper.mapReads(MappedData, Reference)   (remap subset)    Mapper.java:477
 *   MapperMSA.mapAllReads(ArrayList<PairedRead>, ReferenceAlignment)
 *                                                              MapperMSA.java:274
 *   ConsensusBuilderSimple.getConsensusSimple(boolean,int,byte[])
 *                                                              ConsensusBuilderSimple.java:123
 *
 * Each entry point below names the reference code it stands in for. A JNI /
 * Panama / ctypes stub binds exactly these symbols.
 *
 * Conventions
 *   - plain C types only; all host buffers are owned by the caller and are not
 *     retained past the call; device memory is owned by the context.
 *   - every function returns VG_OK (0) or a negative vg_status; the message is
 *     available from vg_last_error(ctx). There is NO CPU fallback: without a
 *     CUDA device every compute entry point fails with VG_ERR_CUDA.
 *   - calls on one context are serialised by the caller (one host thread per
 *     context, like one `Mapper` instance per `mapReads` call).
 *   - coordinates are 0-based offsets into the concatenated reference `seqB`
 *     (Reference.java:207-256), exactly as in MappedRead / Alignment.
 */
#ifndef VIRGENA_GPU_H
#define VIRGENA_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VG_ABI_VERSION 3

typedef enum {
  VG_OK = 0,
  VG_ERR_INVALID = -1,     /* bad argument / unsupported parameter range        */
  VG_ERR_CUDA = -2,        /* CUDA runtime failure, or no device                */
  VG_ERR_STATE = -3,       /* call order (e.g. map before set_reference)        */
  VG_ERR_CAPACITY = -4,    /* caller buffer too small; see vg_last_required()   */
  VG_ERR_ALPHABET = -5,    /* a byte the Java reference would throw on          */
  VG_ERR_INDEX = -6,       /* host-supplied k-mer index inconsistent with seqB  */
  VG_ERR_OVERFLOW = -7     /* a device work buffer overflowed (see message)     */
} vg_status;

/* Pair classes = the six result lists of MappedData (MappedData.java:8-13);
 * the two concordant sides are Mapper.java:226-263. */
enum {
  VG_CLS_CONCORDANT_FR = 0, /* side 0: forward mate1 x reverse mate2 */
  VG_CLS_CONCORDANT_RF = 1, /* side 1: reverse mate1 x forward mate2 */
  VG_CLS_DISCORDANT = 2,
  VG_CLS_LEFT = 3,          /* leftMateMapped  */
  VG_CLS_RIGHT = 4,         /* rightMateMapped */
  VG_CLS_UNMAPPED = 5
};

/* Parameters read by the reference from config.xml at construction
 * (Mapper.java:27-37, KMerCounter.java:14-33, SmithWatermanGotoh.java:7-14).
 * low_coef/high_coef are passed verbatim (Q3: the single-reference Mapper runs
 * with lowCoef == 0.0f). cutoffs[len] is KMerCounterBase.cutoffs (an input:
 * the reference derives it from an unseeded RNG, Q10). */
typedef struct {
  int32_t K;
  float low_coef;
  float high_coef;
  const int32_t* cutoffs; /* n_cutoffs = maxReadLen + 1 entries */
  int32_t n_cutoffs;
  int32_t insert_len;     /* Data/InsertionLength */
  int32_t match, mismatch, gop, gep;
} vg_params;

/* One mapped mate = MappedRead (MappedRead.java:4-16) + its Alignment
 * (Alignment.java:4-35). `present == 0` means the Java field (r1/r2) is null.
 * seq1_offset/length address Alignment.sequence1 inside the caller's pool. */
typedef struct {
  int32_t present;
  int32_t start, end, count;   /* MappedRead.start/end/count (candidate window) */
  int32_t reverse;             /* MappedRead.reverse */
  int32_t fragment_id;         /* MappedRead.fragmentID */
  int32_t score, start1, end1, start2, end2, identity, length; /* Alignment.* */
  int32_t pad_;
  int64_t seq1_offset;         /* -1 when there is no alignment (MSA mode)      */
} vg_mate;

typedef struct {
  int32_t cls;                 /* VG_CLS_* */
  int32_t pad_;
  vg_mate m[2];                /* r1, r2 */
} vg_pair_result;

/* Mapping stats 1)-7) logged by Mapper.java:462-473 (a parity checksum). */
typedef struct {
  int64_t total_pairs, both_exist, reads_total, total_mapped;
  int64_t mapped_forward, mapped_reverse, both_mapped, total_score, total_concordant;
} vg_map_stats;

typedef struct vg_ctx vg_ctx;

/* Library / device probes (no context needed). */
int vg_abi_version(void);
int vg_device_count(void);           /* 0 when no CUDA device is usable */

/* new Mapper(document) [Mapper.java:27] + KMerCounter + SmithWatermanGotoh
 * parameters, bound to one CUDA device. */
int vg_ctx_create(const vg_params* params, int device, vg_ctx** out);
void vg_ctx_destroy(vg_ctx* ctx);
const char* vg_last_error(const vg_ctx* ctx);
int64_t vg_last_required(const vg_ctx* ctx);

/* new Reference(...) [Reference.java:98-205]: the concatenated sequence, contig
 * and fragment ends, and the k-mer index of StringIndexer.buildIndexPara
 * [StringIndexer.java:47-83]. Either pass the host-built index as CSR
 * (n_keys K-byte keys, offsets[n_keys+1], positions) or pass keys == NULL and
 * the thread count the Java would have used (index_thread_num; it determines
 * the duplicated boundary positions, Q1) and the library rebuilds it. */
int vg_set_reference(vg_ctx* ctx, const uint8_t* seq, int32_t G, const int32_t* contig_ends, int32_t n_contigs,
                     const int32_t* fragment_ends, int32_t n_fragments, int32_t is_fragmented,
                     const uint8_t* keys, const int64_t* offsets, const int32_t* positions, int32_t n_keys,
                     int32_t index_thread_num);

/* ReferenceAlignment.refAlnIndex [ReferenceAlignment.java:138-209] + the
 * KMerCounterMSA parameters [KMerCounterMSA.java:14-47]: k-mer -> sorted
 * distinct alignment columns, as CSR. */
int vg_set_msa_index(vg_ctx* ctx, int32_t K, float low_coef, float high_coef, const int32_t* cutoffs,
                     int32_t n_cutoffs, const uint8_t* keys, const int64_t* offsets, const int32_t* columns,
                     int32_t n_keys, int32_t msa_length);

/* DataReader.pairedReads made resident on the device: `bases` is one byte
 * pool; mate k of pair i is bases[off_k[i] .. off_k[i]+len_k[i]). An absent
 * mate is the single byte '*' (Constants.NULL_SEQ). Reads stay resident across
 * mapping passes (iterations 1..3 of ConsensusBuilderWithReassembling). */
int vg_upload_reads(vg_ctx* ctx, const uint8_t* bases, int64_t n_bases, const int64_t* off1, const int32_t* len1,
                    const int64_t* off2, const int32_t* len2, int64_t n_pairs);

/* The same with the bases packed for the transfer (north star: 2-bit reads): 2 bits per base (A=0, C=1, G=2, T=3), base i
 * in bits 2*(i&3) of byte i>>2, plus n_exc exceptions (exc_pos[k], exc_byte[k]) for every byte that is not ACGT ('N',
 * '-', the '*' of an absent mate, ...). off/len address bases exactly as in vg_upload_reads. The device unpacks into
 * the same resident byte pool (the path's alphabet is bytes: any byte takes part in byte equality), so everything
 * downstream is identical; the host-to-device copy is 4x smaller. vg_pack_reads is the host-side packer (plain C loop,
 * no device): packed holds (n_bases+3)/4 bytes; exceptions come out in ascending position order; VG_ERR_CAPACITY with
 * *n_exc = the number needed when exc_cap is too small. */
int vg_upload_reads_packed(vg_ctx* ctx, const uint8_t* packed, int64_t n_bases, const int64_t* exc_pos,
                           const uint8_t* exc_byte, int64_t n_exc, const int64_t* off1, const int32_t* len1,
                           const int64_t* off2, const int32_t* len2, int64_t n_pairs);
int vg_pack_reads(const uint8_t* bases, int64_t n_bases, uint8_t* packed, int64_t* exc_pos, uint8_t* exc_byte,
                  int64_t exc_cap, int64_t* n_exc);

/* Mapper.mapReads(reads, genome) [Mapper.java:420] when pair_idx == NULL
 * (all resident pairs, in order), or Mapper.mapReads(mappedData, genome)
 * [Mapper.java:477] over the subset pair_idx[0..n_idx) (needToRemap).
 * Results stay on the device; fetch them with vg_get_results. */
int vg_map_pairs(vg_ctx* ctx, const int64_t* pair_idx, int64_t n_idx, vg_map_stats* stats);

/* MapperMSA.mapAllReads [MapperMSA.java:274]: same pairing logic in MSA column
 * coordinates, no alignment; score = k-mer count. */
int vg_map_pairs_msa(vg_ctx* ctx, vg_map_stats* stats);

/* Number of results of the last map call and the bytes of all sequence1. */
int vg_result_sizes(vg_ctx* ctx, int64_t* n_results, int64_t* seq1_bytes);

/* Copies the per-pair results (input order) and the sequence1 byte pool of the
 * last map call to the host. seq1_offset values index `seq1_pool`. */
int vg_get_results(vg_ctx* ctx, vg_pair_result* out, int64_t n_out, uint8_t* seq1_pool, int64_t pool_cap);

/* vg_get_results without the sequence1 bytes ("an edit string from which sequence1 is reconstructible"): in
 * the records written to `out`, seq1_offset is the index of the mate's first element in `ops` and pad_ their number; an
 * element is len << 4 | op with op 0 = alignment columns holding a read base (DIAGONAL), 1 = lower-cased read bases (UP:
 * insertion), 2 = '-' (LEFT: deletion) [SmithWatermanGotoh.java:157-222]. sequence1 = the read in the mate's orientation
 * (reverse: Constants.getComplement) from start1 on, copied / lower-cased (Constants.lower) / skipped per element. The
 * device keeps sequence1, so vg_pileup / vg_bam_fields / vg_identity_counting are unaffected. VG_ERR_CAPACITY with
 * vg_last_required = the number of elements when ops_cap is too small (*n_ops is set either way). */
int vg_get_results_compact(vg_ctx* ctx, vg_pair_result* out, int64_t n_out, uint32_t* ops, int64_t ops_cap, int64_t* n_ops);

/* Installs a MappedData held by the host as the context's "last mapping": per-pair records (with seq1_offset / length
 * addressing seq1_pool), and for each record the index of its pair in the resident read set. This is what lets the Java
 * host edit mappedData between a mapping and its consumers the way ConsensusBuilderWithReassembling does — markRemapReads
 * / adjustCoord rewrite the kept reads' coordinates and Mapper.mapReads(MappedData, Reference) [Mapper.java:477-523]
 * appends the remapped pairs to the same lists — and still run getConsensusSimple (vg_pileup), BamPrinter
 * (vg_bam_fields) and ProblemIntervalBuilder (vg_identity_counting) over mappedData.mappedReads on the device.
 * mappedReads order = record order, r1 before r2. Requires vg_set_reference and vg_upload_reads. */
int vg_set_results(vg_ctx* ctx, const vg_pair_result* records, int64_t n, const int64_t* pair_idx,
                   const uint8_t* seq1_pool, int64_t pool_bytes);

/* ConsensusBuilderSimple.ReadCounting [ConsensusBuilderSimple.java:56-75] +
 * the partial-matrix sum [:138-145] over the mapped reads of the context's last mapping (vg_map_pairs or
 * vg_set_results). accumulate != 0 adds to the counts of the previous vg_pileup on the same reference instead of
 * starting from zero; it fails with VG_ERR_STATE when there are none (vg_set_reference discards the counts). Counts
 * are int32[G][5] (A,T,G,C,gap). The device buffer is exposed so that a caller that runs its own collective can
 * all-reduce it across GPUs (the multi-GPU replacement of the serial sum); see also vg_comm_* below. */
int vg_pileup(vg_ctx* ctx, int32_t accumulate);
int vg_pileup_counts_device(vg_ctx* ctx, void** dev_ptr, int64_t* n_int32);
int vg_pileup_counts_host(vg_ctx* ctx, int32_t* out, int64_t n_int32);

/* ConsensusBuildingSimple [ConsensusBuilderSimple.java:98-120] on the current
 * (possibly all-reduced) counts. */
int vg_consensus(vg_ctx* ctx, int32_t save_uncovered, int32_t coverage_threshold, const uint8_t* genome, int32_t G,
                 uint8_t* out);

/* Aligner.align [Aligner.java:14, SmithWatermanGotoh.java:51] over a batch of
 * independent (s1, s2) pairs given as byte pools + offsets (n+1 entries each).
 * rec[i] = {score,start1,end1,start2,end2,identity,length}; sequence1 of task
 * i is written at seq1_pool + seq1_off[i] (caller provides len1+len2 bytes per
 * task). Used by tests and by the "next" callers. */
int vg_sw_align_batch(vg_ctx* ctx, const uint8_t* s1_pool, const int64_t* s1_off, const uint8_t* s2_pool,
                      const int64_t* s2_off, int64_t n, int32_t* rec, uint8_t* seq1_pool, const int64_t* seq1_off);

/* Aligner.findMaxScore [Aligner.java:16, SmithWatermanGotoh.java:240]. */
int vg_sw_max_score_batch(vg_ctx* ctx, const uint8_t* s1_pool, const int64_t* s1_off, const uint8_t* s2_pool,
                          const int64_t* s2_off, int64_t n, int32_t* score);

/* KMerCounter.getNBestRegions [KMerCounter.java:223] /
 * KMerCounterMSA.getNBestRegions [KMerCounterMSA.java:71] for a batch of
 * single reads (no pairing): windows[i] = up to max_windows (start,end,count)
 * triples, n_windows[i] their number. */
int vg_seed_batch(vg_ctx* ctx, int32_t msa, const uint8_t* pool, const int64_t* off, int64_t n, int32_t max_windows,
                  int32_t* windows, int32_t* n_windows);

/* KMerCounter.mapReadToRegion over the whole reference (KMerCounter.java:119-140; from = 0, to = reference / MSA length)
 * and KMerCounterBase.computeKMerCount (:468-487, what RandomModel.java:122-128 scores its random reads with) for a
 * batch of reads in their given orientation: windows[i] = {start, end, count} of the first window of maximal count
 * (no cutoff, no splitLongHits, no pruning), found[i] = 0 when the read has no k-mer hit (mapReadToRegion returns
 * null, computeKMerCount returns 0). SURVEY 8f N1 (seeding part) / N3. */
int vg_best_window_batch(vg_ctx* ctx, int32_t msa, const uint8_t* pool, const int64_t* off, int64_t n, int32_t* windows,
                         int32_t* found);

/* RandomModel.RandomCountsComputing.run (RandomModel.java:113-128): n_reads random reads of read_len bases drawn from the
 * order-`order` Markov model of `genomes` (MarkovModel.java:18-119; genomes = byte pool + n_genomes + 1 offsets: the
 * contigs of a fragmented reference, else the whole sequence, RandomModel.java:88-98; for the MSA model the ungapped
 * references), each scored with KMerCounterBase.computeKMerCount (:468-487) against the context's reference (msa = 0)
 * or MSA index (msa = 1): counts[r] = the score (unsorted). Generation and scoring run on the device; reads_out
 * (optional, n_reads x read_len bytes) receives the reads. RNG policy (Q10): the reference creates an unseeded
 * java.util.Random per read (MarkovModel.java:95), so there is no reference stream to reproduce; read r draws from
 * SplitMix64 started at state seed + (first_draw + r * (2 + read_len - order)) * 0x9E3779B97F4A7C15, i.e. the reads are
 * the ones a single sequential SplitMix64(seed) stream yields when read r starts at draw first_draw + r * (2 + read_len -
 * order). The context's lowCoef / highCoef are used as they are: RandomModel's own counter has lowCoef = 1 / highCoef
 * (KMerCounter.java:51-55, Q3), so the host creates the context it scores with accordingly. SURVEY 8f N3. */
int vg_random_model_counts(vg_ctx* ctx, int32_t msa, const uint8_t* genomes, const int64_t* genome_off, int32_t n_genomes,
                           int32_t order, uint64_t seed, int64_t first_draw, int32_t read_len, int64_t n_reads,
                           int32_t* counts, uint8_t* reads_out);

/* KMerCounter.mapReadToRegion (KMerCounter.java:119-140) for (read, region) pairs where every region is its own small
 * reference with its own StringIndexer.buildIndex index: ContigBuilder.countReads against cluster centroids
 * (ContigBuilder.java:72-104, 219-227), Graph.SWScore against reference sub-sequences (Graph.java:493-550). Regions and
 * reads are byte pools + offsets (n+1 entries); read_region[i] names the region of read i; windows[i] = {start, end,
 * count} inside that region, found[i] = 0 when mapReadToRegion returns null. The alignment of the read against
 * region[start, end) is vg_sw_align_batch. SURVEY 8f N1. Uses K and the coefficients of vg_ctx_create. */
int vg_map_to_regions(vg_ctx* ctx, const uint8_t* region_pool, const int64_t* region_off, int64_t n_regions,
                      const uint8_t* read_pool, const int64_t* read_off, int64_t n_reads, const int32_t* read_region,
                      int32_t* windows, int32_t* found);

/* The pileup loop of ContigBuilder.countReads (ContigBuilder.java:93-99) for a batch of alignments: alignment i (sequence1 at
 * seq1_pool + seq1_off[i], `length[i]` columns) starts at position pos[i] (= region base + MappedRead.start +
 * Alignment.start2 when the regions are laid out back to back) and must stay below end[i] (= the end of its own region: the
 * Java indexes that region's own counts array and throws beyond it; end == NULL means total_positions for all);
 * counts = int32[total_positions][5] (A, T, G, C, gap), zeroed by the call. An alignment that crosses its end, or holds a
 * byte in ('T', 'a'), fails the call with VG_ERR_ALPHABET. */
int vg_pileup_alignments(vg_ctx* ctx, const uint8_t* seq1_pool, const int64_t* seq1_off, const int32_t* length,
                         const int64_t* pos, const int64_t* end, int64_t n, int64_t total_positions, int32_t* counts);

/* ProblemIntervalBuilder.countIdentity (ProblemIntervalBuilder.java:35-160): for every mapped read of the
 * last vg_map_pairs call (mappedData.mappedReads = r1, r2 of every pair in order) whose length is at least window_size,
 * the identity of every window of window_size read bases against the reference the reads were mapped to, added as
 * (float)identity/len to identity[window start] with coverage[window start]++. The sums are binary32 and keep the
 * reference's order: thread t of thread_num sums reads [t*size, (t+1)*size) (size = nReads / thread_num, the last thread
 * takes the rest) in list order, and the per-thread arrays are added in thread order. identity / coverage: G entries. */
int vg_identity_counting(vg_ctx* ctx, int32_t window_size, int32_t thread_num, float* identity, int32_t* coverage);

/* BAM record fields of the last single-reference mapping, replacing the per-read derivations of
 * BamPrinter.setRecordFieldsMapped / setRecordFieldsUnmapped / setCigar (BamPrinter.java:16-171); serialisation stays with
 * htsjdk. Two records per pair result, in (pair, mate) order. ref_index is the index into Reference.contigEnds of the
 * contig holding the alignment start (the shim applies BamPrinter's contigToRefIndex for fragmented references). CIGAR
 * elements are len << 4 | op with the BAM op codes M=0, I=1, D=2, S=4; like the Java, an alignment that starts with an
 * insertion or deletion carries a leading 0M. Unmapped reads have n_cigar = 0 ("*"), ref_index = -1, pos = 0, mapq = 0. */
typedef struct {
  int32_t flag;            /* 0x1 paired | 0x2 proper pair | 0x4 unmapped | 0x8 mate unmapped | 0x10 reverse |
                              0x20 mate reverse | 0x40 first | 0x80 second */
  int32_t ref_index, pos;  /* pos is 1-based inside the contig (record.setAlignmentStart) */
  int32_t mapq;            /* 255 mapped, 0 unmapped */
  int32_t mate_ref_index, mate_pos;
  int32_t xn;              /* XN attribute: aligned read bases - identity */
  int32_t n_cigar;
  int64_t cigar_offset;    /* first element in the pool */
} vg_bam_record;
int vg_bam_fields(vg_ctx* ctx, vg_bam_record* out, int64_t n_records, uint32_t* cigar_pool, int64_t pool_cap, int64_t* n_ops);

/* ---- Multi-GPU ------------------------------------------------------------------------------------------
 * Read pairs shard over the GPUs in contiguous blocks, reference / index / cutoffs are replicated; the only exchange is
 * the integer sum of the pileup counts, which replaces the serial sum of the per-thread matrices at
 * ConsensusBuilderSimple.java:138-145. NCCL (libnccl.so.2) is loaded at first use.
 *
 * (a) one process per GPU: rank 0 calls vg_comm_unique_id, the host ships the VG_COMM_ID_BYTES bytes to the other ranks by
 *     any channel it has, every rank calls vg_comm_init_rank on its context; from then on vg_pileup_allreduce after vg_pileup
 *     leaves the global counts on every rank (in place, on the context's stream). */
#define VG_COMM_ID_BYTES 128
int vg_comm_unique_id(uint8_t* id128);
int vg_comm_init_rank(vg_ctx* ctx, const uint8_t* id128, int rank, int n_ranks);
/* (b) one process, n contexts on n distinct devices (ncclCommInitAll): then vg_pileup_allreduce must be called for all n
 *     contexts concurrently (one host thread each) — or use vg_multi_*, which does exactly that. */
int vg_comm_init_all(vg_ctx** ctxs, int n);
void vg_comm_destroy(vg_ctx* ctx);
int vg_comm_size(const vg_ctx* ctx);             /* 0 when the context has no communicator */
int vg_pileup_allreduce(vg_ctx* ctx);            /* no-op without a communicator */
int vg_last_allreduce_ms(vg_ctx* ctx, double* ms);

/* (c) vg_multi: `new Mapper(document)` bound to n_dev GPUs of this process — the `vg_ctx_create(params, device_ids[], n_dev)`
 *     of SURVEY 8b, for a host that is one process (one JVM). Same call sequence and semantics as the single-context entry
 *     points of the same name; results come back in input order (the order of pair_idx for a subset), the sequence1 pools of
 *     the devices laid out one after the other. vg_multi_pileup = vg_pileup on every device + the all-reduce, so that
 *     vg_multi_consensus / vg_multi_pileup_counts_host see the counts of all reads. vg_multi_ctx(m, i) exposes the i-th
 *     device's context for the per-context queries (timings, kernel launches, BAM fields of its shard, ...). */
typedef struct vg_multi vg_multi;
int vg_multi_create(const vg_params* params, const int32_t* device_ids, int32_t n_dev, vg_multi** out);
void vg_multi_destroy(vg_multi* m);
const char* vg_multi_last_error(const vg_multi* m);
int64_t vg_multi_last_required(const vg_multi* m);
int32_t vg_multi_size(const vg_multi* m);
vg_ctx* vg_multi_ctx(vg_multi* m, int32_t i);
int vg_multi_set_reference(vg_multi* m, const uint8_t* seq, int32_t G, const int32_t* contig_ends, int32_t n_contigs,
                           const int32_t* fragment_ends, int32_t n_fragments, int32_t is_fragmented, const uint8_t* keys,
                           const int64_t* offsets, const int32_t* positions, int32_t n_keys, int32_t index_thread_num);
int vg_multi_set_msa_index(vg_multi* m, int32_t K, float low_coef, float high_coef, const int32_t* cutoffs,
                           int32_t n_cutoffs, const uint8_t* keys, const int64_t* offsets, const int32_t* columns,
                           int32_t n_keys, int32_t msa_length);
int vg_multi_upload_reads(vg_multi* m, const uint8_t* bases, int64_t n_bases, const int64_t* off1, const int32_t* len1,
                          const int64_t* off2, const int32_t* len2, int64_t n_pairs);
int vg_multi_upload_reads_packed(vg_multi* m, const uint8_t* packed, int64_t n_bases, const int64_t* exc_pos,
                                 const uint8_t* exc_byte, int64_t n_exc, const int64_t* off1, const int32_t* len1,
                                 const int64_t* off2, const int32_t* len2, int64_t n_pairs);
int vg_multi_map_pairs(vg_multi* m, const int64_t* pair_idx, int64_t n_idx, vg_map_stats* stats);
int vg_multi_map_pairs_msa(vg_multi* m, vg_map_stats* stats);
int vg_multi_result_sizes(vg_multi* m, int64_t* n_results, int64_t* seq1_bytes);
int vg_multi_get_results(vg_multi* m, vg_pair_result* out, int64_t n_out, uint8_t* seq1_pool, int64_t pool_cap);
int vg_multi_get_results_compact(vg_multi* m, vg_pair_result* out, int64_t n_out, uint32_t* ops, int64_t ops_cap, int64_t* n_ops);
int vg_multi_pileup(vg_multi* m, int32_t accumulate);
int vg_multi_pileup_counts_host(vg_multi* m, int32_t* out, int64_t n_int32);
int vg_multi_consensus(vg_multi* m, int32_t save_uncovered, int32_t coverage_threshold, const uint8_t* genome, int32_t G,
                       uint8_t* out);

/* Timing of the device stages of the last map / pileup call, in milliseconds
 * (CUDA events on the context's stream): [0]=seeding+pair select, [1]=SW fill,
 * [2]=traceback, [3]=pileup, [4]=total, [5]=SW cells (as a double count). */
int vg_last_timings(vg_ctx* ctx, double* out6);
/* Seeding kernels of the last map call (summed over its chunks): [0]=LongHit-stream kernel ms, [1]=window kernel ms,
 * [2]=general (fallback) kernel ms, [3]=LongHits produced by mergeHits phase 1, [4]=LongHits kept by phase 2,
 * [5]=read orientations redone by the general kernel, [6]=windows written, [7]=read orientations seeded. */
int vg_seed_timings(vg_ctx* ctx, double* out8);

/* Measurement helpers (bench.py): a DPX issue-rate microbenchmark that defines
 * the denominator of the SW roofline (packed s16x2 ops/s), and counters of the
 * kernels launched by this context. */
int vg_dpx_peak(vg_ctx* ctx, double* ops_per_s);
int64_t vg_kernel_launches(const vg_ctx* ctx);
/* The CUDA stream (cudaStream_t) all of this context's kernels and copies are issued on, so that a
 * caller can bracket calls with its own CUDA events. */
int vg_ctx_stream(vg_ctx* ctx, void** stream);

#ifdef __cplusplus
}
#endif
#endif /* VIRGENA_GPU_H */

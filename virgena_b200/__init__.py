"""virgena_b200 — B200 (sm_100a) implementation of VirGenA's read-mapping hot path.

Product code: csrc/ (CUDA kernels + the C ABI of include/virgena_gpu.h) and the host-side mirror of
the reference's Mapper / MapperMSA / ConsensusBuilderSimple API (mapper.py). Nothing in this package
imports oracle/ (test infrastructure), and nothing falls back to the CPU.
"""
from ._lib import (CLS_CONCORDANT_FR, CLS_CONCORDANT_RF, CLS_DISCORDANT, CLS_LEFT, CLS_RIGHT, CLS_UNMAPPED,
                   BAM_DTYPE, EXPORTED_SYMBOLS, LIB_PATH, MATE_DTYPE, PAIR_DTYPE, VgError)
from .mapper import (BamPrinter, ConsensusBuilderSimple, ContigBuilder, Context, Graph, MappedData, Mapper, MapperMSA, MultiContext,
                     PairedReads,
                     ProblemIntervalBuilder, RandomModel, Reference, ReferenceAlignment, ReferenceFinder, cigar_string, mark_remap_reads)

__all__ = ["Context", "MultiContext", "Mapper", "MapperMSA", "ConsensusBuilderSimple", "Reference", "ReferenceAlignment",
           "BamPrinter", "ContigBuilder", "Graph", "ReferenceFinder", "ProblemIntervalBuilder", "RandomModel", "BAM_DTYPE", "PairedReads", "MappedData", "VgError", "PAIR_DTYPE", "MATE_DTYPE", "EXPORTED_SYMBOLS", "LIB_PATH", "mark_remap_reads"]

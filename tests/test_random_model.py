"""RandomModel / RandomModelMSA + KMerCounterBase.computeCuttofs. CPU: the mirror's cutoff formula against the
oracle's; GPU: reads generated and scored on the device (vg_random_model_counts) against the oracle's seeded restatement —
the same reads byte for byte, the same sorted scores, the same cutoffs."""
import numpy as np
import pytest

import oracle
from virgena_b200 import synth


def test_compute_cuttofs_mirror_matches_oracle(built):
    import virgena_b200 as vg
    ref, ends = synth.config_reference(1)
    for (mn, mx, step, p, rn) in ((150, 250, 50, 0.01, 300), (100, 250, 10, 0.2, 120), (250, 250, 10, 0.01, 200), (60, 130, 7, 0.5, 150)):
        cut, counts, _ = oracle.random_model(ref, K=5, read_num=rn, min_len=mn, max_len=mx, step=step, p_value=p, seed=3,
                                             index_thread_num=4)
        assert (np.diff(counts, axis=1) >= 0).all()
        mine = vg.RandomModel.computeCuttofs(mn, mx, step, counts, p)
        assert (mine == cut).all(), (mn, mx, step, p)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", [1, 4])
def test_random_model_single_reference(built, cfg):
    import virgena_b200 as vg
    ref, ends = synth.config_reference(cfg)
    frag = len(ends) > 1
    genome = vg.Reference(ref, K=5, contig_ends=ends, fragment_ends=ends, is_fragmented=frag, thread_num=8)
    for seed, (mn, mx, step, rn) in ((1, (150, 250, 50, 1500)), (77, (240, 250, 10, 700))):
        cut, counts, reads = oracle.random_model(ref, K=5, coef=1.25, order=4, read_num=rn, min_len=mn, max_len=mx, step=step,
                                                 p_value=0.01, contig_ends=ends, is_fragmented=frag, index_thread_num=8, seed=seed)
        rm = vg.RandomModel(genome, K=5, order=4, coef=1.25, seed=seed)
        c0, r0 = rm.ctx.random_model_counts(rm.genomes, 4, seed, 0, mx, rn, want_reads=True)
        assert r0.tobytes() == reads.tobytes()
        assert (np.sort(c0) == counts[-1]).all()
        mine = rm.cutoffs(readNum=rn, minReadLen=mn, maxReadLen=mx, step=step, pValue=0.01)
        assert (rm.counts == counts).all()
        assert (mine == cut).all()


@pytest.mark.gpu
def test_random_model_msa(built):
    import virgena_b200 as vg
    rows = synth.config_msa(n_refs=12)
    aln = vg.ReferenceAlignment(rows, K=7)
    mn, mx, step, rn = 200, 250, 25, 800
    cut, counts, reads = oracle.random_model(rows[0], K=7, coef=1.5, order=4, read_num=rn, min_len=mn, max_len=mx, step=step,
                                             p_value=0.01, seed=5, msa_rows=rows)
    rm = vg.RandomModel(aln, K=7, order=4, coef=1.5, seed=5)
    c0, r0 = rm.ctx.random_model_counts(rm.genomes, 4, 5, 0, mx, rn, msa=True, want_reads=True)
    assert r0.tobytes() == reads.tobytes()
    mine = rm.cutoffs(readNum=rn, minReadLen=mn, maxReadLen=mx, step=step, pValue=0.01)
    assert (rm.counts == counts).all()
    assert (mine == cut).all()


@pytest.mark.gpu
def test_random_model_rejects_bad_arguments(built):
    import virgena_b200 as vg
    ref, ends = synth.config_reference(1)
    rm = vg.RandomModel(vg.Reference(ref, K=5, thread_num=1), K=5)
    with pytest.raises(vg.VgError):
        rm.ctx.random_model_counts([b"ACG"], 4, 1, 0, 100, 10)       # genome shorter than the order
    with pytest.raises(vg.VgError):
        rm.ctx.random_model_counts(rm.genomes, 4, 1, 0, 3, 10)       # read shorter than the order
    assert len(rm.ctx.random_model_counts(rm.genomes, 4, 1, 0, 100, 0)) == 0

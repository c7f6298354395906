"""Seeded synthetic inputs for the BASELINE.json configs.

The reference's own data (HIV toy set, HXB2, subtype MSA) ship only in the
GitHub release zip and are not available offline, so every config is restated
as a deterministic generator. Pure numpy; no device code here.
"""
import numpy as np

_ACGT = np.frombuffer(b"ATGC", dtype=np.uint8)
_COMP = np.zeros(256, np.uint8)
for _a, _b in (b"AT", b"TA", b"GC", b"CG", b"NN", b"--"):
    _COMP[_a] = _b


def revcomp(a):
    return _COMP[np.asarray(a, np.uint8)[..., ::-1]]


def markov_reference(G, seed, order=3):
    """An "HIV-like" sequence: A-rich order-`order` Markov chain over ATGC."""
    rng = np.random.default_rng(seed)
    base = np.array([0.36, 0.22, 0.24, 0.18])  # A,T,G,C composition of HIV-1
    trans = rng.dirichlet(base * 12.0, size=4 ** order)
    cum = np.cumsum(trans, axis=1)
    u = rng.random(G)
    out = np.zeros(G, np.int64)
    out[:order] = rng.integers(0, 4, order)
    state = 0
    for i in range(order):
        state = state * 4 + int(out[i])
    mask = 4 ** order
    for i in range(order, G):
        c = int(np.searchsorted(cum[state], u[i]))
        if c > 3:
            c = 3
        out[i] = c
        state = (state * 4 + c) % mask
    return _ACGT[out]


def mutate(seq, sub_rate, indel_rate, rng):
    """Substitutions + short indels over a whole sequence (for strains / MSA rows)."""
    seq = np.asarray(seq, np.uint8)
    out = []
    i = 0
    n = len(seq)
    u = rng.random(n)
    while i < n:
        if u[i] < indel_rate / 2:  # deletion
            i += int(rng.geometric(0.5))
            continue
        if u[i] < indel_rate:  # insertion
            out.extend(_ACGT[rng.integers(0, 4, int(rng.geometric(0.5)))].tolist())
        b = seq[i]
        if rng.random() < sub_rate:
            b = _ACGT[(int(np.where(_ACGT == b)[0][0]) + int(rng.integers(1, 4))) % 4] if b in _ACGT else b
        out.append(int(b))
        i += 1
    return np.array(out, np.uint8)


def _reads_from(src, starts, L, sub_rate, indel_rate, n_rate, rng):
    """Vectorised read sampler: reads of length L from `src` starting at `starts`, each with its own
    substitution rate `sub_rate[i]`; indel events are geometric(0.5)-length."""
    n = len(starts)
    u = rng.random((n, L), dtype=np.float32)
    adv = np.ones((n, L), np.int32)
    ins = np.zeros((n, L), bool)
    if indel_rate > 0:
        ev_del = u < indel_rate / 2
        ev_ins = (u >= indel_rate / 2) & (u < indel_rate)
        glen = rng.geometric(0.5, size=(n, L)).astype(np.int32)
        adv += np.where(ev_del, glen, 0)
        # an insertion of length d at t: bases t..t+d-1 are random and do not consume the source
        ins = ev_ins.copy()
        for d in (1, 2):
            m = ev_ins & (glen > d)
            ins[:, d:] |= m[:, :-d]
        adv = np.where(ins, 0, adv)
    idx = starts[:, None] + np.cumsum(adv, axis=1) - adv[:, :1]
    np.clip(idx, 0, len(src) - 1, out=idx)
    reads = src[idx]
    if ins.any():
        reads = np.where(ins, _ACGT[rng.integers(0, 4, (n, L))], reads)
    sub = rng.random((n, L), dtype=np.float32) < sub_rate[:, None]
    if sub.any():
        code = np.zeros(256, np.uint8)
        code[_ACGT] = np.arange(4, dtype=np.uint8)
        shifted = _ACGT[(code[reads] + rng.integers(1, 4, (n, L), dtype=np.uint8)) % 4]
        reads = np.where(sub, shifted, reads)
    if n_rate > 0:
        reads = np.where(rng.random((n, L), dtype=np.float32) < n_rate, np.uint8(ord("N")), reads)
    return reads.astype(np.uint8)


def make_pairs(sources, n_pairs, L, seed, sub_lo=0.0, sub_hi=0.0, indel_rate=0.0, n_rate=0.0,
               frag_mean=400, frag_sd=80, frag_max=1000, segments=None):
    """Paired reads (n_pairs, 2, L) drawn from `sources` (list of sequences; one is picked per pair).
    Mate 1 is the fragment's left end on a random strand, mate 2 the reverse complement of the other
    end (Illumina FR). `segments` = contig ends: fragments never span a segment boundary."""
    rng = np.random.default_rng(seed)
    m1 = np.zeros((n_pairs, L), np.uint8)
    m2 = np.zeros((n_pairs, L), np.uint8)
    which = rng.integers(0, len(sources), n_pairs)
    for si, src in enumerate(sources):
        sel = np.where(which == si)[0]
        if len(sel) == 0:
            continue
        src = np.asarray(src, np.uint8)
        n = len(sel)
        frag = np.clip(rng.normal(frag_mean, frag_sd, n).astype(np.int64), L, frag_max)
        if segments is None:
            lo = np.zeros(n, np.int64)
            hi = np.full(n, len(src), np.int64)
        else:
            ends = np.asarray(segments, np.int64)
            begs = np.concatenate([[0], ends[:-1]])
            seg = rng.choice(len(ends), n, p=(ends - begs) / float(ends[-1]))
            lo, hi = begs[seg], ends[seg]
        frag = np.minimum(frag, hi - lo)
        frag = np.maximum(frag, np.minimum(L, hi - lo))
        start = lo + (rng.random(n) * (hi - lo - frag + 1)).astype(np.int64)
        sub = rng.uniform(sub_lo, sub_hi, n).astype(np.float32) if sub_hi > 0 else np.zeros(n, np.float32)
        left = _reads_from(src, start, L, sub, indel_rate, n_rate, rng)
        right = revcomp(_reads_from(src, np.maximum(start + frag - L, lo), L, sub, indel_rate, n_rate, rng))
        flip = rng.random(n) < 0.5
        m1[sel] = np.where(flip[:, None], right, left)
        m2[sel] = np.where(flip[:, None], left, right)
    return m1, m2


def pack_pairs(m1, m2):
    """Byte pool + offsets in the layout of vg_upload_reads."""
    n, L1 = m1.shape
    L2 = m2.shape[1]
    bases = np.concatenate([m1.reshape(-1), m2.reshape(-1)])
    off1 = np.arange(n, dtype=np.int64) * L1
    off2 = n * L1 + np.arange(n, dtype=np.int64) * L2
    return bases, off1, np.full(n, L1, np.int32), off2, np.full(n, L2, np.int32)


def pack_ragged(reads1, reads2):
    """Same for lists of byte strings of differing lengths ('*' = absent mate)."""
    parts, off1, len1, off2, len2 = [], [], [], [], []
    pos = 0
    for r1, r2 in zip(reads1, reads2):
        for r, off, ln in ((r1, off1, len1), (r2, off2, len2)):
            b = np.frombuffer(bytes(r), np.uint8)
            parts.append(b)
            off.append(pos)
            ln.append(len(b))
            pos += len(b)
    bases = np.concatenate(parts) if parts else np.zeros(0, np.uint8)
    return bases, np.array(off1, np.int64), np.array(len1, np.int32), np.array(off2, np.int64), np.array(len2, np.int32)


# ---------------------------------------------------------------------------
# BASELINE.json configs restated; seed = 20260101 + config
# ---------------------------------------------------------------------------
SEED0 = 20260101
HIV_G = 9719
FLU_SEGMENTS = [2341, 2341, 2233, 1778, 1565, 1413, 1027, 890]


def config_reference(cfg):
    if cfg in (1, 2, 3):
        return markov_reference(HIV_G, SEED0 + 1), [HIV_G]
    if cfg == 4:
        seq = markov_reference(sum(FLU_SEGMENTS), SEED0 + 4)
        return seq, list(np.cumsum(FLU_SEGMENTS))
    if cfg == 5:
        return markov_reference(9600, SEED0 + 5), [9600]
    raise ValueError(cfg)


def config_pairs(cfg, n_pairs, seed_offset=0, L=None):
    """Reads of config `cfg` (1..5). n_pairs is a parameter so tests can take small samples."""
    ref, ends = config_reference(cfg)
    seed = SEED0 + cfg + 1000 * (seed_offset + 1)
    if cfg in (1, 2):
        rng = np.random.default_rng(SEED0 + 11)
        strains = [mutate(ref, d, 0.005, rng) for d in (0.05, 0.10, 0.15)]
        return make_pairs(strains, n_pairs, L or 250, seed, n_rate=0.001)
    if cfg == 3:
        return make_pairs([ref], n_pairs, L or 250, seed, sub_lo=0.15, sub_hi=0.25, indel_rate=0.01)
    if cfg == 4:
        return make_pairs([ref], n_pairs, L or 150, seed, sub_lo=0.03, sub_hi=0.03, segments=ends)
    if cfg == 5:
        rng = np.random.default_rng(SEED0 + 55)
        genos = [ref] + [mutate(ref, 0.30, 0.002, rng) for _ in range(5)]
        return make_pairs(genos, n_pairs, L or 150, seed)
    raise ValueError(cfg)


def config_msa(n_refs=40, seed=SEED0 + 2):
    """Stand-in subtype MSA for config 2: rows derived from the config-1 reference by a random tree,
    gapped to equal length. Returns uint8 [n_refs, alnLength]."""
    ref, _ = config_reference(1)
    rng = np.random.default_rng(seed)
    # each row = list of (column_key, base); derive rows by mutating parents while tracking columns
    rows = [[(float(i), int(b)) for i, b in enumerate(ref)]]
    while len(rows) < n_refs:
        parent = rows[int(rng.integers(0, len(rows)))]
        div = float(rng.uniform(0.05, 0.20)) / 2
        child = []
        i = 0
        while i < len(parent):
            u = rng.random()
            if u < 0.002:  # deletion
                i += int(rng.geometric(0.4))
                continue
            key, b = parent[i]
            if u < 0.004:  # insertion: new columns right after this one
                child.append((key, b))
                nxt = parent[i + 1][0] if i + 1 < len(parent) else key + 1.0
                d = int(rng.geometric(0.5))
                for t in range(d):
                    child.append((key + (nxt - key) * (t + 1) / (d + 1), int(_ACGT[rng.integers(0, 4)])))
                i += 1
                continue
            if rng.random() < div:
                b = int(_ACGT[rng.integers(0, 4)])
            child.append((key, b))
            i += 1
        rows.append(child)
    keys = sorted({k for r in rows for k, _ in r})
    col = {k: i for i, k in enumerate(keys)}
    aln = np.full((n_refs, len(keys)), ord("-"), np.uint8)
    for r, row in enumerate(rows):
        for k, b in row:
            aln[r, col[k]] = b
    return aln

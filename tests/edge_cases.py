"""Hand-crafted BAM fixtures for what the reference's own fixture never exercises (last bullet)."""
import os

import numpy as np

import bamwriter as W
from bamsignals_b200 import GRanges

REFS = [("chrA", 200000), ("chrB", 50000), ("chrEmpty", 1000)]

CIGARS = ["50M", "5S40M5S", "10H30M", "20M5D25M", "20M5I25M", "10M1000N40M", "25=5X20=", "3S10M2P10M1D5M",
          "*", "100S", "7I", "1M", "30M20000N30M"]


def variety_reads(seed=0, n=3000):
    rng = np.random.default_rng(seed)
    reads = []
    for tid, (_, ln) in enumerate(REFS[:2]):
        pos = np.sort(rng.integers(0, ln - 25000, n))
        for p in pos:
            cg = CIGARS[rng.integers(0, len(CIGARS))]
            flag = int(rng.choice([0, 16, 99, 147, 83, 163, 1024 + 99, 4, 4 + 16, 73, 133, 2048 + 16, 256]))
            tl = int(rng.integers(-600, 600))
            reads.append(dict(tid=tid, pos=int(p), flag=flag, mapq=int(rng.integers(0, 61)), cigar=cg, tlen=tl))
    for _ in range(17):
        reads.append(dict(tid=-1, pos=-1, flag=4, mapq=0, cigar="*", tlen=0))
    return reads


def variety_regions(seed=1, n=120):
    rng = np.random.default_rng(seed)
    names = [REFS[i][0] for i in rng.integers(0, 3, n)]
    start = rng.integers(1, 60000, n)
    start = np.where(np.array(names) == "chrEmpty", rng.integers(1, 900, n), start)
    width = rng.integers(0, 30000, n)
    return GRanges(names, start, width, rng.choice(["+", "-", "*"], n).tolist())


def write_variety(path, **kw):
    reads = variety_reads()
    W.write_bam(path, REFS, reads, **kw)
    return reads


def long_cigar_read():
    """A record whose real CIGAR (> 65535 ops) lives in a CG:B,I tag behind the 'kSmN' placeholder (SAM spec 4.2.2):
    bam_endpos sees only the placeholder, whose N length preserves the reference length."""
    import struct
    n_ops = 70000
    real = [(1, 0), (1, 2)] * (n_ops // 2)          # 1M1D repeated: 70000 reference bases, 35000 query bases
    l_seq = 35000
    aux = b"CGBI" + struct.pack("<i", len(real)) + b"".join(struct.pack("<I", n << 4 | op) for n, op in real)
    return dict(tid=0, pos=1000, flag=0, mapq=40, cigar=[(l_seq, 4), (70000, 3)], tlen=0, l_seq=l_seq, aux=aux)

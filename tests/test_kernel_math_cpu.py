"""Two pieces of integer arithmetic the CUDA kernels rely on, restated in Python and checked exhaustively enough on the
CPU (the kernels themselves are checked end to end by the GPU parity tests):

* k_crc32 (bamsignals_b200/csrc/inflate.cu): a block's CRC32 from 32 independently computed pieces, folded in a
  five-level tree with carry-less multiplications by x^(8 * piece bytes) mod P - must equal zlib.crc32 for any length;
  the 20 constants c_x2n[] are x^(2^k) mod P.
* k_profile_agg (kernels.cu): n / binsize as umulhi(n, m) >> (L - 1) with m = ceil(2^(31+L) / d), L = ceil(log2 d) -
  must be exact for every 0 <= n < 2^31."""
import os
import random
import re
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
POLY = 0xEDB88320


def crc_mul(a, b):
    """a * b mod P, reflected representation (bit 31 = x^0): the kernel's crc_mul / zlib's multmodp"""
    p = 0
    for _ in range(32):
        if a & 0x80000000:
            p ^= b
        a = (a << 1) & 0xFFFFFFFF
        b = (b >> 1) ^ (POLY if b & 1 else 0)
    return p


def x2n_table():
    t = [0x40000000]
    for _ in range(19):
        t.append(crc_mul(t[-1], t[-1]))
    return t


def test_x2n_constants_in_the_kernel_source():
    src = open(os.path.join(ROOT, "bamsignals_b200", "csrc", "inflate.cu")).read()
    body = re.search(r"c_x2n\[20\]\s*=\s*\{([^}]*)\}", src).group(1)
    consts = [int(x.rstrip("u"), 16) for x in re.findall(r"0x[0-9a-fA-F]+u?", body)]
    assert consts == x2n_table()


def test_crc_fold_equals_zlib():
    tab = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ (POLY if c & 1 else 0)
        tab.append(c)

    def reg(init, data):
        for b in data:
            init = tab[(init ^ b) & 0xFF] ^ (init >> 8)
        return init

    x2n = x2n_table()
    rnd = random.Random(7)
    lengths = [0, 1, 3, 31, 32, 127, 128, 129, 131, 255, 4096, 65279, 65280, 65536] + [rnd.randrange(0, 66000) for _ in range(40)]
    for ln in lengths:
        data = bytes(rnd.getrandbits(8) for _ in range(ln))
        seg = (ln // 32) & ~3                       # lanes 1..31: seg bytes each; lane 0: the remaining head
        head = ln - 31 * seg
        regs = [reg(0xFFFFFFFF, data[:head])] + [reg(0, data[head + (k - 1) * seg: head + k * seg]) for k in range(1, 32)]
        X, bits, k = 0x80000000, 8 * seg, 0
        while bits:
            if bits & 1:
                X = crc_mul(x2n[k], X)
            bits >>= 1
            k += 1
        for lvl in range(5):
            st = 1 << lvl
            regs = [crc_mul(regs[i], X) ^ regs[i + st] if i % (2 * st) == 0 else regs[i] for i in range(32)]
            X = crc_mul(X, X)
        assert (~regs[0]) & 0xFFFFFFFF == zlib.crc32(data), ln


def test_magic_division_is_exact():
    rng = np.random.default_rng(1)
    probes = np.concatenate([rng.integers(0, 2 ** 31, 20000, dtype=np.int64), np.arange(0, 3000), np.arange(2 ** 31 - 3000, 2 ** 31)])
    for d in list(range(4, 1200)) + [2 ** k for k in range(2, 31)] + [2 ** k + 1 for k in range(2, 30)] + [2 ** k - 1 for k in range(3, 31)] + \
            [16385, 10 ** 6 + 3, 10 ** 9, 2 ** 31 - 1]:
        L = 0
        while (1 << L) < d:
            L += 1
        m = ((1 << (31 + L)) + d - 1) // d
        assert m < 2 ** 32, d
        n = np.concatenate([probes, (np.arange(1, 2000) * d - 1) % (2 ** 31), (np.arange(1, 2000) * d) % (2 ** 31)])
        q = np.array([((int(x) * m) >> 32) >> (L - 1) for x in n[:400]], dtype=np.int64)
        assert np.array_equal(q, n[:400] // d), d
        # the bulk in float-free numpy: split the 64-bit product
        hi = ((n >> 16) * m + (((n & 0xFFFF) * m) >> 16)) >> 16
        assert np.array_equal(hi >> (L - 1), n // d), d

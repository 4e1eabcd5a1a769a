#!/usr/bin/env python
"""Regenerate the committed fixtures under tests/golden/ from the reference tree (runs only where
/root/reference exists; the GPU box and the CPU test-suite only ever read the committed outputs).

Outputs
  randomBam.bam, randomBam.bam.bai   the reference's own toy BAM + index (inst/extdata; DATA fixtures, config C1's
                                     input — copied byte for byte, not source code)
  randomReads.npz                    the reference's test read table (tests/testthat/randomReads.RData) as numpy
                                     columns rname(code), pos(1-based), qwidth, strand(+1/-1), isize, read1, mapq, flag
  randomAnnot.json                   the 20 regions of inst/extdata/randomAnnot.Rdata (grgenes)
  expected_fixture.npz               known answers on the fixture produced by the numpy restatement of the
                                     reference's R test oracle (tests/spec_r.py, restating tests/testthat/utils.R:178-311)
R is not available offline, so the .RData files are read with the minimal XDR parser below.
"""
import gzip
import json
import os
import shutil
import struct
import sys

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


class XDR:
    """Minimal reader for R's serialize() XDR stream (RDX2 / RDX3 save() files)."""

    def __init__(self, data):
        self.d = data
        self.p = 0
        self.refs = []

    def i32(self):
        v = struct.unpack_from(">i", self.d, self.p)[0]
        self.p += 4
        return v

    def take(self, n):
        b = self.d[self.p:self.p + n]
        self.p += n
        return b

    def item(self):
        flags = self.i32()
        ty = flags & 0xFF
        has_attr = bool(flags & (1 << 9))
        has_tag = bool(flags & (1 << 10))
        if ty == 254:   # NILVALUE
            return None
        if ty == 255:   # REFSXP
            return self.refs[(flags >> 8) - 1]
        if ty == 1:     # SYMSXP
            name = self.item()
            self.refs.append(name)
            return name
        if ty == 2:     # LISTSXP (pairlist) -> dict by tag, in order
            out = {}
            while True:
                attr = self.item() if has_attr else None
                tag = self.item() if has_tag else None
                car = self.item()
                out[tag if tag is not None else len(out)] = car
                nflags = struct.unpack_from(">i", self.d, self.p)[0]
                if (nflags & 0xFF) != 2:
                    self.item()  # the terminating NILVALUE
                    break
                flags = self.i32()
                has_attr = bool(flags & (1 << 9))
                has_tag = bool(flags & (1 << 10))
            return out
        if ty == 9:     # CHARSXP
            n = self.i32()
            return None if n == -1 else self.take(n).decode("latin1")
        attrs = None
        if ty in (10, 13):   # LGLSXP / INTSXP
            n = self.i32()
            val = np.frombuffer(self.take(4 * n), dtype=">i4").astype(np.int32)
        elif ty == 14:       # REALSXP
            n = self.i32()
            val = np.frombuffer(self.take(8 * n), dtype=">f8").astype(np.float64)
        elif ty == 16:       # STRSXP
            n = self.i32()
            val = [self.item() for _ in range(n)]
        elif ty == 19:       # VECSXP
            n = self.i32()
            val = [self.item() for _ in range(n)]
        elif ty == 25:       # S4SXP: attributes only
            val = "S4"
        else:
            raise ValueError(f"unsupported SEXP type {ty} at {self.p}")
        if has_attr:
            attrs = self.item()
        return {"v": val, "a": attrs} if attrs is not None else val


def load_rdata(path):
    raw = gzip.open(path).read()
    assert raw[:5] in (b"RDX2\n", b"RDX3\n"), raw[:5]
    x = XDR(raw)
    x.p = 5
    assert x.take(2) == b"X\n"
    ver = x.i32()
    x.i32()
    x.i32()
    if ver == 3:
        n = x.i32()
        x.take(n)
    return x.item()


def val(o):
    return o["v"] if isinstance(o, dict) and "v" in o else o


def attrs(o):
    return o["a"] if isinstance(o, dict) and "a" in o else {}


def factor_to_codes(o):
    return val(o).astype(np.int32), list(val(attrs(o)["levels"]))


def main():
    if not os.path.isdir(REF):
        sys.exit("reference tree not present; the committed fixtures are authoritative")
    for f in ("randomBam.bam", "randomBam.bam.bai"):
        shutil.copyfile(os.path.join(REF, "inst/extdata", f), os.path.join(HERE, f))

    top = load_rdata(os.path.join(REF, "tests/testthat/randomReads.RData"))
    reads = top["reads"]
    names = list(val(attrs(reads)["names"]))
    cols = dict(zip(names, val(reads)))
    rcode, rlevels = factor_to_codes(cols["rname"])
    scode, slevels = factor_to_codes(cols["strand"])
    strand = np.where(np.array([slevels[c - 1] for c in scode]) == "+", 1, -1).astype(np.int8)
    out = dict(
        rname=(rcode - 1).astype(np.int32), rname_levels=np.array(rlevels),
        pos=val(cols["pos"]).astype(np.int32), qwidth=val(cols["qwidth"]).astype(np.int32),
        strand=strand, isize=val(cols["isize"]).astype(np.int32),
        read1=val(cols["read1"]).astype(bool), mapq=val(cols["mapq"]).astype(np.int32),
        flag=val(cols["flag"]).astype(np.int32),
    )
    np.savez_compressed(os.path.join(HERE, "randomReads.npz"), **out)
    print("reads:", len(out["pos"]), "levels", rlevels)

    top = load_rdata(os.path.join(REF, "inst/extdata/randomAnnot.Rdata"))
    gr = top["grgenes"]
    a = attrs(gr)
    rng = attrs(a["ranges"])
    start = val(rng["start"]).tolist()
    width = val(rng["width"]).tolist()

    def rle(o):
        oa = attrs(o)
        codes, levels = factor_to_codes(oa["values"])
        lens = val(oa["lengths"])
        return [levels[c - 1] for c, n in zip(codes, lens) for _ in range(n)], levels

    seqn, seqlevels = rle(a["seqnames"])
    strd, _ = rle(a["strand"])
    annot = dict(seqnames=seqn, seqlevels=seqlevels, start=start, width=width, strand=strd)
    json.dump(annot, open(os.path.join(HERE, "randomAnnot.json"), "w"), indent=1)
    print("annot regions:", len(start), "seqlevels", seqlevels)

    # known answers from the numpy restatement of the reference's R test oracle
    sys.path.insert(0, os.path.dirname(HERE))
    import spec_r
    spec_r.write_expected(os.path.join(HERE, "expected_fixture.npz"))


if __name__ == "__main__":
    main()

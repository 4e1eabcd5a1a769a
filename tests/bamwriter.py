"""Tiny pure-Python BAM + BAI writer for hand-crafted edge-case fixtures (test infrastructure).
Plays the role of the reference's tests/testthat/utils.R:readsToBam + writeSamAsBamAndIndex for cases its own
fixture never exercises: arbitrary CIGARs, flag-0x4 reads with a position, records straddling BGZF blocks,
several linear-index windows, unsorted or corrupt input."""
import struct
import zlib

CIGAR_OPS = "MIDNSHP=X"


def cigar_ops(cigar):
    """'5S20M3D2I10N30M' -> [(5,4),(20,0),...]; '*' -> []"""
    if cigar in ("*", "", None):
        return []
    out, num = [], ""
    for ch in cigar:
        if ch.isdigit():
            num += ch
        else:
            out.append((int(num), CIGAR_OPS.index(ch)))
            num = ""
    return out


def ref_len(ops, flag):
    if flag & 4:
        return 0
    return sum(n for n, op in ops if op in (0, 2, 3, 7, 8))


def reg2bin(beg, end):
    end -= 1
    for shift, off in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        if beg >> shift == end >> shift:
            return off + (beg >> shift)
    return 0


def record_bytes(tid, pos, flag, mapq, cigar, tlen, name="r", l_seq=0, aux=b"", next_tid=-1, next_pos=-1):
    ops = cigar_ops(cigar) if isinstance(cigar, str) or cigar is None else list(cigar)
    rl = ref_len(ops, flag) or 1
    nm = name.encode() + b"\0"
    n_cig = len(ops)
    body = struct.pack("<iiBBHHHiiii", tid, pos, len(nm), mapq, (reg2bin(pos, pos + rl) & 0xFFFF) if pos >= 0 else 4680,
                       n_cig & 0xFFFF, flag, l_seq, next_tid, next_pos, tlen)
    body += nm + b"".join(struct.pack("<I", n << 4 | op) for n, op in ops)
    body += bytes((l_seq + 1) // 2) + bytes([30] * l_seq) + aux
    return struct.pack("<i", len(body)) + body, rl


def bgzf_block(data, level=6):
    co = zlib.compressobj(level, zlib.DEFLATED, -15)
    comp = co.compress(data) + co.flush()
    hdr = b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", len(comp) + 25)
    return hdr + comp + struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data))


EOF_BLOCK = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def reg2bin_csi(beg, end, min_shift, depth):
    """CSIv1 reg2bin: the smallest bin that contains [beg, end)."""
    end -= 1
    s, t = min_shift, ((1 << depth * 3) - 1) // 7
    for l in range(depth, 0, -1):
        if beg >> s == end >> s:
            return t + (beg >> s)
        s += 3
        t -= 1 << ((l - 1) * 3)
    return 0


def bin_first_window(b, depth):
    """First leaf-level window a bin covers (htslib's hts_bin_bot)."""
    l, first = 0, 0
    while l < depth and b >= first + (1 << 3 * l):
        first += 1 << 3 * l
        l += 1
    return (b - first) << (3 * (depth - l))


def write_bam(path, refs, reads, block_payload=0xFF00, cut_mid_record=False, write_index=True, level=6,
              index="bai", min_shift=14, depth=5):
    """refs: [(name, length)]; reads: dicts(tid,pos,flag,mapq,cigar,tlen[,name,l_seq,aux]) in file order.
    block_payload: uncompressed bytes per BGZF block; cut_mid_record: blocks are cut at exactly block_payload bytes
    (records straddle), otherwise at record boundaries like htslib.
    index: "bai" (<path>.bai), "csi" (<path>.csi, BGZF-compressed CSIv1 with the given min_shift / depth) or "both"."""
    text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join(f"@SQ\tSN:{n}\tLN:{l}\n" for n, l in refs)
    hdr = b"BAM\1" + struct.pack("<i", len(text)) + text.encode() + struct.pack("<i", len(refs))
    for n, l in refs:
        hdr += struct.pack("<i", len(n) + 1) + n.encode() + b"\0" + struct.pack("<i", l)
    out = bytearray(bgzf_block(hdr, level))
    recs = []
    stream = bytearray()
    for i, r in enumerate(reads):
        b, rl = record_bytes(r["tid"], r["pos"], r["flag"], r["mapq"], r.get("cigar"), r.get("tlen", 0),
                             r.get("name", f"q{i}"), r.get("l_seq", 0), r.get("aux", b""))
        recs.append((len(stream), len(b), r["tid"], r["pos"], r["pos"] + rl, r["flag"]))
        stream += b
    # cut into blocks
    cuts = [0]
    if cut_mid_record:
        while cuts[-1] < len(stream):
            cuts.append(min(len(stream), cuts[-1] + block_payload))
    else:
        cur = 0
        for off, ln, *_ in recs:
            if off + ln - cur > block_payload and off > cur:
                cuts.append(off)
                cur = off
            while off + ln - cur > 0xFF00:           # oversize record: forced cut
                cur += 0xFF00
                cuts.append(cur)
        if len(stream) > cuts[-1]:
            cuts.append(len(stream))
    cstart = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        cstart.append(len(out))
        out += bgzf_block(bytes(stream[a:b]), level)
    cstart.append(len(out))
    out += EOF_BLOCK
    open(path, "wb").write(out)
    if not write_index:
        return

    def voff(u):
        import bisect
        k = bisect.bisect_right(cuts, u) - 1
        if k >= len(cuts) - 1:
            return cstart[-1] << 16
        return cstart[k] << 16 | (u - cuts[k])

    if index in ("csi", "both"):
        _write_csi(path + ".csi", refs, recs, voff, min_shift, depth, level)
        if index == "csi":
            return
    idx = [dict(bins={}, lin={}, beg=None, end=None, nm=0, nu=0) for _ in refs]
    cur = (None, None)
    run = None
    n_no_coor = 0
    for off, ln, tid, pos, end, flag in recs:
        if tid < 0:
            n_no_coor += 1
            continue
        vb, ve = voff(off), voff(off + ln)
        ri = idx[tid]
        b = reg2bin(pos, end)
        if (tid, b) != cur:
            if run:
                idx[run[0]]["bins"].setdefault(run[1], []).append((run[2], run[3]))
            cur = (tid, b)
            run = [tid, b, vb, ve]
        else:
            run[3] = ve
        for w in range(pos >> 14, ((end - 1) >> 14) + 1):
            if w not in ri["lin"]:
                ri["lin"][w] = vb
        ri["beg"] = vb if ri["beg"] is None else ri["beg"]
        ri["end"] = ve
        if flag & 4:
            ri["nu"] += 1
        else:
            ri["nm"] += 1
    if run:
        idx[run[0]]["bins"].setdefault(run[1], []).append((run[2], run[3]))
    bai = bytearray(b"BAI\1" + struct.pack("<i", len(refs)))
    for ri in idx:
        nb = len(ri["bins"]) + (1 if ri["beg"] is not None else 0)
        bai += struct.pack("<i", nb)
        for b, chunks in sorted(ri["bins"].items()):
            bai += struct.pack("<Ii", b, len(chunks))
            for c in chunks:
                bai += struct.pack("<QQ", *c)
        if ri["beg"] is not None:
            bai += struct.pack("<IiQQQQ", 37450, 2, ri["beg"], ri["end"], ri["nm"], ri["nu"])
        n_intv = (max(ri["lin"]) + 1) if ri["lin"] else 0
        lin = [ri["lin"].get(w, 0) for w in range(n_intv)]
        for w in range(n_intv - 2, -1, -1):
            if lin[w] == 0:
                lin[w] = lin[w + 1]
        bai += struct.pack("<i", n_intv) + b"".join(struct.pack("<Q", v) for v in lin)
    bai += struct.pack("<Q", n_no_coor)
    open(path + ".bai", "wb").write(bai)


def _write_csi(path, refs, recs, voff, min_shift, depth, level=6):
    """CSIv1 (SAM spec CSIv1.pdf): magic, min_shift, depth, l_aux = 0, n_ref, then per reference the bins with their
    loffset (= linear-index entry of the bin's first window, as htslib's update_loff derives it) and chunks, plus the
    pseudo-bin; the whole file BGZF-compressed."""
    meta = ((1 << (depth + 1) * 3) - 1) // 7 + 1
    idx = [dict(bins={}, lin={}, beg=None, end=None, nm=0, nu=0) for _ in refs]
    cur, run, n_no_coor = (None, None), None, 0
    for off, ln, tid, pos, end, flag in recs:
        if tid < 0:
            n_no_coor += 1
            continue
        vb, ve = voff(off), voff(off + ln)
        ri = idx[tid]
        b = reg2bin_csi(pos, end, min_shift, depth)
        if (tid, b) != cur:
            if run:
                idx[run[0]]["bins"].setdefault(run[1], []).append((run[2], run[3]))
            cur, run = (tid, b), [tid, b, vb, ve]
        else:
            run[3] = ve
        for w in range(pos >> min_shift, ((end - 1) >> min_shift) + 1):
            ri["lin"].setdefault(w, vb)
        ri["beg"] = vb if ri["beg"] is None else ri["beg"]
        ri["end"] = ve
        ri["nu" if flag & 4 else "nm"] += 1
    if run:
        idx[run[0]]["bins"].setdefault(run[1], []).append((run[2], run[3]))
    body = bytearray(b"CSI\1" + struct.pack("<iiii", min_shift, depth, 0, len(refs)))
    for ri in idx:
        n_intv = (max(ri["lin"]) + 1) if ri["lin"] else 0
        lin, prev = [], ri["beg"] or 0          # leading gaps take the reference's first offset, later ones the previous entry
        for w in range(n_intv):
            prev = ri["lin"].get(w, prev)
            lin.append(prev)
        nb = len(ri["bins"]) + (1 if ri["beg"] is not None else 0)
        body += struct.pack("<i", nb)
        for b, chunks in sorted(ri["bins"].items()):
            w = bin_first_window(b, depth)
            body += struct.pack("<IQi", b, lin[w] if w < n_intv else 0, len(chunks))
            for c in chunks:
                body += struct.pack("<QQ", *c)
        if ri["beg"] is not None:
            body += struct.pack("<IQiQQQQ", meta, 0, 2, ri["beg"], ri["end"], ri["nm"], ri["nu"])
    body += struct.pack("<Q", n_no_coor)
    out = bytearray()
    for a in range(0, len(body), 0xFF00):
        out += bgzf_block(bytes(body[a:a + 0xFF00]), level)
    out += EOF_BLOCK
    open(path, "wb").write(out)

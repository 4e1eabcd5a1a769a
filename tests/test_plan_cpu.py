"""The fetch planner on CPU (bsg_debug_plan, host-only): region -> index query -> record-aligned virtual-offset ranges
-> segments.  It replaces what bam_itr_queryi + bam_itr_next select for the reference (src/bamsignals.cpp:267-271), and
the counting path is only exact if the planned ranges hold EVERY record that can reach a region (a
superset is fine, a miss is a wrong count).  Checked against an index-free dump of the file by the oracle."""
import os
import sys

import numpy as np
import pytest

import bamsignals_b200 as B
import bamwriter as W
import edge_cases as E
import oracle_api as O

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import workloads as WL  # noqa: E402


def needed_mask(reads, names, gr, ext):
    """records the reference's iterator would hand to a region: tid == rid, pos < end + ext, endpos > beg - ext"""
    need = np.zeros(len(reads["tid"]), dtype=bool)
    for i in range(len(gr)):
        if gr.width[i] == 0:
            continue
        t = names.index(gr.seqnames[i])
        beg, end = int(gr.start[i]) - 1 - ext, int(gr.start[i]) - 1 + int(gr.width[i]) + ext
        need |= (reads["tid"] == t) & (reads["pos"] < end) & (reads["endpos"] > max(beg, 0))
    return need


def as_multiset(tid, pos):
    keys = tid.astype(np.int64) << 32 | (pos.astype(np.int64) & 0xffffffff)
    u, c = np.unique(keys, return_counts=True)
    return dict(zip(u.tolist(), c.tolist()))


def assert_covers(plan, reads, need):
    have = as_multiset(plan["tid"], plan["pos"])
    want = as_multiset(reads["tid"][need], reads["pos"][need])
    missing = {k: v for k, v in want.items() if have.get(k, 0) < v}
    assert not missing, f"{len(missing)} needed (tid,pos) keys are not inside the planned ranges"


@pytest.mark.parametrize("index,min_shift,depth", [("bai", 14, 5), ("csi", 14, 5), ("csi", 12, 6), ("csi", 16, 4)])
@pytest.mark.parametrize("ext", [0, 150, 20000])
def test_plan_covers_every_needed_record(tmp_path, index, min_shift, depth, ext):
    p = str(tmp_path / "v.bam")
    E.write_variety(p, block_payload=1500, cut_mid_record=True, index=index, min_shift=min_shift, depth=depth)
    reads = O.dump_reads(p)
    names = [r[0] for r in E.REFS]
    gr = E.variety_regions(seed=3, n=40)
    plan = B.debug_plan(p, gr, ext=ext)
    assert_covers(plan, reads, needed_mask(reads, names, gr, ext))
    assert plan["records"] <= len(reads["tid"]) and plan["bytes_inflated"] >= plan["bytes_compressed"] // 4


def test_plan_prunes(tmp_path):
    """a few narrow regions must not pull in the whole file, and regions on an empty contig fetch nothing"""
    bam, info = WL.make_bam("c2", 0.002, str(tmp_path))
    lens = WL.contig_lens("c2", 0.002)
    gr = B.GRanges(["chr1", "chr1", "chr7"], [1000, lens[0] // 2, 5000], [2000, 2000, 2000])
    plan = B.debug_plan(bam, gr, ext=75)
    reads = O.dump_reads(bam)
    assert_covers(plan, reads, needed_mask(reads, WL.NAMES, gr, 75))
    assert 0 < plan["records"] < info["records"] // 20
    p = str(tmp_path / "v.bam")
    E.write_variety(p)
    empty = B.debug_plan(p, B.GRanges(["chrEmpty"] * 3, [1, 100, 500], [50, 50, 50]), ext=0)
    assert empty["records"] == 0 and empty["segments"] == 0


def test_plan_whole_genome_is_the_whole_file(tmp_path):
    """C4-style regions (whole contigs): every placed record is planned exactly once"""
    bam, info = WL.make_bam("c4", 0.001, str(tmp_path), unplaced=5)
    gr, kw, fn = WL.regions("c4", 0.001)
    plan = B.debug_plan(bam, gr, ext=200)
    reads = O.dump_reads(bam)
    placed = reads["tid"] >= 0
    assert plan["records"] == int(placed.sum())
    assert as_multiset(plan["tid"], plan["pos"]) == as_multiset(reads["tid"][placed], reads["pos"][placed])


def test_plan_region_order_does_not_matter(tmp_path):
    """the (rid, loc) order the planner and the tile builder share is a radix sort: shuffled, sorted and reversed
    region sets plan the same ranges"""
    p = str(tmp_path / "v.bam")
    E.write_variety(p, block_payload=1500)
    gr = E.variety_regions(seed=5, n=200)
    base = B.debug_plan(p, gr, ext=10)
    rng = np.random.default_rng(0)
    for perm in (rng.permutation(len(gr)), np.lexsort((gr.start, gr.seq_idx)), np.lexsort((gr.start, gr.seq_idx))[::-1]):
        q = B.debug_plan(p, gr[np.asarray(perm)], ext=10)
        assert q["records"] == base["records"] and q["segments"] == base["segments"]
        assert np.array_equal(q["pos"], base["pos"]) and np.array_equal(q["tid"], base["tid"])


def test_plan_long_contig_csi(tmp_path):
    import test_csi_index as T
    p = str(tmp_path / "big.bam")
    W.write_bam(p, T.BIG_REFS, T.big_reads(), index="csi", min_shift=14, depth=6, block_payload=4000)
    gr = T.big_regions()
    reads = O.dump_reads(p)
    plan = B.debug_plan(p, gr, ext=500)
    assert_covers(plan, reads, needed_mask(reads, [r[0] for r in T.BIG_REFS], gr, 500))
    assert plan["records"] < len(reads["tid"])          # the six clusters are far apart: some are skipped


def test_plan_mixed_long_and_short_ranges(tmp_path):
    """whole contigs (ranges longer than a fetch segment: cut at index entry points) next to narrow regions (never cut,
    never looked up in the entry-point table), in ascending file order and interleaved: nothing needed is lost, the long
    ranges are split, and the narrow regions still prune"""
    bam, info = WL.make_bam("c4", 0.002, str(tmp_path))
    lens = WL.contig_lens("c4", 0.002)
    whole = [0, 2, 6]
    names = [WL.NAMES[i] for i in whole] + ["chr2", "chr2", "chr5", "chr9", "chr9"]
    starts = [1] * len(whole) + [lens[1] // 3, lens[1] // 3 + 30000, lens[4] // 2, 1000, lens[8] - 3000]
    widths = [lens[i] for i in whole] + [2000, 500, 2000, 800, 2000]
    gr = B.GRanges(names, starts, widths)
    reads = O.dump_reads(bam)
    plan = B.debug_plan(bam, gr, ext=100)
    assert_covers(plan, reads, needed_mask(reads, WL.NAMES, gr, 100))
    on_whole = int(np.isin(reads["tid"], whole).sum())
    assert on_whole <= plan["records"] < on_whole + info["records"] // 50
    assert plan["bytes_compressed"] > 3 * (1 << 20) and plan["segments"] >= 3 + plan["bytes_compressed"] // (2 << 20)
    # the same query with the regions shuffled plans the same fetch
    perm = np.random.default_rng(1).permutation(len(gr))
    q = B.debug_plan(bam, gr[perm], ext=100)
    assert q["segments"] == plan["segments"] and np.array_equal(q["pos"], plan["pos"]) and np.array_equal(q["tid"], plan["tid"])


def test_plan_region_order_extreme_locs(tmp_path):
    """the radix keys of the shared (rid, loc) sort are built from loc - min(loc): starts below 1 and near 2^31 must sort
    like any other"""
    p = str(tmp_path / "v.bam")
    E.write_variety(p, block_payload=1500)
    base_gr = E.variety_regions(seed=9, n=60)
    name0 = base_gr.seqnames[0]
    extra = B.GRanges([name0] * 4, [-5000, -3, 2**31 - 5000, 1], [6000, 50, 4000, 1])
    names = list(base_gr.seqnames) + list(extra.seqnames)
    starts = np.concatenate([np.asarray(base_gr.start, dtype=np.int64), np.asarray(extra.start, dtype=np.int64)])
    widths = np.concatenate([np.asarray(base_gr.width, dtype=np.int64), np.asarray(extra.width, dtype=np.int64)])
    gr = B.GRanges(names, starts, widths)
    order = np.lexsort((np.arange(len(gr)), gr.start, gr.seq_idx))
    want = B.debug_plan(p, gr[order], ext=10)             # already sorted: the planner's fast path, no radix sort
    rng = np.random.default_rng(4)
    for perm in (rng.permutation(len(gr)), order[::-1]):
        got = B.debug_plan(p, gr[np.asarray(perm)], ext=10)
        assert got["segments"] == want["segments"]
        assert np.array_equal(got["pos"], want["pos"]) and np.array_equal(got["tid"], want["tid"])

"""Host-side logic that needs no GPU: the C-ABI library loads and exports every declared symbol,
the three name-hash implementations agree, and the text decoders reproduce what the oracle
(hence the reference) extracts from the same files."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, golden_e2e_names, load_golden
from duet_b200 import _lib, namehash, read_file, sv_phasing_fn, synth, write_file
from duet_b200.columnar import from_synth
from oracle import ref_port


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    with open(os.path.join(ROOT, "include", "duet_b200.h")) as f:
        header = f.read()
    declared = set(re.findall(r"\b(duet_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS)
    for sym in declared:
        assert getattr(lib, sym) is not None
    assert lib.duet_abi_version() == 1
    t = _lib.Thresholds()
    lib.duet_default_thresholds(C.byref(t))
    assert (t.svlen_thres, t.suppread_thres, t.pc_max, t.c2_sv_ratio_min, t.c1_totsc_ratio_max) == (50, 2, 8100, 0.72, 9.72)


def test_struct_layouts_match_header():
    # sizes the C side was compiled with (LP64): catches a drifted ctypes mirror
    assert C.sizeof(_lib.Thresholds) == 8 * 4 + 10 * 8
    assert C.sizeof(_lib.PhaseInput) == 8 + 3 * 8 + 13 * 8
    from duet_b200.columnar import TAG_DTYPE
    assert TAG_DTYPE.itemsize == 16 and TAG_DTYPE.fields['chk'][1] == 8 and TAG_DTYPE.fields['hp'][1] == 12
    assert C.sizeof(_lib.PhaseOutput) == 13 * 8 + 8
    assert C.sizeof(_lib.Timings) == 3 * 4 + 8 * 4


def test_no_device_fails_loudly():
    """On a box without a GPU the product path must refuse, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from duet_b200.engine import DuetError, PhaseEngine
    with pytest.raises(DuetError) as ei:
        PhaseEngine(0)
    assert ei.value.code == _lib.ERR_NO_DEVICE


def test_name_hash_three_ways():
    rng = np.random.default_rng(0)
    names = [bytes(rng.integers(33, 127, size=int(n)).astype(np.uint8)) for n in list(range(0, 70)) + [36] * 50]
    lo_py, hi_py = namehash.hash_names(names)
    lens, lo_c, hi_c = sv_phasing_fn.hash_name_lists([[n.decode() for n in names]])
    assert np.array_equal(lo_py, lo_c) and np.array_equal(hi_py, hi_c)
    fixed = synth.names_from_ids(np.arange(1000))
    lo_np, hi_np = namehash.hash128_fixed(fixed)
    lo_py, hi_py = namehash.hash_names([r.tobytes() for r in fixed])
    assert np.array_equal(lo_np, lo_py) and np.array_equal(hi_np, hi_py)
    assert len(set(lo_np.tolist())) == 1000 and (lo_np != np.uint64(namehash.EMPTY_KEY)).all()


@pytest.mark.parametrize("name", golden_e2e_names())
def test_decoders_match_oracle_on_golden_inputs(name, golden_workdir):
    case, home = golden_workdir(name)
    inc = case.get("include_all_ctgs", False)
    vcf, sam_home = home + "/sv_calling/variants.vcf", home + "/snp_phasing/"
    tables = ref_port.haplotag_tables(sam_home, 1, inc)
    cols = sv_phasing_fn.read_hap_bam(sam_home, 1, inc)
    for table, rc in zip(tables, cols):
        # last row per key == dict content
        last = {}
        for i, k in enumerate(rc.key.tolist()):
            last[k] = i
        assert len(last) == len(table)
        want = {namehash.hash128(nm)[0]: tag for nm, tag in table.items()}
        got = {k: (int(rc.hp[i]), int(rc.ps[i]), int(rc.pc[i])) for k, i in last.items()}
        assert got == want
    recs = ref_port.sv_records(vcf, inc)
    svs = read_file.parse_vcf(vcf, inc)
    for rl, cs in zip(recs, svs):
        assert len(rl) == len(cs)
        for i, r in enumerate(rl):
            assert (r.chrom, r.pos, r.ref, r.alt, r.svlen, r.svtype, r.svread, r.names, r.gt, r.refread, r.altread) == \
                   (cs.chrom[i], cs.pos[i], cs.ref[i], cs.alt[i], cs.svlen[i], cs.svtype[i], cs.svread[i],
                    cs.names[i], cs.gt[i], cs.refread[i], cs.altread[i])
    assert write_file.header_text(vcf, inc) == ref_port.header_text(vcf, inc)
    assert write_file.header_text(vcf, inc) + write_file.format_rows(case["rows"]) == case["phased_sv_vcf"]


def test_name_lists_hashed_without_splitting():
    """duet_hash_name_lists splits the unsplit RNAMES strings itself: same lengths and hashes as hashing
    [s.split(',') for s in lists], including empty names and an empty list string."""
    rng = np.random.default_rng(3)
    alphabet = np.array(list("abcdefghijklmnopqrstuvwxyz0123456789-_/:"))
    lists = []
    for _ in range(500):
        k = int(rng.integers(0, 12))
        lists.append(",".join("".join(alphabet[rng.integers(0, len(alphabet), int(rng.integers(0, 40)))]) for _ in range(k)))
    lists += ["", ",", "a,,b", "solo"]
    lens_a, lo_a, hi_a = sv_phasing_fn.hash_name_csv(lists)
    lens_b, lo_b, hi_b = sv_phasing_fn.hash_name_lists([s.split(",") for s in lists])
    assert np.array_equal(lens_a, lens_b) and np.array_equal(lo_a, lo_b) and np.array_equal(hi_a, hi_b)
    assert lens_a[-4:].tolist() == [1, 2, 3, 1]
    lens_e, lo_e, _ = sv_phasing_fn.hash_name_csv([])
    assert lens_e.size == 0 and lo_e.size == 0
    with pytest.raises(ValueError):
        sv_phasing_fn.hash_name_csv(["a\nb"])


def test_info_item_lookup_equals_the_item_scan():
    """read_file._first_with finds 'the first INFO item containing a needle' with str.find; the reference
    scans info.split(';') item by item (read_file.py:34-55).  Same item on adversarial strings: needles
    inside other keys, repeated keys, empty items, no match."""
    rng = np.random.default_rng(7)
    pieces = ["SVLEN=12", "XSVLEN=>7", "SVLEN=.", "SVTYPE=DEL", "RE=4", "SR=9", "SUPPORT=3", "PRE=1", "RNAMES=a,b", "READS=c",
              "AF=0.5", "", "CORE=2", "STRANDS=+-", "FURTHERNAMES=x", "END=5"]
    groups = [("SVLEN=",), ("SVTYPE=",), ("SUPPORT=", "SR=", "RE="), ("RNAMES=", "READS=")]

    def scan(info, needles):
        for it in info.split(";"):
            if any(nd in it for nd in needles):
                return it
        return None

    for _ in range(3000):
        info = ";".join(pieces[i] for i in rng.integers(0, len(pieces), int(rng.integers(0, 9))))
        for needles in groups:
            assert read_file._first_with(info, needles) == scan(info, needles), (info, needles)


def test_text_path_equals_direct_columnar(tmp_path):
    s = synth.make_sample(3, contigs=["1", "5", "X"], n_reads=3000, n_svs=250, bp_per_read=700, block_mean=1e5)
    home = str(tmp_path)
    synth.write_workdir(s, home, "cutesv")
    a = sv_phasing_fn.generate_callinfo(home + "/sv_calling/variants.vcf",
                                        sv_phasing_fn.read_hap_bam(home + "/snp_phasing/", 1, False), False)
    b = from_synth(s)
    shard = {c: i for i, c in enumerate(a.shard_contig)}
    a = a.select_shards([shard[c] for c in b.shard_contig])      # 24 reference contigs -> the 3 that have data
    for k in ("read_off", "sv_off", "read_key", "read_tag", "read_hp", "read_ps", "read_pc", "sv_pos", "sv_svlen",
              "sv_svread", "sv_refread", "sv_flags", "csr_off", "csr_key", "csr_chk"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    assert (a.sv_chrom, a.sv_type, a.sv_alt) == (b.sv_chrom, b.sv_type, b.sv_alt)


@pytest.mark.parametrize("text,exc", [
    (b"r1\t0\t1\t5\tHP:i:1\tPC:i:7\tPS:i:9\n\n", IndexError),                 # blank line
    (b"lonely\n", IndexError),                                                # one field
    (b"PC:i:5\tPS:i:9\n", IndexError),                                        # 'PC:i:' in s[-2] but no s[-3]
    (b"r1\tHP:i:x\tPC:i:7\tPS:i:9\n", ValueError),
    (b"r1\tHP:i:1\tPC:i:7\tPS:i:\n", ValueError),
    (b"r1\tHP:i:1\tPC:i:7\tPS:i:9\xc3\xa9\n", UnicodeDecodeError),
    (b"r1\tHP:i:999\tPC:i:7\tPS:i:9\n", OverflowError),
])
def test_sam_text_error_behaviour(text, exc):
    with pytest.raises(exc):
        sv_phasing_fn.decode_sam_text(text)


def test_sam_text_quirks():
    # tags are positional (last three fields), rows need 'PC:i:' in the second-to-last field,
    # text after the final newline is dropped, any whitespace separates fields
    text = (b"a 0 1 5 XX:i:1 HP:i:2 PC:i:30 PS:i:400\n"
            b"b\t0\tHP:i:1\tPC:i:7\tPS:i:9\tNM:i:0\n"          # tags not last -> s[-2] is PS -> skipped
            b"c\t0\tzz:i:1_0\tPC:i:+7\tqq:i:-9\n"                # 5 characters stripped by position, int() syntax
            b"d\t0\tHP:i:1\tPC:i:7\tPS:i:9")                      # no trailing newline -> dropped
    rc = sv_phasing_fn.decode_sam_text(text)
    assert rc.n_lines == 3 and len(rc) == 2
    assert (rc.hp.tolist(), rc.pc.tolist(), rc.ps.tolist()) == ([2, 10], [30, 7], [400, -9])
    assert rc.key[0] == namehash.hash128("a")[0] and rc.key[1] == namehash.hash128("c")[0]


# ---- native BAM reader vs the SAM-text path (records written by tests/util_bam.py) -------------------
def _bam_and_text(rows):
    from util_bam import record
    recs, text = [], []
    for r in rows:
        b, t = record(*r)
        recs.append(b)
        text.append(t)
    return recs, "".join(text).encode()


def test_bam_reader_equals_text_path(tmp_path):
    from util_bam import write_bam
    s = synth.make_sample(4, contigs=["21"], n_reads=5000, n_svs=10, bp_per_read=700)
    c = s.contigs[0]
    names = synth.name_strings(c.row_id)
    rows = []
    for i, nm in enumerate(names):
        aux = [("NM", "C", 3), ("MD", "Z", "4"), ("AS", "i", 8)]
        if c.row_tagged[i]:
            aux += [("HP", "C", int(c.row_hp[i])), ("PC", "I" if i % 3 else "S", int(min(c.row_pc[i], 60000))), ("PS", "i", int(c.row_ps[i]))]
        rows.append((nm, int(c.row_pos[i]), "ACGT", "IIII", aux))
    recs, text = _bam_and_text(rows)
    path = str(tmp_path / "21.bam")
    write_bam(path, recs, block=20000)
    a = sv_phasing_fn.load_hap_bam(path, 1)                  # gzip magic -> native BAM decode
    b = sv_phasing_fn.decode_sam_text(text)
    assert a.n_lines == b.n_lines == len(rows) and len(a) == len(b) == int(c.row_tagged.sum())
    assert np.array_equal(a.key, b.key) and np.array_equal(a.tag, b.tag)


def test_bam_blocks_inflated_in_parallel(tmp_path):
    """duet_set_decode_threads: the BGZF blocks of one file inflate independently; same columns as the
    single-threaded read, a corrupt block is still reported, and read_hap_bam hands spare threads to
    the blocks when there are fewer files than threads (the one-contig demo)."""
    from util_bam import write_bam
    rows = [(f"read{i:06d}", 10 + i, "ACGT" * 8, "I" * 32, [("HP", "C", 1 + i % 2), ("PC", "i", i % 9000), ("PS", "i", 1000 + i // 50)])
            for i in range(30000)]
    recs, text = _bam_and_text(rows)
    home = tmp_path / "snp_phasing"
    home.mkdir()
    path = str(home / "21.bam")
    write_bam(path, recs, block=8000)                        # a few hundred blocks
    lib = _lib.load()
    one = sv_phasing_fn.load_hap_bam(path, 1)
    prev = lib.duet_set_decode_threads(4)
    try:
        assert prev == 1
        four = sv_phasing_fn.load_hap_bam(path, 1)
        raw = bytearray(open(path, "rb").read())
        raw[len(raw) // 2] ^= 0xFF                           # damage one block in the middle
        with pytest.raises(ValueError):
            sv_phasing_fn.decode_bam(bytes(raw))
    finally:
        assert lib.duet_set_decode_threads(prev) == 4
    assert len(one) == len(four) == len(rows)
    assert np.array_equal(one.key, four.key) and np.array_equal(one.tag, four.tag)
    via_stage = sv_phasing_fn.read_hap_bam(str(home) + "/", 8, False)       # 1 file, 8 threads
    got = [c for c in via_stage if len(c)]
    assert len(got) == 1 and np.array_equal(got[0].key, one.key) and np.array_equal(got[0].tag, one.tag)
    assert lib.duet_set_decode_threads(1) == 1               # the stage restored the setting


def test_bam_reader_text_quirks(tmp_path):
    """The reference looks at the last three WHITESPACE tokens of the text line, whatever they are."""
    from util_bam import write_bam
    rows = [
        ("tagged", 5, "ACGT", "IIII", [("HP", "i", 2), ("PC", "i", 30), ("PS", "i", 400)]),
        ("tags_not_last", 6, "ACGT", "IIII", [("HP", "i", 1), ("PC", "i", 7), ("PS", "i", 9), ("NM", "i", 0)]),
        ("blank_in_string", 7, "ACGT", "IIII", [("HP", "i", 1), ("CO", "Z", "7 PC:i:5 PS:i:6")]),     # tokens: CO:Z:7 PC:i:5 PS:i:6
        ("few_aux", 8, "ACGT", "IIII", [("PS", "i", 9)]),                                            # s[-2] is QUAL
        ("no_aux", 9, "AC", "*", []),
        ("float_hp", 10, "ACGT", "IIII", [("HP", "f", 2.0), ("PC", "i", 1), ("PS", "i", 3)]),        # int('2') works
        ("array_last", 11, "ACGT", "IIII", [("HP", "i", 1), ("PC", "i", 2), ("ZB", "B", ("c", [1, 2]))]),
    ]
    recs, text = _bam_and_text(rows)
    path = str(tmp_path / "x.bam")
    write_bam(path, recs)
    with pytest.raises(ValueError):                          # ZB:B:c,1,2 -> int('c,1,2') fails, in both paths
        sv_phasing_fn.decode_sam_text(text)
    with pytest.raises(ValueError):
        sv_phasing_fn.load_hap_bam(path, 1)
    recs, text = _bam_and_text(rows[:-1])
    write_bam(path, recs)
    a, b = sv_phasing_fn.load_hap_bam(path, 1), sv_phasing_fn.decode_sam_text(text)
    assert len(b) == 3                                       # tagged, blank_in_string (HP 7 from 'CO:Z:7'), float_hp
    assert b.hp.tolist() == [2, 7, 2] and b.pc.tolist() == [30, 5, 1] and b.ps.tolist() == [400, 6, 3]
    assert np.array_equal(a.key, b.key) and np.array_equal(a.tag, b.tag)
    assert a.n_lines == b.n_lines == 6


# ---- a BAM laid out byte by byte from the SAM/BAM specification (not by tests/util_bam.py) ---------------
def test_bam_reader_on_spec_laid_out_fixture():
    """tests/golden/spec_example.bam was assembled field by field from SAMv1 sections 4.1 / 4.2 by
    tests/golden/make_spec_bam.py (the specification's own example alignments r001-r003 plus HP/PC/PS
    tags; one BGZF block is a hand-written STORED deflate block, records straddle block boundaries, every
    aux type appears, the 28-byte EOF marker is the spec's literal).  Python's gzip module -- a third
    party to both writer and reader -- must inflate it, and the native reader must keep exactly the rows
    the reference keeps from the `samtools view` text of the same records (spec_example.sam)."""
    import gzip
    import os
    from conftest import GOLDEN
    from duet_b200.namehash import hash128
    with open(os.path.join(GOLDEN, "spec_example.bam"), "rb") as f:
        data = f.read()
    raw = gzip.decompress(data)                                 # checks every member's CRC32 and ISIZE
    assert raw[:4] == b"BAM\x01" and data.endswith(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))
    got = sv_phasing_fn.decode_bam(data)
    kept = [("r001", 1, 7, 60), ("r003", 2, 100000, 300), ("r001", 2, 7, 10), ("r005", 1, 42, -5)]   # (QNAME, HP, PS, PC)
    assert got.n_lines == 6 and len(got) == len(kept)
    for i, (nm, hp, ps, pc) in enumerate(kept):
        lo, hi = hash128(nm)
        assert int(got.key[i]) == lo and int(got.tag["chk"][i]) == hi & 0xFFFFFFFF
        assert (int(got.hp[i]), int(got.ps[i]), int(got.pc[i])) == (hp, ps, pc)
    with open(os.path.join(GOLDEN, "spec_example.sam"), "rb") as f:
        text = f.read()
    via_text = sv_phasing_fn.decode_sam_text(text)
    assert np.array_equal(got.key, via_text.key) and np.array_equal(got.tag, via_text.tag)
    # and the oracle port's reading of the same text (dict semantics: the later r001 row wins)
    import tempfile
    from oracle import ref_port
    with tempfile.TemporaryDirectory() as home:
        os.makedirs(os.path.join(home, "snp_phasing"))
        with open(os.path.join(home, "snp_phasing", "1.bam"), "wb") as f:
            f.write(text)
        table = ref_port.haplotag_tables(home + "/snp_phasing/", 1, False)[0]
    assert table == {"r001": (2, 7, 10), "r003": (2, 100000, 300), "r005": (1, 42, -5)}


# ---- the native SV-VCF reader (csrc/vcf_decode.cpp) against the general Python reader ---------------------
def _vcf_columns_equal(native, general, vcf_path):
    from duet_b200.columnar import TextColumn
    assert native.sv_off.tolist() == np.concatenate([[0], np.cumsum([len(cs) for cs in general])]).tolist()
    flat = lambda f: [v for cs in general for v in getattr(cs, f)]
    assert native.pos.tolist() == flat("pos") and native.svlen.tolist() == flat("svlen")
    assert native.svread.tolist() == flat("svread") and native.refread.tolist() == flat("refread")
    assert [bool(x) for x in native.flags] == [g == "./." for g in flat("gt")]
    for k, f in enumerate(("chrom", "ref", "alt", "svtype")):
        assert list(TextColumn(native.text, native.str_span[:, k, :])) == flat(f)
    lens, lo, hi = sv_phasing_fn.hash_name_csv(flat("names_csv"))
    assert np.array_equal(np.diff(native.csr_off), lens) and np.array_equal(native.csr_key, lo)
    assert np.array_equal(native.csr_chk, (hi & np.uint64(0xFFFFFFFF)).astype(np.uint32))
    ranks = []
    for cs in general:
        order = {c: i for i, c in enumerate(sorted(set(cs.chrom)))}
        ranks += [order[c] for c in cs.chrom]
    assert native.group.tolist() == ranks and native.has_groups == any(ranks)


@pytest.mark.parametrize("name", [n for n in golden_e2e_names() if n != "all_ctgs"])
@pytest.mark.parametrize("threads", [1, 3, 8])
def test_native_vcf_reader_equals_general_reader(name, threads, golden_workdir):
    """cuteSV / Sniffles2 / SVIM dialects, shuffled records, 'chr' and bare names in one contig, dense lists:
    the one-pass native reader returns the columns the general reader's lists give, for any thread count."""
    from duet_b200 import read_file
    case, home = golden_workdir(name)
    vcf = home + "/sv_calling/variants.vcf"
    native = read_file.decode_sv_vcf(vcf, False, threads)
    assert native is not None
    _vcf_columns_equal(native, read_file.parse_vcf(vcf, False), vcf)


def test_native_vcf_reader_declines_what_it_does_not_reproduce(golden_workdir, tmp_path):
    """A CHROM string claimed by two contigs (`all_ctgs`: '7' and 'chr7' both listed), blank lines, integers
    only Python's int() accepts, needles inside other keys, a dialect change inside a contig, a lone carriage
    return: the native reader returns None and the general reader decides."""
    from duet_b200 import read_file
    case, home = golden_workdir("all_ctgs")
    assert read_file.decode_sv_vcf(home + "/sv_calling/variants.vcf", True, 2) is None
    case = load_golden("e2e_cutesv_3ctg.json.gz")
    text = case["files"]["sv_calling/variants.vcf"]
    lines = text.split("\n")
    first = next(i for i, ln in enumerate(lines) if ln and not ln.startswith("#"))

    def declined(mutated: str) -> bool:
        d = tmp_path / f"case{declined.k}"
        declined.k += 1
        (d / "sv_calling").mkdir(parents=True)
        (d / "sv_calling" / "variants.vcf").write_text(mutated)
        return read_file.decode_sv_vcf(str(d) + "/sv_calling/variants.vcf", False, 2) is None
    declined.k = 0
    assert not declined(text)
    assert declined("\n".join(lines[:first + 1] + [""] + lines[first + 1:]))                      # blank line
    rec = lines[first]
    assert declined(text.replace(rec, rec.replace("RE=", "XRE=", 1), 1))                            # needle not at an item start
    assert declined(text.replace(rec, rec.replace("RNAMES=", "READS=", 1), 1) if "RNAMES=" in lines[first + 1] else text + "\n\n")
    pos = rec.split("\t")[1]
    assert declined(text.replace(rec, rec.replace("\t" + pos + "\t", "\t" + pos[0] + "_" + pos[1:] + "\t", 1), 1) if len(pos) > 1 else text + "\n\n")
    assert declined(text.replace(rec, rec.replace("\t", "\r\t", 1), 1))                             # lone carriage return


# ---- the two-step decode job (rows held by the library, copied into the caller's column slices) ---------------
def _job(data: bytes, kind: int):
    import ctypes as C
    from duet_b200 import _lib
    lib = _lib.load()
    job, n_rows, n_rec, err = C.c_void_p(), C.c_int64(), C.c_int64(), C.c_int64()
    rc = lib.duet_decode_reads(data, len(data), kind, C.byref(job), C.byref(n_rows), C.byref(n_rec), C.byref(err))
    if rc != 0:
        return rc, None, None, err.value
    from duet_b200.columnar import TAG_DTYPE
    key, tag = np.empty(n_rows.value, np.uint64), np.empty(n_rows.value, TAG_DTYPE)
    assert lib.duet_rows_take(job, key.ctypes.data, tag.ctypes.data) == 0
    return 0, key, tag, n_rec.value


def test_decode_job_equals_direct_decoders_and_streams_large_bams(tmp_path):
    """duet_decode_reads on SAM text and on a BAM whose inflated size exceeds one 16 MB batch several times
    (records straddle BGZF blocks and batches): same rows as the whole-buffer decoders."""
    from duet_b200 import _lib
    from util_bam import write_bam
    rng = np.random.default_rng(5)
    rows = []
    for k in range(120_000):
        aux = [("NM", "C", int(rng.integers(0, 9)))]
        if k % 3:
            aux += [("HP", "C", int(rng.integers(1, 3))), ("PC", "S", int(rng.integers(0, 9000))), ("PS", "I", int(rng.integers(1, 10**8)))]
        rows.append((f"read{k:07d}-{int(rng.integers(1 << 30)):x}", int(rng.integers(1, 10**8)), "ACGT" * 60, "*", aux))
    recs, text = _bam_and_text(rows)
    path = str(tmp_path / "big.bam")
    write_bam(path, recs, block=50000)
    with open(path, "rb") as f:
        data = f.read()
    assert sum(len(r) for r in recs) > 3 * (16 << 20)
    rc, key, tag, n_rec = _job(data, _lib.READS_BAM)
    assert rc == 0 and n_rec == len(rows)
    ref = sv_phasing_fn.decode_sam_text(text)
    assert np.array_equal(key, ref.key) and np.array_equal(tag, ref.tag)
    rc, key2, tag2, n_lines = _job(text, _lib.READS_SAM_TEXT)
    assert rc == 0 and np.array_equal(key2, ref.key) and np.array_equal(tag2, ref.tag) and n_lines == len(rows)
    whole = sv_phasing_fn.decode_bam(data)
    assert np.array_equal(whole.key, ref.key) and np.array_equal(whole.tag, ref.tag)


def test_bam_reader_rejects_forged_and_truncated_blocks(tmp_path):
    """ADVICE (round 1): a BGZF header whose XLEN points past the end of the input, an extra subfield that
    overruns XLEN, a forged ISIZE above the 64 KiB a block can hold, a file cut inside a record: an error
    code, never an out-of-bounds read or a giant allocation."""
    from duet_b200 import _lib
    from util_bam import write_bam
    recs, _ = _bam_and_text([(f"r{k}", 10 + k, "ACGT", "*", [("HP", "C", 1), ("PC", "C", 5), ("PS", "C", 9)]) for k in range(50)])
    path = str(tmp_path / "x.bam")
    write_bam(path, recs, block=700)
    good = open(path, "rb").read()
    assert _job(good, _lib.READS_BAM)[0] == 0
    forged_xlen = bytes.fromhex("1f8b08040000000000ffffff4243020011000300")[:18]
    assert _job(forged_xlen, _lib.READS_BAM)[0] == _lib.DECODE_ERR_FORMAT
    bad_sub = bytearray(good)
    bad_sub[14:16] = (60000).to_bytes(2, "little")                   # subfield length overruns XLEN
    assert _job(bytes(bad_sub), _lib.READS_BAM)[0] == _lib.DECODE_ERR_FORMAT
    bsize = int.from_bytes(good[16:18], "little") + 1
    forged_isize = bytearray(good)
    forged_isize[bsize - 4:bsize] = (1 << 30).to_bytes(4, "little")
    assert _job(bytes(forged_isize), _lib.READS_BAM)[0] == _lib.DECODE_ERR_FORMAT
    eof = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
    cut_blocks, pos = [], 0
    while pos < len(good) - len(eof):
        n = int.from_bytes(good[pos + 16:pos + 18], "little") + 1
        cut_blocks.append(good[pos:pos + n])
        pos += n
    truncated = b"".join(cut_blocks[:-1]) + eof                      # the last data block is gone: a record is cut short
    assert _job(truncated, _lib.READS_BAM)[0] == _lib.DECODE_ERR_FORMAT


def test_sam_text_with_header_lines_is_read_like_samtools_view():
    """A SAM file read directly may start with '@' header lines; `samtools view` (what the reference reads)
    does not print them, so they are passed over instead of being indexed as alignments (ADVICE round 1:
    b'@HD...\\n@CO\\n' used to raise IndexError)."""
    text = b"@HD\tVN:1.6\tSO:coordinate\n@CO\nq1\t0\t1\t5\t60\t4M\t*\t0\t0\tACGT\t*\tHP:i:1\tPC:i:7\tPS:i:3\n"
    cols = sv_phasing_fn.decode_sam_text(text)
    assert len(cols) == 1 and (int(cols.hp[0]), int(cols.pc[0]), int(cols.ps[0])) == (1, 7, 3)


def test_header_scan_matches_the_reference_loop(tmp_path):
    """write_file.header_text no longer splits every record (read_file) to find the ##contig lines: same lines, same
    order, same IndexError on a blank line as the reference's `l[0]` (write_file.py:31-40)."""
    from duet_b200 import write_file
    from oracle import ref_port
    home = tmp_path / "h"
    (home / "sv_calling").mkdir(parents=True)
    vcf = home / "sv_calling" / "variants.vcf"
    body = ("##fileformat=VCFv4.2\n##contig=<ID=2,length=5>\n##contig=<ID=chr1,length=9>  trailing tokens\n"
            "##contig=<ID=GL000,length=1>\n##other\t##contig=<ID=3,length=2>\n  ##contig=<ID=X,length=7>\n"
            "#CHROM\tPOS\n1\t5\t##contig=<ID=Y,length=3>\n")
    vcf.write_text(body)
    got = write_file.header_text(str(vcf), False)
    assert got == ref_port.header_text(str(vcf), False)
    assert "##contig=<ID=3," not in got and "##contig=<ID=Y," not in got and "##contig=<ID=X,length=7>\n" in got
    vcf.write_text(body + "\n")                              # a blank line
    with pytest.raises(IndexError):
        write_file.header_text(str(vcf), False)
    with pytest.raises(IndexError):
        ref_port.header_text(str(vcf), False)

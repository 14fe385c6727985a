#!/usr/bin/env python
"""Write tests/golden/spec_example.bam + spec_example.sam: a BAM laid out BYTE BY BYTE from the SAM/BAM
specification (SAMv1, section 4.2 "The BAM format", 4.1 "The BGZF compression format"), NOT with
tests/util_bam.py -- so the native reader (duet_b200/csrc/bam_decode.cpp) is checked against an
independent reading of the spec.  The alignments are the specification's own example (section 1.1:
r001, r002, r003 on `ref`), with WhatsHap's HP / PC / PS tags added; the .sam file is the text
`samtools view` prints for them, written by hand from the same section.

    python tests/golden/make_spec_bam.py

Every multi-byte integer below is spelled out little-endian, as the spec stores it.
"""
import os
import zlib

HERE = os.path.dirname(os.path.abspath(__file__))


def le(value: int, n_bytes: int) -> bytes:
    return bytes((value >> (8 * k)) & 0xFF for k in range(n_bytes))


def seq4(text: str) -> bytes:
    """4-bit packed bases, '=ACMGRSVTWYHKDBN' -> 0..15, high nibble first (SAMv1 4.2)."""
    code = "=ACMGRSVTWYHKDBN"
    out = bytearray()
    for i in range(0, len(text), 2):
        hi = code.index(text[i])
        lo = code.index(text[i + 1]) if i + 1 < len(text) else 0
        out.append((hi << 4) | lo)
    return bytes(out)


def cigar(ops) -> bytes:
    """oplen<<4 | op, op = index in 'MIDNSHP=X' (SAMv1 4.2)."""
    return b"".join(le((n << 4) | "MIDNSHP=X".index(op), 4) for n, op in ops)


def alignment(name, flag, pos0, mapq, bin_, ops, next_pos0, tlen, seq, aux: bytes) -> bytes:
    body = (le(0, 4)                      # refID = 0 ("ref")
            + le(pos0, 4)                 # pos, 0-based leftmost
            + le(len(name) + 1, 1)        # l_read_name, NUL included
            + le(mapq, 1)
            + le(bin_, 2)                 # bin (reg2bin)
            + le(len(ops), 2)             # n_cigar_op
            + le(flag, 2)
            + le(len(seq), 4)             # l_seq
            + le(0 if next_pos0 >= 0 else 0xFFFFFFFF, 4)     # next_refID
            + le(next_pos0 & 0xFFFFFFFF, 4)                  # next_pos
            + le(tlen & 0xFFFFFFFF, 4)                       # tlen
            + name.encode() + b"\x00"
            + cigar(ops)
            + seq4(seq)
            + b"\xff" * len(seq)          # qual: 0xFF = absent ('*')
            + aux)
    return le(len(body), 4) + body        # block_size


# aux fields: two-character tag, one-character type, value (SAMv1 4.2.4)
AUX_R001_FWD = (b"NM" + b"C" + le(1, 1)                       # NM:i:1   (uint8)
                + b"HP" + b"C" + le(1, 1)                     # HP:i:1
                + b"PC" + b"C" + le(60, 1)                    # PC:i:60
                + b"PS" + b"C" + le(7, 1))                    # PS:i:7
AUX_R002 = b"NM" + b"C" + le(0, 1)                            # not haplotagged
AUX_R003 = (b"SA" + b"Z" + b"ref,29,-,6H5M,17,0;" + b"\x00"   # SA:Z:...
            + b"HP" + b"C" + le(2, 1)                         # HP:i:2
            + b"PC" + b"S" + le(300, 2)                       # PC:i:300     (uint16)
            + b"PS" + b"I" + le(100000, 4))                   # PS:i:100000  (uint32)
AUX_R004 = (b"PS" + b"C" + le(7, 1)                           # tags in another order: the last three tokens are
            + b"HP" + b"C" + le(1, 1)                         # PS HP PC -> token [-2] is HP:i:1, not PC:i: -> row
            + b"PC" + b"C" + le(9, 1))                        # NOT kept (sv_phasing_fn.py:28)
AUX_R001_REV = (b"HP" + b"C" + le(2, 1) + b"PC" + b"C" + le(10, 1) + b"PS" + b"C" + le(7, 1))
AUX_R005 = (b"ML" + b"B" + b"C" + le(3, 4) + bytes([1, 2, 3])     # ML:B:C,1,2,3
            + b"XF" + b"f" + le(0x3FC00000, 4)                    # XF:f:1.5
            + b"XA" + b"A" + b"q"                                 # XA:A:q
            + b"XH" + b"H" + b"1AE3" + b"\x00"                    # XH:H:1AE3
            + b"Xs" + b"s" + le(-300 & 0xFFFF, 2)                 # Xs:i:-300  (int16)
            + b"HP" + b"c" + le(1, 1)                             # HP:i:1     (int8)
            + b"PC" + b"c" + le(-5 & 0xFF, 1)                     # PC:i:-5    (int8)
            + b"PS" + b"i" + le(42, 4))                           # PS:i:42    (int32)

RECORDS = [
    alignment("r001", 99, 6, 30, 4681, [(8, "M"), (2, "I"), (4, "M"), (1, "D"), (3, "M")], 36, 39, "TTAGATAAAGGATACTG", AUX_R001_FWD),
    alignment("r002", 0, 8, 30, 4681, [(3, "S"), (6, "M"), (1, "P"), (1, "I"), (4, "M")], -1, 0, "AAAAGATAAGGATA", AUX_R002),
    alignment("r003", 0, 8, 30, 4681, [(5, "S"), (6, "M")], -1, 0, "GCCTAAGCTAA", AUX_R003),
    alignment("r004", 0, 15, 30, 4681, [(6, "M"), (14, "N"), (5, "M")], -1, 0, "ATAGCTTCAGC", AUX_R004),
    alignment("r001", 147, 36, 30, 4681, [(9, "M")], 6, -39, "CAGCGGCAT", AUX_R001_REV),
    alignment("r005", 16, 28, 17, 4681, [(6, "H"), (5, "M")], -1, 0, "TAGGC", AUX_R005),
]

SAM_LINES = [
    "r001\t99\tref\t7\t30\t8M2I4M1D3M\t=\t37\t39\tTTAGATAAAGGATACTG\t*\tNM:i:1\tHP:i:1\tPC:i:60\tPS:i:7",
    "r002\t0\tref\t9\t30\t3S6M1P1I4M\t*\t0\t0\tAAAAGATAAGGATA\t*\tNM:i:0",
    "r003\t0\tref\t9\t30\t5S6M\t*\t0\t0\tGCCTAAGCTAA\t*\tSA:Z:ref,29,-,6H5M,17,0;\tHP:i:2\tPC:i:300\tPS:i:100000",
    "r004\t0\tref\t16\t30\t6M14N5M\t*\t0\t0\tATAGCTTCAGC\t*\tPS:i:7\tHP:i:1\tPC:i:9",
    "r001\t147\tref\t37\t30\t9M\t=\t7\t-39\tCAGCGGCAT\t*\tHP:i:2\tPC:i:10\tPS:i:7",
    "r005\t16\tref\t29\t17\t6H5M\t*\t0\t0\tTAGGC\t*\tML:B:C,1,2,3\tXF:f:1.5\tXA:A:q\tXH:H:1AE3\tXs:i:-300\tHP:i:1\tPC:i:-5\tPS:i:42",
]

# what sv_phasing_fn.py:28-29 keeps, in file order: (QNAME, HP, PS, PC)
KEPT = [("r001", 1, 7, 60), ("r003", 2, 100000, 300), ("r001", 2, 7, 10), ("r005", 1, 42, -5)]


def bgzf(payload: bytes, stored: bool) -> bytes:
    """One BGZF block (SAMv1 4.1): gzip member with the 'BC' extra subfield holding BSIZE = block size - 1."""
    if stored:
        # a single STORED deflate block, written out by hand (RFC 1951 3.2.4): BFINAL=1 BTYPE=00, LEN, ~LEN, bytes
        cdata = b"\x01" + le(len(payload), 2) + le(len(payload) ^ 0xFFFF, 2) + payload
    else:
        co = zlib.compressobj(9, zlib.DEFLATED, -15)
        cdata = co.compress(payload) + co.flush()
    block_size = 12 + 6 + len(cdata) + 8
    return (b"\x1f\x8b"               # ID1 ID2
            + b"\x08"                 # CM = deflate
            + b"\x04"                 # FLG = FEXTRA
            + le(0, 4)                # MTIME
            + b"\x00"                 # XFL
            + b"\xff"                 # OS = unknown
            + le(6, 2)                # XLEN
            + b"BC" + le(2, 2) + le(block_size - 1, 2)
            + cdata
            + le(zlib.crc32(payload) & 0xFFFFFFFF, 4)
            + le(len(payload), 4))    # ISIZE


# the 28-byte end-of-file marker, verbatim from SAMv1 4.1.2
EOF_MARKER = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def main():
    header_text = b"@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:ref\tLN:45\n"
    raw = (b"BAM\x01" + le(len(header_text), 4) + header_text
           + le(1, 4)                                   # n_ref
           + le(4, 4) + b"ref\x00" + le(45, 4))         # l_name, name, l_ref
    raw += b"".join(RECORDS)
    # three data blocks; the cuts fall INSIDE records (mid-header of r002, inside r003's aux string)
    cut1 = raw.index(b"r002") - 20
    cut2 = raw.index(b"6H5M,17") + 3
    blob = bgzf(raw[:cut1], stored=True) + bgzf(raw[cut1:cut2], stored=False) + bgzf(raw[cut2:], stored=True) + EOF_MARKER
    with open(os.path.join(HERE, "spec_example.bam"), "wb") as f:
        f.write(blob)
    with open(os.path.join(HERE, "spec_example.sam"), "w") as f:
        f.write("\n".join(SAM_LINES) + "\n")
    print(f"spec_example.bam: {len(blob)} bytes, {len(RECORDS)} records, {len(KEPT)} kept")


if __name__ == "__main__":
    main()

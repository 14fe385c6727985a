"""Minimal BAM writer for tests (BGZF blocks via zlib) and the `samtools view`-style text of the
same records, so the native BAM decoder can be checked against the SAM-text decoder record by record."""
import struct
import zlib

_SEQ_CODE = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}


def _bgzf_block(data: bytes) -> bytes:
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    comp = co.compress(data) + co.flush()
    bsize = len(comp) + 25
    return (b"\x1f\x8b\x08\x04" + b"\0\0\0\0" + b"\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize)
            + comp + struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))


def _aux(tag: str, typ: str, val) -> bytes:
    head = tag.encode() + typ.encode()
    if typ == "A":
        return head + val.encode()
    if typ in "cCsSiI":
        return head + struct.pack("<" + {"c": "b", "C": "B", "s": "h", "S": "H", "i": "i", "I": "I"}[typ], val)
    if typ == "f":
        return head + struct.pack("<f", val)
    if typ in "ZH":
        return head + val.encode() + b"\0"
    if typ == "B":
        sub, arr = val
        return head + sub.encode() + struct.pack("<i", len(arr)) + b"".join(
            struct.pack("<" + {"c": "b", "C": "B", "s": "h", "S": "H", "i": "i", "I": "I", "f": "f"}[sub], x) for x in arr)
    raise ValueError(typ)


def _aux_text(tag, typ, val) -> str:
    if typ in "cCsSiI":
        return f"{tag}:i:{val}"
    if typ == "f":
        return f"{tag}:f:{val:g}"
    if typ == "B":
        return f"{tag}:B:{val[0]}," + ",".join(str(x) for x in val[1])
    return f"{tag}:{typ}:{val}"


def record(name, pos, seq, qual, aux):
    """aux: list of (tag, type, value) in file order.  Returns (BAM bytes, SAM text line)."""
    l_seq = len(seq)
    packed = bytearray((l_seq + 1) // 2)
    for i, c in enumerate(seq):
        packed[i >> 1] |= _SEQ_CODE[c] << (4 if i % 2 == 0 else 0)
    q = bytes([0xFF] * l_seq) if qual == "*" else bytes(ord(c) - 33 for c in qual)
    cigar = struct.pack("<I", (l_seq << 4) | 0) if l_seq else b""
    body = struct.pack("<iiBBHHHIiii", 0, pos - 1, len(name) + 1, 60, 4680, 1 if l_seq else 0, 0, l_seq, -1, -1, 0)
    body += name.encode() + b"\0" + cigar + bytes(packed) + q + b"".join(_aux(*a) for a in aux)
    text = "\t".join([name, "0", "1", str(pos), "60", f"{l_seq}M" if l_seq else "*", "*", "0", "0", seq or "*",
                      qual if l_seq else "*"] + [_aux_text(*a) for a in aux])
    return struct.pack("<i", len(body)) + body, text + "\n"


def write_bam(path, records, block=60000):
    """records: list of BAM record bytes."""
    hdr_text = b"@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:1\tLN:249250621\n"
    raw = b"BAM\1" + struct.pack("<i", len(hdr_text)) + hdr_text + struct.pack("<i", 1) + \
          struct.pack("<i", 2) + b"1\0" + struct.pack("<i", 249250621) + b"".join(records)
    with open(path, "wb") as f:
        for i in range(0, len(raw), block):
            f.write(_bgzf_block(raw[i:i + block]))
        f.write(_bgzf_block(b""))            # EOF marker

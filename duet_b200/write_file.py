"""phased_sv.vcf writer: same functions and byte-identical output as
/root/reference/src/duet/write_file.py (print_sv :6-17, print_sv_header :19-44)."""
from __future__ import annotations

import logging

from .read_file import init_chrom_list

_FIXED_HEADER = "".join([
    "##fileformat=VCFv4.2\n",
    "##source=Duet\n",
    '##ALT=<ID=INS,Description="Insertion of novel sequence relative to the reference">\n',
    '##ALT=<ID=DEL,Description="Deletion relative to the reference">\n',
    '##FILTER=<ID=PASS,Description="SV calls passed phasing criterion">\n',
    '##INFO=<ID=SVLEN,Number=1,Type=Integer,Description="Estimated length of the variant">\n',
    '##FORMAT=<ID=HP,Number=1,Type=String,Description="Haplotype of the SV call">\n',
    '##FORMAT=<ID=PS,Number=1,Type=String,Description="Phase set which the SV call belongs to">\n',
])
_COLUMNS = "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tVALUE\n"


def format_rows(phased_callset, first_id: int = 1) -> str:
    """Rows 'chrom pos Duet.<idx> ref alt . PASS SVLEN=..;SVTYPE=<..> HP:PS hp:ps' (:14-16).
    `first_id` lets each GPU of a contig-sharded run number its own slice."""
    parts = []
    for idx, c in enumerate(phased_callset, first_id):
        parts.append(f"{c['chrom']}\t{c['pos']}\tDuet.{idx}\t{c['ref']}\t{c['alt']}\t.\tPASS\t"
                     f"SVLEN={c['svlen']};SVTYPE=<{c['svtype']}>\tHP:PS\t{c['hp']}:{c['ps']}\n")
    return "".join(parts)


def print_sv(phased_callset, output_path):
    logging.info("write phased callset into .vcf file")
    with open(output_path, "a") as f:
        f.write(format_rows(phased_callset))


def _contig_tokens(vcf_path) -> list:
    """The first whitespace-separated token of every line whose first token contains '##contig=<ID=', in file
    order -- all the reference's header loop (:31-40) ever matches -- without splitting the 20 MB of RNAMES lists
    the way `read_file` does.  A blank line raises IndexError, as `l[0]` does there."""
    out = []
    with open(vcf_path, "r") as fh:
        for ln in fh:
            if "##contig=<ID=" in ln:
                tok = ln.split(None, 1)[0]
                if "##contig=<ID=" in tok:
                    out.append(tok)
            elif ln.isspace():
                raise IndexError("list index out of range")
    return out


def header_text(vcf_path, include_all_ctgs) -> str:
    tokens = _contig_tokens(vcf_path)
    chrom_list = init_chrom_list(include_all_ctgs, vcf_path[:len(vcf_path) - 24])
    out = [_FIXED_HEADER]
    if not include_all_ctgs:
        for ctg in chrom_list[:24]:                       # contig lines re-ordered to the chrom list (:33-37)
            a, b = "##contig=<ID=chr" + ctg + ",", "##contig=<ID=" + ctg + ","
            out += [t + "\n" for t in tokens if a in t or b in t]
    else:
        out += [t + "\n" for t in tokens]
    out.append(_COLUMNS)
    return "".join(out)


def print_sv_header(vcf_path, output_path, include_all_ctgs):
    with open(output_path, "w") as f:
        f.write(header_text(vcf_path, include_all_ctgs))

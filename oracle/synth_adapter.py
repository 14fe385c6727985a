"""TEST INFRASTRUCTURE: feed a `duet_b200.synth.SynthSample` to the oracle port without the
text round trip (names and tags go straight into the dicts / records the port works on).
Equivalent to ref_port.generate_phased_callset on synth.write_workdir(sample) -- checked by
tests/test_oracle_golden.py::test_adapter_equals_text_path."""
from __future__ import annotations

from duet_b200 import synth as sy

from . import ref_port


def tables_and_records(sample: sy.SynthSample):
    """(contig names, per-contig QNAME dicts, per-contig SvRecord lists) in chrom_list order of
    the sample's contigs."""
    names, tables, per_contig = [], [], []
    for c in sample.contigs:
        names.append(c.name)
        table = {}
        rn = sy.name_strings(c.row_id)
        hp, ps, pc, tagged = c.row_hp.tolist(), c.row_ps.tolist(), c.row_pc.tolist(), c.row_tagged.tolist()
        for i, nm in enumerate(rn):
            if tagged[i]:
                table[nm] = (hp[i], ps[i], pc[i])          # later rows overwrite (sv_phasing_fn.py:29)
        tables.append(table)
        cn = ("chr" if sample.chr_prefix else "") + c.name
        sup = sy.name_strings(c.sup_id)
        recs = []
        for i in range(c.sv_pos.shape[0]):
            t = sy.SVTYPES[int(c.sv_type[i])]
            r = ref_port.SvRecord([cn, str(int(c.sv_pos[i])), ".", "N", "<" + t + ">"])
            r.svlen = int(c.sv_len[i])
            r.svtype = t
            r.svread = int(c.sv_svread[i])
            r.names = sup[int(c.sup_off[i]):int(c.sup_off[i + 1])]
            r.gt = sy.GTS[int(c.sv_gt[i])]
            r.refread = int(c.sv_refread[i])
            r.altread = 0
            recs.append(r)
        per_contig.append(recs)
    return names, tables, per_contig


def phase_sample(sample: sy.SynthSample, svlen_thres=50, suppread_thres=2, trace=None):
    """Oracle rows for one sample; `trace` collects (contig idx, class, record, pred, features)."""
    names, tables, per_contig = tables_and_records(sample)
    flat = ref_port.join_support_reads(per_contig, tables)
    rows = ref_port.phase_records(flat, names, svlen_thres, suppread_thres, trace)
    return rows, flat

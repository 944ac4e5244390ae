"""The CPU restatement (oracle/port.py) against the reference's own outputs (committed fixtures)."""
import pytest

from oracle import port, compare
from tests import util, golden_util as G


@pytest.mark.parametrize("name", G.case_names())
def test_port_matches_reference_files(name):
    c = G.load_case(name)
    kw = G.args_to_kw(c["meta"]["args"])
    got, _ = util.oracle_outputs(c["vcf"], c["sams"], mapq=c["meta"]["mapq"], paired_end=c["meta"]["paired_end"], **kw)
    bad = compare.diff_outputs(c["ref"], got)
    assert not bad, "\n".join(bad)


@pytest.mark.parametrize("name", G.case_names())
def test_port_mapper_matches_reference_tsv(name):
    c = G.load_case(name)
    kw = G.args_to_kw(c["meta"]["args"])
    vt, st, batches, col, fd = util.load_inputs(c["vcf"], c["sams"], mapq=c["meta"]["mapq"], paired_end=c["meta"]["paired_end"],
                                                remove_dups=1, pass_only=kw.get("pass_only", 1),   # the golden TSVs were made without duplicates
                                                id_separator="_", gw_phase_method=0,   # ... and with the default table
                                                include_indels=kw.get("include_indels", 0))
    for b, batch in zip(c["meta"]["bams"], batches):
        tup = port.map_reads(batch, vt, 10, 0.0)      # the golden TSV was made with isize 0
        assert port.tuples_tsv(batch, vt, tup) == c["mapper"][b]

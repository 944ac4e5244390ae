"""The CPU restatement (oracle/port.py) against the reference's own outputs (committed fixtures)."""
import pytest

from oracle import port, compare
from tests import util, golden_util as G


@pytest.mark.parametrize("name", G.case_names())
def test_port_matches_reference_files(name):
    c = G.load_case(name)
    kw = G.args_to_kw(c["meta"]["args"])
    okw = dict(kw)
    if "exclude" in okw:
        okw["haplo_count_bam_exclude"] = okw.pop("exclude")
    got, _ = util.oracle_outputs(c["vcf"], c["sams"], **okw)
    bad = compare.diff_outputs(c["ref"], got)
    assert not bad, "\n".join(bad)


@pytest.mark.parametrize("name", G.case_names())
def test_port_mapper_matches_reference_tsv(name):
    c = G.load_case(name)
    vt, st, batches, col, fd = util.load_inputs(c["vcf"], c["sams"])
    isz = G.args_to_kw(c["meta"]["args"]).get("isize", [0.0])
    for b, batch in zip(c["meta"]["bams"], batches):
        tup = port.map_reads(batch, vt, 10, 0.0)      # the golden TSV was made with isize 0
        assert port.tuples_tsv(batch, vt, tup) == c["mapper"][b]

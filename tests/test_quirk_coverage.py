"""Every engine-side quirk of SURVEY.md section 8a is PINNED by a fixture the unmodified reference produced.

For each quirk the oracle port has a mutant that follows the presumably intended behaviour instead of the reference's
(oracle/port.py: MUTANTS).  A fixture pins a quirk exactly when that mutant no longer reproduces the reference's files --
so a port (and, through the parity tests on the same fixtures, a product) that matches them all provably implements the
quirk and not its repair.  Without a mutant the port reproduces every fixture (tests/test_oracle_golden.py)."""
import pytest

from oracle import compare, port
from tests import util, golden_util as G

PINNED_BY = {
    "Q9": ["q9_shared_qnames"],            # phaser.py:578   read_vars of a QNAME shared by two BAMs is overwritten
    "Q14": ["fuzz_snvs", "fuzz_indels"],   # phaser.py:2152  split_start = used_vars mis-slices the merges that follow
    "Q16": ["q16_tie_glue"],               # phaser.py:708-726 tie edges glue components but carry no allele links
    "Q22": ["q22_unphased_blocks"],        # phaser.py:935-939 phase_concordant = 1 when no phase is known
    "Q23": ["rna_small", "q22_unphased_blocks"],   # phaser.py:962-964 nan objects are distinct set members
    "Q28": ["rna_small", "opt_chr"],       # phaser.py:966, 970-980 gwStat prints 1 or 1.0 depending on the branch taken
}


def _port_files(case):
    c = G.load_case(case)
    kw = G.args_to_kw(c["meta"]["args"])
    got, _ = util.oracle_outputs(c["vcf"], c["sams"], mapq=c["meta"]["mapq"], paired_end=c["meta"]["paired_end"], **kw)
    return c, got


@pytest.mark.parametrize("quirk", sorted(PINNED_BY))
def test_a_port_without_the_quirk_fails_its_fixture(quirk):
    try:
        for case in PINNED_BY[quirk]:
            port.MUTANTS = set()
            c, got = _port_files(case)
            assert not compare.diff_outputs(c["ref"], got), "the faithful port must reproduce " + case
            port.MUTANTS = {quirk}
            c, got = _port_files(case)
            assert compare.diff_outputs(c["ref"], got), "%s does not discriminate %s" % (case, quirk)
    finally:
        port.MUTANTS = set()


def test_shared_qnames_really_occur_in_both_bams():
    c = G.load_case("q9_shared_qnames")
    names = [set(ln.split("\t", 1)[0] for ln in open(p) if ln[0] != "@") for p in c["sams"]]
    assert len(names[0] & names[1]) > 100


def test_duplicate_basenames_get_numbered_display_names():
    c = G.load_case("q26_same_basename")
    assert util.bam_display_names(c["sams"]) == ["x.1", "x.2"]
    assert "\tx.2\t" in c["ref"]["haplotypic_counts"] and "\tx.1\t" not in c["ref"]["haplotypic_counts"]   # BAM 1 is excluded


def test_chr_restricts_the_sites_and_the_output_vcf():
    c = G.load_case("opt_chr")
    contigs = set(ln.split("\t", 1)[0] for ln in c["ref"]["allelic_counts"].splitlines()[1:])
    assert contigs == {"22"}
    # write_vcf re-reads the input through `tabix -h VCF chr:` as well (phaser.py:1679-1680)
    assert {ln.split("\t", 1)[0] for ln in c["ref"]["vcf"].splitlines() if ln[0] != "#"} == {"22"}

"""Feature-level haplotypic counts: oracle/port_gene_ae.py and the product
(phaser_b200/phaser_gene_ae.py over phz_gene_ae_pairs) against outputs of the UNMODIFIED reference script
(tests/golden/gene_ae, made by tests/golden/make_golden.py)."""
import json
import os
import random

import pytest

from oracle import port_gene_ae as pg
from phaser_b200 import phaser_gene_ae as ga
from tests import util, golden_util as G

GA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gene_ae")
NAMES = sorted(os.listdir(GA))


def _load(name):
    d = os.path.join(GA, name)
    m = json.load(open(os.path.join(d, "case.json")))
    kw = {}
    it = iter(m["args"])
    for a in it:
        v = next(it)
        kw[a[2:]] = int(v) if a == "--min_cov" else float(v)
    hc = os.path.join(G.CASES, m["base"], "ref.haplotypic_counts.txt")
    return hc, os.path.join(d, "features.bed"), open(os.path.join(d, "ref.gene_ae.txt")).read(), kw, m["args"]


def _seeded(tmp_path, engine, seed, n_pairs):
    """haplotypic_counts.txt of a seeded 2-BAM run of the product + random features"""
    vcf, sams = util.make_case(tmp_path, seed, 400, n_pairs, n_bams=2, switch_per_base=0.02)
    got, _, _ = util.product_outputs(engine, vcf, sams, max_block_size=6)
    rnd = random.Random(seed)
    feats = []
    for c, L in (("21", 300000), ("22", 200000), ("X", 1000)):
        for _ in range(150):
            a = rnd.randrange(0, L - 10)
            feats.append("%s\t%d\t%d\tg%d\n" % (c, a, min(L, a + rnd.choice([1, 50, 800, 5000, 60000])), len(feats)))
    return got["haplotypic_counts"], "".join(feats)


@pytest.mark.parametrize("name", NAMES)
def test_port_matches_reference_gene_ae(name):
    hc, bed, ref, kw, _ = _load(name)
    assert pg.canon(pg.run(open(hc).read(), open(bed).read(), **kw)) == pg.canon(ref)


@pytest.mark.parametrize("name", NAMES)
def test_hostsim_matches_reference_gene_ae(hostsim, name):
    hc, bed, ref, kw, _ = _load(name)
    assert pg.canon(ga.run_text(hostsim, open(hc).read(), open(bed).read(), **kw)) == pg.canon(ref)


def test_gene_ae_cli_writes_the_reference_file(engine, tmp_path, capsys):
    hc, bed, ref, kw, args = _load("two_bams_maf_cov")
    o = str(tmp_path / "ae.txt")
    ga.run(ga.build_parser().parse_args(["--haplotypic_counts", hc, "--features", bed, "--o", o] + args), engine=engine)
    assert pg.canon(open(o).read()) == pg.canon(ref)
    with pytest.raises(SystemExit) as e:
        ga.run(ga.build_parser().parse_args(["--haplotypic_counts", hc, "--features", bed, "--o", o, "--min_haplo_maf", "0.7"]),
               engine=engine)
    assert e.value.code == 1
    with pytest.raises(SystemExit) as e:        # wrong separator: the reference's ERROR + exit 1 (:181-184)
        ga.run(ga.build_parser().parse_args(["--haplotypic_counts", hc, "--features", bed, "--o", o, "--id_separator", "-"]),
               engine=engine)
    assert e.value.code == 1


def test_hostsim_matches_port_on_a_seeded_run(hostsim, tmp_path):
    hc, feats = _seeded(tmp_path, hostsim, 71, 4000)
    for kw in ({}, {"gw_cutoff": 0.7, "min_cov": 2}, {"min_haplo_maf": 0.3}):
        assert ga.run_text(hostsim, hc, feats, **kw) == pg.run(hc, feats, **kw)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_matches_reference_gene_ae(gpu, name):
    hc, bed, ref, kw, _ = _load(name)
    assert pg.canon(ga.run_text(gpu, open(hc).read(), open(bed).read(), **kw)) == pg.canon(ref)


@pytest.mark.gpu
def test_gpu_matches_port_on_a_seeded_run(gpu, tmp_path):
    hc, feats = _seeded(tmp_path, gpu, 72, 30000)
    for kw in ({}, {"gw_cutoff": 0.7, "min_cov": 2}, {"min_haplo_maf": 0.3}):
        assert ga.run_text(gpu, hc, feats, **kw) == pg.run(hc, feats, **kw)

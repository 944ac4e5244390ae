"""SoA cache file: "parse once" (SURVEY.md section 8f row N2).

The ingest of a BAM (inflate, record decode, QNAME dictionary, SoA build) is the longest stage of the command line.
With `--soa_cache DIR` (or PHZ_SOA_CACHE=DIR) its result -- the packed SoA arrays of include/phz.h::phz_reads plus the
QNAME dictionary -- is kept under DIR, keyed by the BAM's path, size and modification time, the VCF's contig list and
the read filters; a later run on the same inputs (other phasing parameters, another sample column of the same VCF
contigs, a re-run of a batch) maps the arrays back in instead of parsing.  Scope: the FIRST BAM of a run (its fragment
ids are the dictionary's first ids; later BAMs share that namespace and are parsed).  Arrays are plain `.npy` files read
through the page cache (numpy memory maps): nothing is copied before the upload to the device.
"""
import hashlib
import json
import os

import numpy as np

from .layout import ReadBatch

ARRAYS = ("contig_rec_off", "pos", "tlen", "aln_score", "frag", "cigar_off", "cigar", "seq_off", "seq", "qual")
VERSION = 1


def cache_key(bam, contigs, remove_dups, proper_pair, min_mapq):
    st = os.stat(bam)
    h = hashlib.sha1()
    h.update(json.dumps([VERSION, os.path.abspath(bam), st.st_size, st.st_mtime_ns, list(contigs), bool(remove_dups),
                         bool(proper_pair), int(min_mapq)]).encode())
    return h.hexdigest()[:24]


def load(cache_dir, key, fragdict, need_names, threads=0):
    """ReadBatch from the cache, or None.  `fragdict` must be empty; it is filled only when the run needs the names
    (another BAM follows, or --output_read_ids)."""
    d = os.path.join(cache_dir, key)
    if not os.path.isfile(os.path.join(d, "done")) or len(fragdict) != 0:
        return None
    try:
        a = {k: np.load(os.path.join(d, k + ".npy"), mmap_mode="r") for k in ARRAYS}
        meta = json.load(open(os.path.join(d, "meta.json")))
        if need_names:
            fragdict.load(np.load(os.path.join(d, "names.npy"), mmap_mode="r"), np.load(os.path.join(d, "names_off.npy")), threads)
    except (OSError, ValueError):
        return None
    rb = ReadBatch(int(meta["n_contigs"]), np.asarray(a["contig_rec_off"]), a["pos"], a["tlen"], a["aln_score"], a["frag"],
                   a["cigar_off"], a["cigar"], a["seq_off"], a["seq"], a["qual"], None)
    rb.n_fragments = int(meta["n_fragments"])
    return rb


def save(cache_dir, key, rb: ReadBatch, fragdict):
    """Write the batch of the run's FIRST BAM (fragment ids 0 .. n-1 are exactly this BAM's)."""
    d = os.path.join(cache_dir, key)
    os.makedirs(d, exist_ok=True)
    for k in ARRAYS:
        np.save(os.path.join(d, k + ".npy"), np.ascontiguousarray(getattr(rb, k)))
    blob, off = fragdict.export()
    np.save(os.path.join(d, "names.npy"), blob); np.save(os.path.join(d, "names_off.npy"), off)
    json.dump({"n_contigs": rb.n_contigs, "n_fragments": len(fragdict), "records": int(rb.pos.shape[0])}, open(os.path.join(d, "meta.json"), "w"))
    open(os.path.join(d, "done"), "w").close()

"""Minimal stand-in for the `intervaltree` package (absent from this image), TEST INFRASTRUCTURE only: lets the
UNMODIFIED reference phaser_gene_ae.py run when golden fixtures are generated.  Implements what that script uses:
`tree[a:b] = data` and `tree[a:b]` -> set of Interval(begin, end, data) overlapping the half-open range, with the
package's semantics (overlap iff iv.begin < b and iv.end > a; null intervals are rejected)."""
import collections

Interval = collections.namedtuple("Interval", ["begin", "end", "data"])


class IntervalTree:
    def __init__(self):
        self.items = []

    def __setitem__(self, index, data):
        if index.start >= index.stop:
            raise ValueError("IntervalTree: Null Interval objects not allowed in IntervalTree")
        self.items.append(Interval(index.start, index.stop, data))

    def __getitem__(self, index):
        a, b = index.start, index.stop
        if a >= b:
            return set()
        return set(iv for iv in self.items if iv.begin < b and iv.end > a)

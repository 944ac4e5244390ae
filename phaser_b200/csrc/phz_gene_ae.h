// Feature-level haplotypic counts: the join + distinct-read counting of phaser_gene_ae.py
// (phaser_gene_ae/phaser_gene_ae.py:95-101 interval query, :172-219 variant_feature_reads) -- SURVEY.md 8f row N1.
//
// Input: the rows of haplotypic_counts.txt as arrays (per row: contig, start, stop, per-variant position and the two
// per-variant read-id lists, ids renumbered densely per (row, haplotype) by the host parser) and the features sorted by
// (contig, start).  Output: one entry per (row, overlapping feature) with the number of DISTINCT reads on each
// haplotype among the row's variants that lie inside the feature.  The order-dependent fold over rows (phased sums /
// best unphased block, phaser_gene_ae.py:103-141) and the text stay on the host: they are O(pairs) and cheap.
//
// Kernels (generic for_each of the backend, like the rest of the pipeline): count overlaps per row (two binary
// searches over the sorted features: first start >= row stop, first running-max-of-stops > row start-1), scan, emit
// pairs, size one bitmap per pair (ceil(ids/32) words per haplotype), one warp per pair marks the ids of the
// variants inside the feature with atomicOr, one thread per pair pop-counts.  HBM/atomic bound; bytes = lists read
// once per overlapping feature + bitmap words.
#pragma once
#include "phz_backend.h"
#include "../../include/phz.h"

namespace phz {

PHZ_HD u32 popc32(u32 x) {
#if defined(__CUDA_ARCH__)
  return (u32)__popc(x);
#else
  return (u32)__builtin_popcount(x);
#endif
}

template <class B>
struct GeneAE {
  B* be = nullptr;
  Buf<B, u32> cnt, off, ae_row, ae_feat, ae_a, ae_b, words, woff, bits;
  int64_t NPAIR = 0;
  void bind(B* b) {
    be = b;
    cnt.bind(b); off.bind(b); ae_row.bind(b); ae_feat.bind(b); ae_a.bind(b); ae_b.bind(b); words.bind(b); woff.bind(b); bits.bind(b);
  }
  u32 fetch(const u32* p) { u32 v = 0; be->d2h(&v, p, sizeof(u32)); return v; }

  // first j in [lo, hi) with a[j] >= key (a ascending)
  PHZ_HD static int64_t lb_ge(const int32_t* a, int64_t lo, int64_t hi, int32_t key) {
    while (lo < hi) { int64_t m = (lo + hi) >> 1; if (a[m] < key) lo = m + 1; else hi = m; }
    return lo;
  }
  // first j in [lo, hi) with a[j] > key (a non-decreasing)
  PHZ_HD static int64_t lb_gt(const int32_t* a, int64_t lo, int64_t hi, int32_t key) {
    while (lo < hi) { int64_t m = (lo + hi) >> 1; if (a[m] <= key) lo = m + 1; else hi = m; }
    return lo;
  }

  int64_t run(const phz_ae_input& in) {
    const int64_t R = in.n_rows;
    const int32_t* rc = in.row_contig; const int32_t* rs = in.row_start; const int32_t* re = in.row_stop;
    const u32* ra = in.row_a; const u32* rb = in.row_b; const u32* rna = in.row_ids_a; const u32* rnb = in.row_ids_b;
    const u32* voff = in.var_off; const int32_t* vpos = in.var_pos; const u32* ioff = in.id_off; const u32* ids = in.ids;
    const int32_t* fs = in.f_start; const int32_t* fe = in.f_stop; const int32_t* fm = in.f_maxstop; const int64_t* fco = in.f_contig_off;
    be->stage("gene_ae.join");
    u32* c = cnt.ensure(R + 1); u32* o = off.ensure(R + 2);
    // tree[start-1 : stop] (phaser_gene_ae.py:97): features with f.start < stop and f.stop > start-1
    be->for_each(R, PHZ_LAMBDA(int64_t r) {
      int32_t ct = rc[r]; u32 n = 0;
      if (ct >= 0) {
        int32_t a = rs[r] - 1, b = re[r];
        int64_t hi = lb_ge(fs, fco[ct], fco[ct + 1], b);
        int64_t lo = lb_gt(fm, fco[ct], hi, a);
        for (int64_t j = lo; j < hi; ++j) n += fe[j] > a ? 1u : 0u;
      }
      c[r] = n;
    });
    be->exclusive_scan_u32(c, o, R);
    NPAIR = R > 0 ? (int64_t)fetch(o + R) : 0;
    u32* pr = ae_row.ensure(NPAIR); u32* pf = ae_feat.ensure(NPAIR); u32* pa = ae_a.ensure(NPAIR); u32* pb = ae_b.ensure(NPAIR);
    be->for_each(R, PHZ_LAMBDA(int64_t r) {
      int32_t ct = rc[r];
      if (ct < 0 || c[r] == 0) return;
      int32_t a = rs[r] - 1, b = re[r];
      int64_t hi = lb_ge(fs, fco[ct], fco[ct + 1], b);
      int64_t lo = lb_gt(fm, fco[ct], hi, a);
      u32 w = o[r];
      for (int64_t j = lo; j < hi; ++j) if (fe[j] > a) { pr[w] = (u32)r; pf[w] = (u32)j; ++w; }
    });
    be->stage("gene_ae.count");
    // one bitmap per (pair, haplotype); rows with a single variant carry no id lists (phaser_gene_ae.py:196-200)
    u32* wd = words.ensure(NPAIR + 1); u32* wo = woff.ensure(NPAIR + 2);
    be->for_each(NPAIR, PHZ_LAMBDA(int64_t p) {
      u32 r = pr[p];
      wd[p] = (voff[r + 1] - voff[r] == 1) ? 0u : ((rna[r] + 31) / 32 + (rnb[r] + 31) / 32);
    });
    be->exclusive_scan_u32(wd, wo, NPAIR);
    int64_t NW = NPAIR > 0 ? (int64_t)fetch(wo + NPAIR) : 0;
    u32* bm = bits.ensure(NW + 1); be->memset0(bm, (NW + 1) * sizeof(u32));
    // a variant counts for the feature iff (pos-1) - f.start >= 0 and (pos-1) - f.stop <= 0 (phaser_gene_ae.py:191)
    be->for_each_warp(NPAIR, PHZ_LAMBDA(int64_t p, int lane, int nl) {
      u32 r = pr[p]; u32 j = pf[p];
      if (voff[r + 1] - voff[r] == 1) return;
      u32* ba = bm + wo[p]; u32* bb = ba + (rna[r] + 31) / 32;
      for (u32 v = voff[r]; v < voff[r + 1]; ++v) {
        int32_t q = vpos[v] - 1;
        if (q < fs[j] || q > fe[j]) continue;
        for (u32 i = ioff[2 * v] + lane; i < ioff[2 * v + 1]; i += nl) atomic_or(&ba[ids[i] >> 5], 1u << (ids[i] & 31));
        for (u32 i = ioff[2 * v + 1] + lane; i < ioff[2 * v + 2]; i += nl) atomic_or(&bb[ids[i] >> 5], 1u << (ids[i] & 31));
      }
    });
    be->for_each(NPAIR, PHZ_LAMBDA(int64_t p) {
      u32 r = pr[p]; u32 j = pf[p];
      if (voff[r + 1] - voff[r] == 1) {
        int32_t q = vpos[voff[r]] - 1;
        bool used = !(q < fs[j] || q > fe[j]);
        pa[p] = used ? ra[r] : 0u; pb[p] = used ? rb[r] : 0u;       // len(set(range(aCount))), phaser_gene_ae.py:199-200
        return;
      }
      const u32* ba = bm + wo[p]; u32 na = (rna[r] + 31) / 32, nb = (rnb[r] + 31) / 32;
      u32 sa = 0, sb = 0;
      for (u32 k = 0; k < na; ++k) sa += popc32(ba[k]);
      for (u32 k = 0; k < nb; ++k) sb += popc32(ba[na + k]);
      pa[p] = sa; pb[p] = sb;
    });
    be->stage("gene_ae.end");
    return NPAIR;
  }
};

}  // namespace phz

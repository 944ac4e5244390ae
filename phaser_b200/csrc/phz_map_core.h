// K1 per-record logic: CIGAR walk -> spliced segments -> (record, segment, variant, allele class).
//
// Replaces read_variant_map.py of the reference: the streaming join (read_variant_map.py:25-123),
// split_read (:165-234) and identify_allele (:236-258).  For coordinate-sorted input the join is a
// pure predicate per (record, segment, variant) (SURVEY.md "K1 spec"): a het SNV at VCF position p is
// tested against segment s of a record at POS iff 0 <= p-(POS+seg_start) < len(pseudo_read(s)).
// Reference quirks kept on purpose: insertion keys are whole-read reference offsets but are looked
// up with segment-relative offsets (Q3); an insertion right after the SNV base turns the allele into
// a multi-base string (class "other", Q4); deletion placeholders and IUPAC 'D' are stripped (Q5).
#pragma once
#include "phz_backend.h"

namespace phz {

struct ReadsView {
  int64_t n_records;
  int n_contigs;
  const int64_t* contig_rec_off;   // [n_contigs+1] (device)
  const int32_t* pos;
  const int32_t* tlen;
  const int16_t* aln_score;
  const u32* frag;
  const u32* cigar_off;            // [R+1]
  const u32* cigar;
  const u64* seq_off;              // [R+1]
  const u8* seq;                   // nibble packed
  const u8* qual;
  int64_t n_cigar_ops;
  // field accessors (the tile kernel substitutes a view whose slabs live in shared memory)
  PHZ_HD int32_t pos_at(int64_t r) const { return pos[r]; }
  PHZ_HD int32_t tlen_at(int64_t r) const { return tlen[r]; }
  PHZ_HD u32 cig_lo(int64_t r) const { return cigar_off[r]; }
  PHZ_HD u32 cig_hi(int64_t r) const { return cigar_off[r + 1]; }
  PHZ_HD u32 cigar_at(u32 k) const { return cigar[k]; }
  PHZ_HD u64 seq_off_at(int64_t r) const { return seq_off[r]; }
  PHZ_HD int aln_at(int64_t r) const { return aln_score[r]; }
};

// The same record fields, served from the shared-memory slabs a tile's TMA bulk copies filled
// (records [r0, r0+256)); CIGAR words beyond the staged slab fall back to global memory.
template <bool ALL_CIGAR_STAGED>
struct TileRV {
  int64_t r0;
  const int32_t* s_pos; const int32_t* s_tlen; const u32* s_coff; const u32* s_cig; const u64* s_soff; const int16_t* s_as;
  u32 cig_base, cig_n;
  const u32* cigar; const u8* seq; const u8* qual;
  PHZ_HD int32_t pos_at(int64_t r) const { return s_pos[r - r0]; }
  PHZ_HD int32_t tlen_at(int64_t r) const { return s_tlen[r - r0]; }
  PHZ_HD u32 cig_lo(int64_t r) const { return s_coff[r - r0]; }
  PHZ_HD u32 cig_hi(int64_t r) const { return s_coff[r - r0 + 1]; }
  PHZ_HD u32 cigar_at(u32 k) const {
    u32 i = k - cig_base;
    if (ALL_CIGAR_STAGED) return s_cig[i];
    return i < cig_n ? s_cig[i] : cigar[k];
  }
  PHZ_HD u64 seq_off_at(int64_t r) const { return s_soff[r - r0]; }
  PHZ_HD int aln_at(int64_t r) const { return s_as[r - r0]; }
};

struct VariantsView {
  int64_t n_variants;
  int n_contigs;
  const int64_t* contig_var_off;   // [n_contigs+1] (device)
  const int32_t* pos;
  const u8* a0;
  const u8* a1;
  static constexpr bool kIndels = false;
};

// --include_indels 1 (phaser.py:1398-1408): sites whose REF is longer than one base or whose alleles are
// multi-base strings carry a0 == ALLELE_MULTI and are resolved against these side tables.
struct VariantsViewIndels : VariantsView {
  const int32_t* ref_len;          // [V] len(REF) (read_variant_map.py:130, 240)
  const u32* al_off;               // [2V+1] allele strings of site j: [al_off[2j], al_off[2j+1]) and [al_off[2j+1], al_off[2j+2])
  const u8* al_codes;              // 4-bit base code per character, 0xFF = a character no read base can equal
  static constexpr bool kIndels = true;
};
constexpr u8 ALLELE_MULTI = 0xFE;

enum { CLS_A0 = 0, CLS_A1 = 1, CLS_OTHER = 2, CLS_NONE = 3 };
enum { OP_M = 0, OP_I = 1, OP_D = 2, OP_N = 3, OP_S = 4, OP_H = 5, OP_P = 6, OP_EQ = 7, OP_X = 8 };
constexpr u8 BASE_N = 15, BASE_D = 13;

// t_misc layout: bits 0-1 class, bit 2 multi-base string, bits 4-7 first base code, bits 8-15 segment, bits 16-31 AS
PHZ_HD u32 pack_misc(int cls, int multi, int base, int seg, int as16) {
  return (u32)cls | ((u32)multi << 2) | ((u32)(base & 15) << 4) | ((u32)(seg > 255 ? 255 : seg) << 8) | ((u32)(as16 & 0xFFFF) << 16);
}
PHZ_HD int misc_cls(u32 m) { return m & 3; }
PHZ_HD int misc_as(u32 m) { return (int)(int16_t)(m >> 16); }

PHZ_HD bool isize_ok(int32_t tlen, double cutoff) {
  // read_variant_map.py:35,51: abs(TLEN) <= isize_cutoff unless the cutoff is 0
  if (cutoff == 0.0) return true;
  int64_t t = tlen; if (t < 0) t = -t;
  return (double)t <= cutoff;
}

template <class RV>
PHZ_HD u8 masked_base(const RV& rv, u64 base_off, int q, int baseq) {
  u64 i = base_off + (u64)q;
  u8 b = rv.seq[i >> 1];
  b = (i & 1) ? (b & 15) : (b >> 4);
  return ((int)rv.qual[i] < baseq) ? BASE_N : b;    // read_variant_map.py:179-184
}

// Variant-position providers: where the per-segment range search reads the sorted het-site positions.
struct GlobalVP {
  const int32_t* g;
  PHZ_HD int32_t at(int64_t j) const { return g[j]; }
  // lower bounds of lo_key / hi_key inside the contig's range [v0, v1)
  PHZ_HD void range(int32_t lo_key, int32_t hi_key, int64_t v0, int64_t v1, bool, int64_t& lo, int64_t& hi) const {
    lo = lower_bound_i32(g, v0, v1, lo_key);
    hi = lower_bound_i32(g, lo, v1, hi_key);
  }
};

// A slab of the position array staged in shared memory (global indices [wbase, wbase+wn)); searches
// that fall outside it go back to global memory.  32-bit local indices, branch-free halving, short
// linear scan for the upper end.  [hint_lo, hint_hi] (local, inclusive) brackets the lower bound of the
// POS of every record of the tile on the tile's first contig (records are coordinate sorted), so the
// first-segment search of such a record only looks there; hint_hi < 0 disables it.
struct WindowVP {
  const int32_t* g;
  const int32_t* s;
  int64_t wbase;
  int wn;
  int hint_lo, hint_hi;
  PHZ_HD int32_t at(int64_t j) const {
    int64_t k = j - wbase;
    return (k >= 0 && k < wn) ? s[k] : g[j];
  }
  PHZ_HD int lower(int l, int len, int32_t key) const {
    while (len > 0) {
      int half = len >> 1, mid = l + half;
      bool p = s[mid] < key;
      l = p ? mid + 1 : l;
      len = p ? len - half - 1 : half;
    }
    return l;
  }
  PHZ_HD void range(int32_t lo_key, int32_t hi_key, int64_t v0, int64_t v1, bool use_hint, int64_t& lo, int64_t& hi) const {
    int64_t a64 = v0 - wbase, b64 = v1 - wbase;
    int a = a64 > 0 ? (int)a64 : 0, b = b64 < wn ? (int)b64 : wn;
    if (a < b) {
      int l; bool ok;
      if (use_hint && hint_hi >= 0) {
        l = lower(hint_lo, hint_hi - hint_lo, lo_key);      // the true lower bound is inside the bracket
        ok = true;
      } else {
        l = lower(a, b - a, lo_key);
        ok = (l > a || wbase + a == v0) && (l < b || wbase + b == v1);
      }
      if (ok) {
        int h = l;
        while (h < b && s[h] < hi_key) ++h;
        lo = wbase + l;
        hi = (h < b || wbase + b == v1) ? wbase + h : lower_bound_i32(g, wbase + b, v1, hi_key);
        return;
      }
    }
    lo = lower_bound_i32(g, v0, v1, lo_key);
    hi = lower_bound_i32(g, lo, v1, hi_key);
  }
};

// pseudo_read[st] of the segment made of CIGAR ops [kseg, kend) and the insertion whose key equals st
// (read_variant_map.py:191-232).  base: 4-bit code, 16 = deletion placeholder, -1 = outside the segment.
template <class RV>
PHZ_HD void locate_in_segment(const RV& rv, u32 kseg, u32 kend, int32_t seg_start, int32_t q_start, u64 boff, int baseq,
                              int32_t st, int& base, int32_t& ins_q, int32_t& ins_n) {
  int32_t g = seg_start, q = q_start;
  base = -1; ins_q = -1; ins_n = 0;
  for (u32 x = kseg; x < kend; ++x) {
    u32 c = rv.cigar_at(x); int op = c & 15; int32_t n = (int32_t)(c >> 4);
    if (op == OP_M || op == OP_EQ || op == OP_X) {
      int32_t sp = g - seg_start;
      if (st >= sp && st < sp + n) base = masked_base(rv, boff, q + (st - sp), baseq);
      g += n; q += n;
    } else if (op == OP_D) {
      int32_t sp = g - seg_start;
      if (st >= sp && st < sp + n) base = 16;
      g += n;
    } else if (op == OP_I) {
      if (g - 1 == st) { ins_q = q; ins_n = n; }         // whole-read key, segment-relative lookup (Q3)
      q += n;
    } else if (op == OP_S) {
      q += n;
    }
  }
}

// identify_allele (read_variant_map.py:236-258) for a site with multi-base REF / alleles: the string is
// pseudo_read[st : st+ref_len] with the inserted bases spliced in after their key offsets and every 'D'
// removed; it is compared on the fly with the sample's two allele strings.  Returns the packed t_misc word.
template <class RV>
PHZ_HD u32 call_indel_site(const RV& rv, const VariantsViewIndels& vv, int64_t j, u32 kseg, u32 kend, int32_t seg_start,
                           int32_t q_start, int32_t seg_len, u64 boff, int baseq, int32_t st, int seg, int as16) {
  const int32_t L = vv.ref_len[j];
  if ((int64_t)st + L > (int64_t)seg_len) return pack_misc(CLS_NONE, 0, 0, seg, as16);     // read_end > len(pseudo_read)
  const u32 o0 = vv.al_off[2 * j], o1 = vv.al_off[2 * j + 1], o2 = vv.al_off[2 * j + 2];
  const u32 len0 = o1 - o0, len1 = o2 - o1;
  u32 n_chars = 0; int first = 0; bool eq0 = true, eq1 = true;
  auto push = [&](int b) {
    if (b == BASE_D) return;
    if (n_chars == 0) first = b;
    eq0 = eq0 && n_chars < len0 && vv.al_codes[o0 + n_chars] == (u8)b;
    eq1 = eq1 && n_chars < len1 && vv.al_codes[o1 + n_chars] == (u8)b;
    n_chars++;
  };
  for (int32_t x = st; x < st + L; ++x) {
    int base; int32_t ins_q, ins_n;
    locate_in_segment(rv, kseg, kend, seg_start, q_start, boff, baseq, x, base, ins_q, ins_n);
    if (base >= 0 && base != 16) push(base);
    for (int32_t z = 0; z < ins_n; ++z) push(masked_base(rv, boff, ins_q + z, baseq));
  }
  eq0 = eq0 && n_chars == len0; eq1 = eq1 && n_chars == len1;
  int cls;
  if (n_chars == 0) cls = CLS_NONE;
  else if (n_chars == 1 && first == BASE_N) cls = CLS_NONE;
  else if (eq0) cls = CLS_A0;
  else if (eq1) cls = CLS_A1;
  else cls = CLS_OTHER;
  return pack_misc(cls, n_chars > 1 ? 1 : 0, first, seg, as16);
}

// identify_allele (read_variant_map.py:236-258) for a single-base site: pseudo_read[st] of the segment made of CIGAR ops
// [kseg, kend) (a closing N ends it early) plus the insertion keyed st, minus every 'D'.  Keys of insertions are
// whole-read offsets looked up segment-relative (Q3).  Returns the packed t_misc word.
template <class RV>
PHZ_HD u32 call_snv_site(const RV& rv, u8 a0, u8 a1, u32 kseg, u32 kend, int32_t seg_start, int32_t q_start, u64 boff,
                         int baseq, int32_t st, int seg, int as16) {
  int32_t g = seg_start, q = q_start;
  int base = -1;            // 16: deletion placeholder
  int32_t ins_q = -1, ins_n = 0;
  for (u32 x = kseg; x < kend; ++x) {
    u32 c = rv.cigar_at(x); int op = c & 15; int32_t n = (int32_t)(c >> 4);
    if (op == OP_N) break;
    if (op == OP_M || op == OP_EQ || op == OP_X) {
      int32_t sp = g - seg_start;
      if (st >= sp && st < sp + n) base = masked_base(rv, boff, q + (st - sp), baseq);
      g += n; q += n;
    } else if (op == OP_D) {
      int32_t sp = g - seg_start;
      if (st >= sp && st < sp + n) base = 16;
      g += n;
    } else if (op == OP_I) {
      if (g - 1 == st) { ins_q = q; ins_n = n; }       // later insertion with the same key wins
      q += n;
    } else if (op == OP_S) {
      q += n;
    }
  }
  // string = [base] + inserted bases, minus every 'D'
  int n_chars = 0, first = 0;
  if (base != 16 && base != BASE_D && base >= 0) { first = base; n_chars = 1; }
  for (int32_t z = 0; z < ins_n; ++z) {
    int b = masked_base(rv, boff, ins_q + z, baseq);
    if (b != BASE_D) { if (n_chars == 0) first = b; n_chars++; }
  }
  int cls, multi = 0;
  if (n_chars == 0) cls = CLS_NONE;                     // "" -> nothing written
  else if (n_chars == 1) {
    if (first == BASE_N) cls = CLS_NONE;                // "N" -> nothing written
    else if (first == a0) cls = CLS_A0;
    else if (first == a1) cls = CLS_A1;
    else cls = CLS_OTHER;
  } else { cls = CLS_OTHER; multi = 1; }
  return pack_misc(cls, multi, first, seg, as16);
}

// Walks one record.
// MODE 0 (count): returns the number of candidate (segment, variant) pairs.
// MODE 1 (emit):  writes one tuple per candidate starting at out index `o` (class CLS_NONE when the
//                 reference would print nothing) and returns the number written.
// MODE 2 (k-th):  `o` is the ordinal of ONE candidate of this record; writes that tuple at out index 0.
template <int MODE, class RV, class VP, class VV>
PHZ_HD u32 map_record(const RV& rv, const VV& vv, const VP& vp, int64_t r, int contig, int baseq,
                      double isize_cutoff, u64 o, u32* t_rec, u32* t_var, u32* t_misc) {
  constexpr bool EMIT = MODE != 0;
  if (!isize_ok(rv.tlen_at(r), isize_cutoff)) return 0;
  const int64_t v0 = vv.contig_var_off[contig], v1 = vv.contig_var_off[contig + 1];
  if (v0 == v1) return 0;
  const int32_t rpos = rv.pos_at(r);
  const u32 c0 = rv.cig_lo(r), c1 = rv.cig_hi(r);
  u32 n_out = 0;
  int seg = 0;
  u32 k = c0;
  int32_t gp = 0;      // whole-read reference offset (genome_pos)
  int32_t qp = 0;      // query offset (read_pos)
  while (true) {
    // one segment: ops [kseg, kend) up to (not including) the closing N or the end
    const u32 kseg = k;
    const int32_t seg_start = gp, q_start = qp;
    int32_t g_end = gp, q_end = qp;
    for (; k < c1; ++k) {
      u32 c = rv.cigar_at(k); int op = c & 15; int32_t n = (int32_t)(c >> 4);
      if (op == OP_N) break;
      if (op == OP_M || op == OP_EQ || op == OP_X) { g_end += n; q_end += n; }
      else if (op == OP_D) g_end += n;
      else if (op == OP_I || op == OP_S) q_end += n;
    }
    const u32 kend = k;
    const int32_t seg_len = g_end - seg_start;           // len(pseudo_read): M/=/X and D
    if (seg_len > 0) {
      // candidates: het sites with 0 <= pos-(rpos+seg_start) < seg_len
      const int64_t lo_pos = (int64_t)rpos + seg_start, hi_pos = lo_pos + seg_len;
      if (lo_pos < 2147483647LL) {
        const int32_t hi32 = hi_pos > 2147483647LL ? 2147483647 : (int32_t)hi_pos;
        int64_t lo, hi;
        vp.range((int32_t)lo_pos, hi32, v0, v1, seg == 0, lo, hi);
        if (!EMIT) {
          n_out += (u32)(hi - lo);
        } else if (MODE == 2 && (u64)n_out + (u64)(hi - lo) <= o) {
          n_out += (u32)(hi - lo);                 // the wanted candidate is in a later segment
        } else if (hi > lo) {
          const u64 boff = rv.seq_off_at(r);
          const int as16 = rv.aln_at(r);
          if (MODE == 2) { lo += (int64_t)(o - n_out); hi = lo + 1; }
          for (int64_t j = lo; j < hi; ++j) {
            const int32_t st = (int32_t)((int64_t)vp.at(j) - lo_pos);       // offset in pseudo_read
            if constexpr (VV::kIndels) {
              if (vv.a0[j] == ALLELE_MULTI) {
                const u64 w = (MODE == 2) ? 0 : o + n_out;
                t_rec[w] = (u32)r;
                t_var[w] = (u32)j;
                t_misc[w] = call_indel_site(rv, vv, j, kseg, kend, seg_start, q_start, seg_len, boff, baseq, st, seg, as16);
                n_out++;
                continue;
              }
            }
            const u64 w = (MODE == 2) ? 0 : o + n_out;
            t_rec[w] = (u32)r;
            t_var[w] = (u32)j;
            t_misc[w] = call_snv_site(rv, vv.a0[j], vv.a1[j], kseg, kend, seg_start, q_start, boff, baseq, st, seg, as16);
            n_out++;
          }
          if (MODE == 2) return 1;
        }
      }
    }
    if (kend >= c1) break;
    // closing N: advance the reference, open the next segment
    gp = g_end + (int32_t)(rv.cigar_at(kend) >> 4);
    qp = q_end;
    k = kend + 1;
    seg++;
  }
  return n_out;
}

}  // namespace phz

// Native host VCF ingest and VCF output (no GPU involved) -- SURVEY.md section 8f rows N2 / N3.
//
// phz_vcf_open / phz_vcf_parse replace the text pipeline + parser of the reference for one sample:
//   `gunzip -c VCF | cut -f 1-9,<col> | grep -v '0|0\|1|1'`   phaser/phaser.py:205-225
//   the per-line het / PASS test                              phaser/phaser.py:396-434
//   the per-contig mapping table rows, indel exclusion        phaser/phaser.py:1355-1413
// phz_vcf_write replaces write_vcf (phaser/phaser.py:1661-1845): it is designed from the OUTPUT FORMAT -- every data
// line of the input keeps its first nine columns, gets the six phASER FORMAT keys appended once, and a sample field
// whose values come from per-variant / per-block arrays the caller hands over -- not from the reference's control flow.
// The decompressed input is kept in memory between the two calls, so the VCF is inflated and split into lines once.
// Lines are classified and formatted on n_threads host threads over fixed line ranges; results are merged in file order.
#include <string>
#include <vector>
#include <cstring>
#include <algorithm>
#include <map>

namespace phzvcf {

using phz::u64; using phz::u32; using phz::u8; using phz::PhzError;

struct Span { const char* p; size_t n; };

static inline bool eq(Span s, const char* lit) { size_t m = std::strlen(lit); return s.n == m && std::memcmp(s.p, lit, m) == 0; }
static inline bool contains(const char* p, size_t n, const char* lit) {
  size_t m = std::strlen(lit);
  if (n < m) return false;
  for (size_t i = 0; i + m <= n; ++i) if (p[i] == lit[0] && std::memcmp(p + i, lit, m) == 0) return true;
  return false;
}

// columns 0..8 and the sample column of one line; false when the line has fewer than nine columns or no sample column
struct Cut { Span c[10]; };
static bool cut_line(const char* p, const char* e, int sample_column, Cut& out) {
  int col = 0; const char* s = p;
  bool have_sample = false;
  for (const char* q = p;; ++q) {
    if (q == e || *q == '\t') {
      if (col < 9) out.c[col] = Span{s, (size_t)(q - s)};
      if (col == sample_column) { out.c[9] = Span{s, (size_t)(q - s)}; have_sample = true; }
      ++col; s = q + 1;
      if (q == e || (col > sample_column && col >= 9)) break;
    }
  }
  return col >= 9 && have_sample;
}

// k-th ':'-separated field of s (false when there are fewer)
static bool colon_field(Span s, int k, Span& out) {
  const char* p = s.p; const char* e = s.p + s.n; int i = 0; const char* st = p;
  for (const char* q = p;; ++q) {
    if (q == e || *q == ':') {
      if (i == k) { out = Span{st, (size_t)(q - st)}; return true; }
      ++i; st = q + 1;
      if (q == e) return false;
    }
  }
}
static int colon_count(Span s) { int n = 1; for (size_t i = 0; i < s.n; ++i) if (s.p[i] == ':') ++n; return n; }
// index of the field that equals `name` exactly, -1 when absent
static int colon_index(Span s, const char* name) {
  int n = colon_count(s);
  for (int k = 0; k < n; ++k) { Span f; if (colon_field(s, k, f) && eq(f, name)) return k; }
  return -1;
}

static const u8 ALLELE_NONE = 0xFF, VCF_ALLELE_MULTI = 0xFE;
static u8 allele_code(Span a) {
  static const char* alpha = "=ACMGRSVTWYHKDBN";
  if (a.n != 1) return ALLELE_NONE;
  const char* f = std::strchr(alpha, a.p[0]);
  return (f && a.p[0]) ? (u8)(f - alpha) : ALLELE_NONE;
}

// the reference's reading of a genotype string: its characters minus the first '|' and the first '/'
struct Geno { char ch[16]; int n; bool dot, unphased, usable, overflow; };
static Geno read_geno(Span g) {
  Geno o; o.n = 0; o.dot = false; o.unphased = false; o.usable = false; o.overflow = false;
  bool bar_gone = false, slash_gone = false;
  for (size_t i = 0; i < g.n; ++i) if (g.p[i] == '.') o.dot = true;
  if (o.dot) return o;
  for (size_t i = 0; i < g.n; ++i) {
    char ch = g.p[i];
    if (ch == '|' && !bar_gone) { bar_gone = true; continue; }
    if (o.n < 16) o.ch[o.n++] = ch; else o.overflow = true;
  }
  int m = 0;
  for (int i = 0; i < o.n; ++i) {                    // xgeno.remove("/") after the '|' was removed
    if (o.ch[i] == '/' && !slash_gone) { slash_gone = true; o.unphased = true; continue; }
    o.ch[m++] = o.ch[i];
  }
  o.n = m;
  for (int i = 1; i < o.n; ++i) if (o.ch[i] != o.ch[0]) o.usable = true;
  return o;
}

enum LineKind : u8 { L_HEADER = 0, L_GREPPED, L_OTHER_CHROM, L_NO_GT, L_UNUSABLE, L_FILTERED, L_KEPT };

struct Vcf {
  phzio::Bytes text;
  std::vector<u64> line_off;                 // n_lines + 1
  bool has_cr = false;
  // ---- result of the last parse
  int sample_column = -1;
  std::vector<u8> kind;                      // per line
  std::vector<int32_t> line_var;             // per line: het-table index, -1 otherwise
  std::vector<std::string> contigs, seen;
  std::string contigs_blob, seen_blob;
  std::vector<int64_t> contig_var_off;
  std::vector<int32_t> pos, ref_len; std::vector<u8> a0, a1;
  std::vector<int64_t> var_off; std::vector<int32_t> var_len;
  int64_t stats[4] = {0, 0, 0, 0};           // het_count, filter_count, indels_excluded, unphased_count
  // ---- result of the last write
  std::vector<char> out_text;
  std::vector<int32_t> rec_chrom; std::vector<int64_t> rec_beg, rec_end, rec_off;
  std::vector<std::string> rec_names; std::string rec_names_blob;
  size_t n_lines() const { return line_off.size() - 1; }
  const char* lp(size_t i) const { return (const char*)text.data() + line_off[i]; }
  const char* le(size_t i) const {          // end of line i without its '\n'
    u64 e = line_off[i + 1];
    if (e > line_off[i] && text[e - 1] == '\n') --e;
    return (const char*)text.data() + e;
  }
};

static void index_lines(Vcf& v, int n_threads) {
  const size_t n = v.text.size();
  const size_t chunk = 1 << 22;
  const size_t nch = (n + chunk - 1) / chunk;
  std::vector<std::vector<u64>> parts(nch ? nch : 1);
  std::atomic<int> cr(0);
  phzio::parallel_for(nch, n_threads, [&](size_t c) {
    size_t a = c * chunk, b = a + chunk < n ? a + chunk : n;
    auto& out = parts[c];
    for (size_t i = a; i < b; ++i) { u8 ch = v.text[i]; if (ch == '\n') out.push_back(i + 1); else if (ch == '\r') cr = 1; }
  });
  v.line_off.clear(); v.line_off.push_back(0);
  for (auto& p : parts) v.line_off.insert(v.line_off.end(), p.begin(), p.end());
  if (v.line_off.back() != n) v.line_off.push_back(n);       // last line without a newline
  if (n == 0) { v.line_off.assign(1, 0); }
  v.has_cr = cr != 0;
}

}  // namespace phzvcf

struct phz_vcf { phzvcf::Vcf v; };

extern "C" {

phz_vcf* phz_vcf_open(const char* path, int n_threads) {
  try {
    phzio::FileMap raw;
    if (!raw.open(path)) throw PhzError(std::string("cannot read ") + path);
    phz_vcf* h = new phz_vcf();
    bool gz = raw.size() >= 18 && raw[0] == 0x1f && raw[1] == 0x8b;
    bool bgzf = gz && (raw[3] & 4) && raw[12] == 'B' && raw[13] == 'C';
    if (bgzf) phzio::inflate_bgzf(raw, h->v.text, n_threads);
    else if (gz) phzio::inflate_gzip_stream(raw, h->v.text);
    else h->v.text.assign(raw.data(), raw.data() + raw.size());
    phzvcf::index_lines(h->v, n_threads);
    return h;
  } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}

void phz_vcf_close(phz_vcf* h) { delete h; }

int phz_vcf_text(phz_vcf* h, const char** text, int64_t* n_bytes, int64_t* n_lines, int* has_carriage_returns) {
  PHZ_TRY
  *text = (const char*)h->v.text.data(); *n_bytes = (int64_t)h->v.text.size(); *n_lines = (int64_t)h->v.n_lines();
  *has_carriage_returns = h->v.has_cr ? 1 : 0;
  PHZ_CATCH
}

// first line that contains "#CHR" (phaser.py:2326-2342 sample_column_map): byte offset and length without the newline
int phz_vcf_chrom_line(phz_vcf* h, int64_t* off, int64_t* len) {
  PHZ_TRY
  auto& v = h->v;
  *off = -1; *len = 0;
  for (size_t i = 0; i < v.n_lines(); ++i) {
    const char* p = v.lp(i); const char* e = v.le(i);
    if (phzvcf::contains(p, (size_t)(e - p), "#CHR")) { *off = (int64_t)v.line_off[i]; *len = (int64_t)(e - p); break; }
    if (p < e && *p != '#') break;            // the column header precedes the data
  }
  PHZ_CATCH
}

int phz_vcf_parse(phz_vcf* h, int sample_column, int pass_only, const char* chrom_of_interest, int include_indels,
                  int n_threads, phz_vcf_table* out) {
  PHZ_TRY
  using namespace phzvcf;
  auto& v = h->v;
  const size_t NL = v.n_lines();
  const std::string coi = chrom_of_interest ? chrom_of_interest : "";
  if (sample_column < 9) throw PhzError("phz_vcf_parse: the sample column must be 9 or larger");
  v.sample_column = sample_column;
  v.kind.assign(NL, L_HEADER); v.line_var.assign(NL, -1);
  // ---- pass 1 (parallel over line ranges): classify every line; per range the chromosome names in order of appearance
  const size_t RANGE = 1 << 15;
  const size_t NR = (NL + RANGE - 1) / RANGE;
  struct RangeOut { std::vector<std::string> seen, pooled; std::vector<u32> kept_line; std::vector<int> kept_chrom; int64_t filtered = 0, unphased = 0; std::string err; };
  std::vector<RangeOut> ro(NR ? NR : 1);
  phzio::parallel_for(NR, n_threads, [&](size_t r) {
    RangeOut& o = ro[r];
    auto local_id = [](std::vector<std::string>& names, Span s) {
      for (size_t i = names.size(); i-- > 0;) if (names[i].size() == s.n && std::memcmp(names[i].data(), s.p, s.n) == 0) return (int)i;
      names.emplace_back(s.p, s.n); return (int)names.size() - 1;
    };
    const size_t a = r * RANGE, b = a + RANGE < NL ? a + RANGE : NL;
    for (size_t i = a; i < b; ++i) {
      const char* p = v.lp(i); const char* e = v.le(i);
      if (p < e && *p == '#') { v.kind[i] = L_HEADER; continue; }
      Cut c;
      if (!cut_line(p, e, sample_column, c)) { if (o.err.empty()) o.err = "malformed VCF line " + std::to_string(i + 1) + " (fewer columns than the sample's)"; return; }
      // grep -v '0|0\|1|1' on the cut line: columns 0-8 are one stretch of the line, the sample column another
      const char* c8e = c.c[8].p + c.c[8].n;
      if (contains(p, (size_t)(c8e - p), "0|0") || contains(p, (size_t)(c8e - p), "1|1") ||
          contains(c.c[9].p, c.c[9].n, "0|0") || contains(c.c[9].p, c.c[9].n, "1|1")) { v.kind[i] = L_GREPPED; continue; }
      local_id(o.seen, c.c[0]);
      if (!coi.empty() && !(c.c[0].n == coi.size() && std::memcmp(c.c[0].p, coi.data(), coi.size()) == 0)) { v.kind[i] = L_OTHER_CHROM; continue; }
      const int cid = local_id(o.pooled, c.c[0]);
      const int gi = colon_index(c.c[8], "GT");
      if (gi < 0) { v.kind[i] = L_NO_GT; continue; }
      Span gs;
      if (!colon_field(c.c[9], gi, gs)) { if (o.err.empty()) o.err = "malformed VCF line " + std::to_string(i + 1) + " (no GT value in the sample column)"; return; }
      const Geno g = read_geno(gs);
      if (!g.usable) { v.kind[i] = L_UNUSABLE; continue; }
      bool pass = false;
      { const char* fp = c.c[6].p; const char* fe = fp + c.c[6].n; const char* st = fp;
        for (const char* q = fp;; ++q) if (q == fe || *q == ';') { if (q - st == 4 && std::memcmp(st, "PASS", 4) == 0) pass = true; st = q + 1; if (q == fe) break; } }
      if (pass_only == 0 || pass) {
        v.kind[i] = L_KEPT; o.kept_line.push_back((u32)i); o.kept_chrom.push_back(cid);
        if (g.unphased) o.unphased++;
      } else { v.kind[i] = L_FILTERED; o.filtered++; }
    }
  });
  for (auto& o : ro) if (!o.err.empty()) throw PhzError(o.err);
  // ---- merge: contig order = first appearance among the pooled lines; kept rows grouped by contig, file order inside
  v.contigs.clear(); v.seen.clear();
  auto global_id = [](std::vector<std::string>& names, const std::string& s) {
    for (size_t i = 0; i < names.size(); ++i) if (names[i] == s) return (int)i;
    names.push_back(s); return (int)names.size() - 1;
  };
  std::vector<std::vector<u32>> rows;
  v.stats[0] = v.stats[1] = v.stats[2] = v.stats[3] = 0;
  for (auto& o : ro) {
    for (auto& s : o.seen) global_id(v.seen, s);
    std::vector<int> map(o.pooled.size());
    for (size_t k = 0; k < o.pooled.size(); ++k) { map[k] = global_id(v.contigs, o.pooled[k]); if ((size_t)map[k] >= rows.size()) rows.resize(map[k] + 1); }
    for (size_t k = 0; k < o.kept_line.size(); ++k) rows[map[o.kept_chrom[k]]].push_back(o.kept_line[k]);
    v.stats[1] += o.filtered; v.stats[3] += o.unphased;
  }
  rows.resize(v.contigs.size());
  // ---- pass 2 (parallel over kept rows): alleles, indel rule, codes
  std::vector<u32> flat; std::vector<int> flat_contig;
  for (size_t c = 0; c < rows.size(); ++c) for (u32 ln : rows[c]) { flat.push_back(ln); flat_contig.push_back((int)c); }
  const size_t NK = flat.size();
  std::vector<u8> keep(NK, 0), ka0(NK), ka1(NK); std::vector<int32_t> kpos(NK), krl(NK);
  std::string fatal; std::mutex fm;
  phzio::parallel_for((NK + 4095) / 4096, n_threads, [&](size_t blk) {
    for (size_t k = blk * 4096; k < NK && k < (blk + 1) * 4096; ++k) {
      const size_t i = flat[k];
      const char* p = v.lp(i); const char* e = v.le(i);
      Cut c; cut_line(p, e, sample_column, c);
      Span gs; colon_field(c.c[9], colon_index(c.c[8], "GT"), gs);
      const Geno g = read_geno(gs);
      // alleles: REF + ALT.split(",")
      Span al[64]; int na = 0; bool all_single = c.c[3].n == 1;
      al[na++] = c.c[3];
      { const char* st = c.c[4].p; const char* fe = st + c.c[4].n;
        for (const char* q = st;; ++q) if (q == fe || *q == ',') { if (na < 64) al[na++] = Span{st, (size_t)(q - st)}; if ((size_t)(q - st) != 1) all_single = false; st = q + 1; if (q == fe) break; } }
      if (!(all_single || include_indels == 1)) { keep[k] = 2; continue; }          // indel excluded (phaser.py:1398-1408)
      Span ind[3]; int ni = 0;
      for (int a = 0; a < na && a < 10; ++a) {            // str(a) in xgeno: single characters only
        bool in = false;
        for (int x = 0; x < g.n; ++x) if (g.ch[x] == (char)('0' + a)) in = true;
        if (in) { if (ni < 3) ind[ni] = al[a]; ++ni; }
      }
      int64_t pv = 0; bool okp = c.c[1].n > 0;
      for (size_t x = 0; x < c.c[1].n; ++x) { char ch = c.c[1].p[x]; if (ch < '0' || ch > '9') okp = false; else pv = pv * 10 + (ch - '0'); if (pv > 2147483647LL) okp = false; }
      if (ni != 2 || g.n != 2 || g.overflow || !okp) {
        std::lock_guard<std::mutex> lk(fm);
        if (fatal.empty()) fatal = !okp ? "VCF line " + std::to_string(i + 1) + ": POS is not a number"
                                        : "Variant " + std::string(c.c[0].p, c.c[0].n) + ":" + std::string(c.c[1].p, c.c[1].n) +
                                          ": only diploid genotypes with two distinct alleles are supported.";
        continue;
      }
      const bool indel_site = c.c[3].n != 1 || ind[0].n != 1 || ind[1].n != 1;
      if (!all_single && indel_site) { ka0[k] = VCF_ALLELE_MULTI; ka1[k] = VCF_ALLELE_MULTI; }
      else { ka0[k] = allele_code(ind[0]); ka1[k] = allele_code(ind[1]); }
      kpos[k] = (int32_t)pv; krl[k] = (int32_t)c.c[3].n; keep[k] = 1;
    }
  });
  if (!fatal.empty()) throw PhzError(fatal);
  v.pos.clear(); v.ref_len.clear(); v.a0.clear(); v.a1.clear(); v.var_off.clear(); v.var_len.clear();
  v.contig_var_off.assign(1, 0);
  size_t k = 0;
  for (size_t c = 0; c < rows.size(); ++c) {
    for (size_t j = 0; j < rows[c].size(); ++j, ++k) {
      if (keep[k] == 2) { v.stats[2]++; continue; }
      const size_t i = flat[k];
      v.line_var[i] = (int32_t)v.pos.size();
      v.pos.push_back(kpos[k]); v.ref_len.push_back(krl[k]); v.a0.push_back(ka0[k]); v.a1.push_back(ka1[k]);
      v.var_off.push_back((int64_t)v.line_off[i]); v.var_len.push_back((int32_t)(v.le(i) - v.lp(i)));
    }
    v.contig_var_off.push_back((int64_t)v.pos.size());
  }
  v.stats[0] = (int64_t)v.pos.size();
  v.contigs_blob.clear(); for (auto& s : v.contigs) { v.contigs_blob += s; v.contigs_blob.push_back('\0'); }
  v.seen_blob.clear(); for (auto& s : v.seen) { v.seen_blob += s; v.seen_blob.push_back('\0'); }
  out->n_variants = (int64_t)v.pos.size(); out->n_contigs = (int)v.contigs.size();
  out->contig_var_off = v.contig_var_off.data(); out->pos = v.pos.data(); out->a0 = v.a0.data(); out->a1 = v.a1.data();
  out->ref_len = v.ref_len.data(); out->var_line_off = v.var_off.data(); out->var_line_len = v.var_len.data();
  out->contig_names = v.contigs_blob.data(); out->n_seen = (int)v.seen.size(); out->seen_names = v.seen_blob.data();
  for (int s = 0; s < 4; ++s) out->stats[s] = v.stats[s];
  PHZ_CATCH
}


// ---------------------------------------------------------------------------------------------- output VCF
// What every output line is made of (phaser.py:1661-1845, restated from the format):
//   header      the line cut to columns 1-9 + the sample's; before #CHROM the definitions of PG PB PI PM PW PC (and PS with
//               --gw_phase_vcf 2) that the input does not define already
//   data line   columns 1-8 as they are; FORMAT with the keys PG PB PI PW PC PM appended where absent; the sample field padded
//               to the FORMAT's length and, key by key:
//               site in a phased block    PG = the block's two alleles as VCF allele indices "a|b", PB = the block's rsids,
//                                         PI = block index, PM = block maf, PW = genome-wide phase, PC = its confidence;
//                                         GT rewritten per --gw_phase_vcf (and PS = block index when the anchoring is weak)
//               any other site            PG = the genotype's characters sorted and joined by '/', PW = the GT text, rest '.'
//   lines of other chromosomes are dropped under --chr; lines without GT pass through cut.
int phz_vcf_write(phz_vcf* h, const phz_vcf_annot* A, int n_threads, const char** text, int64_t* n_bytes, int64_t* counts) {
  PHZ_TRY
  using namespace phzvcf;
  auto& v = h->v;
  if (v.sample_column < 9 || v.kind.size() != v.n_lines()) throw PhzError("phz_vcf_write: phz_vcf_parse must run first");
  const int sc = v.sample_column;
  const size_t NL = v.n_lines();
  const std::string coi = A->chrom_of_interest ? A->chrom_of_interest : "";
  const int64_t NB = A->n_blocks;
  if (A->n_variants != (int64_t)v.pos.size()) throw PhzError("phz_vcf_write: annotation arrays do not match the parsed table");
  // per-block strings handed over back to back
  std::vector<const char*> stat_s(NB), maf_s(NB);
  { const char* p = A->blk_stat; for (int64_t b = 0; b < NB; ++b) { stat_s[b] = p; p += std::strlen(p) + 1; }
    p = A->blk_maf; for (int64_t b = 0; b < NB; ++b) { maf_s[b] = p; p += std::strlen(p) + 1; } }
  // PB of every block: rsids of its members, ':' -> '_', comma-joined
  const std::string sep = A->id_separator ? A->id_separator : "_", prefix = A->chr_prefix ? A->chr_prefix : "";
  std::vector<std::string> pb(NB);
  phzio::parallel_for((size_t)((NB + 1023) / 1024), n_threads, [&](size_t blk) {
    for (int64_t b = (int64_t)blk * 1024; b < NB && b < (int64_t)(blk + 1) * 1024; ++b) {
      std::string& o = pb[b];
      for (int64_t k = 0; k < A->blk_len[b]; ++k) {
        const int32_t m = A->blk_members[A->blk_first[b] + k];
        const char* p = (const char*)v.text.data() + v.var_off[m]; const char* e = p + v.var_len[m];
        Cut c; cut_line(p, e, sc, c);
        if (k) o.push_back(',');
        if (c.c[2].n == 0 || (c.c[2].n == 1 && c.c[2].p[0] == '.')) {
          // no rsid: the site's own id (phaser.py:1448-1449) = contig SEP pos SEP alleles joined by SEP
          std::string id = prefix; id.append(c.c[0].p, c.c[0].n); id += sep; id.append(c.c[1].p, c.c[1].n); id += sep; id.append(c.c[3].p, c.c[3].n);
          const char* st = c.c[4].p; const char* fe = st + c.c[4].n;
          for (const char* q = st;; ++q) if (q == fe || *q == ',') { id += sep; id.append(st, q - st); st = q + 1; if (q == fe) break; }
          for (char ch : id) o.push_back(ch == ':' ? '_' : ch);
        } else
          for (size_t x = 0; x < c.c[2].n; ++x) o.push_back(c.c[2].p[x] == ':' ? '_' : c.c[2].p[x]);
      }
    }
  });
  // a table id names the LAST line that spells it (the id dictionary of phaser.py:1690 keeps the last): duplicates of one
  // (position, REF, ALT) spelling among the table's sites are redirected
  std::vector<int32_t> redirect(v.pos.size());
  for (size_t i = 0; i < redirect.size(); ++i) redirect[i] = (int32_t)i;
  for (size_t c = 0; c + 1 < v.contig_var_off.size(); ++c)
    for (int64_t a = v.contig_var_off[c]; a < v.contig_var_off[c + 1];) {
      int64_t b = a + 1;
      while (b < v.contig_var_off[c + 1] && v.pos[b] == v.pos[a]) ++b;
      if (b - a > 1)
        for (int64_t i = a; i < b; ++i)
          for (int64_t j = b - 1; j > i; --j) {
            Cut ci, cj;
            cut_line((const char*)v.text.data() + v.var_off[i], (const char*)v.text.data() + v.var_off[i] + v.var_len[i], sc, ci);
            cut_line((const char*)v.text.data() + v.var_off[j], (const char*)v.text.data() + v.var_off[j] + v.var_len[j], sc, cj);
            if (ci.c[3].n == cj.c[3].n && ci.c[4].n == cj.c[4].n && !std::memcmp(ci.c[3].p, cj.c[3].p, ci.c[3].n) &&
                !std::memcmp(ci.c[4].p, cj.c[4].p, ci.c[4].n)) { redirect[i] = (int32_t)j; break; }
          }
      a = b;
    }
  // ---- header lines in order (they carry state: the FORMAT definitions seen so far)
  std::vector<std::string> head_out(NL);
  std::string fmt_text;
  auto cut_text = [&](size_t i) {
    std::string o; const char* p = v.lp(i); const char* e = v.le(i);
    int col = 0; const char* st = p; bool first = true;
    for (const char* q = p;; ++q)
      if (q == e || *q == '\t') {
        if (col < 9 || col == sc) { if (!first) o.push_back('\t'); o.append(st, q - st); first = false; }
        ++col; st = q + 1;
        if (q == e) break;
      }
    o.push_back('\n');
    return o;
  };
  for (size_t i = 0; i < NL; ++i) {
    if (v.kind[i] != L_HEADER) {
      if (contains(v.lp(i), (size_t)(v.le(i) - v.lp(i)), "##FORMAT")) throw PhzError("unsupported: a data line mentions ##FORMAT");
      continue;
    }
    std::string line = cut_text(i);
    if (contains(line.data(), line.size(), "##FORMAT")) { fmt_text += line; head_out[i] = line; }
    else if (line.compare(0, 6, "#CHROM") == 0) {
      static const char* defs[6][2] = {{"PG", "phASER Local Genotype"}, {"PB", "phASER Local Block"},
                                       {"PI", "phASER Local Block Index (unique for each block)"},
                                       {"PM", "phASER Local Block Maximum Variant MAF"}, {"PW", "phASER Genome Wide Genotype"},
                                       {"PC", "phASER Genome Wide Confidence"}};
      std::string o;
      for (auto& d : defs) {
        std::string key = std::string("##FORMAT=<ID=") + d[0] + ",";
        if (!contains(fmt_text.data(), fmt_text.size(), key.c_str()))
          o += std::string("##FORMAT=<ID=") + d[0] + ",Number=1,Type=String,Description=\"" + d[1] + "\">\n";
      }
      if (A->gw_phase_vcf == 2 && !contains(fmt_text.data(), fmt_text.size(), "##FORMAT=<ID=PS,"))
        o += "##FORMAT=<ID=PS,Number=1,Type=String,Description=\"Phase Set\">\n";
      Cut c;
      if (!cut_line(v.lp(i), v.le(i), sc, c)) throw PhzError("unsupported: #CHROM line without the sample column");
      head_out[i] = o + line;
    } else head_out[i] = line;
  }
  // ---- data lines, in parallel over line ranges
  const size_t RANGE = 1 << 14;
  const size_t NR = (NL + RANGE - 1) / RANGE;
  struct RangeOut { std::string text; std::vector<int32_t> chrom; std::vector<int64_t> beg, end, off; std::vector<std::string> names;
                    int64_t corrections = 0, unphased_phased = 0; std::string err; };
  std::vector<RangeOut> ro(NR ? NR : 1);
  static const char* TAGS[6] = {"PG", "PB", "PI", "PW", "PC", "PM"};
  phzio::parallel_for(NR, n_threads, [&](size_t r) {
    RangeOut& o = ro[r];
    std::vector<std::string> sf;           // the sample's fields
    std::vector<Span> ff;                  // FORMAT fields (names)
    const size_t a = r * RANGE, b = a + RANGE < NL ? a + RANGE : NL;
    for (size_t i = a; i < b; ++i) {
      if (v.kind[i] == L_HEADER) { o.text += head_out[i]; continue; }
      const char* p = v.lp(i); const char* e = v.le(i);
      Cut c;
      if (!cut_line(p, e, sc, c)) { o.err = "unsupported: short data line"; return; }
      if (!coi.empty() && !(c.c[0].n == coi.size() && !std::memcmp(c.c[0].p, coi.data(), coi.size()))) continue;
      int64_t pos = 0; bool okp = c.c[1].n > 0;
      for (size_t x = 0; x < c.c[1].n; ++x) { char ch = c.c[1].p[x]; if (ch < '0' || ch > '9') okp = false; else pos = pos * 10 + (ch - '0'); }
      if (!okp) { o.err = "unsupported: POS is not a number"; return; }
      const size_t line_start = o.text.size();
      o.text.append(p, c.c[8].p - p);                     // columns 1-8 and the tab before FORMAT
      const Span fmt = c.c[8];
      if (contains(fmt.p, fmt.n, "GT")) {
        const int n_fields = colon_count(fmt);
        const int gt_index = colon_index(fmt, "GT");
        if (gt_index < 0) { o.err = "unsupported: GT inside another FORMAT key"; return; }
        ff.clear();
        for (int k = 0; k < n_fields; ++k) { Span f; colon_field(fmt, k, f); ff.push_back(f); }
        int tag_idx[6]; std::string fmt_out(fmt.p, fmt.n);
        for (int t = 0; t < 6; ++t) {
          int at = -1;
          for (size_t k = 0; k < ff.size(); ++k) if (eq(ff[k], TAGS[t])) { at = (int)k; break; }
          if (at < 0) { at = (int)ff.size(); ff.push_back(Span{TAGS[t], 2}); fmt_out.push_back(':'); fmt_out += TAGS[t]; }
          tag_idx[t] = at;
        }
        // the sample's fields, padded to the FORMAT's original length
        sf.clear();
        { const char* st = c.c[9].p; const char* fe = st + c.c[9].n;
          for (const char* q = st;; ++q) if (q == fe || *q == ':') { sf.emplace_back(st, q - st); st = q + 1; if (q == fe) break; } }
        if ((int)sf.size() <= gt_index) { o.err = "unsupported: sample column without a GT value"; return; }
        const std::string gt_orig = sf[gt_index];
        while ((int)sf.size() < n_fields) sf.emplace_back();
        // genotype characters: minus the first '|' and the first '/'
        std::string geno;
        { bool bar = false; for (char ch : gt_orig) { if (ch == '|' && !bar) { bar = true; continue; } geno.push_back(ch); }
          size_t sl = geno.find('/'); if (sl != std::string::npos) geno.erase(sl, 1); }
        int32_t tv = A->ids_match ? v.line_var[i] : -1;
        if (tv >= 0) tv = redirect[tv];
        const int64_t blk = tv >= 0 ? A->v_block[tv] : -1;
        if (blk >= 0) {
          // alleles of the record and the two the sample carries (allele-index order)
          Span al[64]; int na = 0;
          al[na++] = c.c[3];
          { const char* st = c.c[4].p; const char* fe = st + c.c[4].n;
            for (const char* q = st;; ++q) if (q == fe || *q == ',') { if (na < 64) al[na++] = Span{st, (size_t)(q - st)}; st = q + 1; if (q == fe) break; } }
          // GT of the table's own line (the redirect target may be another line with the same spelling)
          Span mine[2]; int nm = 0;
          { const char* tp = (const char*)v.text.data() + v.var_off[tv]; Cut tc; cut_line(tp, tp + v.var_len[tv], sc, tc);
            Span tg; colon_field(tc.c[9], colon_index(tc.c[8], "GT"), tg);
            Span tal[64]; int tna = 0; tal[tna++] = tc.c[3];
            { const char* st = tc.c[4].p; const char* fe = st + tc.c[4].n;
              for (const char* q = st;; ++q) if (q == fe || *q == ',') { if (tna < 64) tal[tna++] = Span{st, (size_t)(q - st)}; st = q + 1; if (q == fe) break; } }
            for (int x = 0; x < tna && x < 10 && nm < 2; ++x) {
              bool in = false;
              for (size_t y = 0; y < tg.n; ++y) if (tg.p[y] == (char)('0' + x)) in = true;
              if (in) mine[nm++] = tal[x];
            } }
          if (nm != 2) { o.err = "unsupported: table site without two alleles"; return; }
          const int hap = A->v_hap[tv] & 1;
          std::string alleles_out[2], gw_out[2];
          for (int k = 0; k < 2; ++k) {
            const int ai = k == 0 ? hap : 1 - hap;
            int vidx = -1;
            for (int x = 0; x < na; ++x) if (al[x].n == mine[ai].n && !std::memcmp(al[x].p, mine[ai].p, al[x].n)) { vidx = x; break; }
            if (vidx < 0) { o.err = "unsupported: allele of the table site not on this line"; return; }
            const int g = A->v_gw[2 * (int64_t)tv + ai];
            if (g == 0 || g == 1) gw_out[g] = std::to_string(vidx);
            alleles_out[k] = std::to_string(vidx);
          }
          const std::string new_phase = gw_out[0] + "|" + gw_out[1];
          const bool conf = A->blk_confident[blk] != 0;
          std::string& gt = sf[gt_index];
          if (conf) {
            if (gt.find('|') != std::string::npos && gt != new_phase) o.corrections++;
            if (gt.find('/') != std::string::npos && gt != "./." && gt != new_phase) o.unphased_phased++;
            if (A->gw_phase_vcf == 1 || A->gw_phase_vcf == 2) gt = new_phase;
          }
          if (A->gw_phase_vcf == 2 && !conf) gt = alleles_out[0] + "|" + alleles_out[1];
          while (sf.size() < ff.size()) sf.emplace_back();
          sf[tag_idx[0]] = alleles_out[0] + "|" + alleles_out[1];
          sf[tag_idx[1]] = pb[blk];
          sf[tag_idx[2]] = std::to_string(A->blk_index[blk]);
          sf[tag_idx[5]] = maf_s[blk];
          sf[tag_idx[3]] = new_phase;
          sf[tag_idx[4]] = stat_s[blk];
          if (A->gw_phase_vcf == 2 && !conf) {
            int ps = -1;
            for (size_t k = 0; k < ff.size(); ++k) if (eq(ff[k], "PS")) { ps = (int)k; break; }
            if (ps < 0) { fmt_out += ":PS"; ps = (int)ff.size(); sf.emplace_back(); }
            if ((size_t)ps >= sf.size()) { o.err = "unsupported: PS beyond the sample's fields"; return; }
            sf[ps] = std::to_string(A->blk_index[blk]);
          }
        } else {
          while (sf.size() < ff.size()) sf.emplace_back();
          std::string sorted = geno; std::sort(sorted.begin(), sorted.end());
          std::string pg;
          for (size_t x = 0; x < sorted.size(); ++x) { if (x) pg.push_back('/'); pg.push_back(sorted[x]); }
          sf[tag_idx[0]] = pg; sf[tag_idx[1]] = "."; sf[tag_idx[2]] = "."; sf[tag_idx[5]] = ".";
          sf[tag_idx[3]] = gt_orig; sf[tag_idx[4]] = ".";
        }
        o.text += fmt_out; o.text.push_back('\t');
        for (size_t k = 0; k < sf.size(); ++k) { if (k) o.text.push_back(':'); o.text += sf[k]; }
      } else {
        o.text.append(fmt.p, fmt.n); o.text.push_back('\t'); o.text.append(c.c[9].p, c.c[9].n);
      }
      o.text.push_back('\n');
      // reference span of the record (for the index): [POS-1, POS-1+len(REF)), stretched by INFO END=
      int64_t end = pos - 1 + (int64_t)c.c[3].n;
      if (contains(c.c[7].p, c.c[7].n, "END=")) {
        const char* st = c.c[7].p; const char* fe = st + c.c[7].n;
        for (const char* q = st;; ++q) if (q == fe || *q == ';') {
          if (q - st > 4 && !std::memcmp(st, "END=", 4)) {
            int64_t val = 0; bool ok = true; bool neg = false; const char* d = st + 4;
            if (d < q && (*d == '-' || *d == '+')) { neg = *d == '-'; ++d; }
            if (d == q) ok = false;
            for (; d < q; ++d) { if (*d < '0' || *d > '9') { ok = false; break; } val = val * 10 + (*d - '0'); }
            if (ok) { if (neg) val = -val; if (val > end) end = val; }
          }
          st = q + 1; if (q == fe) break;
        }
      }
      int cid = -1;
      for (size_t k = o.names.size(); k-- > 0;) if (o.names[k].size() == c.c[0].n && !std::memcmp(o.names[k].data(), c.c[0].p, c.c[0].n)) { cid = (int)k; break; }
      if (cid < 0) { o.names.emplace_back(c.c[0].p, c.c[0].n); cid = (int)o.names.size() - 1; }
      o.chrom.push_back(cid); o.beg.push_back(pos - 1); o.end.push_back(end); o.off.push_back((int64_t)line_start);
    }
  });
  for (auto& o : ro) if (!o.err.empty()) throw PhzError(o.err);
  size_t total = 0, nrec = 0;
  for (auto& o : ro) { total += o.text.size(); nrec += o.chrom.size(); }
  v.out_text.resize(total);
  v.rec_chrom.clear(); v.rec_beg.clear(); v.rec_end.clear(); v.rec_off.clear(); v.rec_names.clear();
  v.rec_chrom.reserve(nrec); v.rec_beg.reserve(nrec); v.rec_end.reserve(nrec); v.rec_off.reserve(nrec);
  size_t at = 0; counts[0] = 0; counts[1] = 0;
  for (auto& o : ro) {
    std::memcpy(v.out_text.data() + at, o.text.data(), o.text.size());
    std::vector<int> map(o.names.size());
    for (size_t k = 0; k < o.names.size(); ++k) {
      int g = -1;
      for (size_t x = 0; x < v.rec_names.size(); ++x) if (v.rec_names[x] == o.names[k]) { g = (int)x; break; }
      if (g < 0) { v.rec_names.push_back(o.names[k]); g = (int)v.rec_names.size() - 1; }
      map[k] = g;
    }
    for (size_t k = 0; k < o.chrom.size(); ++k) {
      v.rec_chrom.push_back(map[o.chrom[k]]); v.rec_beg.push_back(o.beg[k]); v.rec_end.push_back(o.end[k]);
      v.rec_off.push_back((int64_t)at + o.off[k]);
    }
    at += o.text.size();
    counts[0] += o.unphased_phased; counts[1] += o.corrections;
  }
  v.rec_names_blob.clear(); for (auto& s : v.rec_names) { v.rec_names_blob += s; v.rec_names_blob.push_back('\0'); }
  *text = v.out_text.data(); *n_bytes = (int64_t)total;
  PHZ_CATCH
}

// the data lines phz_vcf_write produced, in file order: chromosome (index into the NUL-separated names), reference span
int phz_vcf_records(phz_vcf* h, int64_t* n, const int32_t** chrom, const int64_t** beg, const int64_t** end,
                    const int64_t** text_off, const char** names, int32_t* n_names) {
  PHZ_TRY
  auto& v = h->v;
  *n = (int64_t)v.rec_chrom.size(); *chrom = v.rec_chrom.data(); *beg = v.rec_beg.data(); *end = v.rec_end.data();
  *text_off = v.rec_off.data(); *names = v.rec_names_blob.data(); *n_names = (int32_t)v.rec_names.size();
  PHZ_CATCH
}


// ---------------------------------------------------------------------------------------------- bgzip + tabix of the output
// What `bgzip -f` and `tabix -f -p vcf [--csi]` leave behind (phaser.py:1847-1853): the text of the last phz_vcf_write cut
// into 0xff00-byte BGZF blocks (deflated in parallel) and the index of its data lines -- UCSC binning with min_shift 14 and
// depth 5, one linear-index entry per 16 kb window, the pseudo-bin 37450 per reference, exactly as htslib writes them.
}  // extern "C"
namespace phzvcf {
static void bgzf_block(const u8* raw, size_t len, std::vector<u8>& o, int level) {
  o.resize(len + len / 8 + 1024);
  z_stream zs; std::memset(&zs, 0, sizeof(zs));
  if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) throw PhzError("zlib deflate init failed");
  zs.next_in = (Bytef*)raw; zs.avail_in = (uInt)len; zs.next_out = o.data() + 18; zs.avail_out = (uInt)(o.size() - 18 - 8);
  if (deflate(&zs, Z_FINISH) != Z_STREAM_END) { deflateEnd(&zs); throw PhzError("BGZF deflate failed"); }
  const size_t clen = zs.total_out; deflateEnd(&zs);
  const u8 hdr[18] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0, 0};
  std::memcpy(o.data(), hdr, 18);
  const size_t bsize = 18 + clen + 8 - 1;
  o[16] = (u8)bsize; o[17] = (u8)(bsize >> 8);
  const u32 crc = (u32)crc32(crc32(0L, Z_NULL, 0), raw, (uInt)len);
  u8* t = o.data() + 18 + clen;
  t[0] = (u8)crc; t[1] = (u8)(crc >> 8); t[2] = (u8)(crc >> 16); t[3] = (u8)(crc >> 24);
  t[4] = (u8)len; t[5] = (u8)(len >> 8); t[6] = (u8)(len >> 16); t[7] = (u8)(len >> 24);
  o.resize(18 + clen + 8);
}
static const u8 BGZF_EOF[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
static void write_bgzf_file(const char* path, const u8* data, size_t n, int n_threads, std::vector<size_t>* sizes) {
  const size_t BS = 0xff00, nblk = (n + BS - 1) / BS;
  std::vector<std::vector<u8>> blocks(nblk);
  phzio::parallel_for(nblk, n_threads, [&](size_t b) { bgzf_block(data + b * BS, std::min(BS, n - b * BS), blocks[b], 6); });
  FILE* f = std::fopen(path, "wb");
  if (!f) throw PhzError(std::string("cannot write ") + path);
  for (auto& b : blocks) std::fwrite(b.data(), 1, b.size(), f);
  std::fwrite(BGZF_EOF, 1, 28, f);
  std::fclose(f);
  if (sizes) { sizes->clear(); for (auto& b : blocks) sizes->push_back(b.size()); }
}
struct RefIndex { std::map<u32, std::vector<std::pair<u64, u64>>> bins; std::vector<u64> lin; u64 first = 0, last = 0, n = 0; int64_t last_bin = -1; };
static u32 reg2bin14(int64_t beg, int64_t end) {
  --end;
  int s = 14; u32 t = ((1u << 15) - 1) / 7;
  for (int level = 5; level > 0; --level) { if (beg >> s == end >> s) return t + (u32)(beg >> s); s += 3; t -= 1u << (3 * (level - 1)); }
  return 0;
}
static u64 bin_start_window(u32 b) {
  u32 t = 0;
  for (int level = 0; level <= 5; ++level) { u32 n = 1u << (3 * level); if (b < t + n) return (u64)(b - t) << (3 * (5 - level)); t += n; }
  return 0;
}
template <class T> static void put(std::vector<u8>& o, T v) { for (size_t i = 0; i < sizeof(T); ++i) o.push_back((u8)((u64)v >> (8 * i))); }
}  // namespace phzvcf
extern "C" {

int phz_vcf_save(phz_vcf* h, const char* path_vcf_gz, int csi, int n_threads) {
  PHZ_TRY
  using namespace phzvcf;
  auto& v = h->v;
  const u8* data = (const u8*)v.out_text.data(); const size_t n = v.out_text.size();
  std::vector<size_t> sizes;
  write_bgzf_file(path_vcf_gz, data, n, n_threads, &sizes);
  std::vector<u64> cstart(sizes.size() + 1, 0);
  for (size_t b = 0; b < sizes.size(); ++b) cstart[b + 1] = cstart[b] + sizes[b];
  auto voff = [&](u64 u) { const u64 b = u / 0xff00; return (cstart[b] << 16) | (u - b * 0xff00); };
  std::vector<RefIndex> refs(v.rec_names.size());
  std::vector<int> order; std::vector<char> seen(v.rec_names.size(), 0);
  const size_t NR = v.rec_chrom.size();
  for (size_t k = 0; k < NR; ++k) {
    const int c = v.rec_chrom[k];
    if (!seen[c]) { seen[c] = 1; order.push_back(c); }
    RefIndex& r = refs[c];
    int64_t beg = v.rec_beg[k], end = v.rec_end[k];
    if (end <= beg) end = beg + 1;
    const u64 o0 = (u64)v.rec_off[k];
    const u8* nl = (const u8*)std::memchr(data + o0, '\n', n - o0);
    const u64 o1 = nl ? (u64)(nl - data) + 1 : n;
    const u64 v0 = voff(o0), v1 = voff(o1);
    const u32 b = reg2bin14(beg, end);
    auto& chunks = r.bins[b];
    if (r.last_bin == (int64_t)b && !chunks.empty() && chunks.back().second == v0) chunks.back().second = v1;
    else chunks.emplace_back(v0, v1);
    r.last_bin = b;
    const u64 w0 = (u64)beg >> 14, w1 = (u64)(end - 1) >> 14;
    if (r.lin.size() <= w1) r.lin.resize(w1 + 1, 0);
    for (u64 w = w0; w <= w1; ++w) if (r.lin[w] == 0) r.lin[w] = v0;
    if (r.n == 0) r.first = v0;
    r.last = v1; r.n++;
  }
  std::vector<u8> aux;
  { std::string names; for (int c : order) { names += v.rec_names[c]; names.push_back('\0'); }
    for (int32_t x : {2, 1, 2, 0, (int32_t)'#', 0, (int32_t)names.size()}) put<int32_t>(aux, x);
    aux.insert(aux.end(), names.begin(), names.end()); }
  const u32 META = ((1u << 18) - 1) / 7 + 1;
  std::vector<u8> out;
  if (!csi) { out.insert(out.end(), {'T', 'B', 'I', 1}); put<int32_t>(out, (int32_t)order.size()); out.insert(out.end(), aux.begin(), aux.end()); }
  else { out.insert(out.end(), {'C', 'S', 'I', 1}); put<int32_t>(out, 14); put<int32_t>(out, 5); put<int32_t>(out, (int32_t)aux.size());
         out.insert(out.end(), aux.begin(), aux.end()); put<int32_t>(out, (int32_t)order.size()); }
  for (int c : order) {
    RefIndex& r = refs[c];
    for (size_t i = 1; i < r.lin.size(); ++i) if (r.lin[i] == 0) r.lin[i] = r.lin[i - 1];      // empty windows point at the previous one
    put<int32_t>(out, (int32_t)r.bins.size() + 1);
    for (auto& kv : r.bins) {
      put<u32>(out, kv.first);
      if (csi) { const u64 w = bin_start_window(kv.first); put<u64>(out, w < r.lin.size() ? r.lin[w] : (r.lin.empty() ? 0 : r.lin.back())); }
      put<int32_t>(out, (int32_t)kv.second.size());
      for (auto& ch : kv.second) { put<u64>(out, ch.first); put<u64>(out, ch.second); }
    }
    put<u32>(out, META); if (csi) put<u64>(out, 0); put<int32_t>(out, 2);
    put<u64>(out, r.first); put<u64>(out, r.last); put<u64>(out, r.n); put<u64>(out, 0);
    if (!csi) { put<int32_t>(out, (int32_t)r.lin.size()); for (u64 x : r.lin) put<u64>(out, x); }
  }
  put<u64>(out, 0);
  const std::string idx = std::string(path_vcf_gz) + (csi ? ".csi" : ".tbi");
  write_bgzf_file(idx.c_str(), out.data(), out.size(), n_threads, nullptr);
  PHZ_CATCH
}


// ---------------------------------------------------------------------------------------------- aReads / bReads columns
// haplotypic_counts.txt, columns 17-18 (phaser.py:1105-1115): per row (final block, BAM, haplotype) and per listed variant
// the reads that carry the haplotype's allele there, as indices into the row's list of distinct reads.  The reference's
// numbering follows CPython's set order; the canonical form is first-occurrence order inside the row, which is
// what the only consumer needs (ids are opaque per row).  Input: the read-list triples of phz_read_lists (sorted by row,
// then variant rank, then tuple order); for every requested row its key and the variants to print, in order.  Output: one
// string per row, variants ';'-joined, indices ','-joined (a variant without reads in the row gives an empty field).
static thread_local std::string g_site_text;

// Text-side view of the requested het sites for the table writers (what generate_variant_dict keeps per variant,
// phaser.py:1418-1462), one line per site: POS, ID, REF, ALT as in the file, then the two alleles the genotype names in
// allele-index order, then the two alleles in genotype order when the genotype is phased ("-" twice otherwise).
// A site whose genotype is not two single digits gets the line "?" and is left to the caller's general reading.
int phz_vcf_site_text(phz_vcf* h, const int64_t* sites, int64_t n, int n_threads, const char** text, int64_t* n_bytes) {
  PHZ_TRY
  using namespace phzvcf;
  Vcf& v = h->v;
  if (v.sample_column < 0) throw PhzError("phz_vcf_site_text: phz_vcf_parse first");
  const int64_t V = (int64_t)v.pos.size();
  const size_t CH = 4096, nch = ((size_t)n + CH - 1) / CH;
  std::vector<std::string> parts(nch);
  std::atomic<int> bad{0};
  const char* base = (const char*)v.text.data();
  const int sc = v.sample_column;
  const int64_t* voff = v.var_off.data(); const int32_t* vlen = v.var_len.data();
  phzio::parallel_for(nch, n_threads, [&](size_t c) {
    std::string& o = parts[c];
    const size_t i1 = std::min((size_t)n, (c + 1) * CH);
    o.reserve((i1 - c * CH) * 40);
    for (size_t i = c * CH; i < i1; ++i) {
      const int64_t s = sites[i];
      if (s < 0 || s >= V) { bad = 1; return; }
      const char* lp = base + voff[s]; const char* le = lp + vlen[s];
      while (le > lp && (le[-1] == '\n' || le[-1] == '\r')) --le;
      Cut cut; bool ok = cut_line(lp, le, sc, cut);
      Span g{nullptr, 0}; int gi = ok ? colon_index(cut.c[8], "GT") : -1;
      ok = ok && gi >= 0 && colon_field(cut.c[9], gi, g);
      Geno q; if (ok) q = read_geno(g);
      ok = ok && !q.dot && !q.overflow && q.n == 2 && q.ch[0] >= '0' && q.ch[0] <= '9' && q.ch[1] >= '0' && q.ch[1] <= '9' &&
           q.ch[0] != q.ch[1];
      Span al[10]; int na = 0;
      if (ok) {
        al[na++] = cut.c[3];
        const char* st = cut.c[4].p; const char* e = st + cut.c[4].n;
        for (const char* t = st;; ++t)
          if (t == e || *t == ',') { if (na < 10) al[na] = Span{st, (size_t)(t - st)}; ++na; st = t + 1; if (t == e) break; }
        const int d0 = q.ch[0] - '0', d1 = q.ch[1] - '0';
        if (na > 10 || d0 >= na || d1 >= na) ok = false;      // an 11th allele would be named "10": not this simple reading
      }
      if (!ok) { o += "?\n"; continue; }
      const int d0 = q.ch[0] - '0', d1 = q.ch[1] - '0';
      const int lo = d0 < d1 ? d0 : d1, hi = d0 < d1 ? d1 : d0;
      bool phased = false;
      for (size_t t = 0; t < g.n; ++t) if (g.p[t] == '|') phased = true;
      auto add = [&](Span x) { o.append(x.p, x.n); o.push_back('\t'); };
      add(cut.c[1]); add(cut.c[2]); add(cut.c[3]); add(cut.c[4]); add(al[lo]); add(al[hi]);
      if (phased) { add(al[d0]); o.append(al[d1].p, al[d1].n); } else o += "-\t-";
      o.push_back('\n');
    }
  });
  if (bad) throw PhzError("phz_vcf_site_text: site index out of range");
  g_site_text.clear();
  size_t tot = 0; for (auto& x : parts) tot += x.size();
  g_site_text.reserve(tot);
  for (auto& x : parts) g_site_text += x;
  *text = g_site_text.data(); *n_bytes = (int64_t)g_site_text.size();
  PHZ_CATCH
}

static thread_local std::string g_rl_text;
static thread_local std::vector<int64_t> g_rl_off;

int phz_format_read_lists(int64_t n, const uint32_t* rl_row, const uint32_t* rl_var, const uint32_t* rl_frag, int64_t n_rows,
                          const uint32_t* row_key, const int64_t* row_var_off, const uint32_t* row_vars, int n_threads,
                          const char** text, const int64_t** row_text_off) {
  PHZ_TRY
  std::vector<std::string> out((size_t)n_rows);
  phzio::parallel_for((size_t)((n_rows + 255) / 256), n_threads, [&](size_t blk) {
    std::vector<u32> hk; std::vector<u32> hv;            // open addressing: fragment -> label, rebuilt per row
    char num[16];
    for (int64_t r = (int64_t)blk * 256; r < n_rows && r < (int64_t)(blk + 1) * 256; ++r) {
      const u32 key = row_key[r];
      const uint32_t* lo = std::lower_bound(rl_row, rl_row + n, key);
      const uint32_t* hi = std::upper_bound(lo, rl_row + n, key);
      int64_t a = lo - rl_row; const int64_t b = hi - rl_row;
      size_t cap = 16; while (cap < (size_t)(b - a) * 2 + 2) cap <<= 1;
      hk.assign(cap, 0xFFFFFFFFu); hv.assign(cap, 0);
      u32 next = 0;
      std::string& o = out[r];
      for (int64_t k = row_var_off[r]; k < row_var_off[r + 1]; ++k) {
        if (k > row_var_off[r]) o.push_back(';');
        const u32 var = row_vars[k];
        // the row's entries come in the order of the block's variants; skip variants that are not asked for
        while (a < b && rl_var[a] != var) {
          bool later = false;
          for (int64_t kk = k + 1; kk < row_var_off[r + 1]; ++kk) if (row_vars[kk] == rl_var[a]) { later = true; break; }
          if (later) break;
          ++a;
        }
        bool first = true;
        while (a < b && rl_var[a] == var) {
          const u32 f = rl_frag[a];
          size_t p = (size_t)(f * 2654435761u) & (cap - 1);
          while (hk[p] != 0xFFFFFFFFu && hk[p] != f) p = (p + 1) & (cap - 1);
          if (hk[p] == 0xFFFFFFFFu) { hk[p] = f; hv[p] = next++; }
          if (!first) o.push_back(',');
          first = false;
          int len = 0; u32 x = hv[p];
          do { num[len++] = (char)('0' + x % 10); x /= 10; } while (x);
          while (len) o.push_back(num[--len]);
          ++a;
        }
      }
    }
  });
  g_rl_off.assign((size_t)n_rows + 1, 0);
  for (int64_t r = 0; r < n_rows; ++r) g_rl_off[r + 1] = g_rl_off[r] + (int64_t)out[r].size();
  g_rl_text.resize((size_t)g_rl_off[n_rows]);
  char* dst = g_rl_text.data(); const int64_t* off = g_rl_off.data();      // (thread_local: the workers must not name them)
  phzio::parallel_for((size_t)((n_rows + 255) / 256), n_threads, [&, dst, off](size_t blk) {
    for (int64_t r = (int64_t)blk * 256; r < n_rows && r < (int64_t)(blk + 1) * 256; ++r)
      if (!out[r].empty()) std::memcpy(dst + off[r], out[r].data(), out[r].size());
  });
  *text = g_rl_text.data(); *row_text_off = g_rl_off.data();
  PHZ_CATCH
}

}  // extern "C"

// extern "C" surface of include/phz.h over Pipeline<PHZ_BACKEND>.  Included by phz_api.cu (product,
// DeviceBackend) and by tests/hostsim/hostsim.cpp (logic-test double, HostSimBackend).
#include "../../include/phz.h"
#include "phz_pipeline.h"
#include "phz_gene_ae.h"
#include <climits>
#include <map>

using namespace phz;

// What one sample's host -> device copy lands in.  Two slots: while K1 and the rest of the path work on one, the
// copy stream can already fill the other with the next sample (phz_prefetch_packed).
struct TransportSlot {
  // the packed fields as they arrive (include/phz.h: phz_packed_reads)
  Buf<PHZ_BACKEND, uint16_t> pos_d, lsq; Buf<PHZ_BACKEND, int16_t> tl16, as_tab; Buf<PHZ_BACKEND, u32> pos_xi, tl_xi, cig_tab;
  Buf<PHZ_BACKEND, int32_t> pos_xv, tl_xv; Buf<PHZ_BACKEND, u8> as_raw, ncg, cig_raw, seq2, qualp, exc, qtab; Buf<PHZ_BACKEND, u64> exi;
  Buf<PHZ_BACKEND, u32> frag;       // the ids (copied as they are, or rebuilt from the implicit coding); read again by phz_commit_bam
  Buf<PHZ_BACKEND, u32> frag_fb, frag_xi, frag_xv; Buf<PHZ_BACKEND, uint16_t> frag_bk;      // implicit coding (phz.h: frag_first / frag_back)
  const void* tag = nullptr;      // host buffer the staged copy came from
  bool staged = false;
  u64 seq = 0;
  void* ready = nullptr;          // backend event: all copies of the staged sample have landed
  void bind(PHZ_BACKEND* b) {
    pos_d.bind(b); lsq.bind(b); tl16.bind(b); as_tab.bind(b); pos_xi.bind(b); tl_xi.bind(b); cig_tab.bind(b); pos_xv.bind(b);
    tl_xv.bind(b); as_raw.bind(b); ncg.bind(b); cig_raw.bind(b); seq2.bind(b); qualp.bind(b); exc.bind(b); qtab.bind(b);
    exi.bind(b); frag.bind(b); frag_fb.bind(b); frag_xi.bind(b); frag_xv.bind(b); frag_bk.bind(b);
  }
};

struct phz_ctx {
  Pipeline<PHZ_BACKEND> p;
  GeneAE<PHZ_BACKEND> ae;
  TransportSlot slot[2];
  int last_slot = 1;
  u64 stage_seq = 0;
  const u32* cur_frag = nullptr;   // fragment ids of the sample mapped last through a host entry point
  int64_t last_R = 0, last_NC = 0, last_NB = 0;     // sizes of the arrays expanded last (phz_array "st_*")
  // expanded arrays (device layout of phz_reads) that the copies / expansion kernels produce for K1
  Buf<PHZ_BACKEND, u32> st_coff, st_cig, st_tmp; Buf<PHZ_BACKEND, u64> st_soff; Buf<PHZ_BACKEND, u8> st_seq, st_qual;
  Buf<PHZ_BACKEND, int32_t> st_pos, st_tlen; Buf<PHZ_BACKEND, int16_t> st_as;
  phz_ctx() {
    PHZ_BACKEND* b = &p.be;
    ae.bind(b); slot[0].bind(b); slot[1].bind(b);
    st_coff.bind(b); st_cig.bind(b); st_tmp.bind(b); st_soff.bind(b); st_seq.bind(b); st_qual.bind(b);
    st_pos.bind(b); st_tlen.bind(b); st_as.bind(b);
  }
  ~phz_ctx() { for (auto& s : slot) p.be.free_event(s.ready); }
};

static thread_local std::string g_err;

#define PHZ_TRY try {
#define PHZ_CATCH                                                     \
  }                                                                   \
  catch (const std::exception& e) { g_err = e.what(); return -1; }    \
  catch (...) { g_err = "unknown error"; return -2; }                 \
  return 0;

extern "C" {

const char* phz_last_error(void) { return g_err.c_str(); }
const char* phz_backend_name(void) { return PHZ_BACKEND_NAME; }

phz_ctx* phz_create(int device, void* stream) {
  try {
#ifdef __CUDACC__
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) { g_err = "no CUDA device available (this library has no CPU path)"; return nullptr; }
    PHZ_CUDA(cudaSetDevice(device));
#endif
    phz_ctx* c = new phz_ctx();
    c->p.be.device = device;
#ifdef __CUDACC__
    c->p.be.stream = (cudaStream_t)stream;
#else
    (void)stream;
#endif
    return c;
  } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}

void phz_destroy(phz_ctx* ctx) { delete ctx; }

int phz_sync(phz_ctx* ctx) { PHZ_TRY ctx->p.be.sync(); PHZ_CATCH }

int phz_set_variants(phz_ctx* ctx, int n_contigs, const int64_t* h_off, const int32_t* d_pos, const uint8_t* d_a0,
                     const uint8_t* d_a1, int64_t n_variants) {
  PHZ_TRY ctx->p.set_variants(n_contigs, h_off, d_pos, d_a0, d_a1, n_variants); PHZ_CATCH
}

int phz_set_indel_alleles(phz_ctx* ctx, const int32_t* d_ref_len, const uint32_t* d_al_off, const uint8_t* d_al_codes) {
  PHZ_TRY
  if (d_ref_len && (!d_al_off || !d_al_codes)) throw PhzError("phz_set_indel_alleles: all three tables are needed");
  ctx->p.v_ref_len = d_ref_len; ctx->p.v_al_off = d_ref_len ? d_al_off : nullptr; ctx->p.v_al_codes = d_ref_len ? d_al_codes : nullptr;
  PHZ_CATCH
}

int phz_set_haplo_blacklist(phz_ctx* ctx, const uint8_t* d_flags) { PHZ_TRY ctx->p.vblack = d_flags; PHZ_CATCH }

static ReadsView view_of(const phz_reads* r, int nc) {
  ReadsView v;
  v.n_records = r->n_records; v.n_contigs = nc; v.contig_rec_off = nullptr;
  v.pos = r->pos; v.tlen = r->tlen; v.aln_score = r->aln_score; v.frag = r->frag;
  v.cigar_off = r->cigar_off; v.cigar = r->cigar; v.seq_off = (const u64*)r->seq_off; v.seq = r->seq; v.qual = r->qual;
  v.n_cigar_ops = r->n_cigar_ops;
  return v;
}

int phz_map_reads(phz_ctx* ctx, const phz_reads* reads, int baseq, double isize_cutoff, int64_t* n_candidates) {
  PHZ_TRY
  ReadsView v = view_of(reads, ctx->p.nc);
  *n_candidates = ctx->p.map_reads(v, reads->h_contig_rec_off, baseq, isize_cutoff);
  PHZ_CATCH
}

int phz_map_reads_host(phz_ctx* ctx, const phz_reads* h, int baseq, double isize_cutoff, int64_t* n_candidates) {
  PHZ_TRY
  auto& be = ctx->p.be;
  int64_t R = h->n_records;
  phz_reads d = *h;
  TransportSlot& S = ctx->slot[ctx->slot[0].staged ? 1 : 0];
  if (S.staged) throw PhzError("phz_map_reads_host: both transport slots hold prefetched samples");
  be.h2d(ctx->st_pos.ensure(R), h->pos, R * 4); d.pos = ctx->st_pos.p;
  be.h2d(ctx->st_tlen.ensure(R), h->tlen, R * 4); d.tlen = ctx->st_tlen.p;
  be.h2d(ctx->st_as.ensure(R), h->aln_score, R * 2); d.aln_score = ctx->st_as.p;
  be.h2d(S.frag.ensure(R), h->frag, R * 4); d.frag = S.frag.p; ctx->cur_frag = S.frag.p;
  be.h2d(ctx->st_coff.ensure(R + 1), h->cigar_off, (R + 1) * 4); d.cigar_off = ctx->st_coff.p;
  be.h2d(ctx->st_cig.ensure(h->n_cigar_ops), h->cigar, h->n_cigar_ops * 4); d.cigar = ctx->st_cig.p;
  be.h2d(ctx->st_soff.ensure(R + 1), h->seq_off, (R + 1) * 8); d.seq_off = (const uint64_t*)ctx->st_soff.p;
  be.h2d(ctx->st_seq.ensure((h->n_bases + 1) / 2), h->seq, (h->n_bases + 1) / 2); d.seq = ctx->st_seq.p;
  be.h2d(ctx->st_qual.ensure(h->n_bases), h->qual, h->n_bases); d.qual = ctx->st_qual.p;
  ReadsView v = view_of(&d, ctx->p.nc);
  *n_candidates = ctx->p.map_reads(v, h->h_contig_rec_off, baseq, isize_cutoff);
  PHZ_CATCH
}

static void check_packed(const phz_packed_reads* h) {
  const int bits = h->qual_bits;
  if (bits != 1 && bits != 2 && bits != 4 && bits != 8) throw PhzError("packed reads: qual_bits must be 1, 2, 4 or 8");
  if (h->as_bits != 8 && h->as_bits != 16) throw PhzError("packed reads: as_bits must be 8 or 16");
  if (h->n_cigar_bits != 8 && h->n_cigar_bits != 16) throw PhzError("packed reads: n_cigar_bits must be 8 or 16");
  if (h->cigar_bits != 16 && h->cigar_bits != 32) throw PhzError("packed reads: cigar_bits must be 16 or 32");
  if (h->l_seq_const < 0 && !h->l_seq && h->n_records > 0) throw PhzError("packed reads: l_seq missing");
  if (h->frag_bits != 32 && h->frag_bits != 16) throw PhzError("packed reads: frag_bits must be 32 or 16");
}

// sizes the slot's buffers for `h` (may reallocate: call before any copy is enqueued)
static void size_slot(TransportSlot& S, const phz_packed_reads* h) {
  const int64_t R = h->n_records, NB = h->n_bases, NC = h->n_cigar_ops, NX = h->n_exceptions;
  S.ncg.ensure(R * 2 + 16); S.lsq.ensure(R); S.qualp.ensure((NB * h->qual_bits + 7) / 8 + 16); S.qtab.ensure(256);
  S.seq2.ensure((NB + 3) / 4 + 16); S.exi.ensure(NX); S.exc.ensure(NX);
  S.cig_raw.ensure(NC * 4 + 16); S.cig_tab.ensure(65536);
  S.pos_d.ensure(R); S.pos_xi.ensure(h->n_pos_exc); S.pos_xv.ensure(h->n_pos_exc);
  S.tl16.ensure(R); S.tl_xi.ensure(h->n_tlen_exc); S.tl_xv.ensure(h->n_tlen_exc);
  S.as_raw.ensure(R * 2 + 16); S.as_tab.ensure(256); S.frag.ensure(R);
  if (h->frag_bits == 16) { S.frag_fb.ensure((R + 31) / 32 + 1); S.frag_bk.ensure(h->n_frag_back + 1); S.frag_xi.ensure(h->n_frag_exc); S.frag_xv.ensure(h->n_frag_exc); }
}

// the four copy groups of one sample, in the order the expansion consumes them
static void copy_group(PHZ_BACKEND& be, TransportSlot& S, const phz_packed_reads* h, int group) {
  const int64_t R = h->n_records, NB = h->n_bases, NC = h->n_cigar_ops, NX = h->n_exceptions;
  if (group == 0) {
    be.h2d_copy(S.ncg.p, h->n_cigar, R * (h->n_cigar_bits / 8));
    if (h->l_seq_const < 0) be.h2d_copy(S.lsq.p, h->l_seq, R * 2);
    if (h->frag_bits == 16) {         // implicit fragment ids travel with the counts: they are rebuilt under the copies that follow
      be.h2d_copy(S.frag_fb.p, h->frag_first, ((R + 31) / 32) * 4); be.h2d_copy(S.frag_bk.p, h->frag_back, h->n_frag_back * 2);
      be.h2d_copy(S.frag_xi.p, h->frag_exc_index, h->n_frag_exc * 4); be.h2d_copy(S.frag_xv.p, h->frag_exc_value, h->n_frag_exc * 4);
    }
  } else if (group == 1) {
    be.h2d_copy(S.qualp.p, h->qualp, (NB * h->qual_bits + 7) / 8); be.h2d_copy(S.qtab.p, h->qual_table, 256);
  } else if (group == 2) {
    be.h2d_copy(S.seq2.p, h->seq2, (NB + 3) / 4); be.h2d_copy(S.exi.p, h->exc_index, NX * 8); be.h2d_copy(S.exc.p, h->exc_code, NX);
  } else {
    be.h2d_copy(S.cig_raw.p, h->cigar, NC * (h->cigar_bits / 8));
    if (h->cigar_bits == 16) be.h2d_copy(S.cig_tab.p, h->cigar_table, (size_t)h->n_cigar_table * 4);
    be.h2d_copy(S.pos_d.p, h->pos_delta, R * 2);
    be.h2d_copy(S.pos_xi.p, h->pos_exc_index, h->n_pos_exc * 4); be.h2d_copy(S.pos_xv.p, h->pos_exc_delta, h->n_pos_exc * 4);
    be.h2d_copy(S.tl16.p, h->tlen16, R * 2);
    be.h2d_copy(S.tl_xi.p, h->tlen_exc_index, h->n_tlen_exc * 4); be.h2d_copy(S.tl_xv.p, h->tlen_exc_value, h->n_tlen_exc * 4);
    be.h2d_copy(S.as_raw.p, h->as_data, R * (h->as_bits / 8));
    if (h->as_bits == 8) be.h2d_copy(S.as_tab.p, h->as_table, 512);
    if (h->frag_bits == 32) be.h2d_copy(S.frag.p, h->frag, R * 4);
  }
}

int phz_prefetch_packed(phz_ctx* ctx, const phz_packed_reads* h) {
  PHZ_TRY
  auto& be = ctx->p.be;
  check_packed(h);
  int si = ctx->slot[0].staged ? 1 : (ctx->slot[1].staged ? 0 : 1 - ctx->last_slot);
  TransportSlot& S = ctx->slot[si];
  if (S.staged) throw PhzError("phz_prefetch_packed: both transport slots are already staged");
  size_slot(S, h);
  if (!S.ready) S.ready = be.new_event();
  be.copy_begin();        // the slot may still be read by kernels queued on the main stream
  for (int g = 0; g < 4; ++g) copy_group(be, S, h, g);
  be.copy_record(S.ready);
  S.tag = (const void*)h->seq2; S.staged = true; S.seq = ++ctx->stage_seq;
  PHZ_CATCH
}

int phz_map_reads_packed(phz_ctx* ctx, const phz_packed_reads* h, int baseq, double isize_cutoff, int64_t* n_candidates) {
  PHZ_TRY
  auto& be = ctx->p.be;
  check_packed(h);
  const int64_t R = h->n_records, NB = h->n_bases, NC = h->n_cigar_ops, NX = h->n_exceptions;
  const int bits = h->qual_bits;
  phz_reads d; std::memset(&d, 0, sizeof(d));
  d.n_records = R; d.n_cigar_ops = NC; d.n_bases = NB; d.h_contig_rec_off = h->h_contig_rec_off;
  // a sample staged by phz_prefetch_packed from these very buffers?  (oldest first)
  int si = -1;
  for (int k = 0; k < 2; ++k)
    if (ctx->slot[k].staged && ctx->slot[k].tag == (const void*)h->seq2 && (si < 0 || ctx->slot[k].seq < ctx->slot[si].seq)) si = k;
  const bool prefetched = si >= 0;
  if (!prefetched) {
    si = ctx->slot[0].staged ? 1 : (ctx->slot[1].staged ? 0 : 1 - ctx->last_slot);
    if (ctx->slot[si].staged) throw PhzError("phz_map_reads_packed: both transport slots hold other prefetched samples");
    size_slot(ctx->slot[si], h);
  }
  TransportSlot& S = ctx->slot[si];
  const u8* qp = S.qualp.p; const u8* qt = S.qtab.p; const u8* s2 = S.seq2.p; const u64* exi = S.exi.p; const u8* exc = S.exc.p;
  u32* coff = ctx->st_coff.ensure(R + 1); u64* soff = ctx->st_soff.ensure(R + 1); u32* tmp = ctx->st_tmp.ensure(R + 1);
  const int64_t nwq = (NB + 7) / 8, nws = (NB + 15) / 16;
  u64* qout = (u64*)ctx->st_qual.ensure((size_t)nwq * 8 + 16);
  u64* sout = (u64*)ctx->st_seq.ensure((size_t)nws * 8 + 16);
  u32* cig = ctx->st_cig.ensure(NC); int32_t* pos = ctx->st_pos.ensure(R); int32_t* tlen = ctx->st_tlen.ensure(R);
  int16_t* as = ctx->st_as.ensure(R);
  // ---- host -> device on the copy stream, in the order the expansion needs it: counts, qualities, bases, then the
  // per-record fields.  Each expansion kernel starts as soon as ITS input has landed and runs under the copies that
  // follow, so only the PCIe time is on the critical path.  With a prefetched sample the copies are already under
  // way (or done): the main stream just waits for them.
  be.stage("h2d+unpack");
  if (prefetched) be.wait_event(S.ready);
  else { be.copy_begin(); copy_group(be, S, h, 0); be.copy_fence(); copy_group(be, S, h, 1); }
  {   // counts -> offsets
    const u8* n8 = S.ncg.p; const uint16_t* n16 = (const uint16_t*)S.ncg.p; const int nbits = h->n_cigar_bits;
    be.for_each(R, PHZ_LAMBDA(int64_t r) { tmp[r] = nbits == 8 ? (u32)n8[r] : (u32)n16[r]; });
    be.exclusive_scan_u32(tmp, coff, R); d.cigar_off = coff;
    if (h->l_seq_const >= 0) { const u64 L = (u64)h->l_seq_const; be.for_each(R + 1, PHZ_LAMBDA(int64_t r) { soff[r] = (u64)r * L; }); }
    else be.exclusive_scan_u16_to_u64(S.lsq.p, soff, R);
    d.seq_off = (const uint64_t*)soff;
  }
  if (h->frag_bits == 16) {   // fragment ids from the first-appearance bitmap: one scan (set bits before r), one pass, the exceptions
    const u32* fb = S.frag_fb.p; const uint16_t* bk = S.frag_bk.p; const u32* fxi = S.frag_xi.p; const u32* fxv = S.frag_xv.p;
    const u32 base = h->frag_base; u32* fr = S.frag.p;
    auto opens = PHZ_LAMBDA(int64_t r) -> u32 { return (fb[r >> 5] >> (r & 31)) & 1u; };
    be.exclusive_scan_fn_u32(opens, tmp, R);
    be.for_each(R, PHZ_LAMBDA(int64_t r) {
      const u32 nf = tmp[r];
      fr[r] = ((fb[r >> 5] >> (r & 31)) & 1u) ? base + nf : base + nf - (u32)bk[r - nf];
    });
    be.for_each(h->n_frag_exc, PHZ_LAMBDA(int64_t e) { fr[fxi[e]] = fxv[e]; });
  }
  if (!prefetched) { be.copy_fence(); copy_group(be, S, h, 2); }
  {   // base qualities: one logical thread per 8 bases = `bits` packed bytes in, 8 phred bytes out
    const u32 mask = (1u << bits) - 1;
    be.for_each(nwq, PHZ_LAMBDA(int64_t w) {
      u64 in = 0;
      for (int j = 0; j < bits; ++j) in |= (u64)qp[w * bits + j] << (8 * j);
      u64 o = 0;
      for (int k = 0; k < 8; ++k) o |= (u64)qt[(in >> (k * bits)) & mask] << (8 * k);
      qout[w] = o;
    });
    d.qual = ctx->st_qual.p;
  }
  if (!prefetched) { be.copy_fence(); copy_group(be, S, h, 3); }
  {   // bases: one logical thread per 16 bases = 4 packed bytes in, 8 bytes out (A C G T -> 1 2 4 8, even index = high nibble)
    const u32* in32 = (const u32*)s2;
    be.for_each(nws, PHZ_LAMBDA(int64_t w) {
      u32 in = in32[w]; u64 o = 0;
      for (int k = 0; k < 16; ++k) {
        u64 nib = (u64)1 << ((in >> (2 * k)) & 3);
        o |= nib << ((k >> 1) * 8 + ((k & 1) ? 0 : 4));
      }
      sout[w] = o;
    });
    u32* words = (u32*)ctx->st_seq.p;
    be.for_each(NX, PHZ_LAMBDA(int64_t e) {      // every other code (N, IUPAC, '='): patch the nibble
      u64 i = exi[e]; u32 sh = (u32)(((i >> 1) & 3) * 8 + ((i & 1) ? 0 : 4));
      atomic_and(&words[i >> 3], ~(0xFu << sh));
      atomic_or(&words[i >> 3], (u32)exc[e] << sh);
    });
    d.seq = ctx->st_seq.p;
  }
  if (!prefetched) be.copy_fence();
  {   // per-record fields: table look-ups, exception scatters, one inclusive scan for pos
    const uint16_t* c16 = (const uint16_t*)S.cig_raw.p; const u32* c32 = (const u32*)S.cig_raw.p; const u32* ctab = S.cig_tab.p;
    const int cbits = h->cigar_bits;
    be.for_each(NC, PHZ_LAMBDA(int64_t i) { cig[i] = cbits == 16 ? ctab[c16[i]] : c32[i]; });
    const uint16_t* pd = S.pos_d.p; const int16_t* t16 = S.tl16.p;
    const u8* a8 = S.as_raw.p; const int16_t* a16 = (const int16_t*)S.as_raw.p; const int16_t* atab = S.as_tab.p; const int abits = h->as_bits;
    be.for_each(R, PHZ_LAMBDA(int64_t r) {
      pos[r] = pd[r] == 65535 ? 0 : (int32_t)pd[r];
      tlen[r] = (int32_t)t16[r];
      as[r] = abits == 8 ? atab[a8[r]] : a16[r];
    });
    const u32* pxi = S.pos_xi.p; const int32_t* pxv = S.pos_xv.p; const u32* txi = S.tl_xi.p; const int32_t* txv = S.tl_xv.p;
    be.for_each(h->n_pos_exc, PHZ_LAMBDA(int64_t e) { pos[pxi[e]] = pxv[e]; });
    be.for_each(h->n_tlen_exc, PHZ_LAMBDA(int64_t e) { tlen[txi[e]] = txv[e]; });
    be.inclusive_scan_i32_inplace(pos, R);
    d.cigar = cig; d.pos = pos; d.tlen = tlen; d.aln_score = as; d.frag = S.frag.p;
  }
  S.staged = false; ctx->last_slot = si; ctx->cur_frag = S.frag.p;
  ctx->last_R = R; ctx->last_NC = NC; ctx->last_NB = NB;
  be.stage("h2d+unpack.end");
  ReadsView v = view_of(&d, ctx->p.nc);
  *n_candidates = ctx->p.map_reads(v, h->h_contig_rec_off, baseq, isize_cutoff);
  PHZ_CATCH
}

int phz_as_histogram(phz_ctx* ctx, uint64_t* d_hist) { PHZ_TRY ctx->p.as_histogram((u64*)d_hist); PHZ_CATCH }

int phz_commit_bam(phz_ctx* ctx, int bam_index, int32_t as_cutoff, const uint32_t* d_frag, int64_t* n_kept) {
  PHZ_TRY
  const u32* f = d_frag ? d_frag : ctx->cur_frag;
  if (!f && ctx->p.n_cand > 0) throw PhzError("phz_commit_bam: no fragment ids");
  *n_kept = ctx->p.commit_bam(bam_index, as_cutoff, f);
  PHZ_CATCH
}

int phz_variant_stats(phz_ctx* ctx, uint64_t* h_noise) {
  PHZ_TRY
  u64 nz[2] = {0, 0};
  void* ev = ctx->p.noise_event; ctx->p.noise_event = nullptr;        // synchronous form: wait here
  try { ctx->p.variant_stats(nz); } catch (...) { ctx->p.noise_event = ev; throw; }
  ctx->p.noise_event = ev;
  h_noise[0] = nz[0]; h_noise[1] = nz[1];
  PHZ_CATCH
}

int phz_variant_stats_async(phz_ctx* ctx) {
  PHZ_TRY
  if (!ctx->p.noise_event) ctx->p.noise_event = ctx->p.be.new_event();
  u64 unused[2];
  ctx->p.variant_stats(unused);
  PHZ_CATCH
}

int phz_variant_stats_device(phz_ctx* ctx, uint64_t* d_noise) {
  PHZ_TRY
  u64 unused[2];
  ctx->p.noise_dev_out = (u64*)d_noise;
  try { ctx->p.variant_stats(unused); } catch (...) { ctx->p.noise_dev_out = nullptr; throw; }
  ctx->p.noise_dev_out = nullptr;
  PHZ_CATCH
}

int phz_noise_publish(phz_ctx* ctx, const uint64_t* d_noise) {
  PHZ_TRY
  if (!ctx->p.noise_event) ctx->p.noise_event = ctx->p.be.new_event();
  ctx->p.noise_publish((const u64*)d_noise);
  PHZ_CATCH
}

int phz_noise_wait(phz_ctx* ctx, uint64_t* h_noise) {
  PHZ_TRY
  u64 nz[2] = {0, 0};
  ctx->p.noise_wait(nz);
  h_noise[0] = nz[0]; h_noise[1] = nz[1];
  PHZ_CATCH
}

int phz_build_graph(phz_ctx* ctx, uint64_t n_fragments, uint64_t excl, int64_t* n_edges, uint32_t* max_c_total) {
  PHZ_TRY
  ctx->p.build_graph(n_fragments, excl);
  *n_edges = ctx->p.E; *max_c_total = ctx->p.max_tot;
  PHZ_CATCH
}

int phz_set_big_critical_values(phz_ctx* ctx, const uint32_t* h_n, const uint32_t* h_k, int64_t count) {
  PHZ_TRY
  for (int64_t i = 1; i < count; ++i) if (h_n[i] <= h_n[i - 1]) throw PhzError("phz_set_big_critical_values: totals must ascend");
  ctx->p.set_big_critical_values(h_n, h_k, count);
  PHZ_CATCH
}

int phz_phase(phz_ctx* ctx, const uint32_t* h_kstar, int64_t kstar_len, int max_block_size, uint64_t excl,
              int64_t* n_final_blocks, int* status_flags) {
  PHZ_TRY
  int err = 0;
  ctx->p.phase(h_kstar, kstar_len, max_block_size, excl, &err);
  *n_final_blocks = ctx->p.NF; *status_flags = err;
  PHZ_CATCH
}

int phz_gene_ae_pairs(phz_ctx* ctx, const phz_ae_input* in, int64_t* n_pairs) {
  PHZ_TRY *n_pairs = ctx->ae.run(*in); PHZ_CATCH
}

int phz_read_lists(phz_ctx* ctx, uint64_t excl, int64_t* n_entries) {
  PHZ_TRY *n_entries = ctx->p.read_lists(excl); PHZ_CATCH
}

struct ArrRef { const void* p; int64_t n; int eb; };

static bool find_array(phz_ctx* ctx, const std::string& name, ArrRef* out) {
  auto& p = ctx->p;
  if (name == "t_rec" || name == "t_var" || name == "t_misc") p.ensure_canonical();
  const int64_t nb = p.n_bams > 0 ? p.n_bams : 1;
#define A(nm, buf, cnt) if (name == nm) { *out = ArrRef{(const void*)p.buf.p, (int64_t)(cnt), (int)sizeof(*p.buf.p)}; return true; }
  A("t_rec", t_rec, p.n_cand) A("t_var", t_var, p.n_cand) A("t_misc", t_misc, p.n_cand)
  A("g_frag", g_frag, p.n_tuples) A("g_var", g_var, p.n_tuples) A("g_cb", g_cb, p.n_tuples)
  A("vfirst", vfirst, p.V) A("ncls", ncls, p.V * 3) A("setsize", setsize, p.V * 3) A("vb_cnt", vb_cnt, p.V * nb * 2)
  A("vrank", vrank, p.V) A("cfirst", cfirst, p.nc) A("crank", crank, p.nc)
  A("e_key", e_key, p.NE) A("e_bam", e_bam, p.NE) A("e_mask", e_mask, p.NE) A("e_tmin", e_tmin, p.NE) A("grp_off", grp_off, p.NG + 1)
  A("ed_a", ed_a, p.E) A("ed_b", ed_b, p.E) A("ed_sup", ed_sup, p.E) A("ed_tot", ed_tot, p.E) A("ed_n9", ed_n9, p.E * 9)
  A("ed_cfg", ed_cfg, p.E) A("ed_keep", ed_keep, p.E) A("big_tot", big_tot, p.NBT)
  A("members", members, p.NM) A("blk_off", blk_off, p.NB + 1) A("blk_order", blk_order, p.NB) A("blk_status", blk_status, p.NB)
  A("blk_nfinal", blk_nfinal, p.NB) A("blk_rank", blk_rank, p.NB)
  A("fb_first", fb_first, p.NF) A("fb_len", fb_len, p.NF) A("fb_blk", fb_blk, p.NF) A("fb_sup", fb_sup, p.NF) A("fb_tot", fb_tot, p.NF)
  A("fb_cnt", fb_cnt, p.NF * 2) A("fb_bcnt", fb_bcnt, p.NF * nb * 2) A("v_final", v_final, p.V) A("v_hap", v_hap, p.V)
  A("rl_frag", rl_frag, p.NRL) A("rl_var", rl_var, p.NRL) A("rl_row", rl_row, p.NRL)
#undef A
  // the arrays phz_map_reads_packed expanded last (device layout of phz_reads): for round-trip checks
#define S(nm, buf, cnt) if (name == nm) { *out = ArrRef{(const void*)ctx->buf.p, (int64_t)(cnt), (int)sizeof(*ctx->buf.p)}; return true; }
  S("st_pos", st_pos, ctx->last_R) S("st_tlen", st_tlen, ctx->last_R) S("st_as", st_as, ctx->last_R) S("st_cig", st_cig, ctx->last_NC)
  S("st_coff", st_coff, ctx->last_R + 1) S("st_soff", st_soff, ctx->last_R + 1) S("st_seq", st_seq, (ctx->last_NB + 1) / 2)
  S("st_qual", st_qual, ctx->last_NB)
#undef S
  if (name == "st_frag") { *out = ArrRef{(const void*)ctx->cur_frag, ctx->cur_frag ? ctx->last_R : 0, (int)sizeof(u32)}; return true; }
#define G(nm, buf) if (name == nm) { *out = ArrRef{(const void*)ctx->ae.buf.p, ctx->ae.NPAIR, (int)sizeof(u32)}; return true; }
  G("ae_row", ae_row) G("ae_feat", ae_feat) G("ae_a", ae_a) G("ae_b", ae_b)
#undef G
  return false;
}

int phz_array(phz_ctx* ctx, const char* name, const void** d_ptr, int64_t* count, int* elem_bytes) {
  PHZ_TRY
  ArrRef r;
  if (!find_array(ctx, name, &r)) throw PhzError(std::string("unknown array: ") + name);
  *d_ptr = r.p; *count = r.n; *elem_bytes = r.eb;
  PHZ_CATCH
}

int phz_download(phz_ctx* ctx, const char* name, void* h_dst, int64_t dst_bytes) {
  PHZ_TRY
  ArrRef r;
  if (!find_array(ctx, name, &r)) throw PhzError(std::string("unknown array: ") + name);
  if (r.n * r.eb > dst_bytes) throw PhzError(std::string("destination too small for array ") + name);
  if (r.n > 0) ctx->p.be.d2h(h_dst, r.p, (size_t)(r.n * r.eb)); else ctx->p.be.sync();
  PHZ_CATCH
}

int phz_download_async(phz_ctx* ctx, const char* name, void* h_dst, int64_t dst_bytes) {
  PHZ_TRY
  ArrRef r;
  if (!find_array(ctx, name, &r)) throw PhzError(std::string("unknown array: ") + name);
  if (r.n * r.eb > dst_bytes) throw PhzError(std::string("destination too small for array ") + name);
  if (r.n > 0) ctx->p.be.d2h_async(h_dst, r.p, (size_t)(r.n * r.eb));
  PHZ_CATCH
}

int phz_copy_array(phz_ctx* ctx, const char* name, void* dst, int64_t dst_bytes) {
  PHZ_TRY
  ArrRef r;
  if (!find_array(ctx, name, &r)) throw PhzError(std::string("unknown array: ") + name);
  if (r.n * r.eb > dst_bytes) throw PhzError(std::string("destination too small for array ") + name);
  if (r.n > 0) ctx->p.be.copy_out_async(dst, r.p, (size_t)(r.n * r.eb));
  PHZ_CATCH
}

int phz_expand_runs(phz_ctx* ctx, int64_t n_runs, const int64_t* rf, const int64_t* rd, const int64_t* rr, const int64_t* rb,
                    int64_t n, const uint32_t* site, const uint32_t* frag, const int64_t* map, uint32_t* o_row, uint32_t* o_site,
                    uint32_t* o_frag) {
  PHZ_TRY
  if (n_runs <= 0 || n <= 0) return 0;
  ctx->p.be.for_each(n, PHZ_LAMBDA(int64_t i) {
    int64_t lo = 0, hi = n_runs;                 // last run with first <= i
    while (hi - lo > 1) { const int64_t mid = (lo + hi) >> 1; if (rf[mid] <= i) lo = mid; else hi = mid; }
    const int64_t d = rd[lo] + (i - rf[lo]);
    o_row[d] = (u32)rr[lo];
    o_site[d] = (u32)map[(int64_t)site[i] + rb[lo]];
    o_frag[d] = frag[i];
  });
  PHZ_CATCH
}

int phz_counters(phz_ctx* ctx, int64_t* c) {
  PHZ_TRY
  auto& p = ctx->p;
  int64_t v[16] = {p.n_tuples, p.NE, p.NG, p.NP, p.NX, p.E, (int64_t)p.n_dropped, p.NM, p.NB, p.NH, p.NF, p.NRL,
                   p.n_cand, p.n_bams, p.n_runs_resorted, p.full_sort_fallback ? 1 : 0};
  for (int i = 0; i < 16; ++i) c[i] = v[i];
  PHZ_CATCH
}

int phz_set_option(phz_ctx* ctx, const char* name, int64_t value) {
  PHZ_TRY
  std::string n(name);
  if (n == "k1_mode") ctx->p.k1_mode = (int)value;
  else if (n == "k1_min_ctas") ctx->p.k1_min_ctas = (int)value;
  else if (n == "k1_staged_emit") ctx->p.k1_staged_emit = (int)value;
  else if (n == "big_total_threshold") ctx->p.big_total_thr = (u32)value;
  else if (n == "lazy_canonical") ctx->p.lazy_canonical = (int)value;
  else if (n == "two_pass_read_lists") ctx->p.two_pass_read_lists = (int)value;
  else if (n == "wide_pair_keys") ctx->p.wide_pair_keys = (int)value;
  else if (n == "window_agg") ctx->p.window_agg = (int)value;
  else if (n == "graph_mode") ctx->p.graph_mode = (int)value;
  else if (n == "frag_stage") ctx->p.frag_stage = (int)value;
  else if (n == "n_fragments") ctx->p.n_frag_hint = value > 0 ? value : 0;
  else if (n == "pair_table_slots") { u64 s = 16; while (s < (u64)value) s <<= 1; ctx->p.pair_table_slots = s; }
  else if (n == "frag_run_limit") ctx->p.frag_run_limit = value < 1 ? 1 : value;
  else throw PhzError("unknown option: " + n);
  PHZ_CATCH
}

int phz_get_option(phz_ctx* ctx, const char* name, int64_t* value) {
  PHZ_TRY
  std::string n(name);
  if (n == "graph_ranked_in_commit") *value = ctx->p.graph_ranked_in_commit ? 1 : 0;
  else if (n == "pair_table_slots") *value = (int64_t)ctx->p.pair_table_slots;
  else if (n == "graph_mode") *value = ctx->p.graph_mode;
  else if (n == "frag_stage") *value = ctx->p.frag_stage;
  else if (n == "fragments_deferred") *value = (int64_t)ctx->p.n_frag_deferred;
  else if (n == "k1_mode") *value = ctx->p.k1_mode;
  else if (n == "n_fragments") *value = ctx->p.n_frag_hint;
  else throw PhzError("unknown option: " + n);
  PHZ_CATCH
}

int phz_set_profiling(phz_ctx* ctx, int on) { PHZ_TRY ctx->p.be.profiling = on; PHZ_CATCH }

int phz_map_times(phz_ctx* ctx, float* ms) {
  PHZ_TRY
  ms[0] = ctx->p.be.elapsed(0, 1); ms[1] = ctx->p.be.elapsed(1, 2); ms[2] = ctx->p.be.elapsed(2, 3);
  PHZ_CATCH
}

int phz_stage_report(phz_ctx* ctx, char* buf, int64_t buf_len) {
  PHZ_TRY
  std::string r = ctx->p.be.stage_report();
  if ((int64_t)r.size() + 1 > buf_len) r.resize(buf_len > 0 ? buf_len - 1 : 0);
  std::memcpy(buf, r.c_str(), r.size() + 1);
  PHZ_CATCH
}

int phz_sync_count(phz_ctx* ctx, uint64_t* n) { PHZ_TRY *n = ctx->p.be.syncs; PHZ_CATCH }

int phz_launch_counts(phz_ctx* ctx, uint64_t* own, uint64_t* library) {
  PHZ_TRY *own = ctx->p.be.launches; *library = ctx->p.be.lib_launches; PHZ_CATCH
}

}  // extern "C"

#include "phz_io.inl"
#include "phz_vcf.inl"

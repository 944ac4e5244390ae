// The read -> variant -> haplotype pipeline on packed SoA arrays (see DESIGN.md for the data layout).
//
// Stage map (reference file:line each stage replaces, all in /root/reference/phaser/):
//   map_reads      read_variant_map.py:25-123,165-258  K1: one logical thread per record walks the CIGAR
//   as_histogram   phaser.py:545-553                    exact AS histogram -> percentile on host
//   commit_bam     phaser.py:1304                       AS cutoff filter, append to the run-wide tuple store
//   build_graph    phaser.py:1287-1328, 558-581, 610-640, 1265-1285, 667-678, 1594-1642
//                  per-variant lists/sets, noise sums, fragment->variant groups, pair table, edge table
//   phase          phaser.py:686-726, 1861-1887, 2107-2324, 865-931, 1048-1095
//                  edge drop (integer critical value), components, block order, phasing, counts
//   read_lists     phaser.py:1105-1115                  per (block, BAM, haplotype, variant) read lists
#pragma once
#include <type_traits>
#include <cmath>
#include "phz_map_core.h"
#include "phz_phase_core.h"
#include "phz_graph.h"

#ifdef __CUDACC__
#define PHZ_LAMBDA [=] __host__ __device__
#define PHZ_LAMBDA_WARP [=] __host__ __device__
#else
#define PHZ_LAMBDA [=]
#define PHZ_LAMBDA_WARP [=]
#endif

namespace phz {

constexpr u32 NONE32 = 0xFFFFFFFFu;
constexpr u64 NONE64 = 0xFFFFFFFFFFFFFFFFull;
constexpr int AS_BINS = 65536;

inline int ceil_log2_host(u64 x) { int b = 0; while (((u64)1 << b) < x && b < 63) b++; return b < 1 ? 1 : b; }

#ifdef __CUDACC__
// AS histogram with a shared-memory window: alignment scores cluster in a few dozen values, so
// global atomics on them would serialise in L2.
__global__ void __launch_bounds__(256) as_hist_kernel(const u32* __restrict__ t_misc, int64_t n, u64* __restrict__ hist) {
  constexpr int W = 4096, LO = 32768 - 2048;        // window: AS in [-2048, 2047]
  __shared__ u32 sh[W];
  for (int i = threadIdx.x; i < W; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    u32 m = t_misc[i];
    if ((m & 3) == CLS_NONE) continue;
    int bin = (int)(int16_t)(m >> 16) + 32768;
    int w = bin - LO;
    if (w >= 0 && w < W) atomicAdd(&sh[w], 1u); else atomicAdd((unsigned long long*)&hist[bin], 1ull);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < W; i += blockDim.x) if (sh[i]) atomicAdd((unsigned long long*)&hist[LO + i], (unsigned long long)sh[i]);
}
#endif

#ifdef __CUDACC__
// Per-variant counters with a shared-memory window.  Consecutive tuples (record order) and consecutive entries
// (fragment order, fragments numbered by first appearance) belong to one locus, and expression is heavy-tailed: the
// counters of the hottest loci would otherwise take one global reduction per warp and serialise in a single L2 slice
// (ncu: that slice's atomic unit 36-38 % busy, the average slice 2.5 %; profiles/r01_k2_ncu_summary.txt).  A CTA
// therefore counts into a window of AGG_W variant indices around its first item in shared memory and flushes each
// non-zero slot once; items outside the window (sparse regions: no contention there) go straight to global memory.
constexpr int AGG_W = 512;            // variant indices per window
constexpr int AGG_ITEMS = 8;          // items per thread

// first-seen rank + list lengths with duplicates (phaser.py:1310, Q17): vf[v] = min t, nl[v*3+cls] += 1
__global__ void __launch_bounds__(256) variant_lists_kernel(const u32* __restrict__ gv, const u8* __restrict__ gc, int64_t n,
                                                            u32* __restrict__ vf, u32* __restrict__ nl) {
  __shared__ u32 s_cnt[AGG_W * 3];
  __shared__ u32 s_first[AGG_W];
  const int64_t t0 = (int64_t)blockIdx.x * (256 * AGG_ITEMS);
  const u32 v0 = gv[t0];
  const u32 base = v0 > AGG_W / 4 ? v0 - AGG_W / 4 : 0;
  for (int i = threadIdx.x; i < AGG_W * 3; i += 256) s_cnt[i] = 0;
  for (int i = threadIdx.x; i < AGG_W; i += 256) s_first[i] = NONE32;
  __syncthreads();
  for (int k = 0; k < AGG_ITEMS; ++k) {
    const int64_t t = t0 + (int64_t)k * 256 + threadIdx.x;
    if (t >= n) break;
    const u32 v = gv[t]; const u32 cls = gc[t] & 3; const u32 d = v - base;
    if (d < (u32)AGG_W) { atomicAdd(&s_cnt[d * 3 + cls], 1u); atomicMin(&s_first[d], (u32)t); }
    else { atomicAdd(&nl[(int64_t)v * 3 + cls], 1u); atomicMin(&vf[v], (u32)t); }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < AGG_W; i += 256) {
    if (s_first[i] == NONE32) continue;
    const int64_t v = (int64_t)base + i;
    atomicMin(&vf[v], s_first[i]);
    for (int x = 0; x < 3; ++x) if (s_cnt[i * 3 + x]) atomicAdd(&nl[v * 3 + x], s_cnt[i * 3 + x]);
  }
}

// Cell sums of the sorted pair table (see build_graph): every thread owns PAIR_CH consecutive pairs, accumulates its
// runs in registers and flushes one reduction per (run, non-zero cell).  The CTA first stages run ids and cell words
// in shared memory with coalesced loads (row stride PAIR_CH + 1: conflict-free when each thread then walks its own
// row) -- read straight from global memory, 32 lanes x 128-byte stride thrash L1 (ncu: lg_throttle).
constexpr int PAIR_CH = 32;
__global__ void __launch_bounds__(256) pair_cells_kernel(const u32* __restrict__ ps, const u32* __restrict__ pf,
                                                         const u32* __restrict__ pv2, int64_t np, u32* __restrict__ xacc) {
  extern __shared__ u32 sh_pairs[];
  u32* sx = sh_pairs; u32* scell = sh_pairs + 256 * (PAIR_CH + 1);
  const int64_t base = (int64_t)blockIdx.x * (256 * PAIR_CH);
  for (int k = 0; k < PAIR_CH; ++k) {
    const int l = k * 256 + threadIdx.x; const int64_t i = base + l;
    const int slot = (l / PAIR_CH) * (PAIR_CH + 1) + (l % PAIR_CH);
    if (i < np) { sx[slot] = ps[i] + pf[i] - 1; scell[slot] = pv2[i]; }
  }
  __syncthreads();
  const int64_t i0 = base + (int64_t)threadIdx.x * PAIR_CH;
  if (i0 >= np) return;
  const int n = (int)((np - i0) < PAIR_CH ? (np - i0) : PAIR_CH);
  const u32* rx = sx + threadIdx.x * (PAIR_CH + 1); const u32* rc = scell + threadIdx.x * (PAIR_CH + 1);
  u32 acc[10];
#pragma unroll
  for (int k = 0; k < 10; ++k) acc[k] = 0;
  u32 cur = rx[0];
  for (int j = 0; j < n; ++j) {
    const u32 x = rx[j];
    if (x != cur) {
#pragma unroll
      for (int k = 0; k < 10; ++k) if (acc[k]) { atomicAdd(&xacc[(int64_t)cur * 10 + k], acc[k]); acc[k] = 0; }
      cur = x;
    }
    const u32 cells = rc[j];
#pragma unroll
    for (int k = 0; k < 10; ++k) acc[k] += (cells >> k) & 1u;
  }
#pragma unroll
  for (int k = 0; k < 10; ++k) if (acc[k]) atomicAdd(&xacc[(int64_t)cur * 10 + k], acc[k]);
}

// unique-set sizes and per-BAM allele counts, one item per (fragment, variant, bam) entry (see build_graph)
__global__ void __launch_bounds__(256) entry_stats_kernel(const u64* __restrict__ ek, const u32* __restrict__ em, const u8* __restrict__ eb,
                                                          int64_t ne, u64 vmask, u64 excl_mask, int nb,
                                                          u32* __restrict__ sz, u32* __restrict__ vbc) {
  __shared__ u32 s_sz[AGG_W * 3];
  __shared__ u32 s_vb[AGG_W * 2 * 4];   // per-BAM allele counts, used when there are at most 4 BAMs
  const int64_t j0 = (int64_t)blockIdx.x * (256 * AGG_ITEMS);
  const u32 v0 = (u32)(ek[j0] & vmask);
  const u32 base = v0 > AGG_W / 2 ? v0 - AGG_W / 2 : 0;
  for (int i = threadIdx.x; i < AGG_W * 3; i += 256) s_sz[i] = 0;
  for (int i = threadIdx.x; i < AGG_W * 2 * 4; i += 256) s_vb[i] = 0;
  __syncthreads();
  for (int k = 0; k < AGG_ITEMS; ++k) {
    const int64_t j = j0 + (int64_t)k * 256 + threadIdx.x;
    if (j >= ne) break;
    const u64 key = ek[j];
    const u32 v = (u32)(key & vmask); const u32 d = v - base;
    const bool first = (j == 0) || (key != ek[j - 1]);
    u32 mask = 0;
    if (first) for (int64_t jj = j; jj < ne && ek[jj] == key; ++jj) mask |= em[jj];
    const u32 m = em[j]; const u32 bam = eb[j];
    const bool counted = !((excl_mask >> bam) & 1);          // haplo_reads, phaser.py:1320-1322 (Q25)
    const bool in_win = d < (u32)AGG_W;
    for (int x = 0; x < 3; ++x)
      if ((mask >> x) & 1) { if (in_win) atomicAdd(&s_sz[d * 3 + x], 1u); else atomicAdd(&sz[(int64_t)v * 3 + x], 1u); }
    if (counted)
      for (int a = 0; a < 2; ++a)
        if ((m >> a) & 1) {
          if (in_win && nb <= 4) atomicAdd(&s_vb[(d * nb + bam) * 2 + a], 1u);
          else atomicAdd(&vbc[((int64_t)v * nb + bam) * 2 + a], 1u);
        }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < AGG_W; i += 256) {
    const int64_t v = (int64_t)base + i;
    for (int x = 0; x < 3; ++x) if (s_sz[i * 3 + x]) atomicAdd(&sz[v * 3 + x], s_sz[i * 3 + x]);
    if (nb <= 4)
      for (int a = 0; a < nb * 2; ++a) if (s_vb[i * nb * 2 + a]) atomicAdd(&vbc[v * nb * 2 + a], s_vb[i * nb * 2 + a]);
  }
}
#endif

// ----------------------------------------------------------------------------- K1, windowed
// Records are coordinate sorted, so the het sites a tile of consecutive records can touch form a
// short contiguous slab of the position array.  A CTA walks a strip of K1_TILES tiles of K1_THREADS
// records; the slab [wbase, wbase+wn) lives in shared memory (one TMA bulk copy, re-armed only when
// the strip has advanced past the middle of the slab), and the per-segment range searches run there.
constexpr int K1_THREADS = 256;
constexpr int K1_TILES = 8;
constexpr int K1_WIN = 2048;          // het-site positions per slab (8 KB)

struct WindowState { int contig; int64_t wbase; int wn; };

// Decides whether the slab must be (re)loaded for the tile starting at record r0.  `win` is the
// current slab (shared memory on the device).  Returns true when st was changed.
PHZ_HD bool window_for_tile(const ReadsView& rv, const VariantsView& vv, int64_t r0, const int32_t* win, WindowState& st) {
  int c = upper_slot_i64(rv.contig_rec_off, rv.n_contigs, r0);
  int64_t v0 = vv.contig_var_off[c], v1 = vv.contig_var_off[c + 1];
  int32_t p0 = rv.pos[r0];
  if (c == st.contig && st.wn > 0) {
    bool reaches_end = st.wbase + st.wn >= v1;
    if (reaches_end || p0 <= win[st.wn / 2]) return false;      // still in the first half: keep
  }
  int64_t wlo = lower_bound_i32(vv.pos, v0, v1, p0);
  int64_t wbase = (wlo > 0 ? wlo - 1 : 0) & ~(int64_t)3;         // one element of slack, 16-byte aligned source
  int64_t wn = vv.n_variants - wbase; if (wn > K1_WIN) wn = K1_WIN;
  st.contig = c; st.wbase = wbase; st.wn = (int)wn;
  return true;
}

#ifdef __CUDACC__
__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }

template <bool EMIT>
__global__ void __launch_bounds__(K1_THREADS) k1_window_kernel(ReadsView rv, VariantsView vv, int baseq, double isize_cutoff,
                                                             u32* __restrict__ cnt, const u32* __restrict__ off,
                                                             u32* __restrict__ t_rec, u32* __restrict__ t_var,
                                                             u32* __restrict__ t_misc) {
  __shared__ __align__(128) int32_t win[K1_WIN];
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ WindowState st;
  __shared__ int s_bulk;                  // bytes in flight for this tile's reload (0 = no reload)
  const int tid = threadIdx.x;
  const int64_t R = rv.n_records;
  const int64_t strip0 = (int64_t)blockIdx.x * (K1_THREADS * K1_TILES);
  if (tid == 0) {
    st.contig = -1; st.wbase = 0; st.wn = 0; s_bulk = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  u32 parity = 0;
  for (int t = 0; t < K1_TILES; ++t) {
    const int64_t r0 = strip0 + (int64_t)t * K1_THREADS;
    if (r0 >= R) break;
    __syncthreads();                      // everyone is done with the slab of the previous tile
    if (tid == 0) {
      WindowState ns = st;
      int bytes = 0;
      if (window_for_tile(rv, vv, r0, win, ns)) {
        st = ns;
        int nb = ns.wn & ~3;              // elements moved by the bulk copy (16-byte granules)
        bytes = nb * 4;
        if (bytes > 0) {
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(bytes) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(smem_u32(win)), "l"(vv.pos + ns.wbase), "r"(bytes), "r"(smem_u32(&mbar)) : "memory");
        }
        for (int i = nb; i < ns.wn; ++i) win[i] = vv.pos[ns.wbase + i];     // < 4 tail elements
      }
      s_bulk = bytes;
    }
    __syncthreads();
    if (s_bulk > 0) {
      u32 done = 0;
      while (!done) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(smem_u32(&mbar)), "r"(parity) : "memory");
      }
      parity ^= 1;
    }
    const int64_t r = r0 + tid;
    if (r < R) {
      WindowVP vp{vv.pos, win, st.wbase, st.wn, 0, -1};
      int c = st.contig;
      if (r >= rv.contig_rec_off[c + 1]) c = upper_slot_i64(rv.contig_rec_off, rv.n_contigs, r);
      if (!EMIT) {
        cnt[r] = map_record<0>(rv, vv, vp, r, c, baseq, isize_cutoff, 0, nullptr, nullptr, nullptr);
      } else if (cnt[r] != 0) {
        map_record<1>(rv, vv, vp, r, c, baseq, isize_cutoff, off[r], t_rec, t_var, t_misc);
      }
    }
  }
}
#endif

// ----------------------------------------------------------------------------- K1, fused single pass
// One CTA = one tile of KF_THREADS consecutive records.  A tiny pre-pass stores, per tile, where its
// het-site slab starts; the CTA's first thread turns that into one TMA bulk copy while the others are
// already loading their records.  Every thread counts its candidates in the slab, the CTA scans the
// counts, a decoupled look-back over per-tile status words (aggregate / inclusive prefix) yields the
// tile's global output offset, and the threads that have candidates emit them -- in (record, segment,
// variant) order, without a second kernel, without per-record count/offset arrays in HBM.
constexpr int KF_THREADS = 256;
constexpr int KF_WIN = 1024;                      // het-site positions per slab (4 KB)
constexpr int KT_CIG = 1024;                      // CIGAR words staged per tile of the tile kernel (4 KB); tiles with more fall back to global loads
constexpr u64 KF_FLAG_AGG = 1ull << 62, KF_FLAG_PREFIX = 2ull << 62, KF_VALUE_MASK = (1ull << 62) - 1;

// contig << 16 | wn ; hint_lo << 16 | bracket length (0xFFFF = none) ; first CIGAR word of the tile ; the sites of the
// tile's first contig inside the slab are slab[a, b): a | b << 16 ; geo bit 0: slab[a] is the contig's first site, bit 1:
// slab[b - 1] its last, bit 2: the contig has no site at all (tile-uniform geometry the tile kernel would otherwise
// re-derive per thread in 64-bit arithmetic)
// ; rec_cig: records of the tile | CIGAR words staged for it << 16 ; n_first: its records on the first contig
struct alignas(16) TileInfo { u32 wbase; u32 contig_wn; u32 hint; u32 cig0; u32 ab; u32 geo; u32 rec_cig; u32 n_first; };

PHZ_HD TileInfo tile_info_for(const ReadsView& rv, const VariantsView& vv, int64_t r0) {
  int c = upper_slot_i64(rv.contig_rec_off, rv.n_contigs, r0);
  int64_t v0 = vv.contig_var_off[c], v1 = vv.contig_var_off[c + 1];
  int64_t wlo = lower_bound_i32(vv.pos, v0, v1, rv.pos[r0]);
  // one element of slack in front (so an in-slab result equal to the slab start is never ambiguous),
  // rounded down to a 16-byte aligned source for the bulk copy
  int64_t wbase = (wlo > 0 ? wlo - 1 : 0) & ~(int64_t)3;
  int64_t wn = vv.n_variants - wbase; if (wn > KF_WIN) wn = KF_WIN;
  // bracket of lower_bound(POS) over the tile's records that lie on contig c
  int64_t r_last = r0 + KF_THREADS - 1;
  if (r_last >= rv.n_records) r_last = rv.n_records - 1;
  if (r_last >= rv.contig_rec_off[c + 1]) r_last = rv.contig_rec_off[c + 1] - 1;
  int64_t wlast = lower_bound_i32(vv.pos, wlo, v1, rv.pos[r_last]);
  int64_t hint_lo = wlo - wbase, nfirst = wlast - wlo;
  TileInfo ti; ti.wbase = (u32)wbase; ti.contig_wn = ((u32)c << 16) | (u32)wn; ti.cig0 = rv.cigar_off[r0];
  ti.hint = (hint_lo + nfirst <= wn && nfirst < 0xFFFF) ? (((u32)hint_lo << 16) | (u32)nfirst) : 0xFFFFu;
  const int64_t a64 = v0 - wbase, b64 = v1 - wbase;
  const int64_t a = a64 > 0 ? a64 : 0, b = b64 < wn ? b64 : wn;
  ti.ab = (u32)(a < b ? a : 0) | ((u32)(a < b ? b : 0) << 16);
  ti.geo = ((a < b && wbase + a == v0) ? 1u : 0u) | ((a < b && wbase + b == v1) ? 2u : 0u) | (v0 == v1 ? 4u : 0u);
  const int64_t nrec = (rv.n_records - r0) < KF_THREADS ? (rv.n_records - r0) : KF_THREADS;
  const int64_t cig_left = rv.n_cigar_ops - (int64_t)(ti.cig0 & ~3u);
  ti.rec_cig = (u32)nrec | ((u32)(cig_left < KT_CIG ? cig_left : KT_CIG) << 16);
  ti.n_first = (u32)(r_last - r0 + 1);
  return ti;
}

PHZ_HD WindowVP window_of(const TileInfo& ti, const int32_t* g, const int32_t* slab, bool same_contig) {
  WindowVP vp; vp.g = g; vp.s = slab; vp.wbase = (int64_t)ti.wbase; vp.wn = (int)(ti.contig_wn & 0xFFFF);
  u32 n = ti.hint & 0xFFFFu;
  vp.hint_lo = (int)(ti.hint >> 16); vp.hint_hi = (n != 0xFFFFu && same_contig) ? vp.hint_lo + (int)n : -1;
  return vp;
}

#ifdef __CUDACC__
__global__ void __launch_bounds__(KF_THREADS) k1_fused_kernel(ReadsView rv, VariantsView vv, const TileInfo* __restrict__ tiles,
                                                             int baseq, double isize_cutoff, u32* __restrict__ t_rec,
                                                             u32* __restrict__ t_var, u32* __restrict__ t_misc, u64 capacity,
                                                             unsigned long long* tile_status, u32* ticket,
                                                             unsigned long long* total_out, int64_t n_tiles) {
  __shared__ __align__(128) int32_t win[KF_WIN];
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ u32 warp_sum[KF_THREADS / 32];
  __shared__ unsigned long long s_base;
  __shared__ u32 s_tile; __shared__ TileInfo s_ti;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    u32 t = atomicAdd(ticket, 1u);                // tiles are taken in launch order: predecessors are running
    TileInfo ti = tiles[t];
    s_tile = t; s_ti = ti;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    int wn = (int)(ti.contig_wn & 0xFFFF), nb = wn & ~3, bytes = nb * 4;
    if (bytes > 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(win)), "l"(vv.pos + ti.wbase), "r"(bytes), "r"(smem_u32(&mbar)) : "memory");
    }
    for (int i = nb; i < wn; ++i) win[i] = vv.pos[(int64_t)ti.wbase + i];
  }
  __syncthreads();
  const int64_t tile = s_tile;
  const TileInfo ti = s_ti;
  const int wn = (int)(ti.contig_wn & 0xFFFF);
  const int contig0 = (int)(ti.contig_wn >> 16);
  int contig = contig0;
  const int64_t r = tile * KF_THREADS + tid;
  const bool live = r < rv.n_records;
  if (live && r >= rv.contig_rec_off[contig + 1]) contig = upper_slot_i64(rv.contig_rec_off, rv.n_contigs, r);
  if ((wn & ~3) > 0) {
    u32 done = 0;
    while (!done) {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0) : "memory");
    }
  }
  const WindowVP vp = window_of(ti, vv.pos, win, contig == contig0);
  u32 cnt = live ? map_record<0>(rv, vv, vp, r, contig, baseq, isize_cutoff, 0, nullptr, nullptr, nullptr) : 0u;
  // ---- CTA exclusive scan of the counts
  u32 incl = cnt;
  #pragma unroll
  for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  if (lane == 31) warp_sum[warp] = incl;
  __syncthreads();
  u32 warp_base = 0, cta_total = 0;
  #pragma unroll
  for (int w = 0; w < KF_THREADS / 32; ++w) { u32 v = warp_sum[w]; if (w < warp) warp_base += v; cta_total += v; }
  const u32 excl_in_cta = warp_base + incl - cnt;
  // ---- decoupled look-back: global offset of this tile
  if (warp == 0) {
    volatile unsigned long long* st = tile_status;
    if (lane == 0) st[tile] = (tile == 0 ? KF_FLAG_PREFIX : KF_FLAG_AGG) | (unsigned long long)cta_total;
    unsigned long long excl = 0;
    if (tile > 0) {
      int64_t idx = tile - 1;
      while (true) {
        int64_t j = idx - lane;
        unsigned long long sw = KF_FLAG_PREFIX;                  // before tile 0: prefix 0
        if (j >= 0) {
          u32 spins = 0;
          do {
            sw = st[j];
            if (++spins > (1u << 24)) { atomicOr(&ticket[1], 1u); sw = KF_FLAG_PREFIX; }   // never hang the GPU
          } while ((sw >> 62) == 0);
        }
        unsigned pm = __ballot_sync(0xffffffffu, (sw >> 62) == 2);
        unsigned long long val = sw & KF_VALUE_MASK;
        if (pm) { int first = __ffs(pm) - 1; if (lane > first) val = 0; }
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
        excl += val;
        if (pm) break;
        idx -= 32;
      }
      if (lane == 0) st[tile] = KF_FLAG_PREFIX | (excl + cta_total);
    }
    if (lane == 0) {
      s_base = excl;
      if (tile == n_tiles - 1) *total_out = excl + cta_total;
    }
  }
  __syncthreads();
  if (cnt != 0) {
    const u64 o = s_base + excl_in_cta;
    if (o + cnt <= capacity) map_record<1>(rv, vv, vp, r, contig, baseq, isize_cutoff, o, t_rec, t_var, t_misc);
  }
}
#endif

// ----------------------------------------------------------------------------- K1, tile kernel
// Same slab / count / CTA scan as the fused kernel, but (a) a tile takes its output range from ONE
// atomic cursor (no waiting on other tiles) and records (base, count) in a tile table, and (b) the
// emission is DENSE: candidate i of the tile goes to thread i (binary search over the tile's 256
// running offsets in shared memory), so all lanes gather SEQ/QUAL bytes and the stores are fully
// coalesced.  A scan over the tile table then gives the canonical offsets and a streaming permute
// moves each tile's block into (record, segment, variant) order.
// Count pass of ONE record served entirely from the tile's slabs (shared memory on the device), in 32-bit arithmetic: the
// CIGAR words, POS and the het-site slab are all staged, the record lies on the tile's first contig, so the general walk's
// 64-bit offsets, contig look-ups and global fall-backs are not needed.  Same predicate as map_record<0>
// (read_variant_map.py:191-232, 239-242).  Returns false when the slab does not pin a search down (the caller takes the
// general walk).  state_out tells the dense emission where the record's candidates are, so that it does not have to find
// them again: (slab index of the first candidate) | (CIGAR op, counted from the record's first, that opens their segment)
// << 16 when all candidates lie in ONE segment, EMIT_COMPLEX otherwise (the emission re-derives those).
constexpr u32 EMIT_COMPLEX = 0xFFFFFFFFu;
PHZ_HD bool tile_count_fast(const int32_t* __restrict__ s_pos, const u32* __restrict__ s_coff,
                            const u32* __restrict__ s_cig, u32 cig_al, int i, const int32_t* __restrict__ win,
                            int a, int b, bool at_start, bool at_end, int hint_lo, int hint_hi, u32& cnt_out, u32& state_out) {
  const int32_t rpos = s_pos[i];
  const u32 k0 = s_coff[i] - cig_al; u32 k = k0; const u32 kend = s_coff[i + 1] - cig_al;
  u32 cnt = 0, state = EMIT_COMPLEX; int32_t gp = 0; bool first = true;
  while (true) {
    const u32 kseg = k;
    int32_t g_end = gp;
    for (; k < kend; ++k) {
      const u32 c = s_cig[k]; const u32 op = c & 15u;
      if (op == OP_N) break;
      if (op == OP_M || op == OP_D || op == OP_EQ || op == OP_X) g_end += (int32_t)(c >> 4);
    }
    const int32_t seg_len = g_end - gp;
    if (seg_len > 0) {
      const int64_t lo64 = (int64_t)rpos + gp;
      if (lo64 + seg_len > 2147483647LL) return false;
      const int32_t lo_key = (int32_t)lo64, hi_key = lo_key + seg_len;
      int l, len;
      if (first && hint_hi >= 0) { l = hint_lo; len = hint_hi - hint_lo; } else { l = a; len = b - a; }
      while (len > 0) {                          // branch-free halving inside the bracket
        const int half = len >> 1, mid = l + half;
        const bool p = win[mid] < lo_key;
        l = p ? mid + 1 : l;
        len = p ? len - half - 1 : half;
      }
      if (!(first && hint_hi >= 0) && !((l > a || at_start) && (l < b || at_end))) return false;
      int h = l;
      while (h < b && win[h] < hi_key) ++h;
      if (h == b && !at_end) return false;
      if (h > l) {
        state = (cnt == 0 && kseg - k0 < 65536u) ? ((u32)l | ((kseg - k0) << 16)) : EMIT_COMPLEX;
        cnt += (u32)(h - l);
      }
    }
    if (k >= kend) break;
    gp = g_end + (int32_t)(s_cig[k] >> 4); ++k; first = false;
  }
  cnt_out = cnt; state_out = state;
  return true;
}

// Dense emission of ONE candidate whose place the count pass recorded (state != EMIT_COMPLEX): candidate j of record i of
// the tile sits at slab index (state & 0xFFFF) + j, in the segment opened by the record's CIGAR op (state >> 16).  A short
// walk over the ops in front of that segment gives its reference / query offsets and its number; no search.  Same tuple
// as map_record<2> (the tests compare the two ways array by array).
template <class RV>
PHZ_HD void tile_emit_simple(const RV& trv, const int32_t* __restrict__ s_pos, const u32* __restrict__ s_coff,
                             const u32* __restrict__ s_cig, u32 cig_al, const int32_t* __restrict__ win, u32 wbase,
                             const u8* __restrict__ a0, const u8* __restrict__ a1, int64_t r0, int i, u32 j, u32 state, int baseq,
                             u32* t_rec, u32* t_var, u32* t_misc) {
  const u32 c0 = s_coff[i] - cig_al, c1 = s_coff[i + 1] - cig_al, kseg = c0 + (state >> 16);
  int32_t gp = 0, qp = 0; int seg = 0;
  for (u32 x = c0; x < kseg; ++x) {
    const u32 c = s_cig[x]; const u32 op = c & 15u; const int32_t n = (int32_t)(c >> 4);
    if (op == OP_M || op == OP_EQ || op == OP_X) { gp += n; qp += n; }
    else if (op == OP_D) gp += n;
    else if (op == OP_N) { gp += n; ++seg; }
    else if (op == OP_I || op == OP_S) qp += n;
  }
  const u32 widx = (state & 0xFFFFu) + j;
  const int64_t vj = (int64_t)wbase + widx;
  const int32_t st = (int32_t)((int64_t)win[widx] - ((int64_t)s_pos[i] + gp));       // offset in pseudo_read
  const int64_t r = r0 + i;
  *t_rec = (u32)r;
  *t_var = (u32)vj;
  *t_misc = call_snv_site(trv, a0[vj], a1[vj], cig_al + kseg, cig_al + c1, gp, qp, trv.seq_off_at(r), baseq, st, seg, trv.aln_at(r));
}

#ifdef __CUDACC__
constexpr int KT_OWN = 1024;          // candidate slots with a direct owner entry; beyond that a search over the offsets

__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, int bytes, unsigned long long* mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(mbar)) : "memory");
}

template <int MIN_CTAS, class VV>
__global__ void __launch_bounds__(KF_THREADS, MIN_CTAS) k1_tile_kernel(ReadsView rv, VV vv, const TileInfo* __restrict__ tiles,
                                                            int baseq, double isize_cutoff, long long isize_floor, int isize_on,
                                                            u32* __restrict__ s_rec,
                                                            u32* __restrict__ s_var, u32* __restrict__ s_misc, u64 capacity,
                                                            unsigned long long* cursor, u32* __restrict__ tile_base,
                                                            u32* __restrict__ tile_cnt, int k1_staged_emit) {
  __shared__ __align__(128) int32_t win[KF_WIN];
  __shared__ __align__(128) int32_t sh_pos[KF_THREADS];
  __shared__ __align__(128) int32_t sh_tlen[KF_THREADS];
  __shared__ __align__(128) u32 sh_coff[KF_THREADS + 4];
  __shared__ __align__(128) u32 sh_cig[KT_CIG];
  __shared__ __align__(128) u64 sh_soff[KF_THREADS];
  __shared__ __align__(128) int16_t sh_as[KF_THREADS];
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ u32 warp_sum[KF_THREADS / 32];
  __shared__ u32 excl_of[KF_THREADS + 1];
  __shared__ u8 sh_owner[KT_OWN];                     // candidate slot -> record of the tile that owns it
  __shared__ u32 sh_state[KF_THREADS];                // where the count pass found the record's candidates (tile_count_fast)
  __shared__ unsigned long long s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t tile = blockIdx.x;
  const TileInfo ti = tiles[tile];
  const int wn = (int)(ti.contig_wn & 0xFFFF);
  // abs(TLEN) fits 32 unsigned bits: a floor beyond that never gates, a negative one gates everything but is kept exact
  const u32 isize_floor32 = isize_floor >= 0xFFFFFFFFll ? 0xFFFFFFFFu : (isize_floor < 0 ? 0u : (u32)isize_floor);
  if (isize_floor >= 0xFFFFFFFFll) isize_on = 0;
  const int64_t r0 = tile * KF_THREADS;
  const int nrec = (int)(ti.rec_cig & 0xFFFFu);
  // CIGAR slab: words [cig_al, cig_al + cig_n) with a 16-byte aligned start
  const u32 cig_al = ti.cig0 & ~3u;
  const u32 cig_n = ti.rec_cig >> 16;
  const bool gate_all = isize_on && isize_floor < 0;       // abs(TLEN) <= a negative cutoff: never
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // every slab is [multiple of 16 bytes] by TMA + a short tail by plain loads
    const int n4 = nrec & ~3, n2 = nrec & ~1, n8 = nrec & ~7, nw = wn & ~3, nc4 = (int)(cig_n & ~3u);
    const int bytes = nw * 4 + n4 * 4 + n4 * 4 + n4 * 4 + nc4 * 4 + n2 * 8 + n8 * 2;
    if (bytes > 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(bytes) : "memory");
      if (nw) tma_load_1d(win, vv.pos + ti.wbase, nw * 4, &mbar);
      if (n4) {
        tma_load_1d(sh_pos, rv.pos + r0, n4 * 4, &mbar);
        tma_load_1d(sh_tlen, rv.tlen + r0, n4 * 4, &mbar);
        tma_load_1d(sh_coff, rv.cigar_off + r0, n4 * 4, &mbar);
      }
      if (nc4) tma_load_1d(sh_cig, rv.cigar + cig_al, nc4 * 4, &mbar);
      if (n2) tma_load_1d(sh_soff, rv.seq_off + r0, n2 * 8, &mbar);
      if (n8) tma_load_1d(sh_as, rv.aln_score + r0, n8 * 2, &mbar);
    }
    for (int i = nw; i < wn; ++i) win[i] = vv.pos[(int64_t)ti.wbase + i];
    for (int i = n4; i < nrec; ++i) { sh_pos[i] = rv.pos[r0 + i]; sh_tlen[i] = rv.tlen[r0 + i]; sh_coff[i] = rv.cigar_off[r0 + i]; }
    sh_coff[nrec] = rv.cigar_off[r0 + nrec];
    for (u32 i = (u32)nc4; i < cig_n; ++i) sh_cig[i] = rv.cigar[cig_al + i];
    for (int i = n2; i < nrec; ++i) sh_soff[i] = rv.seq_off[r0 + i];
    for (int i = n8; i < nrec; ++i) sh_as[i] = rv.aln_score[r0 + i];
    s_base = (unsigned long long)bytes;
  }
  const int64_t r = r0 + tid;
  const bool live = tid < nrec;
  const int contig0 = (int)(ti.contig_wn >> 16);
  const int n_first = (int)ti.n_first;                 // records [0, n_first) of the tile lie on its first contig
  int contig = contig0;
  if (live && tid >= n_first) contig = upper_slot_i64(rv.contig_rec_off, rv.n_contigs, r);
  __syncthreads();                                   // mbarrier initialised, tail elements visible
  if (s_base > 0) {
    u32 done = 0;
    while (!done) {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0) : "memory");
    }
  }
  // Everything after the slabs have landed, instantiated twice: for all but pathological tiles the whole
  // CIGAR range of the tile is in shared memory and is read unconditionally.
  auto body = [&](auto staged) {
    const TileRV<decltype(staged)::value> trv{r0, sh_pos, sh_tlen, sh_coff, sh_cig, sh_soff, sh_as, cig_al, cig_n, rv.cigar, rv.seq, rv.qual};
    u32 cnt = 0, state = EMIT_COMPLEX;
    if (live) {
      bool done = false;
      if constexpr (!VV::kIndels && decltype(staged)::value) {
        if (contig == contig0) {
          // tile-uniform slab geometry (from the pre-pass): the contig's sites inside the slab are win[a, b)
          const int a = (int)(ti.ab & 0xFFFFu), b = (int)(ti.ab >> 16);
          const int32_t tl = sh_tlen[tid]; const u32 atl = tl < 0 ? 0u - (u32)tl : (u32)tl;
          if (isize_on && (gate_all || atl > isize_floor32)) done = true;        // read_variant_map.py:35,51
          else if (a < b) {
            const u32 hn = ti.hint & 0xFFFFu; const int hlo = (int)(ti.hint >> 16);
            done = tile_count_fast(sh_pos, sh_coff, sh_cig, cig_al, tid, win, a, b, (ti.geo & 1u) != 0, (ti.geo & 2u) != 0, hlo,
                                   hn != 0xFFFFu ? hlo + (int)hn : -1, cnt, state);
            if (!done) state = EMIT_COMPLEX;
          } else if (ti.geo & 4u) done = true;                                    // a contig without het sites
        }
      }
      if (!done) {
        const WindowVP vp = window_of(ti, vv.pos, win, contig == contig0);
        cnt = map_record<0>(trv, vv, vp, r, contig, baseq, isize_cutoff, 0, nullptr, nullptr, nullptr);
      }
    }
    // ---- CTA exclusive scan of the counts
    u32 incl = cnt;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    static_assert(KF_THREADS / 32 == 8, "the warp sums are scanned by three shuffle steps");
    u32 wv = lane < KF_THREADS / 32 ? warp_sum[lane] : 0u, wi = wv;
    #pragma unroll
    for (int o = 1; o < KF_THREADS / 32; o <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
    const u32 warp_base = __shfl_sync(0xffffffffu, wi - wv, warp), cta_total = __shfl_sync(0xffffffffu, wi, KF_THREADS / 32 - 1);
    const u32 my_excl = warp_base + incl - cnt;
    excl_of[tid] = my_excl;
    if (k1_staged_emit == 0) state = EMIT_COMPLEX;
    sh_state[tid] = state;
    for (u32 k = 0; k < cnt && my_excl + k < KT_OWN; ++k) sh_owner[my_excl + k] = (u8)tid;
    if (tid == 0) {
      excl_of[KF_THREADS] = cta_total;
      unsigned long long b = cta_total ? atomicAdd(cursor, (unsigned long long)cta_total) : 0ull;
      s_base = b;
      tile_base[tile] = (u32)b; tile_cnt[tile] = cta_total;
    }
    __syncthreads();
    const u64 base = s_base;
    if (cta_total == 0 || base + cta_total > capacity) return;
    // ---- dense emission: candidate i of the tile -> thread i
    for (u32 i = (u32)tid; i < cta_total; i += KF_THREADS) {
      int lo;
      if (i < KT_OWN) lo = sh_owner[i];
      else {                                            // last record index with excl_of <= i
        lo = 0; int hi = KF_THREADS;
        while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (excl_of[mid] <= i) lo = mid; else hi = mid; }
      }
      const u32 stt = sh_state[lo];
      if constexpr (!VV::kIndels && decltype(staged)::value) {
        if (stt != EMIT_COMPLEX) {          // the count pass left the candidates' place behind: no second search
          tile_emit_simple(trv, sh_pos, sh_coff, sh_cig, cig_al, win, ti.wbase, vv.a0, vv.a1, r0, lo, i - excl_of[lo], stt, baseq,
                           s_rec + base + i, s_var + base + i, s_misc + base + i);
          continue;
        }
      }
      const int64_t rr = r0 + lo;
      int c = contig0;
      if (lo >= n_first) c = upper_slot_i64(rv.contig_rec_off, rv.n_contigs, rr);
      const WindowVP vpk = window_of(ti, vv.pos, win, c == contig0);
      map_record<2>(trv, vv, vpk, rr, c, baseq, isize_cutoff, (u64)(i - excl_of[lo]), s_rec + base + i, s_var + base + i,
                    s_misc + base + i);
    }
  };
  if ((u64)sh_coff[nrec] <= (u64)cig_al + cig_n) body(std::true_type{}); else body(std::false_type{});
}
#endif

template <class B>
struct Pipeline {
  B be;
  ~Pipeline() {
    if (noise_host) B::host_free(noise_host, noise_host_pinned);
    if (noise_event) be.free_event(noise_event);
  }
  bool stats_done = false;
  int k1_min_ctas = 8;        // register budget of the tile kernel: 8 -> 32 regs, 6 -> 40 regs, else unconstrained
  int k1_mode = 3;            // 3: tile kernel + permute (default), 2: fused look-back, 1: windowed two-pass, 0: generic two-pass
  int k1_staged_emit = 1;     // tile kernel: 1 the dense emission takes the candidates' place from the count pass, 0 it re-derives every candidate
  Buf<B, u32> s_rec, s_var, s_misc, tile_base, tile_cnt, tile_canon;
  Buf<B, TileInfo> tile_info; Buf<B, u64> tile_status; Buf<B, u32> k1_ticket;
  // ------------------------------------------------------------------ variants
  int nc = 0; int64_t V = 0; int vbits = 1;
  std::vector<int64_t> h_cvoff;
  Buf<B, int64_t> d_cvoff, d_croff;
  const int32_t* vpos = nullptr; const u8* va0 = nullptr; const u8* va1 = nullptr;
  const int32_t* v_ref_len = nullptr; const u32* v_al_off = nullptr; const u8* v_al_codes = nullptr;   // --include_indels side tables
  const u8* vblack = nullptr;      // per het site: 1 = left out of the haplotypic counts (phaser.py:1070)
  Buf<B, u32> vcontig;
  // ------------------------------------------------------------------ K1 candidates of the current BAM
  Buf<B, u32> cand_cnt, cand_off, t_rec, t_var, t_misc, keep_flag, keep_off;
  int64_t n_cand = 0;
  // ------------------------------------------------------------------ run-wide kept tuples
  Buf<B, u32> g_frag, g_var; Buf<B, u8> g_cb;      // g_cb = class | bam << 2
  int64_t n_tuples = 0; int n_bams = 0;
  // ------------------------------------------------------------------ per-variant
  Buf<B, u32> vfirst, ncls, setsize, vb_cnt, cfirst, crank;
  Buf<B, u64> vrank, noise;
  // ------------------------------------------------------------------ fragment x variant entries
  Buf<B, u64> s_key, s_key2; Buf<B, u32> s_val, s_val2, s_flag, s_scan;
  Buf<B, u64> e_key; Buf<B, u8> e_bam; Buf<B, u32> e_mask, e_tmin, e_flag, e_scan;
  Buf<B, u32> grp_off, pair_cnt, pair_off;
  int64_t NE = 0, NG = 0, NP = 0;
  // ------------------------------------------------------------------ pairs / edges
  Buf<B, u64> p_key, p_key2; Buf<B, u32> p_val, p_val2, p_flag, p_scan, pe_start, p_k32, p_k32b, pair_dmax;
  Buf<B, u32> s_f32, s_f32b, v_member, big_n_d, big_k_d; int64_t n_big_k = 0;
  u32 max_final_len = 0; int two_pass_read_lists = 0;
  int wide_pair_keys = 0;           // 1: always sort the pair table on 64-bit keys (A/B switch)
  Buf<B, u32> x_flag, x_scan, x_acc;
  Buf<B, u32> ed_a, ed_b, ed_sup, ed_tot, ed_n9; Buf<B, u8> ed_cfg, ed_keep;
  Buf<B, u32> scalars, kstar_d;                     // scalars: [0]=max_tot [1]=err flags [2]=dropped [3]=n_big_tot
  Buf<B, u32> big_tot;                              // c_total values above big_total_thr (unordered, with repeats)
  int64_t NX = 0, E = 0; u32 max_tot = 0; int64_t NBT = 0; u32 big_total_thr = 2048;
  int lazy_canonical = 1;           // 1: the tile kernel's output stays in arrival order until somebody needs canonical order
  bool tile_order = false, canonical_valid = true; int64_t tiles_cur = 0;
  int graph_mode = 1;               // 1: fragment-table graph stage (phz_graph.h), 0: sort-based stage (A/B switch and fallback)
  u64* noise_dev_out = nullptr; void* noise_event = nullptr; u64* noise_host = nullptr; bool noise_host_pinned = false, noise_pending = false;      // asynchronous variant_stats
  int frag_stage = 1;               // fragment kernel of the fragment-table stage: 1 slot chunks staged in shared memory, 0 ranges of fragment ids
  u64 n_frag_deferred = 0;          // fragments of the last graph stage that the slot-chunk kernel left to its second pass
  Buf<B, u32> frag_deferred, v_packed, v_rank_in_final;
  bool frag_entries = false;        // which form of the (fragment, variant, BAM) entries the last build_graph left behind
  Buf<B, u32> f_cnt, f_off; Buf<B, uint16_t> t_rank, f_info; Buf<B, u64> f_key, pt_keys, pt_cnt3; Buf<B, u32> pt_vals, pt_flags, pt_slot, rank_flag;
  int64_t n_frag_cur = 0; u64 pair_table_slots = 1ull << 20; int64_t pair_table_grown = 0;
  // the fragment count announced before the commits (option "n_fragments"): the commit then ranks every kept tuple inside
  // its fragment while it writes it, and the graph stage skips its own ranking pass over the tuples
  int64_t n_frag_hint = 0; bool ranks_valid = false; bool graph_ranked_in_commit = false;
  int window_agg = 1;               // 1: shared-memory window kernels for the per-variant counters, 0: warp-aggregated global atomics
  int64_t frag_run_limit = 1024; int64_t n_runs_resorted = 0; bool full_sort_fallback = false;
  // ------------------------------------------------------------------ blocks
  Buf<B, u32> parent, deg, root, m_flag, m_scan, m_list, m_key, m_key2, m_val2, members;
  Buf<B, u32> b_flag, b_scan, blk_off, blk_of, pos_in_blk, blk_contig_rank, blk_order, blk_pos;
  Buf<B, u64> blk_rank, bs_key, bs_key2; Buf<B, u32> bs_val, bs_val2, bs_k32, bs_k32b;
  Buf<B, u64> d_key, d_key2; Buf<B, u32> d_sign, d_sign2, adj_off, adj_list;
  Buf<B, u8> color, v_hap, blk_status; Buf<B, u32> bfsq, run_start, run_len, blk_nfinal, v_fin_local;
  Buf<B, u32> ebk_cnt, ebk_off, ebk_key, ebk_key2, ebk_val, ebk_list;
  Buf<B, u32> h_flag, h_scan, h_list, h_words, h_woff, h_scratch;
  Buf<B, u32> nf_ord, fb_base, fb_first, fb_len, fb_blk, v_final, fb_sup, fb_tot, fb_cnt, fb_bcnt;
  int64_t NM = 0, NB = 0, NH = 0, NF = 0; u32 n_dropped = 0;
  // ------------------------------------------------------------------ read lists
  Buf<B, u32> rl_flag, rl_scan, rl_k32, rl_k32b, rl_t, rl_t2; Buf<B, u64> rl_k64, rl_k64b;
  Buf<B, u32> rl_frag, rl_var, rl_row;
  int64_t NRL = 0;

  Pipeline() {
    B* b = &be;
    d_cvoff.bind(b); d_croff.bind(b); vcontig.bind(b); tile_info.bind(b); tile_status.bind(b); k1_ticket.bind(b);
    s_rec.bind(b); s_var.bind(b); s_misc.bind(b); tile_base.bind(b); tile_cnt.bind(b); tile_canon.bind(b);
    cand_cnt.bind(b); cand_off.bind(b); t_rec.bind(b); t_var.bind(b); t_misc.bind(b); keep_flag.bind(b); keep_off.bind(b);
    g_frag.bind(b); g_var.bind(b); g_cb.bind(b);
    vfirst.bind(b); ncls.bind(b); setsize.bind(b); vb_cnt.bind(b); cfirst.bind(b); crank.bind(b); vrank.bind(b); noise.bind(b);
    s_key.bind(b); s_key2.bind(b); s_val.bind(b); s_val2.bind(b); s_flag.bind(b); s_scan.bind(b);
    e_key.bind(b); e_bam.bind(b); e_mask.bind(b); e_tmin.bind(b); e_flag.bind(b); e_scan.bind(b);
    grp_off.bind(b); pair_cnt.bind(b); pair_off.bind(b);
    p_k32.bind(b); p_k32b.bind(b); pair_dmax.bind(b); s_f32.bind(b); s_f32b.bind(b); v_member.bind(b); big_n_d.bind(b); big_k_d.bind(b);
    p_key.bind(b); p_key2.bind(b); p_val.bind(b); p_val2.bind(b); p_flag.bind(b); p_scan.bind(b); pe_start.bind(b);
    x_flag.bind(b); x_scan.bind(b); x_acc.bind(b);
    ed_a.bind(b); ed_b.bind(b); ed_sup.bind(b); ed_tot.bind(b); ed_n9.bind(b); ed_cfg.bind(b); ed_keep.bind(b); scalars.bind(b); kstar_d.bind(b); big_tot.bind(b);
    parent.bind(b); deg.bind(b); root.bind(b); m_flag.bind(b); m_scan.bind(b); m_list.bind(b); m_key.bind(b); m_key2.bind(b);
    m_val2.bind(b); members.bind(b);
    b_flag.bind(b); b_scan.bind(b); blk_off.bind(b); blk_of.bind(b); pos_in_blk.bind(b); blk_contig_rank.bind(b);
    blk_order.bind(b); blk_pos.bind(b); blk_rank.bind(b); bs_key.bind(b); bs_key2.bind(b); bs_val.bind(b); bs_val2.bind(b);
    bs_k32.bind(b); bs_k32b.bind(b);
    d_key.bind(b); d_key2.bind(b); d_sign.bind(b); d_sign2.bind(b); adj_off.bind(b); adj_list.bind(b);
    color.bind(b); v_hap.bind(b); blk_status.bind(b); bfsq.bind(b); run_start.bind(b); run_len.bind(b); blk_nfinal.bind(b);
    v_fin_local.bind(b);
    ebk_cnt.bind(b); ebk_off.bind(b); ebk_key.bind(b); ebk_key2.bind(b); ebk_val.bind(b); ebk_list.bind(b);
    h_flag.bind(b); h_scan.bind(b); h_list.bind(b); h_words.bind(b); h_woff.bind(b); h_scratch.bind(b);
    nf_ord.bind(b); fb_base.bind(b); fb_first.bind(b); fb_len.bind(b); fb_blk.bind(b); v_final.bind(b); fb_sup.bind(b);
    fb_tot.bind(b); fb_cnt.bind(b); fb_bcnt.bind(b);
    rl_flag.bind(b); rl_scan.bind(b); rl_k32.bind(b); rl_k32b.bind(b); rl_t.bind(b); rl_t2.bind(b); rl_k64.bind(b);
    rl_k64b.bind(b); rl_frag.bind(b); rl_var.bind(b); rl_row.bind(b);
    f_cnt.bind(b); f_off.bind(b); t_rank.bind(b); rank_flag.bind(b); f_info.bind(b); f_key.bind(b); pt_keys.bind(b); pt_cnt3.bind(b); pt_vals.bind(b);
    pt_flags.bind(b); pt_slot.bind(b); frag_deferred.bind(b); v_packed.bind(b); v_rank_in_final.bind(b);
  }

  u32 fetch_u32(const u32* p) { u32 v = 0; be.d2h(&v, p, sizeof(u32)); return v; }

  // =================================================================== variants
  void set_variants(int n_contigs, const int64_t* contig_var_off_host, const int32_t* d_pos, const u8* d_a0,
                    const u8* d_a1, int64_t n_variants) {
    nc = n_contigs; V = n_variants;
    if (V >= (int64_t)0x7FFFFFFF) throw PhzError("too many variants for 32-bit variant ids");
    vbits = ceil_log2_host((u64)(V > 1 ? V : 2));
    h_cvoff.assign(contig_var_off_host, contig_var_off_host + nc + 1);
    be.h2d(d_cvoff.ensure(nc + 1), h_cvoff.data(), (nc + 1) * sizeof(int64_t));
    vpos = d_pos; va0 = d_a0; va1 = d_a1; vblack = nullptr;
    v_ref_len = nullptr; v_al_off = nullptr; v_al_codes = nullptr;
    u32* vc = vcontig.ensure(V);
    const int64_t* off = d_cvoff.p; int n = nc;
    be.for_each(V, PHZ_LAMBDA(int64_t v) { vc[v] = (u32)upper_slot_i64(off, n, v); });
    n_tuples = 0; n_bams = 0; n_cand = 0; ranks_valid = false;
  }

  // =================================================================== K1
  // Returns the number of candidate (record, segment, variant) triples of this BAM.
  int64_t map_reads(ReadsView rv, const int64_t* contig_rec_off_host, int baseq, double isize_cutoff) {
    if (rv.n_contigs != nc) throw PhzError("reads and variants disagree on the number of contigs");
    be.h2d(d_croff.ensure(nc + 1), contig_rec_off_host, (nc + 1) * sizeof(int64_t));
    rv.contig_rec_off = d_croff.p;
    VariantsView vv{V, nc, d_cvoff.p, vpos, va0, va1};
    const int64_t R = rv.n_records;
    if (R >= (int64_t)0x7FFFFFFF) throw PhzError("more than 2^31-1 records in one map_reads call; split the BAM");
    if (v_ref_len) {                    // sites with multi-base alleles: tile kernel, indel instantiation
      VariantsViewIndels vi; static_cast<VariantsView&>(vi) = vv;
      vi.ref_len = v_ref_len; vi.al_off = v_al_off; vi.al_codes = v_al_codes;
      return map_reads_tiles(rv, vi, baseq, isize_cutoff);
    }
    if (k1_mode == 3) return map_reads_tiles(rv, vv, baseq, isize_cutoff);
    tile_order = false; canonical_valid = true; tiles_cur = 0;
    if (k1_mode == 2) return map_reads_fused(rv, vv, baseq, isize_cutoff);
    u32* cnt = cand_cnt.ensure(R + 1);
    u32* off = cand_off.ensure(R + 2);
    be.mark(0);
    run_k1<false>(rv, vv, baseq, isize_cutoff, cnt, off, nullptr, nullptr, nullptr);
    be.mark(1);
    be.exclusive_scan_u32(cnt, off, R);
    n_cand = R > 0 ? (int64_t)fetch_u32(off + R) : 0;
    u32* tr = t_rec.ensure(n_cand); u32* tv = t_var.ensure(n_cand); u32* tm = t_misc.ensure(n_cand);
    be.mark(2);
    run_k1<true>(rv, vv, baseq, isize_cutoff, cnt, off, tr, tv, tm);
    be.mark(3);
    return n_cand;
  }

  // Tile kernel + permute (see k1_tile_kernel).  Scratch capacity is what the buffers already hold (or a
  // first guess); if the exact total turns out larger the buffers grow and the kernel runs once more.
  template <class VV>
  int64_t map_reads_tiles(const ReadsView& rv, const VV& vv, int baseq, double isize_cutoff) {
    const int64_t R = rv.n_records;
    tile_order = false; tiles_cur = 0; canonical_valid = true;
    if (R <= 0) { n_cand = 0; be.mark(0); be.mark(1); be.mark(2); be.mark(3); return 0; }
    const int64_t n_tiles = (R + KF_THREADS - 1) / KF_THREADS;
    TileInfo* ti = tile_info.ensure(n_tiles);
    u32* tb = tile_base.ensure(n_tiles + 1); u32* tc = tile_cnt.ensure(n_tiles + 1); u32* tn = tile_canon.ensure(n_tiles + 2);
    be.mark(0);
    be.for_each(n_tiles, PHZ_LAMBDA(int64_t t) { ti[t] = tile_info_for(rv, vv, t * KF_THREADS); });
    be.mark(1);
    size_t cap = s_rec.cap < s_var.cap ? s_rec.cap : s_var.cap; if (s_misc.cap < cap) cap = s_misc.cap;
    if (cap == 0) cap = (size_t)(R / 2 + (1 << 20));
    u64 total = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
      u32* sr = s_rec.ensure(cap); u32* sv = s_var.ensure(cap); u32* sm = s_misc.ensure(cap);
      total = 0;
      // abs(TLEN) <= cutoff with an integer TLEN  <=>  abs(TLEN) <= floor(cutoff)  (read_variant_map.py:35,51; 0 = no gate)
      const int isz_on = isize_cutoff != 0.0 ? 1 : 0;
      const long long isz_floor = !isz_on ? 0 : (isize_cutoff >= 9.0e18 ? (long long)9e18 : (isize_cutoff <= -9.0e18 ? -(long long)9e18 : (long long)std::floor(isize_cutoff)));
#ifdef __CUDACC__
      u64* cur = tile_status.ensure(2); be.memset0(cur, 2 * sizeof(u64));
      if constexpr (VV::kIndels)      // the indel instantiation is not register-capped: its string walk would spill at 32 registers
        k1_tile_kernel<1, VV><<<(unsigned)n_tiles, KF_THREADS, 0, be.stream>>>(rv, vv, ti, baseq, isize_cutoff, isz_floor, isz_on, sr, sv, sm, (u64)cap, cur, tb, tc, k1_staged_emit);
      else if (k1_min_ctas >= 8)
        k1_tile_kernel<8, VV><<<(unsigned)n_tiles, KF_THREADS, 0, be.stream>>>(rv, vv, ti, baseq, isize_cutoff, isz_floor, isz_on, sr, sv, sm, (u64)cap, cur, tb, tc, k1_staged_emit);
      else if (k1_min_ctas >= 6)
        k1_tile_kernel<6, VV><<<(unsigned)n_tiles, KF_THREADS, 0, be.stream>>>(rv, vv, ti, baseq, isize_cutoff, isz_floor, isz_on, sr, sv, sm, (u64)cap, cur, tb, tc, k1_staged_emit);
      else
        k1_tile_kernel<1, VV><<<(unsigned)n_tiles, KF_THREADS, 0, be.stream>>>(rv, vv, ti, baseq, isize_cutoff, isz_floor, isz_on, sr, sv, sm, (u64)cap, cur, tb, tc, k1_staged_emit);
      PHZ_CUDA(cudaGetLastError());
      be.launches++;
      be.d2h(&total, cur, sizeof(u64));
#else
      // host simulation: same tile descriptors, slab logic, count pass and candidate emission (the slabs are the global
      // arrays themselves), tiles in sequence
      for (int64_t t = 0; t < n_tiles; ++t) {
        const int c0 = (int)(ti[t].contig_wn >> 16);
        const int64_t r0 = t * KF_THREADS;
        const u32 cig_al = ti[t].cig0 & ~3u;
        const int32_t* win = vv.pos + ti[t].wbase;
        const int a = (int)(ti[t].ab & 0xFFFFu), b = (int)(ti[t].ab >> 16);
        const u32 hn = ti[t].hint & 0xFFFFu; const int hlo = (int)(ti[t].hint >> 16);
        u32 tile_total = 0;
        for (int64_t r = r0; r < r0 + KF_THREADS && r < R; ++r) {
          int c = c0;
          if (r >= rv.contig_rec_off[c + 1]) c = upper_slot_i64(rv.contig_rec_off, rv.n_contigs, r);
          const WindowVP vp = window_of(ti[t], vv.pos, win, c == c0);
          u32 cnt = 0, state = EMIT_COMPLEX; bool done = false;
          if constexpr (!VV::kIndels) {
            if (c == c0) {
              const int64_t tl = rv.tlen[r]; const int64_t atl = tl < 0 ? -tl : tl;
              if (isz_on && atl > isz_floor) done = true;
              else if (a < b) {
                done = tile_count_fast(rv.pos + r0, rv.cigar_off + r0, rv.cigar + cig_al, cig_al, (int)(r - r0), win, a, b,
                                       (ti[t].geo & 1u) != 0, (ti[t].geo & 2u) != 0, hlo,
                                       hn != 0xFFFFu ? hlo + (int)hn : -1, cnt, state);
                if (!done) state = EMIT_COMPLEX;
              } else if (ti[t].geo & 4u) done = true;
            }
          }
          if (!done) cnt = map_record<0>(rv, vv, vp, r, c, baseq, isize_cutoff, 0, nullptr, nullptr, nullptr);
          if (k1_staged_emit == 0) state = EMIT_COMPLEX;
          if (total + tile_total + cnt <= cap)
            for (u32 k = 0; k < cnt; ++k) {
              u64 w = total + tile_total + k;
              if constexpr (!VV::kIndels) {
                if (state != EMIT_COMPLEX) {
                  tile_emit_simple(rv, rv.pos + r0, rv.cigar_off + r0, rv.cigar + cig_al, cig_al, win, ti[t].wbase, vv.a0, vv.a1, r0,
                                   (int)(r - r0), k, state, baseq, sr + w, sv + w, sm + w);
                  continue;
                }
              }
              map_record<2>(rv, vv, vp, r, c, baseq, isize_cutoff, k, sr + w, sv + w, sm + w);
            }
          tile_total += cnt;
        }
        tb[t] = (u32)total; tc[t] = tile_total; total += tile_total;
      }
      be.launches++;
#endif
      if (total >= 0xFFFFFFF0ull) throw PhzError("more than 2^32 candidate tuples in one map_reads call; split the BAM");
      if (total <= cap) break;
      cap = (size_t)total;
    }
    n_cand = (int64_t)total;
    be.mark(2);
    // Tile t's block sits at tile_base[t] (arrival order); its canonical place is the running sum of the counts.
    // The stages that follow (AS histogram, commit) read the blocks where they are, through the tile table; the
    // canonical arrays t_rec / t_var / t_misc are only materialised when somebody asks for them (ensure_canonical).
    be.exclusive_scan_u32(tc, tn, n_tiles);
    tiles_cur = n_tiles; tile_order = lazy_canonical != 0; canonical_valid = false;
    if (!tile_order) ensure_canonical();
    be.mark(3);
    return n_cand;
  }

  // canonical (record, segment, variant) order of the candidate tuples of the last tile-kernel run
  void ensure_canonical() {
    if (canonical_valid || tiles_cur == 0) { canonical_valid = true; return; }
    const u32* tb = tile_base.p; const u32* tc = tile_cnt.p; const u32* tn = tile_canon.p;
    u32* tr = t_rec.ensure(n_cand); u32* tv = t_var.ensure(n_cand); u32* tm = t_misc.ensure(n_cand);
    const u32* sr = s_rec.p; const u32* sv = s_var.p; const u32* sm = s_misc.p;
    be.for_each_warp(tiles_cur, PHZ_LAMBDA_WARP(int64_t t, int lane, int nlanes) {
      u32 n = tc[t]; u32 src = tb[t], dst = tn[t];
      for (u32 i = (u32)lane; i < n; i += (u32)nlanes) { tr[dst + i] = sr[src + i]; tv[dst + i] = sv[src + i]; tm[dst + i] = sm[src + i]; }
    });
    canonical_valid = true;
  }

  // Fused single-pass K1.  Output capacity is what the buffers already hold (or a first guess); if the
  // exact total turns out larger the buffers grow and the kernel runs once more.
  int64_t map_reads_fused(const ReadsView& rv, const VariantsView& vv, int baseq, double isize_cutoff) {
    const int64_t R = rv.n_records;
    tile_order = false; tiles_cur = 0; canonical_valid = true;
    if (R <= 0) { n_cand = 0; be.mark(0); be.mark(1); be.mark(2); be.mark(3); return 0; }
    const int64_t n_tiles = (R + KF_THREADS - 1) / KF_THREADS;
    TileInfo* ti = tile_info.ensure(n_tiles);
    be.mark(0);
    be.for_each(n_tiles, PHZ_LAMBDA(int64_t t) { ti[t] = tile_info_for(rv, vv, t * KF_THREADS); });
    be.mark(1); be.mark(2);
    size_t cap = t_rec.cap < t_var.cap ? t_rec.cap : t_var.cap; if (t_misc.cap < cap) cap = t_misc.cap;
    if (cap == 0) cap = (size_t)(R / 2 + (1 << 20));
    for (int attempt = 0; attempt < 2; ++attempt) {
      u32* tr = t_rec.ensure(cap); u32* tv = t_var.ensure(cap); u32* tm = t_misc.ensure(cap);
      u64 total = 0;
#ifdef __CUDACC__
      u64* st = tile_status.ensure(n_tiles + 1); be.memset0(st, (n_tiles + 1) * sizeof(u64));
      u32* tk = k1_ticket.ensure(4); be.memset0(tk, 4 * sizeof(u32));
      k1_fused_kernel<<<(unsigned)n_tiles, KF_THREADS, 0, be.stream>>>(rv, vv, ti, baseq, isize_cutoff, tr, tv, tm, (u64)cap,
                                                                      st, tk, st + n_tiles, n_tiles);
      PHZ_CUDA(cudaGetLastError());
      be.launches++;
      be.mark(3);
      be.d2h(&total, st + n_tiles, sizeof(u64));
      if (fetch_u32(tk + 1) != 0) throw PhzError("K1 look-back timed out (a predecessor tile never published its prefix)");
#else
      // host simulation: same tile descriptors, slab logic and emission order, tiles in sequence
      for (int64_t t = 0; t < n_tiles; ++t) {
        int c0 = (int)(ti[t].contig_wn >> 16);
        for (int64_t r = t * KF_THREADS; r < (t + 1) * KF_THREADS && r < R; ++r) {
          int c = c0;
          if (r >= rv.contig_rec_off[c + 1]) c = upper_slot_i64(rv.contig_rec_off, rv.n_contigs, r);
          const WindowVP vp = window_of(ti[t], vv.pos, vv.pos + ti[t].wbase, c == c0);
          u32 cnt = map_record<0>(rv, vv, vp, r, c, baseq, isize_cutoff, 0, nullptr, nullptr, nullptr);
          if (cnt && total + cnt <= cap) map_record<1>(rv, vv, vp, r, c, baseq, isize_cutoff, total, tr, tv, tm);
          total += cnt;
        }
      }
      be.launches++;
#endif
      if (total >= 0xFFFFFFF0ull) throw PhzError("more than 2^32 candidate tuples in one map_reads call; split the BAM");
      n_cand = (int64_t)total;
      if (total <= cap) break;
      cap = (size_t)total;
    }
    return n_cand;
  }

  template <bool EMIT>
  void run_k1(const ReadsView& rv, const VariantsView& vv, int baseq, double isize_cutoff, u32* cnt, const u32* off,
              u32* tr, u32* tv, u32* tm) {
    const int64_t R = rv.n_records;
    if (R <= 0) return;
    if (k1_mode == 1) {
#ifdef __CUDACC__
      int64_t strips = (R + (int64_t)K1_THREADS * K1_TILES - 1) / ((int64_t)K1_THREADS * K1_TILES);
      k1_window_kernel<EMIT><<<(unsigned)strips, K1_THREADS, 0, be.stream>>>(rv, vv, baseq, isize_cutoff, cnt, off, tr, tv, tm);
      PHZ_CUDA(cudaGetLastError());
      be.launches++;
#else
      // host simulation of the same tiling / slab logic (the slab is a view of the array itself)
      WindowState st{-1, 0, 0};
      const int64_t strip = (int64_t)K1_THREADS * K1_TILES;
      for (int64_t s0 = 0; s0 < R; s0 += strip) {
        st = WindowState{-1, 0, 0};
        for (int t = 0; t < K1_TILES; ++t) {
          int64_t r0 = s0 + (int64_t)t * K1_THREADS;
          if (r0 >= R) break;
          window_for_tile(rv, vv, r0, vv.pos + st.wbase, st);
          for (int64_t r = r0; r < r0 + K1_THREADS && r < R; ++r) {
            WindowVP vp{vv.pos, vv.pos + st.wbase, st.wbase, st.wn, 0, -1};
            int c = st.contig;
            if (r >= rv.contig_rec_off[c + 1]) c = upper_slot_i64(rv.contig_rec_off, rv.n_contigs, r);
            if (!EMIT) cnt[r] = map_record<0>(rv, vv, vp, r, c, baseq, isize_cutoff, 0, nullptr, nullptr, nullptr);
            else if (cnt[r] != 0) map_record<1>(rv, vv, vp, r, c, baseq, isize_cutoff, off[r], tr, tv, tm);
          }
        }
      }
      be.launches++;
#endif
      return;
    }
    int ncg = nc;
    be.for_each(R, PHZ_LAMBDA(int64_t r) {
      if (EMIT && cnt[r] == 0) return;
      int c = upper_slot_i64(rv.contig_rec_off, ncg, r);
      GlobalVP vp{vv.pos};
      if (!EMIT) cnt[r] = map_record<0>(rv, vv, vp, r, c, baseq, isize_cutoff, 0, nullptr, nullptr, nullptr);
      else map_record<1>(rv, vv, vp, r, c, baseq, isize_cutoff, off[r], tr, tv, tm);
    });
  }

  // hist: u64[65536] on the device, bin = AS + 32768, counts the tuples the reference would print
  void as_histogram(u64* hist) {
    be.memset0(hist, AS_BINS * sizeof(u64));
    if (n_cand == 0) return;
    const u32* tm = tile_order ? s_misc.p : t_misc.p; int64_t n = n_cand;      // order does not matter here
#ifdef __CUDACC__
    int blocks = (int)((n + 256 * 16 - 1) / (256 * 16)); if (blocks > 148 * 8) blocks = 148 * 8; if (blocks < 1) blocks = 1;
    as_hist_kernel<<<blocks, 256, 0, be.stream>>>(tm, n, hist);
    PHZ_CUDA(cudaGetLastError());
    be.launches++;
#else
    be.for_each(n, PHZ_LAMBDA(int64_t i) {
      u32 m = tm[i];
      if (misc_cls(m) != CLS_NONE) atomic_add((unsigned long long*)&hist[misc_as(m) + 32768], 1ull);
    });
#endif
  }

  // commit straight from the tile kernel's output: per tile the number of kept tuples, a scan over the tile table
  // (canonical order = tile order), then every warp compacts its own tile to its canonical place in the run-wide
  // store.  Same result as permute + flag scan + scatter over all candidates, without moving the candidates twice.
  int64_t commit_bam_tiles(int bam, int32_t as_cutoff, const u32* frag) {
    const int64_t nt = tiles_cur;
    be.stage("commit_bam");
    const u32* tb = tile_base.p; const u32* tc = tile_cnt.p;
    const u32* sr = s_rec.p; const u32* sv = s_var.p; const u32* sm = s_misc.p;
    u32* kc = keep_flag.ensure(nt + 1); u32* ko = keep_off.ensure(nt + 2);
    be.for_each_warp(nt, PHZ_LAMBDA_WARP(int64_t t, int lane, int nlanes) {
      u32 n = tc[t]; u32 src = tb[t]; u32 k = 0;
      for (u32 i0 = 0; i0 < n; i0 += (u32)nlanes) {
        u32 i = i0 + (u32)lane; bool keep = false;
        if (i < n) { u32 m = sm[src + i]; keep = misc_cls(m) != CLS_NONE && misc_as(m) >= as_cutoff; }
        k += popc_u32(warp_ballot(keep));
      }
      if (lane == 0) kc[t] = k;
    });
    be.exclusive_scan_u32(kc, ko, nt);
    int64_t nk = nt > 0 ? (int64_t)fetch_u32(ko + nt) : 0;
    if (n_tuples + nk >= (int64_t)0xFFFFFFF0ull) throw PhzError("more than 2^32 tuples");
    u32* gf = g_frag.grow(n_tuples + nk, n_tuples); u32* gv = g_var.grow(n_tuples + nk, n_tuples);
    u8* gc = g_cb.grow(n_tuples + nk, n_tuples);
    int64_t base = n_tuples;
    // rank inside the fragment (what the graph stage's counting sort needs), taken while the tuple is written
    const bool rank_here = n_frag_hint > 0 && (bam == 0 || ranks_valid);
    u32* fcn = nullptr; uint16_t* rk = nullptr; u32* rf = nullptr;
    const int64_t F = n_frag_hint;
    if (rank_here) {
      fcn = f_cnt.ensure(F + 1); rk = t_rank.grow(n_tuples + nk, n_tuples); rf = rank_flag.ensure(2);
      if (bam == 0) { be.memset0(fcn, (F + 1) * sizeof(u32)); be.memset0(rf, 2 * sizeof(u32)); }
    }
    ranks_valid = rank_here;
    be.for_each_warp(nt, PHZ_LAMBDA_WARP(int64_t t, int lane, int nlanes) {
      u32 n = tc[t]; u32 src = tb[t]; int64_t o = base + ko[t];
      for (u32 i0 = 0; i0 < n; i0 += (u32)nlanes) {
        u32 i = i0 + (u32)lane; bool keep = false; u32 m = 0;
        if (i < n) { m = sm[src + i]; keep = misc_cls(m) != CLS_NONE && misc_as(m) >= as_cutoff; }
        u32 b = warp_ballot(keep);
        if (keep) {
          int64_t w = o + popc_u32(b & ((1u << lane) - 1u));
          const u32 f = frag[sr[src + i]];
          gf[w] = f; gv[w] = sv[src + i]; gc[w] = (u8)(misc_cls(m) | (bam << 2));
          if (rank_here) {
            if ((int64_t)f < F) {
              const u32 r = atomic_add(&fcn[f], 1u);
              if (r >= 65535u) { atomic_or(&rf[0], 8u); rk[w] = 65535; } else rk[w] = (uint16_t)r;
            } else { atomic_or(&rf[0], 16u); rk[w] = 65535; }         // id beyond the announced count: sort-based graph stage
          }
        }
        o += popc_u32(b);
      }
    });
    n_tuples += nk; n_bams = bam + 1; n_cand = 0; tile_order = false; tiles_cur = 0; canonical_valid = true;
    be.stage("commit_bam.end");
    return nk;
  }

  // keep tuples with a printed allele and AS >= cutoff (phaser.py:1304); cutoff = INT32_MIN disables it
  int64_t commit_bam(int bam, int32_t as_cutoff, const u32* frag) {
    if (bam != n_bams) throw PhzError("commit_bam: BAMs must be committed in order");
    if (bam >= 64) throw PhzError("at most 64 BAMs");
    int64_t n = n_cand;
    if (tile_order) return commit_bam_tiles(bam, as_cutoff, frag);
    u32* ko = keep_off.ensure(n + 2);
    be.stage("commit_bam");
    const u32* tr = t_rec.p; const u32* tv = t_var.p; const u32* tm = t_misc.p;
    // the keep test is one compare on t_misc: evaluated inside the scan and again in the scatter, never stored
    auto keep = PHZ_LAMBDA(int64_t i) -> u32 { u32 m = tm[i]; return (misc_cls(m) != CLS_NONE && misc_as(m) >= as_cutoff) ? 1u : 0u; };
    be.exclusive_scan_fn_u32(keep, ko, n);
    int64_t nk = n > 0 ? (int64_t)fetch_u32(ko + n) : 0;
    if (n_tuples + nk >= (int64_t)0xFFFFFFF0ull) throw PhzError("more than 2^32 tuples");
    u32* gf = g_frag.grow(n_tuples + nk, n_tuples); u32* gv = g_var.grow(n_tuples + nk, n_tuples);
    u8* gc = g_cb.grow(n_tuples + nk, n_tuples);
    int64_t base = n_tuples;
    be.for_each(n, PHZ_LAMBDA(int64_t i) {
      u32 m = tm[i];
      if (!(misc_cls(m) != CLS_NONE && misc_as(m) >= as_cutoff)) return;
      int64_t o = base + ko[i];
      gf[o] = frag[tr[i]]; gv[o] = tv[i]; gc[o] = (u8)(misc_cls(m) | (bam << 2));
    });
    n_tuples += nk; n_bams = bam + 1; n_cand = 0;
    be.stage("commit_bam.end");
    return nk;
  }

  // =================================================================== graph
  // n_frag: number of fragment ids (max id + 1).  excl_mask: bit b set = BAM b excluded from haplotypic counts.
  // Per-variant lists and the two noise sums (phaser.py:610-632).  Split from build_graph so that the host
  // can start on the critical-value table (which only needs the noise level) while the graph is being built.
  void variant_stats(u64* noise_out /*host [2]: match, mismatch*/) {
    const int64_t n = n_tuples; const int64_t Vn = V;
    const u32* gv = g_var.p; const u8* gc = g_cb.p; const u32* vc = vcontig.p;
    stats_done = true;
    be.stage("graph.variant_lists");
    // ---- per-variant lists: first-seen rank (phaser.py:1310), list lengths with duplicates (Q17)
    u32* vf = vfirst.ensure(Vn); be.memset_ff(vf, Vn * sizeof(u32));
    u32* nl = ncls.ensure(Vn * 3); be.memset0(nl, Vn * 3 * sizeof(u32));
#ifdef __CUDACC__
    if (window_agg && n > 0) {
      const int64_t per = 256 * AGG_ITEMS;
      variant_lists_kernel<<<(unsigned)((n + per - 1) / per), 256, 0, be.stream>>>(gv, gc, n, vf, nl);
      PHZ_CUDA(cudaGetLastError());
      be.launches++;
    } else
#endif
    be.for_each(n, PHZ_LAMBDA(int64_t t) {
      u32 v = gv[t]; u32 cls = gc[t] & 3;
#if defined(__CUDA_ARCH__)
      // reads of one locus are neighbours in tuple order: combine the lanes that hit the same variant
      unsigned act = __activemask();
      unsigned peers = __match_any_sync(act, v);
      unsigned b0 = __ballot_sync(act, cls == 0) & peers, b1 = __ballot_sync(act, cls == 1) & peers,
               b2 = __ballot_sync(act, cls == 2) & peers;
      if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) {          // lowest lane = smallest t of the group
        atomic_min(&vf[v], (u32)t);
        if (b0) atomic_add(&nl[(int64_t)v * 3], (u32)__popc(b0));
        if (b1) atomic_add(&nl[(int64_t)v * 3 + 1], (u32)__popc(b1));
        if (b2) atomic_add(&nl[(int64_t)v * 3 + 2], (u32)__popc(b2));
      }
#else
      atomic_min(&vf[v], (u32)t);
      atomic_add(&nl[(int64_t)v * 3 + cls], 1u);
#endif
    });
    u32* cf = cfirst.ensure(nc + 1); be.memset_ff(cf, (nc + 1) * sizeof(u32));
    u64* nz = noise.ensure(2); be.memset0(nz, 2 * sizeof(u64));
    be.for_each(Vn, PHZ_LAMBDA(int64_t v) {
      // (captures are named here, in one order, before the host/device split below: the closure type must not depend on it)
      const u32* vf_ = vf; const u32* nl_ = nl; const u32* vc_ = vc; u32* cf_ = cf; u64* nz_ = nz;
      const bool seen = vf_[v] != NONE32;
      // noise estimate, phaser.py:614-624
      u32 mis = 0, mat = 0;
      if (seen) { mis = nl_[v * 3 + 2]; mat = nl_[v * 3] + nl_[v * 3 + 1]; }
      const bool counts = seen && mat > 0 && ((double)mis / (double)(mis + mat)) < 0.05;
#if defined(__CUDA_ARCH__)
      // all covered sites would otherwise hammer two counters and one slot per contig: reduce inside the warp first
      const unsigned act = __activemask();
      const u32 c = seen ? vc_[v] : 0xFFFFFFFFu;
      const unsigned peers = __match_any_sync(act, c);
      const u32 mn = __reduce_min_sync(peers, seen ? vf_[v] : NONE32);
      if (seen && (int)(threadIdx.x & 31) == __ffs(peers) - 1) atomic_min(&cf_[c], mn);
      const u32 a = counts ? mat : 0u, b = counts ? mis : 0u;          // 16-bit halves: 32 lanes cannot overflow 32 bits
      const u32 alo = __reduce_add_sync(act, a & 0xFFFFu), ahi = __reduce_add_sync(act, a >> 16);
      const u32 blo = __reduce_add_sync(act, b & 0xFFFFu), bhi = __reduce_add_sync(act, b >> 16);
      if ((int)(threadIdx.x & 31) == __ffs(act) - 1) {
        const unsigned long long ta = ((unsigned long long)ahi << 16) + alo, tb = ((unsigned long long)bhi << 16) + blo;
        if (ta) atomic_add((unsigned long long*)&nz_[0], ta);
        if (tb) atomic_add((unsigned long long*)&nz_[1], tb);
      }
#else
      if (seen) atomic_min(&cf_[vc_[v]], vf_[v]);
      if (counts) {
        atomic_add((unsigned long long*)&nz_[0], (unsigned long long)mat);
        atomic_add((unsigned long long*)&nz_[1], (unsigned long long)mis);
      }
#endif
    });
    {   // contig order of first appearance (read_vars key order, phaser.py:573-574)
      u32* cr = crank.ensure(nc + 1); int ncg = nc;
      be.for_each(nc, PHZ_LAMBDA(int64_t c) {
        u32 r = 0;
        for (int o = 0; o < ncg; ++o) if (cf[o] < cf[c] || (cf[o] == cf[c] && o < c)) r++;
        cr[c] = r;
      });
    }
    if (noise_dev_out) {    // no wait, no host copy: the two sums go to the caller's device buffer (a collective sums them over the ranks)
      be.d2d(noise_dev_out, nz, 2 * sizeof(u64));
    } else if (noise_event) {      // no wait here: the two sums travel to a page-locked slot, an event tells when (noise_wait)
      if (!noise_host) { bool pinned = false; noise_host = (u64*)B::host_alloc(2 * sizeof(u64), &pinned); noise_host_pinned = pinned; }
      be.d2h_async(noise_host, nz, 2 * sizeof(u64));
      be.record_event(noise_event);
      noise_pending = true;
    } else be.d2h(noise_out, nz, 2 * sizeof(u64));
    be.stage("graph.variant_lists.end");
  }
  // the (summed) noise counters of a device buffer -> the page-locked slot + event of noise_wait
  void noise_publish(const u64* d_in) {
    if (!noise_host) { bool pinned = false; noise_host = (u64*)B::host_alloc(2 * sizeof(u64), &pinned); noise_host_pinned = pinned; }
    be.d2h_async(noise_host, d_in, 2 * sizeof(u64));
    be.record_event(noise_event);
    noise_pending = true;
  }
  // second half of an asynchronous variant_stats: may be called from another host thread while this one queues the graph stage
  void noise_wait(u64* noise_out) {
    if (!noise_pending) throw PhzError("noise_wait without a pending variant_stats");
    be.host_wait_event(noise_event);
    noise_out[0] = noise_host[0]; noise_out[1] = noise_host[1];
    noise_pending = false;
  }

  // ------------------------------------------------------------------ fragment-table graph stage (phz_graph.h)
  // Returns false when the input does not fit the stage (a fragment with more than 65535 tuples, more than 2^32 - 1
  // fragment ids): the caller then takes the sort-based stage.
  bool build_graph_frag(u64 n_frag, u64 excl_mask) {
    const int64_t n = n_tuples; const int64_t Vn = V; const int nb = n_bams > 0 ? n_bams : 1;
    if (n_frag >= 0xFFFFFFF0ull) return false;
    const int64_t F = (int64_t)(n_frag > 0 ? n_frag : 1);
    const u32* gf = g_frag.p; const u32* gv = g_var.p; const u8* gc = g_cb.p; const u32* vc = vcontig.p;
    be.stage("graph.frag_rank");
    u32* sc = scalars.ensure(8); be.memset0(sc, 8 * sizeof(u32));
    // rank of every tuple inside its fragment (arrival order: the per-fragment sort below makes the result canonical);
    // already taken by the commits when the fragment count was announced to them
    const bool ranked = ranks_valid && n_frag_hint == F && n > 0;
    ranks_valid = false;                  // the fragment kernel reuses the count array: a second call ranks again
    graph_ranked_in_commit = ranked;
    u32* fc = f_cnt.ensure(F + 1);
    u32* fo = f_off.ensure(F + 2);
    uint16_t* rk = ranked ? t_rank.p : t_rank.ensure(n);
    if (ranked) {
      const u32* rf = rank_flag.p;
      be.for_each(1, PHZ_LAMBDA(int64_t) { if (rf[0]) sc[1] |= 8u; });
    } else {
      be.memset0(fc, (F + 1) * sizeof(u32));
      be.for_each(n, PHZ_LAMBDA(int64_t t) {
        const u32 r = atomic_add(&fc[gf[t]], 1u);
        if (r >= 65535u) { atomic_or(&sc[1], 8u); rk[t] = 65535; } else rk[t] = (uint16_t)r;
      });
    }
    be.exclusive_scan_u32(fc, fo, F);
    u64* fk = f_key.ensure(n); uint16_t* fi = f_info.ensure(n);
    u32* sz = setsize.ensure(Vn * 3); u32* vbc = vb_cnt.ensure(Vn * nb * 2); u64* vr = vrank.ensure(Vn);
    u64* c3 = pt_cnt3.ensure(8);
    u32 hsc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int attempt = 0;; ++attempt) {
      const u64 S = pair_table_slots;
      be.stage("graph.frag_scatter");
      be.for_each(n, PHZ_LAMBDA(int64_t t) {
        const u32 r = rk[t]; if (r == 65535u) return;
        const u32 slot = fo[gf[t]] + r;
        fk[slot] = ((u64)gv[t] << 32) | (u64)(u32)t; fi[slot] = (uint16_t)(gc[t] | (r == 0 ? INFO_HEAD : 0));
      });
      be.memset0(sz, Vn * 3 * sizeof(u32)); be.memset0(vbc, Vn * nb * 2 * sizeof(u32)); be.memset_ff(vr, Vn * sizeof(u64));
      u64* pk = pt_keys.ensure(S); u32* pv = pt_vals.ensure(S * PAIR_CELLS); u32* pf = pt_flags.ensure(4);
      be.memset_ff(pk, S * sizeof(u64)); be.memset0(pv, S * PAIR_CELLS * sizeof(u32)); be.memset0(pf, 4 * sizeof(u32));
      be.memset0(c3, 8 * sizeof(u64));
      PairTable pt{pk, pv, (u32)(S - 1), pf};
      FragCtx fx{vc, excl_mask, vr, sc + 1};
      u32* fne = fc;                      // the counts are not needed any more: the slot becomes the entry count
      be.stage("graph.fragments");
#ifdef __CUDACC__
      {
        const int64_t per = (int64_t)FRAG_CTA * FRAG_PER_THREAD;
        if (frag_stage == 1 && n > 0) {
          // slot-chunk form: a CTA owns a chunk of tuple slots, staged in shared memory; fragments it cannot stage go to a
          // side list that one more small pass works off from global memory
          const int64_t slots = nb == 1 ? FRAG_SLOTS : FRAG_SLOTS * 3 / 4;
          const int64_t n_cta = (n + slots - 1) / slots;
          u32* dfl = frag_deferred.ensure(2 * n_cta + 2); u32* dfn = dfl + 2 * n_cta;
          be.memset0(dfn, 2 * sizeof(u32));
          if (nb == 1)
            fragment_slots_kernel<true><<<(unsigned)n_cta, FRAG_CTA, 0, be.stream>>>(fx, n, fk, fi, gf, fo, fne, nb, sz, vbc, pt,
                                                                                   (unsigned long long*)c3, dfl, dfn);
          else
            fragment_slots_kernel<false><<<(unsigned)n_cta, FRAG_CTA, 0, be.stream>>>(fx, n, fk, fi, gf, fo, fne, nb, sz, vbc, pt,
                                                                                    (unsigned long long*)c3, dfl, dfn);
          PHZ_CUDA(cudaGetLastError());
          be.launches++;
          const u32* abortp = sc + 1;
          be.for_each(n_cta, PHZ_LAMBDA(int64_t i) {
            if ((u32)i >= dfn[0] || (*abortp & 8u)) return;
            const u32 o0 = dfl[2 * i], f = dfl[2 * i + 1]; const u32 cnt = fo[f + 1] - fo[f];
            DirectSink sink{sz, vbc, nb, pt};
            u32 ng = 0, np = 0;
            fi[o0] &= 0x7FFFu;
            const u32 ne = nb == 1 ? process_fragment<true>(fx, fk + o0, fi + o0, cnt, sink, ng, np)
                                   : process_fragment<false>(fx, fk + o0, fi + o0, cnt, sink, ng, np);
            fi[o0] |= INFO_HEAD;
            if (ne != cnt) fne[f] = ne;
            atomic_add((unsigned long long*)&c3[0], (unsigned long long)ne); atomic_add((unsigned long long*)&c3[1], (unsigned long long)ng);
            atomic_add((unsigned long long*)&c3[2], (unsigned long long)np); atomic_add((unsigned long long*)&c3[7], 1ull);
          });
        } else
        if (nb == 1)
          fragment_kernel<true><<<(unsigned)((F + per - 1) / per), FRAG_CTA, 0, be.stream>>>(fx, fo, F, fk, fi, fne, nb, sz, vbc, pt,
                                                                                          (unsigned long long*)c3);
        else
          fragment_kernel<false><<<(unsigned)((F + per - 1) / per), FRAG_CTA, 0, be.stream>>>(fx, fo, F, fk, fi, fne, nb, sz, vbc, pt,
                                                                                           (unsigned long long*)c3);
        PHZ_CUDA(cudaGetLastError());
        be.launches++;
      }
#else
      {
        DirectSink sink{sz, vbc, nb, pt};
        u64 ne_sum = 0, ng_sum = 0, np_sum = 0;
        for (int64_t f = 0; f < F && !(sc[1] & 8u); ++f) {
          const u32 o0 = fo[f], cnt = fo[f + 1] - o0; u32 ng = 0, np = 0, ne = 0;
          if (cnt) {
            fi[o0] &= 0x7FFFu;            // the head mark is not part of the tuple's class | bam
            ne = nb == 1 ? process_fragment<true>(fx, fk + o0, fi + o0, cnt, sink, ng, np)
                         : process_fragment<false>(fx, fk + o0, fi + o0, cnt, sink, ng, np);
            fi[o0] |= INFO_HEAD;
          }
          if (ne != cnt) fne[f] = ne;
          ne_sum += ne; ng_sum += ng; np_sum += np;
        }
        c3[0] = ne_sum; c3[1] = ng_sum; c3[2] = np_sum;
        be.launches++;
      }
#endif
      be.stage("graph.edge_table");
      // ---- non-empty slots = distinct pairs; eligible ones (phaser.py:667-678) become edges, sorted by (va, vb)
      u32* xf = x_flag.ensure(S + 1); u32* xs = x_scan.ensure(S + 2);
      be.for_each((int64_t)S, PHZ_LAMBDA(int64_t h) {
        const u64 key = pk[h];
        const bool used = key != PAIR_EMPTY;
        if (used) atomic_add(&pf[1], 1u);
        const bool edge = used && pv[h * PAIR_CELLS + 9];
        xf[h] = edge ? 1u : 0u;
        // the two sites of a pair are neighbours: the largest index difference sizes the sort key of the edge table
        if (edge) { const u32 d = (u32)key - (u32)(key >> 32); if (d > load_volatile(&pf[2])) atomic_max(&pf[2], d); }
      });
      be.exclusive_scan_u32(xf, xs, (int64_t)S);
      // ONE wait for everything the host needs from the stage so far: table-full flag, distinct pairs, edges, the ranking
      // pass's verdict, and the entry / group / pair totals
      { const u32* xs_c = xs; int64_t ss = (int64_t)S;
        be.for_each(1, PHZ_LAMBDA(int64_t) { c3[3] = pf[0]; c3[4] = pf[1]; c3[5] = xs_c[ss]; c3[6] = sc[1] | ((u64)pf[2] << 32); }); }
      u64 h7[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      be.d2h(h7, c3, sizeof(h7));
      n_frag_deferred = h7[7];
      const int dbits = ceil_log2_host((h7[6] >> 32) + 1) > 0 ? ceil_log2_host((h7[6] >> 32) + 1) : 1;
      if (h7[6] & 8u) return false;            // a fragment beyond the 16-bit rank: sort-based stage
      if (h7[3] & 1u) {                        // pair table full: grow it and run the fragments again
        if (attempt >= 8) throw PhzError("pair table keeps overflowing");
        pair_table_slots *= 4; pair_table_grown++;
        continue;
      }
      NX = (int64_t)h7[4]; E = (int64_t)h7[5];
      if ((u64)NX * 2 > S) { pair_table_slots *= 2; }     // keep the load factor low for the next sample (no re-run needed now)
      NE = (int64_t)h7[0]; NG = (int64_t)h7[1]; NP = (int64_t)h7[2];
      // ---- edge table
      u64* ek = d_key.ensure(E); u64* ek2 = d_key2.ensure(E); u32* es = pt_slot.ensure(2 * E + 2); u32* es2 = es + E + 1;
      // sort key: first site << dbits | index difference (vbits + dbits bits: four radix passes instead of seven)
      const u64 dmask = (((u64)1) << dbits) - 1;
      be.for_each((int64_t)S, PHZ_LAMBDA(int64_t h) {
        if (xf[h]) { const u64 key = pk[h]; ek[xs[h]] = ((key >> 32) << dbits) | (u64)((u32)key - (u32)(key >> 32)); es[xs[h]] = (u32)h; }
      });
      be.sort_pairs(ek, ek2, es, es2, E, 0, vbits + dbits);
      u32* ea_ = ed_a.ensure(E); u32* eb_ = ed_b.ensure(E); u32* esup = ed_sup.ensure(E); u32* etot = ed_tot.ensure(E);
      u32* en9 = ed_n9.ensure(E * 9); u8* ecfg = ed_cfg.ensure(E); ed_keep.ensure(E);
      be.memset0(sc, 8 * sizeof(u32));
      u32* bt = big_tot.ensure(E); const u32 bthr = big_total_thr;
      be.for_each(E, PHZ_LAMBDA(int64_t e) {
        const u64 key = ek2[e]; const u32 h = es2[e];
        u32 n9[9]; for (int c = 0; c < 9; ++c) n9[c] = pv[(int64_t)h * PAIR_CELLS + c];
        ea_[e] = (u32)(key >> dbits); eb_[e] = (u32)(key >> dbits) + (u32)(key & dmask);
        for (int c = 0; c < 9; ++c) en9[e * 9 + c] = n9[c];
        u32 cis = n9[0] + n9[4], trans = n9[3] + n9[1];          // n[x][y] at x*3+y
        u32 other = n9[6] + n9[7] + n9[2] + n9[5] + n9[8];
        u32 sup = cis > trans ? cis : trans, tot = cis + trans + other;
        esup[e] = sup; etot[e] = tot;
        ecfg[e] = (u8)(cis > trans ? EDGE_CIS : (cis < trans ? EDGE_TRANS : EDGE_TIE));
        if (tot > load_volatile(&sc[0])) atomic_max(&sc[0], tot);
        if (tot > bthr) bt[atomic_add(&sc[3], 1u)] = tot;
      });
      break;
    }
    if (E > 0) be.d2h(hsc, sc, 4 * sizeof(u32));
    max_tot = hsc[0]; NBT = hsc[3];
    n_frag_cur = F; frag_entries = true; n_runs_resorted = 0; full_sort_fallback = false;
    be.stage("graph.end");
    return true;
  }

  void build_graph(u64 n_frag, u64 excl_mask) {
    if (!stats_done) { u64 tmp[2]; variant_stats(tmp); }
    stats_done = false;
    graph_ranked_in_commit = false;
    if (graph_mode == 1 && build_graph_frag(n_frag, excl_mask)) return;
    ranks_valid = false;
    frag_entries = false;
    const int64_t n = n_tuples; const int64_t Vn = V; const int nb = n_bams > 0 ? n_bams : 1;
    const u32* gf = g_frag.p; const u32* gv = g_var.p; const u8* gc = g_cb.p; const u32* vc = vcontig.p;
    be.stage("graph.sort_tuples");
    // ---- (fragment, variant, bam) entries: sort tuples by (fragment, variant); t ascending inside
    const int fb = ceil_log2_host(n_frag > 1 ? n_frag : 2); const int vb = vbits;
    if (fb + vb > 64) throw PhzError("fragment/variant id space too large");
    u64* k2 = s_key2.ensure(n); u32* x1 = s_val.ensure(n); u32* x2 = s_val2.ensure(n);
    u32* f1 = s_f32.ensure(n); u32* f2 = s_f32b.ensure(n);
    be.for_each(n, PHZ_LAMBDA(int64_t t) { f1[t] = gf[t]; x1[t] = (u32)t; });
    // Tuples arrive in (record, variant) order, so a STABLE sort on the fragment bits alone (4 radix passes instead
    // of 6) already leaves a fragment's tuples variant-sorted unless its records interleave (overlapping mates, a
    // spliced mate jumping over the other).  The first tuple of each fragment checks its run and insertion-sorts it
    // in place when needed (stable: equal keys keep tuple order).  Runs too long for that raise a flag and the
    // whole array is sorted on the full key instead.
    {
      u32* sc = scalars.ensure(8); be.memset0(sc, 8 * sizeof(u32));
      // (32-bit fragment keys: 8-byte pairs through the radix passes; the 64-bit (fragment, variant) keys the
      // following stages compare are assembled afterwards -- the sorted tuple indices are nearly ascending, so the
      // gather of the variant ids stays in cache)
      be.sort_pairs32(f1, f2, x1, x2, n, 0, fb);
      be.for_each(n, PHZ_LAMBDA(int64_t i) { k2[i] = ((u64)f2[i] << vb) | (u64)gv[x2[i]]; });
      const int64_t nn = n; const int64_t MAXRUN = frag_run_limit;
      be.for_each(n, PHZ_LAMBDA(int64_t i) {
        const u64 f = k2[i] >> vb;
        if (i > 0 && (load_volatile(&k2[i - 1]) >> vb) == f) return;           // not the first tuple of its fragment
        int64_t end = i + 1; bool sorted = true;
        while (end < nn && (load_volatile(&k2[end]) >> vb) == f) {
          if (end - i >= MAXRUN) { atomic_max(&sc[4], 1u); return; }
          if (k2[end] < k2[end - 1]) sorted = false;
          ++end;
        }
        if (sorted) return;
        atomic_add(&sc[5], 1u);
        for (int64_t j = i + 1; j < end; ++j) {
          u64 key = k2[j]; u32 val = x2[j]; int64_t m = j;
          while (m > i && k2[m - 1] > key) { k2[m] = k2[m - 1]; x2[m] = x2[m - 1]; --m; }
          k2[m] = key; x2[m] = val;
        }
      });
      u32 h2[2] = {0, 0};
      if (n > 0) be.d2h(h2, sc + 4, sizeof(h2));
      n_runs_resorted = h2[1]; full_sort_fallback = h2[0] != 0;
      if (full_sort_fallback) {
        u64* k1 = s_key.ensure(n);
        be.for_each(n, PHZ_LAMBDA(int64_t t) { k1[t] = ((u64)gf[t] << vb) | (u64)gv[t]; x1[t] = (u32)t; });
        be.sort_pairs(k1, k2, x1, x2, n, 0, fb + vb);
      }
    }
    be.stage("graph.entries");
    u32* sf = s_flag.ensure(n + 1); u32* ss = s_scan.ensure(n + 2);
    be.for_each(n, PHZ_LAMBDA(int64_t i) {
      sf[i] = (i == 0 || k2[i] != k2[i - 1] || (gc[x2[i]] >> 2) != (gc[x2[i - 1]] >> 2)) ? 1u : 0u;
    });
    be.exclusive_scan_u32(sf, ss, n);
    NE = n > 0 ? (int64_t)fetch_u32(ss + n) : 0;
    u64* ek = e_key.ensure(NE); u8* eb = e_bam.ensure(NE); u32* em = e_mask.ensure(NE); u32* et = e_tmin.ensure(NE);
    be.memset0(em, NE * sizeof(u32)); be.memset_ff(et, NE * sizeof(u32));
    be.for_each(n, PHZ_LAMBDA(int64_t i) {
      u32 e = ss[i] + sf[i] - 1; u32 t = x2[i]; u32 cls = gc[t] & 3;
      if (sf[i]) { ek[e] = k2[i]; eb[e] = gc[t] >> 2; }
      atomic_or(&em[e], 1u << cls);
      if (cls < 2) atomic_min(&et[e], t);
    });
    be.stage("graph.groups");
    // ---- groups = (fragment, contig) runs of entries
    u32* ef = e_flag.ensure(NE + 1); u32* es = e_scan.ensure(NE + 2);
    const u64 vmask = (((u64)1) << vb) - 1;
    be.for_each(NE, PHZ_LAMBDA(int64_t j) {
      bool head = (j == 0) || ((ek[j] >> vb) != (ek[j - 1] >> vb)) || (vc[ek[j] & vmask] != vc[ek[j - 1] & vmask]);
      ef[j] = head ? 1u : 0u;
    });
    be.exclusive_scan_u32(ef, es, NE);
    NG = NE > 0 ? (int64_t)fetch_u32(es + NE) : 0;
    u32* go = grp_off.ensure(NG + 1);
    { int64_t ne = NE, ng = NG;
      be.for_each(NE + 1, PHZ_LAMBDA(int64_t j) { if (j == ne) go[ng] = (u32)ne; else if (ef[j]) go[es[j]] = (u32)j; }); }
    be.stage("graph.group_stats");
    // ---- per group: sets, per-BAM allele counts, overlap rank, pair count
    u32* sz = setsize.ensure(Vn * 3); be.memset0(sz, Vn * 3 * sizeof(u32));
    u32* vbc = vb_cnt.ensure(Vn * nb * 2); be.memset0(vbc, Vn * nb * 2 * sizeof(u32));
    u64* vr = vrank.ensure(Vn); be.memset_ff(vr, Vn * sizeof(u64));
    u32* pc = pair_cnt.ensure(NG + 1); u32* po = pair_off.ensure(NG + 2);
    // (a) one logical thread per entry: unique-set sizes (at the first entry of a (fragment, variant) run) and
    //     per-BAM allele counts, warp-aggregated because neighbouring entries sit on the same locus
#ifdef __CUDACC__
    if (window_agg && NE > 0) {
      const int64_t per = 256 * AGG_ITEMS;
      entry_stats_kernel<<<(unsigned)((NE + per - 1) / per), 256, 0, be.stream>>>(ek, em, eb, NE, vmask, excl_mask, nb, sz, vbc);
      PHZ_CUDA(cudaGetLastError());
      be.launches++;
    } else
#endif
    { int64_t ne = NE;
      be.for_each(NE, PHZ_LAMBDA(int64_t j) {
        u32 v = (u32)(ek[j] & vmask);
        bool first = (j == 0) || (ek[j] != ek[j - 1]);
        u32 mask = 0;
        if (first) for (int64_t jj = j; jj < ne && ek[jj] == ek[j]; ++jj) mask |= em[jj];
        bool counted = !((excl_mask >> eb[j]) & 1);          // haplo_reads, phaser.py:1320-1322 (Q25)
        u32 kb = (v * (u32)nb + eb[j]) * 2;
#if defined(__CUDA_ARCH__)
        // one match per key for the five counters of this entry (every lane of the launch gets here: no divergence above)
        const unsigned act = __activemask();
        const unsigned pv = __match_any_sync(act, v);
        const unsigned pk = nb == 1 ? pv : __match_any_sync(act, kb);
        const bool lead_v = (int)(threadIdx.x & 31) == __ffs(pv) - 1, lead_k = (int)(threadIdx.x & 31) == __ffs(pk) - 1;
        for (int x = 0; x < 3; ++x) {
          const unsigned b = __ballot_sync(act, first && ((mask >> x) & 1)) & pv;
          if (lead_v && b) atomicAdd(&sz[v * 3 + (u32)x], (u32)__popc(b));
        }
        const unsigned b0 = __ballot_sync(act, counted && (em[j] & 1)) & pk, b1 = __ballot_sync(act, counted && (em[j] & 2)) & pk;
        if (lead_k && b0) atomicAdd(&vbc[kb], (u32)__popc(b0));
        if (lead_k && b1) atomicAdd(&vbc[kb + 1], (u32)__popc(b1));
#else
        for (int x = 0; x < 3; ++x) warp_agg_inc(sz, v * 3 + (u32)x, first && ((mask >> x) & 1));
        warp_agg_inc(vbc, kb, counted && (em[j] & 1));
        warp_agg_inc(vbc, kb + 1, counted && (em[j] & 2));
#endif
      }); }
    // (b) one logical thread per (fragment, contig) group: effective BAM, overlap rank, pair count
    u32* dm = pair_dmax.ensure(4); be.memset0(dm, 4 * sizeof(u32));
    be.for_each(NG, PHZ_LAMBDA(int64_t g) {
      u32 j0 = go[g], j1 = go[g + 1];
      if (j1 - j0 < 2) { pc[g] = 0; return; }          // one entry = one variant: no pair, no overlap rank (most groups)
      int effbam = -1; u32 first_t = NONE32;
      for (u32 j = j0; j < j1; ++j) if (em[j] & 3) { if ((int)eb[j] > effbam) effbam = eb[j]; if (et[j] < first_t) first_t = et[j]; }
      u32 k = 0, kelig = 0;
      for (u32 j = j0; j < j1;) {
        u32 v = (u32)(ek[j] & vmask); bool elig = false; u32 jj = j;
        for (; jj < j1 && (u32)(ek[jj] & vmask) == v; ++jj)
          if ((em[jj] & 3) && (int)eb[jj] == effbam) elig = true;
        k++; if (elig) kelig++;
        j = jj;
      }
      pc[g] = k * (k - 1) / 2;
      if (k >= 2) {         // widest variant-index distance inside a fragment: sizes the pair keys below
        u32 d = (u32)(ek[j1 - 1] & vmask) - (u32)(ek[j0] & vmask);
        if (d > load_volatile(&dm[0])) atomic_max(&dm[0], d);
      }
      if (kelig >= 2) {     // insertion order of dict_variant_overlap, phaser.py:1271-1283
        for (u32 j = j0; j < j1; ++j)
          if ((em[j] & 3) && (int)eb[j] == effbam) {
            unsigned long long key = ((unsigned long long)first_t << 32) | et[j];
            unsigned long long* slot = (unsigned long long*)&vr[ek[j] & vmask];
            if (key < *(volatile unsigned long long*)slot) atomic_min(slot, key);     // the minimum settles early
          }
      }
    });
    be.exclusive_scan_u32(pc, po, NG);
    NP = NG > 0 ? (int64_t)fetch_u32(po + NG) : 0;
    be.stage("graph.emit_pairs");
    // ---- pairs: key (va, vb), value = 9 co-occurrence cells + eligibility.  The key is (va << dbits) | (vb - va)
    // with dbits sized by the widest distance seen above: same order as (va, vb), but usually <= 32 bits, i.e. 4
    // radix passes over 8-byte pairs instead of 6 over 12-byte ones.
    const u32 dmax = NP > 0 ? fetch_u32(dm) : 0;
    const int dbits = ceil_log2_host((u64)dmax + 1) > 0 ? ceil_log2_host((u64)dmax + 1) : 1;
    const int kbits = vb + dbits;
    const bool k32 = kbits <= 32 && !wide_pair_keys;
    u64* pk = p_key.ensure(k32 ? 1 : NP); u64* pk2 = p_key2.ensure(k32 ? 1 : NP);
    u32* pk32 = p_k32.ensure(k32 ? NP : 1); u32* pk32b = p_k32b.ensure(k32 ? NP : 1);
    u32* pv = p_val.ensure(NP); u32* pv2 = p_val2.ensure(NP);
    const u64 dmask = (((u64)1) << dbits) - 1;
    be.for_each(NG, PHZ_LAMBDA(int64_t g) {
      if (pc[g] == 0) return;
      u32 j0 = go[g], j1 = go[g + 1];
      int effbam = -1;
      for (u32 j = j0; j < j1; ++j) if ((em[j] & 3) && (int)eb[j] > effbam) effbam = eb[j];
      u64 o = po[g];
      for (u32 a = j0; a < j1;) {
        u32 va = (u32)(ek[a] & vmask); u32 ma = 0; bool ea = false; u32 a1 = a;
        for (; a1 < j1 && (u32)(ek[a1] & vmask) == va; ++a1) { ma |= em[a1]; if ((em[a1] & 3) && (int)eb[a1] == effbam) ea = true; }
        for (u32 b = a1; b < j1;) {
          u32 vbb = (u32)(ek[b] & vmask); u32 mb = 0; bool ebb = false; u32 b1 = b;
          for (; b1 < j1 && (u32)(ek[b1] & vmask) == vbb; ++b1) { mb |= em[b1]; if ((em[b1] & 3) && (int)eb[b1] == effbam) ebb = true; }
          u32 cells = 0;
          for (int x = 0; x < 3; ++x) for (int y = 0; y < 3; ++y) if (((ma >> x) & 1) && ((mb >> y) & 1)) cells |= 1u << (x * 3 + y);
          if (ea && ebb) cells |= 1u << 9;
          if (k32) pk32[o] = (va << dbits) | (vbb - va); else pk[o] = ((u64)va << dbits) | (u64)(vbb - va);
          pv[o] = cells; o++;
          b = b1;
        }
        a = a1;
      }
    });
    be.stage("graph.sort_pairs");
    if (k32) be.sort_pairs32(pk32, pk32b, pv, pv2, NP, 0, kbits); else be.sort_pairs(pk, pk2, pv, pv2, NP, 0, kbits);
    be.stage("graph.edge_table");
    u32* pf = p_flag.ensure(NP + 1); u32* ps = p_scan.ensure(NP + 2);
    be.for_each(NP, PHZ_LAMBDA(int64_t i) {
      pf[i] = (i == 0 || (k32 ? pk32b[i] != pk32b[i - 1] : pk2[i] != pk2[i - 1])) ? 1u : 0u;
    });
    be.exclusive_scan_u32(pf, ps, NP);
    NX = NP > 0 ? (int64_t)fetch_u32(ps + NP) : 0;
    u32* pst = pe_start.ensure(NX + 1);
    { int64_t np = NP, nx = NX;
      be.for_each(NP + 1, PHZ_LAMBDA(int64_t i) { if (i == np) pst[nx] = (u32)np; else if (pf[i]) pst[ps[i]] = (u32)i; }); }
    // ---- distinct pairs: 9 cell sums + eligibility.  One logical thread per fixed chunk of the sorted
    // pair array accumulates its runs locally and flushes one atomic per (run, non-zero cell), so a pair
    // supported by a million fragments costs the same per thread as any other.
    u32* xacc = x_acc.ensure((NX + 1) * 10); be.memset0(xacc, (NX + 1) * 10 * sizeof(u32));
#ifdef __CUDACC__
    if (window_agg && NP > 0) {
      const size_t smem = 2 * 256 * (PAIR_CH + 1) * sizeof(u32);
      static bool attr_set = false;
      if (!attr_set) { PHZ_CUDA(cudaFuncSetAttribute(pair_cells_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_set = true; }
      const int64_t per = 256 * PAIR_CH;
      pair_cells_kernel<<<(unsigned)((NP + per - 1) / per), 256, smem, be.stream>>>(ps, pf, pv2, NP, xacc);
      PHZ_CUDA(cudaGetLastError());
      be.launches++;
    } else
#endif
    {
      const int64_t CH = 32; int64_t np = NP; int64_t nchunks = (NP + CH - 1) / CH;
      be.for_each(nchunks, PHZ_LAMBDA(int64_t c) {
        int64_t i0 = c * CH, i1 = i0 + CH; if (i1 > np) i1 = np;
        u32 acc[10]; for (int k = 0; k < 10; ++k) acc[k] = 0;
        u32 cur = ps[i0] + pf[i0] - 1;
        for (int64_t i = i0; i < i1; ++i) {
          u32 x = ps[i] + pf[i] - 1;
          if (x != cur) {
            for (int k = 0; k < 10; ++k) if (acc[k]) { atomic_add(&xacc[(int64_t)cur * 10 + k], acc[k]); acc[k] = 0; }
            cur = x;
          }
          u32 cells = pv2[i];
          for (int k = 0; k < 10; ++k) acc[k] += (cells >> k) & 1u;
        }
        for (int k = 0; k < 10; ++k) if (acc[k]) atomic_add(&xacc[(int64_t)cur * 10 + k], acc[k]);
      });
    }
    // ---- eligible pairs -> edge table (phaser.py:667-678, 1594-1642)
    u32* xf = x_flag.ensure(NX + 1); u32* xs = x_scan.ensure(NX + 2);
    be.for_each(NX, PHZ_LAMBDA(int64_t x) { xf[x] = xacc[x * 10 + 9] ? 1u : 0u; });
    be.exclusive_scan_u32(xf, xs, NX);
    E = NX > 0 ? (int64_t)fetch_u32(xs + NX) : 0;
    u32* ea_ = ed_a.ensure(E); u32* eb_ = ed_b.ensure(E); u32* esup = ed_sup.ensure(E); u32* etot = ed_tot.ensure(E);
    u32* en9 = ed_n9.ensure(E * 9); u8* ecfg = ed_cfg.ensure(E); ed_keep.ensure(E);
    u32* sc = scalars.ensure(8); be.memset0(sc, 8 * sizeof(u32));
    // totals above the threshold go to a side list: the host needs a critical value for each distinct one
    u32* bt = big_tot.ensure(E); const u32 bthr = big_total_thr;
    be.for_each(NX, PHZ_LAMBDA(int64_t x) {
      if (!xf[x]) return;
      u32 e = xs[x];
      u32 n9[9]; for (int c = 0; c < 9; ++c) n9[c] = xacc[x * 10 + c];
      u64 key = k32 ? (u64)pk32b[pst[x]] : pk2[pst[x]];
      ea_[e] = (u32)(key >> dbits); eb_[e] = (u32)(key >> dbits) + (u32)(key & dmask);
      for (int c = 0; c < 9; ++c) en9[(int64_t)e * 9 + c] = n9[c];
      u32 cis = n9[0] + n9[4], trans = n9[3] + n9[1];          // n[x][y] at x*3+y
      u32 other = n9[6] + n9[7] + n9[2] + n9[5] + n9[8];
      u32 sup = cis > trans ? cis : trans, tot = cis + trans + other;
      esup[e] = sup; etot[e] = tot;
      ecfg[e] = (u8)(cis > trans ? EDGE_CIS : (cis < trans ? EDGE_TRANS : EDGE_TIE));
      if (tot > load_volatile(&sc[0])) atomic_max(&sc[0], tot);
      if (tot > bthr) bt[atomic_add(&sc[3], 1u)] = tot;
    });
    u32 hsc[4] = {0, 0, 0, 0};
    if (E > 0) be.d2h(hsc, sc, sizeof(hsc));
    max_tot = hsc[0]; NBT = hsc[3];
    be.stage("graph.end");
  }

  // =================================================================== drop, blocks, phasing, counts
  // kstar[n] (host, n in [0, max_tot]): smallest k with binom.cdf(k, n, p) >= cc_threshold; an edge is
  // dropped iff c_supporting == 0 or (c_total > c_supporting and c_supporting < kstar[c_total])
  // (phaser.py:1645-1652, 696).
  // Totals beyond the dense table are looked up in the sparse list set by set_big_critical_values (ascending n).
  void set_big_critical_values(const u32* n_host, const u32* k_host, int64_t count) {
    n_big_k = count;
    u32* bn = big_n_d.ensure(count + 1); u32* bk = big_k_d.ensure(count + 1);
    be.h2d(bn, n_host, count * sizeof(u32)); be.h2d(bk, k_host, count * sizeof(u32));
  }

  void phase(const u32* kstar_host, int64_t kstar_len, int max_block_size, u64 excl_mask, int* err_out) {
    if ((int64_t)max_tot >= kstar_len && E > 0 && n_big_k == 0) throw PhzError("critical-value table shorter than max c_total");
    const int64_t Vn = V; const int nb = n_bams > 0 ? n_bams : 1; const int vb = vbits;
    const u32* vc = vcontig.p;
    be.stage("phase.drop_union");
    u32* ks = kstar_d.ensure(kstar_len);
    be.h2d(ks, kstar_host, kstar_len * sizeof(u32));
    const int64_t klen = kstar_len; const int64_t nbig = n_big_k; const u32* bign = big_n_d.p; const u32* bigk = big_k_d.p;
    n_big_k = 0;                      // the sparse list belongs to this call only
    const u32* ea_ = ed_a.p; const u32* eb_ = ed_b.p; const u32* esup = ed_sup.p; const u32* etot = ed_tot.p;
    const u8* ecfg = ed_cfg.p; u8* keep = ed_keep.p;
    u32* par = parent.ensure(Vn); u32* dg = deg.ensure(Vn + 1); be.memset0(dg, (Vn + 1) * sizeof(u32));
    u32* sc = scalars.p; be.memset0(sc, 8 * sizeof(u32));
    be.for_each(Vn, PHZ_LAMBDA(int64_t v) { par[v] = (u32)v; });
    be.for_each(E, PHZ_LAMBDA(int64_t e) {
      u32 sup = esup[e], tot = etot[e];
      u32 kcrit;
      if ((int64_t)tot < klen) kcrit = ks[tot];
      else {                          // sparse tail: exact match required (the host computed one value per distinct total)
        int64_t lo = 0, hi = nbig;
        while (lo < hi) { int64_t m = (lo + hi) >> 1; if (bign[m] < tot) lo = m + 1; else hi = m; }
        if (lo < nbig && bign[lo] == tot) kcrit = bigk[lo]; else { kcrit = 0; atomic_or(&sc[1], 4u); }
      }
      bool drop = (sup == 0) || (tot > sup && sup < kcrit);
      keep[e] = drop ? 0 : 1;
      if (drop) { atomic_add(&sc[2], 1u); return; }
      u32 a = ea_[e], b = eb_[e];
      atomic_add(&dg[a], 1u); atomic_add(&dg[b], 1u);
      // lock-free union: hook the larger root under the smaller one
      while (true) {
        while (true) { u32 p = load_volatile(&par[a]); if (p == a) break; a = p; }
        while (true) { u32 p = load_volatile(&par[b]); if (p == b) break; b = p; }
        if (a == b) break;
        if (a > b) { u32 t = a; a = b; b = t; }
        if (atomic_cas(&par[b], b, a) == b) break;
      }
    });
    be.stage("phase.members");
    // ---- members of the kept-edge graph, grouped by component root, ascending variant index inside.  Their number is
    // read with the number of dropped edges in ONE wait.
    u32* rt = root.ensure(Vn); u32* mf = m_flag.ensure(Vn + 1); u32* ms = m_scan.ensure(Vn + 2);
    be.for_each(Vn, PHZ_LAMBDA(int64_t v) {
      if (dg[v] == 0) { mf[v] = 0; rt[v] = NONE32; return; }
      u32 a = (u32)v; while (true) { u32 p = par[a]; if (p == a) break; a = p; }
      rt[v] = a; mf[v] = 1;
    });
    be.exclusive_scan_u32(mf, ms, Vn);
    u32* af = x_flag.ensure(E + 1 > NX + 1 ? E + 1 : NX + 1); u32* as_ = x_scan.ensure(E + 2 > NX + 2 ? E + 2 : NX + 2);
    { const int64_t vn = Vn;
      be.for_each(1, PHZ_LAMBDA(int64_t) { sc[4] = vn > 0 ? ms[vn] : 0u; }); }
    u32 h3[6] = {0, 0, 0, 0, 0, 0};
    be.d2h(h3, sc, sizeof(h3));
    n_dropped = h3[2]; NM = h3[4];
    u32* ml = m_list.ensure(NM); u32* mk = m_key.ensure(NM); u32* mk2 = m_key2.ensure(NM); u32* mem = members.ensure(NM);
    be.for_each(Vn, PHZ_LAMBDA(int64_t v) { if (mf[v]) { ml[ms[v]] = (u32)v; mk[ms[v]] = rt[v]; } });
    be.sort_pairs32(mk, mk2, ml, mem, NM, 0, vb);
    u32* bf = b_flag.ensure(NM + 1); u32* bs = b_scan.ensure(NM + 2);
    be.for_each(NM, PHZ_LAMBDA(int64_t i) { bf[i] = (i == 0 || mk2[i] != mk2[i - 1]) ? 1u : 0u; });
    be.exclusive_scan_u32(bf, bs, NM);
    // ---- adjacency (kept, non-tie) by source variant: a counting sort (degree, scan, scatter through per-variant cursors;
    // the 2-colouring below does not depend on the order inside a variant's list).  Independent of the block count, so
    // it is queued before that is read.  adj[k] = neighbour | configuration << 31.
    u32* adj = adj_list.ensure(2 * E + 1);
    u32* ao = adj_off.ensure(Vn + 2);
    u32* dcnt = m_flag.p;      // reuse (the member flags were consumed above): per-variant directed degree, then cursor
    be.memset0(dcnt, (Vn + 1) * sizeof(u32));
    be.for_each(E, PHZ_LAMBDA(int64_t e) {
      if (!keep[e] || ecfg[e] == EDGE_TIE) return;
      atomic_add(&dcnt[ea_[e]], 1u); atomic_add(&dcnt[eb_[e]], 1u);
    });
    be.exclusive_scan_u32(dcnt, ao, Vn);
    be.memset0(dcnt, (Vn + 1) * sizeof(u32));
    be.for_each(E, PHZ_LAMBDA(int64_t e) {
      if (!keep[e] || ecfg[e] == EDGE_TIE) return;
      const u32 a = ea_[e], b = eb_[e], sgn = (u32)ecfg[e] << 31;
      adj[ao[a] + atomic_add(&dcnt[a], 1u)] = b | sgn;
      adj[ao[b] + atomic_add(&dcnt[b], 1u)] = a | sgn;
    });
    NB = NM > 0 ? (int64_t)fetch_u32(bs + NM) : 0;
    u32* bo = blk_off.ensure(NB + 1); u32* bof = blk_of.ensure(Vn); u32* pib = pos_in_blk.ensure(Vn);
    be.memset_ff(bof, Vn * sizeof(u32));
    { int64_t nm = NM, nbk = NB;
      be.for_each(NM + 1, PHZ_LAMBDA(int64_t i) { if (i == nm) bo[nbk] = (u32)nm; else if (bf[i]) bo[bs[i]] = (u32)i; }); }
    u64* br = blk_rank.ensure(NB); be.memset_ff(br, NB * sizeof(u64));
    const u64* vr = vrank.p;
    be.for_each(NM, PHZ_LAMBDA(int64_t i) {
      u32 b = bs[i] + bf[i] - 1; u32 v = mem[i];
      bof[v] = b;
      atomic_min((unsigned long long*)&br[b], (unsigned long long)vr[v]);
    });
    be.for_each(NM, PHZ_LAMBDA(int64_t i) { u32 v = mem[i]; pib[v] = (u32)i - bo[bof[v]]; });
    be.stage("phase.block_order");
    // ---- block output order: contig first-appearance rank, then first remaining key (phaser.py:1870)
    u64* bk = bs_key.ensure(NB); u64* bk2 = bs_key2.ensure(NB); u32* bv = bs_val.ensure(NB); u32* bv2 = bs_val2.ensure(NB);
    u32* b32 = bs_k32.ensure(NB); u32* b32b = bs_k32b.ensure(NB); u32* bord = blk_order.ensure(NB); u32* bpos = blk_pos.ensure(NB);
    const u32* cr = crank.p;
    be.for_each(NB, PHZ_LAMBDA(int64_t b) { bk[b] = br[b]; bv[b] = (u32)b; });
    be.sort_pairs(bk, bk2, bv, bv2, NB, 0, 64);
    be.for_each(NB, PHZ_LAMBDA(int64_t i) { b32[i] = cr[vc[mem[bo[bv2[i]]]]]; });
    be.sort_pairs32(b32, b32b, bv2, bord, NB, 0, ceil_log2_host((u64)(nc > 1 ? nc : 2)));
    be.for_each(NB, PHZ_LAMBDA(int64_t i) { bpos[bord[i]] = (u32)i; });
    be.stage("phase.fast_bfs");
    // ---- fast path: 2-colouring from the leftmost variant (resolve_phase, phaser.py:2172-2207)
    u8* col = color.ensure(Vn); be.memset_ff(col, Vn);
    u8* vh = v_hap.ensure(Vn); be.memset0(vh, Vn);
    u32* q = bfsq.ensure(NM); u32* rs = run_start.ensure(NM); u32* rl = run_len.ensure(NM);
    u8* bst = blk_status.ensure(NB); u32* bnf = blk_nfinal.ensure(NB + 1);
    u32* vfl = v_fin_local.ensure(Vn); be.memset_ff(vfl, Vn * sizeof(u32));
    const u64 vmask = (((u64)1) << vb) - 1;
    be.for_each(NB, PHZ_LAMBDA(int64_t b) {
      u32 o0 = bo[b], o1 = bo[b + 1]; u32 n = o1 - o0;
      u32 seed = mem[o0];
      col[seed] = 0; q[o0] = seed; u32 qh = 0, qt = 1; bool conflict = false;
      while (qh < qt) {
        u32 v = q[o0 + qh++]; u8 cv = col[v];
        for (u32 k = ao[v]; k < ao[v + 1]; ++k) {
          const u32 x = adj[k]; u32 w = x & 0x7FFFFFFFu; u8 want = cv ^ (u8)(x >> 31);
          if (col[w] == 0xFF) { col[w] = want; q[o0 + qt++] = w; }
          else if (col[w] != want) conflict = true;
        }
      }
      u32 m = qt;
      if (!conflict && m == n) {
        bst[b] = 0; bnf[b] = 1; rs[o0] = 0; rl[o0] = n;
        for (u32 i = o0; i < o1; ++i) { u32 v = mem[i]; vh[v] = col[v]; vfl[v] = 0; }
      } else if (conflict && 2 * m == n) {       // reaches n alleles on m variants: short all-zero string
        bst[b] = 1; bnf[b] = 1; rs[o0] = 0; rl[o0] = m;
        for (u32 i = o0; i < o0 + m; ++i) { u32 v = mem[i]; vh[v] = 0; vfl[v] = 0; }
      } else { bst[b] = 2; bnf[b] = 0; atomic_add(&sc[7], n); }       // sc[7]: members of hard blocks (sizes their scratch)
    });
    be.stage("phase.hard_prepare");
    // ---- hard path: phase_v3 proper, one logical thread per block
    u32* hf = h_flag.ensure(NB + 1); u32* hs = h_scan.ensure(NB + 2);
    be.for_each(NB, PHZ_LAMBDA(int64_t b) { hf[b] = bst[b] == 2 ? 1u : 0u; });
    be.exclusive_scan_u32(hf, hs, NB);
    u32 h8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (NB > 0) {
      const int64_t nbk = NB;
      be.for_each(1, PHZ_LAMBDA(int64_t) { sc[3] = hs[nbk]; });
      be.d2h(h8, sc, sizeof(h8));            // ONE wait: number of hard blocks + their members (kept edges = E - dropped is known)
    }
    NH = h8[3];
    if (NH > 0) {
      // per-block lists of kept edges (ties included: they count as connections in find_weak_points)
      u32* ec = ebk_cnt.ensure(NB + 1); u32* eo = ebk_off.ensure(NB + 2); be.memset0(ec, (NB + 1) * sizeof(u32));
      be.for_each(E, PHZ_LAMBDA(int64_t e) { af[e] = keep[e] ? 1u : 0u; if (keep[e]) atomic_add(&ec[bof[ea_[e]]], 1u); });
      be.exclusive_scan_u32(ec, eo, NB);
      be.exclusive_scan_u32(af, as_, E);
      const int64_t NK = E - (int64_t)n_dropped;
      u32* kk = ebk_key.ensure(NK); u32* kk2 = ebk_key2.ensure(NK); u32* kv = ebk_val.ensure(NK); u32* kl = ebk_list.ensure(NK);
      be.for_each(E, PHZ_LAMBDA(int64_t e) { if (af[e]) { kk[as_[e]] = bof[ea_[e]]; kv[as_[e]] = (u32)e; } });
      be.sort_pairs32(kk, kk2, kv, kl, NK, 0, ceil_log2_host((u64)(NB > 1 ? NB : 2)));
      u32* hl = h_list.ensure(NH); u32* hw = h_words.ensure(NH + 1); u32* hwo = h_woff.ensure(NH + 2);
      be.for_each(NB, PHZ_LAMBDA(int64_t b) {
        if (!hf[b]) return;
        hl[hs[b]] = (u32)b; hw[hs[b]] = (u32)hard_scratch_words(bo[b + 1] - bo[b]);
      });
      be.exclusive_scan_u32(hw, hwo, NH);
      const size_t words = hard_scratch_words(0) * (size_t)NH + (hard_scratch_words(1) - hard_scratch_words(0)) * (size_t)h8[7];
      if (words >= 0xFFFFFFF0ull) throw PhzError("hard-block scratch beyond 2^32 words");
      u32* scr = h_scratch.ensure(words);
      int mbs = max_block_size;
      be.stage("phase.hard_kernel");
      be.for_each_warp(NH, PHZ_LAMBDA_WARP(int64_t h, int lane, int nlanes) {
        u32 b = hl[h]; u32 o0 = bo[b]; int n = (int)(bo[b + 1] - o0);
        BlockEdges bed{kl + eo[b], eo[b + 1] - eo[b], ea_, eb_, ecfg, pib};
        Coop cp{lane, nlanes};
        // hap / fin_local are written through small local views over the member segment
        u8* hap_loc = (u8*)(q + o0);           // n bytes inside this block's queue segment (n u32 words)
        u32* fin_loc = scr + hwo[h] + hard_scratch_words(n) - n;   // tail of this block's scratch
        for (int i = lane; i < n; i += nlanes) { fin_loc[i] = NONE32; hap_loc[i] = 0; }
        coop_sync(cp);
        int err = 0;
        int nr = phase_block_hard(bed, n, mbs, scr + hwo[h], rs + o0, rl + o0, hap_loc, fin_loc, &err, cp);
        if (lane == 0) bnf[b] = (u32)nr;
        for (int i = lane; i < n; i += nlanes) { u32 v = mem[o0 + i]; vfl[v] = fin_loc[i]; vh[v] = hap_loc[i]; }
        if (err && lane == 0) atomic_or(&sc[1], (u32)err);
      });
    }
    // ---- final blocks in output order (block_index of phaser.py:863-867)
    u32* nfo = nf_ord.ensure(NB + 1); u32* fbb = fb_base.ensure(NB + 2);
    be.for_each(NB, PHZ_LAMBDA(int64_t i) { nfo[i] = bnf[bord[i]]; });
    be.exclusive_scan_u32(nfo, fbb, NB);
    // the number of final blocks is read with the closing counters: until then every array is sized by its bound (a final
    // block has at least one member) and no launch depends on it
    const int64_t NFmax = NM > 0 ? NM : 1;
    { const int64_t nbk = NB; be.for_each(1, PHZ_LAMBDA(int64_t) { sc[0] = nbk > 0 ? fbb[nbk] : 0u; }); }
    u32* ff = fb_first.ensure(NFmax); u32* fl = fb_len.ensure(NFmax); u32* fbk = fb_blk.ensure(NFmax);
    u32* vfin = v_final.ensure(Vn); be.memset_ff(vfin, Vn * sizeof(u32));
    u32* vmi = v_member.ensure(Vn);
    be.for_each(NB, PHZ_LAMBDA(int64_t i) {
      u32 b = bord[i]; u32 o0 = bo[b]; u32 longest = 0;
      for (u32 r = 0; r < bnf[b]; ++r) {
        u32 f = fbb[i] + r; ff[f] = o0 + rs[o0 + r]; fl[f] = rl[o0 + r]; fbk[f] = b;
        if (rl[o0 + r] > longest) longest = rl[o0 + r];
      }
      if (longest > load_volatile(&sc[6])) atomic_max(&sc[6], longest);      // sizes the rank field of the read-list keys
    });
    be.for_each(NM, PHZ_LAMBDA(int64_t i) {
      u32 v = mem[i]; vmi[v] = (u32)i;
      if (vfl[v] != NONE32) vfin[v] = fbb[bpos[bof[v]]] + vfl[v];
    });
    be.stage("phase.edge_support");
    // ---- edge support per final block (phaser.py:876-895)
    u32* fsup = fb_sup.ensure(NFmax); u32* ftot = fb_tot.ensure(NFmax);
    be.memset0(fsup, NFmax * sizeof(u32)); be.memset0(ftot, NFmax * sizeof(u32));
    be.for_each(E, PHZ_LAMBDA(int64_t e) {
      if (!keep[e] || ecfg[e] == EDGE_TIE) return;
      u32 a = ea_[e], b = eb_[e]; u32 fa = vfin[a];
      if (fa == NONE32 || fa != vfin[b]) return;
      atomic_add(&ftot[fa], 1u);
      if ((vh[a] ^ vh[b]) == ecfg[e]) atomic_add(&fsup[fa], 1u);
    });
    be.stage("phase.hap_counts");
    // ---- unique-fragment counts per final block x haplotype (all BAMs, and per counted BAM)
    u32* fc = fb_cnt.ensure(NFmax * 2); u32* fbc = fb_bcnt.ensure(NFmax * nb * 2);
    be.memset0(fc, NFmax * 2 * sizeof(u32)); be.memset0(fbc, NFmax * nb * 2 * sizeof(u32));
    const u8* vbl = vblack;
    // per site: final block << 1 | haplotype in ONE word (NONE32 outside the final blocks), and the site's rank inside
    // its final block -- what the haplotypic counts and the read lists would otherwise gather from three arrays
    u32* pv = v_packed.ensure(Vn); u32* pr = v_rank_in_final.ensure(Vn);
    be.for_each(Vn, PHZ_LAMBDA(int64_t v) {
      const u32 f = vfin[v];
      pv[v] = f == NONE32 ? NONE32 : ((f << 1) | (u32)(vh[v] & 1));
      pr[v] = f == NONE32 ? 0u : vmi[v] - ff[f];
    });
    if (frag_entries) {
      // entries as the fragment-table stage left them, walked by SLOT: every fragment's entries sit at the front of its
      // slots, the slots behind them carry an empty class mask, the first slot carries the head mark -- one thread per
      // slot looks back over the earlier entries of its fragment (a handful); neither the fragment table nor the entry
      // counts are read.  Per site one packed word (final block << 1 | haplotype) instead of two gathers.
      const u64* fk = f_key.p; const uint16_t* fi = f_info.p;
      be.for_each(n_tuples, PHZ_LAMBDA(int64_t j) {
        const u32 ij = fi[j]; const u32 mj = ij & 7u;
        if (!mj) return;
        const u32 v = (u32)(fk[j] >> 32); const u32 pj = pv[v];
        if (pj == NONE32) return;
        const u32 f = pj >> 1, hv = pj & 1u, bj = (ij & 0x7FFFu) >> 3;
        for (u32 h = 0; h < 2; ++h) {
          if (!((mj >> (hv ^ h)) & 1)) continue;
          bool seen_any = false, seen_bam = false;
          if (!(ij & INFO_HEAD))
            for (int64_t i = j - 1;; --i) {
              const u32 ii = fi[i]; const u32 mi = ii & 7u;
              if (mi) {
                const u32 w = (u32)(fk[i] >> 32); const u32 pw = pv[w];
                if (pw != NONE32 && (pw >> 1) == f && ((mi >> ((pw & 1u) ^ h)) & 1)) {
                  seen_any = true; if (((ii & 0x7FFFu) >> 3) == bj && !(vbl && vbl[w])) seen_bam = true;
                }
              }
              if ((ii & INFO_HEAD) || i == 0) break;
            }
          if (!seen_any) converged_inc(&fc[(int64_t)f * 2 + h]);
          if (!seen_bam && !((excl_mask >> bj) & 1) && !(vbl && vbl[v]))
            converged_inc(&fbc[((int64_t)f * nb + bj) * 2 + h]);
        }
      });
    } else {
    const u64* ek = e_key.p; const u8* eb = e_bam.p; const u32* em = e_mask.p; const u32* go = grp_off.p;
    be.for_each(NG, PHZ_LAMBDA(int64_t g) {
      u32 j0 = go[g], j1 = go[g + 1];
      for (u32 j = j0; j < j1; ++j) {
        u32 v = (u32)(ek[j] & vmask); u32 f = vfin[v];
        if (f == NONE32) continue;
        for (int h = 0; h < 2; ++h) {
          if (!((em[j] >> (vh[v] ^ h)) & 1)) continue;
          bool seen_any = false, seen_bam = false;
          for (u32 i = j0; i < j; ++i) {
            u32 w = (u32)(ek[i] & vmask);
            if (vfin[w] != f || !((em[i] >> (vh[w] ^ h)) & 1)) continue;
            seen_any = true; if (eb[i] == eb[j] && !(vbl && vbl[w])) seen_bam = true;
          }
          // fragments of one locus sit in neighbouring lanes and hit the same block counter: combine them
          if (!seen_any) converged_inc(&fc[(int64_t)f * 2 + h]);
          if (!seen_bam && !((excl_mask >> eb[j]) & 1) && !(vbl && vbl[v]))
            converged_inc(&fbc[((int64_t)f * nb + eb[j]) * 2 + h]);
        }
      }
    });
    }
    u32 hsc8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    be.d2h(hsc8, sc, sizeof(hsc8));
    *err_out = (int)hsc8[1]; max_final_len = hsc8[6]; NF = hsc8[0];
    be.stage("phase.end");
  }

  // =================================================================== per-variant read lists of the rows
  // Tuples (reference/alternative calls of counted BAMs) of variants inside final blocks, ordered by
  // (final block, BAM, haplotype, variant, tuple order): the lists behind aReads/bReads (phaser.py:1105-1115).
  int64_t read_lists(u64 excl_mask) {
    const int64_t n = n_tuples; const int nb = n_bams > 0 ? n_bams : 1; const int vb = vbits;
    const u32* gf = g_frag.p; const u32* gv = g_var.p; const u8* gc = g_cb.p; const u32* vfin = v_final.p; const u8* vh = v_hap.p;
    const u8* vbl = vblack;
    const u32* pv = v_packed.p; const u32* pr = v_rank_in_final.p;
    be.stage("read_lists");
    // a tuple is listed iff it is a reference / alternative call of a counted BAM at a site inside a final block.  The
    // selection keeps tuple order without a per-tuple flag or offset array: one warp per tile of RL_TILE tuples counts its
    // listed tuples, a scan over the tile counts places the tiles, and the same warps write their keys in order
    // (ballot prefix) -- the test is evaluated twice, its one gather is the packed per-site word.
    auto listed = PHZ_LAMBDA(int64_t t) -> bool {
      const u32 c = gc[t];
      if ((c & 3u) >= 2u || ((excl_mask >> (c >> 2)) & 1)) return false;
      const u32 v = gv[t];
      return pv[v] != NONE32 && !(vbl && vbl[v]);
    };
    constexpr int64_t RL_TILE = 1024;
    const int64_t n_rt = (n + RL_TILE - 1) / RL_TILE;
    u32* tcn = rl_flag.ensure(n_rt + 1); u32* tof = rl_scan.ensure(n_rt + 2);
    be.for_each_warp(n_rt, PHZ_LAMBDA_WARP(int64_t tile, int lane, int nlanes) {
      const int64_t t0 = tile * RL_TILE, t1 = t0 + RL_TILE < n ? t0 + RL_TILE : n;
      u32 k = 0;
      for (int64_t i0 = t0; i0 < t1; i0 += nlanes) {
        const int64_t t = i0 + lane;
        k += popc_u32(warp_ballot(t < t1 && listed(t)));
      }
      if (lane == 0) tcn[tile] = k;
    });
    be.exclusive_scan_u32(tcn, tof, n_rt);
    NRL = n_rt > 0 ? (int64_t)fetch_u32(tof + n_rt) : 0;
    u32* k32 = rl_k32.ensure(NRL); u32* k32b = rl_k32b.ensure(NRL); u32* tt = rl_t.ensure(NRL); u32* tt2 = rl_t2.ensure(NRL);
    int bb = ceil_log2_host((u64)(nb > 1 ? nb : 2));
    int fbits = ceil_log2_host((u64)(NF > 1 ? NF : 2));
    if (fbits + bb + 1 > 32) throw PhzError("row key of read_lists does not fit 32 bits");
    const int rbits = ceil_log2_host((u64)max_final_len + 1) > 0 ? ceil_log2_host((u64)max_final_len + 1) : 1;
    const u32* kres = k32b;
    int shift = 0;
    const bool one_sort = fbits + bb + 1 + rbits <= 32 && !two_pass_read_lists;
    // one sort: key = (block, BAM, haplotype, rank of the variant inside its block), tuple order kept by stability;
    // otherwise first by variant, then by (block, BAM, haplotype)
    u32* tsel = one_sort ? tt2 : tt;
    be.for_each_warp(n_rt, PHZ_LAMBDA_WARP(int64_t tile, int lane, int nlanes) {
      const int64_t t0 = tile * RL_TILE, t1 = t0 + RL_TILE < n ? t0 + RL_TILE : n;
      u32 o = tof[tile];
      for (int64_t i0 = t0; i0 < t1; i0 += nlanes) {
        const int64_t t = i0 + lane;
        const bool keep = t < t1 && listed(t);
        const u32 bal = warp_ballot(keep);
        if (keep) {
          const u32 w = o + popc_u32(bal & ((1u << lane) - 1u));
          const u32 v = gv[t];
          if (one_sort) {
            const u32 p = pv[v]; const u32 c = gc[t];
            const u32 row = ((((p >> 1) << bb) | (c >> 2)) << 1) | ((c & 3u) ^ (p & 1u));
            k32[w] = (row << rbits) | pr[v];
          } else k32[w] = v;
          tsel[w] = (u32)t;
        }
        o += popc_u32(bal);
      }
    });
    if (one_sort) {
      be.sort_pairs32(k32, k32b, tt2, tt, NRL, 0, fbits + bb + 1 + rbits);
      shift = rbits;
    } else {
      be.sort_pairs32(k32, k32b, tt, tt2, NRL, 0, vb);          // by variant, tuple order kept
      be.for_each(NRL, PHZ_LAMBDA(int64_t i) {
        u32 t = tt2[i]; u32 v = gv[t]; u32 hap = (gc[t] & 3) ^ vh[v];
        k32[i] = (((vfin[v] << bb) | (u32)(gc[t] >> 2)) << 1) | hap;
      });
      be.sort_pairs32(k32, k32b, tt2, tt, NRL, 0, fbits + bb + 1);      // by (block, BAM, haplotype), variant order kept
    }
    u32* of = rl_frag.ensure(NRL); u32* ov = rl_var.ensure(NRL); u32* orow = rl_row.ensure(NRL);
    be.for_each(NRL, PHZ_LAMBDA(int64_t i) { u32 t = tt[i]; of[i] = gf[t]; ov[i] = gv[t]; orow[i] = kres[i] >> shift; });
    be.stage("read_lists.end");
    return NRL;
  }
};

}  // namespace phz

// K2 graph stage, fragment-table form (graph_mode 1, the default).
//
// Replaces phaser.py:1287-1328, 558-581 (tuples -> per-variant lists, Q9 effective BAM), :636-640 (unique read sets),
// :1265-1285 (connectivity map / insertion order), :667-678 (pair enumeration), :1594-1642 (3x3 co-occurrence cells).
//
// Fragment ids are dense (one per distinct QNAME, numbered by first appearance in the coordinate-sorted input), so
// "group the tuples by fragment" needs no sort: a counting pass ranks every tuple inside its fragment, a scan over
// the fragment table gives each fragment its slot range, a scatter drops (variant << 32 | tuple) keys there.  One
// logical thread then owns one fragment: it sorts its few keys in place, folds them into (variant, BAM) entries,
// and -- with everything of the fragment at hand -- adds to the per-variant counters, to the insertion ranks and to
// the pair table.  Consecutive fragments sit on the same locus and expression is heavy-tailed, so a CTA first sums
// in shared memory (a window of variant indices for the per-variant counters, a small hash table for the pairs) and
// flushes each touched slot once; the run-wide pair table is an open-addressing hash in global memory (L2-resident:
// a few 100 k distinct pairs) whose non-empty slots are compacted and sorted by (va, vb) into the edge table.
// No tuple is ever sorted globally, no pair instance is ever written to HBM.
#pragma once
#include "phz_backend.h"

namespace phz {

constexpr u64 PAIR_EMPTY = ~0ull;
constexpr int PAIR_CELLS = 10;            // 9 co-occurrence cells n[x][y] at x*3+y, [9] = eligible fragments
constexpr int FRAG_CTA = 256;             // threads per CTA of the fragment kernel
#ifndef PHZ_FRAG_PER_THREAD
#define PHZ_FRAG_PER_THREAD 8
#endif
#ifndef PHZ_FRAG_HS
#define PHZ_FRAG_HS 512
#endif
#ifndef PHZ_FRAG_W
#define PHZ_FRAG_W 256
#endif
constexpr int FRAG_PER_THREAD = PHZ_FRAG_PER_THREAD;   // consecutive fragment ids per CTA = FRAG_CTA * FRAG_PER_THREAD
constexpr int FRAG_W = PHZ_FRAG_W;        // variant indices per shared-memory window
constexpr int FRAG_HS = PHZ_FRAG_HS;      // slots of the CTA's pair hash
#ifndef PHZ_FRAG_SLOTS
#define PHZ_FRAG_SLOTS 2048
#endif
#ifndef PHZ_FRAG_SLOT_CTAS
#define PHZ_FRAG_SLOT_CTAS 4
#endif
constexpr int FRAG_SLOTS = PHZ_FRAG_SLOTS;   // slot-chunk form: tuple slots per CTA (single-BAM instantiation; 3/4 of it otherwise)
constexpr int FRAG_EXTRA = 128;           // slots staged beyond the chunk, for the tail of its last fragment
constexpr uint16_t INFO_HEAD = 0x8000;    // f_info bit: first slot of a fragment (set by the scatter pass, kept for good)

PHZ_HD u32 pair_hash(u64 k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 29;
  return (u32)k;
}

struct PairTable {
  u64* keys;          // [S], PAIR_EMPTY = free
  u32* vals;          // [S * PAIR_CELLS]
  u32 mask;           // S - 1
  u32* flags;         // [0] bit 0: table full (the host grows it and runs the stage again)
};

// slot of `key`, inserting it when absent; -1 when the probe sequence is exhausted
PHZ_HD int pair_slot(const PairTable& pt, u64 key) {
  u32 h = pair_hash(key) & pt.mask;
  for (int probe = 0; probe < 512; ++probe) {
    u64 k = load_volatile((const unsigned long long*)&pt.keys[h]);
    if (k == key) return (int)h;
    if (k == PAIR_EMPTY) {
      u64 old = atomic_cas((unsigned long long*)&pt.keys[h], (unsigned long long)PAIR_EMPTY, (unsigned long long)key);
      if (old == PAIR_EMPTY || old == key) return (int)h;
    }
    h = (h + 1) & pt.mask;
  }
  atomic_or(&pt.flags[0], 1u);
  return -1;
}

// 3 x 3 outer product of two class masks: bit x*3+y set iff class x at the first site and class y at the second
PHZ_HD u32 pair_cells_of(u32 ma, u32 mb) {
  return ((ma & 1u) ? mb : 0u) | ((ma & 2u) ? (mb << 3) : 0u) | ((ma & 4u) ? (mb << 6) : 0u);
}

struct FragCtx {
  const u32* vc;      // contig of every het site
  u64 excl_mask;      // BAMs left out of the haplotypic counts (phaser.py:1320, Q25)
  u64* vrank;         // per site: insertion key into the overlap dict (phaser.py:1271-1283), min wins
  const u32* abort;   // bit 3 set by the ranking pass: a fragment does not fit the stage, its slots hold nothing valid
};

// One fragment.  k[0..n): keys (variant << 32 | tuple index) in any order, info[0..n): class | bam << 2 of the same tuples
// (written next to the keys by the scatter pass, so no dependent look-up per tuple is needed here).  On return
// k[0..ne) / info[0..ne) hold the fragment's (variant, BAM) entries sorted by (variant, BAM):
// k = variant << 32 | first tuple with a reference / alternative call (NONE32 if none), info = bam << 3 | class mask;
// info[ne..n) = 0 (an empty class mask), so that a pass over the SLOTS needs neither the fragment table nor ne.
// Returns ne; adds groups / pairs to ng / np.
template <bool ONE_BAM, class Sink>
PHZ_HD u32 process_fragment(const FragCtx& c, u64* k, uint16_t* info, u32 n, Sink& sink, u32& ng, u32& np) {
  // tuples arrive nearly sorted (record order = position order): insertion sort, stable because keys are unique
  for (u32 j = 1; j < n; ++j) {
    u64 key = k[j]; uint16_t cbj = info[j]; u32 m = j;
    while (m > 0 && k[m - 1] > key) { k[m] = k[m - 1]; info[m] = info[m - 1]; --m; }
    k[m] = key; info[m] = cbj;
  }
  if (ONE_BAM) {
    // one BAM: an entry is a distinct variant; no effective-BAM bookkeeping, no re-scans of equal-variant runs
    u32 ne = 0, cur_v = 0xFFFFFFFFu;
    for (u32 i = 0; i < n; ++i) {
      const u64 key = k[i]; const u32 v = (u32)(key >> 32), t = (u32)key;
      const u32 cls = info[i] & 3u;          // read before slot ne <= i is overwritten below
      if (v != cur_v) { k[ne] = ((u64)v << 32) | 0xFFFFFFFFu; info[ne] = 0; ++ne; cur_v = v; }
      info[ne - 1] = (uint16_t)(info[ne - 1] | (1u << cls));
      if (cls < 2 && (u32)k[ne - 1] == 0xFFFFFFFFu) k[ne - 1] = ((u64)v << 32) | t;
    }
    const bool counted = !(c.excl_mask & 1);
    for (u32 g0 = 0; g0 < ne;) {
      const u32 v0 = (u32)(k[g0] >> 32); const u32 cg = c.vc[v0]; u32 g1 = g0;
      u32 kelig = 0, first_t = 0xFFFFFFFFu;
      for (; g1 < ne; ++g1) {
        const u32 v = (u32)(k[g1] >> 32);
        if (g1 > g0 && c.vc[v] != cg) break;
        const u32 m = info[g1];
        if (m & 1) sink.set_size(v, 0);
        if (m & 2) sink.set_size(v, 1);
        if (m & 4) sink.set_size(v, 2);
        if (counted) { if (m & 1) sink.bam_count(v, 0, 0); if (m & 2) sink.bam_count(v, 0, 1); }
        if (m & 3) { ++kelig; const u32 t = (u32)k[g1]; if (t < first_t) first_t = t; }
      }
      ++ng;
      const u32 kk = g1 - g0;
      if (kk >= 2) {
        np += kk * (kk - 1) / 2;
        if (kelig >= 2)
          for (u32 j = g0; j < g1; ++j)
            if (info[j] & 3) {
              const unsigned long long key = ((unsigned long long)first_t << 32) | (u32)k[j];
              unsigned long long* slot = (unsigned long long*)&c.vrank[(u32)(k[j] >> 32)];
              if (key < load_volatile(slot)) atomic_min(slot, key);
            }
        for (u32 a = g0; a + 1 < g1; ++a) {
          const u32 va = (u32)(k[a] >> 32), ma = info[a];
          for (u32 b = a + 1; b < g1; ++b) {
            const u32 mb = info[b];
            u32 cells = pair_cells_of(ma, mb);
            if ((ma & 3) && (mb & 3)) cells |= 1u << 9;
            sink.pair(((u64)va << 32) | (u32)(k[b] >> 32), cells);
          }
        }
      }
      g0 = g1;
    }
    for (u32 x = ne; x < n; ++x) info[x] = 0;       // slots behind the entries: no class, inert for whoever walks the slots
    return ne;
  }
  // (variant, BAM) entries, in place: tuple order inside a variant is BAM-major (commit order)
  u32 ne = 0, cur_v = 0xFFFFFFFFu, cur_b = 0xFFFFFFFFu;
  for (u32 i = 0; i < n; ++i) {
    const u64 key = k[i]; const u32 v = (u32)(key >> 32), t = (u32)key;
    const u32 cb = info[i]; const u32 cls = cb & 3, bam = cb >> 2;          // read before slot ne <= i is overwritten below
    if (v != cur_v || bam != cur_b) { k[ne] = ((u64)v << 32) | 0xFFFFFFFFu; info[ne] = (uint16_t)(bam << 3); ++ne; cur_v = v; cur_b = bam; }
    info[ne - 1] = (uint16_t)(info[ne - 1] | (1u << cls));
    if (cls < 2 && (u32)k[ne - 1] == 0xFFFFFFFFu) k[ne - 1] = ((u64)v << 32) | t;
  }
  // unique-set sizes (phaser.py:636-640) and per-BAM allele counts (haplo_reads, phaser.py:1320-1322)
  for (u32 j = 0; j < ne;) {
    const u32 v = (u32)(k[j] >> 32); u32 mask = 0, j1 = j;
    for (; j1 < ne && (u32)(k[j1] >> 32) == v; ++j1) mask |= info[j1] & 7u;
    for (int x = 0; x < 3; ++x) if ((mask >> x) & 1) sink.set_size(v, x);
    for (u32 i = j; i < j1; ++i) {
      const u32 bam = info[i] >> 3;
      if ((c.excl_mask >> bam) & 1) continue;
      if (info[i] & 1) sink.bam_count(v, bam, 0);
      if (info[i] & 2) sink.bam_count(v, bam, 1);
    }
    j = j1;
  }
  // (fragment, contig) groups: effective BAM (Q9), insertion ranks, pairs
  for (u32 g0 = 0; g0 < ne;) {
    const u32 cg = c.vc[(u32)(k[g0] >> 32)]; u32 g1 = g0 + 1;
    while (g1 < ne && c.vc[(u32)(k[g1] >> 32)] == cg) ++g1;
    ++ng;
    if (g1 - g0 >= 2) {
      int effbam = -1; u32 first_t = 0xFFFFFFFFu;
      for (u32 j = g0; j < g1; ++j)
        if (info[j] & 3) { const int b = (int)(info[j] >> 3); if (b > effbam) effbam = b; const u32 t = (u32)k[j]; if (t < first_t) first_t = t; }
      u32 kk = 0, kelig = 0;
      for (u32 j = g0; j < g1;) {
        const u32 v = (u32)(k[j] >> 32); bool elig = false; u32 jj = j;
        for (; jj < g1 && (u32)(k[jj] >> 32) == v; ++jj) if ((info[jj] & 3) && (int)(info[jj] >> 3) == effbam) elig = true;
        ++kk; if (elig) ++kelig;
        j = jj;
      }
      np += kk * (kk - 1) / 2;
      if (kelig >= 2)         // insertion order of dict_variant_overlap, phaser.py:1271-1283
        for (u32 j = g0; j < g1; ++j)
          if ((info[j] & 3) && (int)(info[j] >> 3) == effbam) {
            const unsigned long long key = ((unsigned long long)first_t << 32) | (u32)k[j];
            unsigned long long* slot = (unsigned long long*)&c.vrank[(u32)(k[j] >> 32)];
            if (key < load_volatile(slot)) atomic_min(slot, key);          // the minimum settles early
          }
      if (kk >= 2)
        for (u32 a = g0; a < g1;) {
          const u32 va = (u32)(k[a] >> 32); u32 ma = 0; bool ea = false; u32 a1 = a;
          for (; a1 < g1 && (u32)(k[a1] >> 32) == va; ++a1) { ma |= info[a1] & 7u; if ((info[a1] & 3) && (int)(info[a1] >> 3) == effbam) ea = true; }
          for (u32 b = a1; b < g1;) {
            const u32 vb = (u32)(k[b] >> 32); u32 mb = 0; bool eb = false; u32 b1 = b;
            for (; b1 < g1 && (u32)(k[b1] >> 32) == vb; ++b1) { mb |= info[b1] & 7u; if ((info[b1] & 3) && (int)(info[b1] >> 3) == effbam) eb = true; }
            u32 cells = pair_cells_of(ma, mb);
            if (ea && eb) cells |= 1u << 9;
            sink.pair(((u64)va << 32) | vb, cells);
            b = b1;
          }
          a = a1;
        }
    }
    g0 = g1;
  }
  for (u32 x = ne; x < n; ++x) info[x] = 0;         // slots behind the entries: no class, inert for whoever walks the slots
  return ne;
}

// sink of the host simulation (and of fragments the device kernel cannot stage): straight to the run-wide arrays
struct DirectSink {
  u32* sz; u32* vbc; int nb; PairTable pt;
  PHZ_HD void set_size(u32 v, int x) { atomic_add(&sz[(int64_t)v * 3 + x], 1u); }
  PHZ_HD void bam_count(u32 v, u32 bam, int a) { atomic_add(&vbc[((int64_t)v * nb + bam) * 2 + a], 1u); }
  PHZ_HD void pair(u64 key, u32 cells) {
    const int s = pair_slot(pt, key);
    if (s < 0) return;
    for (int c = 0; c < PAIR_CELLS; ++c) if ((cells >> c) & 1) atomic_add(&pt.vals[(int64_t)s * PAIR_CELLS + c], 1u);
  }
};


#ifdef __CUDACC__
// counter += v on a shared-memory word, result not needed: the shared-space reduction (the pointers reach the sink as generic
// addresses, which would make these generic atomics)
#ifndef PHZ_FRAG_RED_SHARED
#define PHZ_FRAG_RED_SHARED 1
#endif
__device__ __forceinline__ void red_shared_add(u32* p, u32 v) {
#if PHZ_FRAG_RED_SHARED
  asm volatile("red.shared.add.u32 [%0], %1;" ::"r"((u32)__cvta_generic_to_shared(p)), "r"(v) : "memory");
#else
  atomicAdd(p, v);
#endif
}

template <int VB_BAMS>
struct CtaSink {
  u32* s_sz; u32* s_vb; u32 base; int nb; u32* sz; u32* vbc;
  unsigned long long* h_keys; u32* h_vals;       // shared: FRAG_HS keys, FRAG_HS * 5 words (two 16-bit counts per word)
  PairTable pt;
  __device__ __forceinline__ void set_size(u32 v, int x) {
    const u32 d = v - base;
    if (d < (u32)FRAG_W) red_shared_add(&s_sz[d * 3 + x], 1u); else atomicAdd(&sz[(int64_t)v * 3 + x], 1u);
  }
  __device__ __forceinline__ void bam_count(u32 v, u32 bam, int a) {
    const u32 d = v - base;
    if (d < (u32)FRAG_W && nb <= VB_BAMS) red_shared_add(&s_vb[(d * nb + bam) * 2 + a], 1u);
    else atomicAdd(&vbc[((int64_t)v * nb + bam) * 2 + a], 1u);
  }
  __device__ __forceinline__ void pair(u64 key, u32 cells) {
    u32 h = pair_hash(key) & (FRAG_HS - 1);
    int found = -1;
    for (int probe = 0; probe < 8; ++probe) {
      const unsigned long long kk = *(volatile unsigned long long*)&h_keys[h];
      if (kk == key) { found = (int)h; break; }
      if (kk == PAIR_EMPTY) {
        const unsigned long long old = atomicCAS(&h_keys[h], (unsigned long long)PAIR_EMPTY, (unsigned long long)key);
        if (old == PAIR_EMPTY || old == key) { found = (int)h; break; }
      }
      h = (h + 1) & (FRAG_HS - 1);
    }
    if (found >= 0) {
      while (cells) { const int c = __ffs(cells) - 1; cells &= cells - 1; red_shared_add(&h_vals[found * 5 + (c >> 1)], 1u << (16 * (c & 1))); }
    } else {        // the CTA's table is crowded here: straight to the run-wide table
      const int s = pair_slot(pt, key);
      if (s >= 0) while (cells) { const int c = __ffs(cells) - 1; cells &= cells - 1; atomicAdd(&pt.vals[(int64_t)s * PAIR_CELLS + c], 1u); }
    }
  }
};

// f_off[f] .. f_off[f + 1]: slot range of fragment f in fk / info.  f_ne[f] holds the fragment's tuple count on entry
// (0 for most fragments: those slots are never touched) and receives its number of entries.
// cnt3: [0] entries, [1] groups, [2] pairs (64-bit sums).
// Most fragments own no tuple (most reads touch no het site) and most of the others own exactly one, so the CTA first
// sorts its range into two lists in shared memory -- single-tuple fragments (no sort, no pairs: a few instructions
// each, all lanes alike) and the rest -- and its warps then draw 32 fragments at a time from a shared cursor: every
// lane of a round has work of the same kind, and a warp stuck on a deep fragment does not hold up the others.
constexpr int FRAG_RANGE = FRAG_CTA * FRAG_PER_THREAD;
static_assert(FRAG_RANGE <= 65536, "fragment indices inside a range are 16-bit, cell counts of a CTA's pair hash too");

template <bool ONE_BAM>
__global__ void __launch_bounds__(FRAG_CTA, 6) fragment_kernel(FragCtx c, const u32* __restrict__ f_off, int64_t n_frag,
                                                               u64* __restrict__ fk, uint16_t* __restrict__ info,
                                                               u32* __restrict__ f_ne, int nb, u32* __restrict__ sz,
                                                               u32* __restrict__ vbc, PairTable pt,
                                                               unsigned long long* __restrict__ cnt3) {
  constexpr int VB_BAMS = ONE_BAM ? 1 : 4;      // BAMs whose per-allele counts are staged in shared memory
  __shared__ u32 s_sz[FRAG_W * 3];
  __shared__ u32 s_vb[FRAG_W * 2 * VB_BAMS];
  __shared__ unsigned long long h_keys[FRAG_HS];
  __shared__ u32 h_vals[FRAG_HS * 5];
  __shared__ u32 s_base;
  // work lists (fragment index inside the range): single-tuple fragments from the front, the others from the back
  __shared__ uint16_t s_list[FRAG_RANGE];
  __shared__ u32 s_n1, s_nm, s_next1, s_nextm;
  const int tid = threadIdx.x, lane = tid & 31;
  const int64_t f0 = (int64_t)blockIdx.x * FRAG_RANGE;
  const int64_t f1 = (f0 + FRAG_RANGE < n_frag) ? f0 + FRAG_RANGE : n_frag;
  const u32 o_first = f_off[f0], o_last = f_off[f1];
  if (*c.abort & 8u) return;        // the host takes the sort-based stage instead
  if (o_first == o_last) return;    // no tuple in this range of fragments (f_ne aliases the tuple counts: already 0)
  for (int i = tid; i < FRAG_W * 3; i += FRAG_CTA) s_sz[i] = 0;
  for (int i = tid; i < FRAG_W * 2 * VB_BAMS; i += FRAG_CTA) s_vb[i] = 0;
  for (int i = tid; i < FRAG_HS; i += FRAG_CTA) h_keys[i] = PAIR_EMPTY;
  for (int i = tid; i < FRAG_HS * 5; i += FRAG_CTA) h_vals[i] = 0;
  if (tid == 0) {
    const u32 v0 = (u32)(fk[o_first] >> 32);
    s_base = v0 > FRAG_W / 4 ? v0 - FRAG_W / 4 : 0;
    s_n1 = 0; s_nm = 0; s_next1 = 0; s_nextm = 0;
  }
  __syncthreads();
  // ---- sort the range into the two lists (one shared-memory reduction per warp and list)
  for (int64_t b = f0 + (tid & ~31); b < f1; b += FRAG_CTA) {
    const int64_t f = b + lane;
    u32 n = 0, o0 = 0;
    if (f < f1) { o0 = f_off[f]; n = f_off[f + 1] - o0; }
    const u32 m1 = __ballot_sync(0xFFFFFFFFu, n == 1), mm = __ballot_sync(0xFFFFFFFFu, n > 1);
    u32 b1 = 0, bm = 0;
    if (lane == 0) { if (m1) b1 = atomicAdd(&s_n1, (u32)__popc(m1)); if (mm) bm = atomicAdd(&s_nm, (u32)__popc(mm)); }
    b1 = __shfl_sync(0xFFFFFFFFu, b1, 0); bm = __shfl_sync(0xFFFFFFFFu, bm, 0);
    const u32 lt = (1u << lane) - 1u;
    if (n == 1) s_list[b1 + __popc(m1 & lt)] = (uint16_t)(f - f0);
    else if (n > 1) s_list[FRAG_RANGE - 1 - (bm + __popc(mm & lt))] = (uint16_t)(f - f0);
  }
  __syncthreads();
  CtaSink<VB_BAMS> sink{s_sz, s_vb, s_base, nb, sz, vbc, h_keys, h_vals, pt};
  u32 ne_sum = 0, ng = 0, np = 0;
  const u32 n1 = s_n1, nm = s_nm;
  // ---- single-tuple fragments: one entry, one group, no pair; f_ne = 1 is the tuple count already in place
  for (;;) {
    u32 start = 0;
    if (lane == 0) start = atomicAdd(&s_next1, 32u);
    start = __shfl_sync(0xFFFFFFFFu, start, 0);
    if (start >= n1) break;
    const u32 i = start + lane;
    if (i < n1) {
      const u32 o0 = f_off[f0 + s_list[i]];
      const u64 key = fk[o0]; const u32 v = (u32)(key >> 32);
      const u32 cb = info[o0] & 0x7FFFu; const u32 cls = cb & 3, bam = cb >> 2;
      if (cls >= 2) fk[o0] = key | 0xFFFFFFFFull;
      info[o0] = (uint16_t)((bam << 3) | (1u << cls) | INFO_HEAD);
      sink.set_size(v, (int)cls);
      if (cls < 2 && !((c.excl_mask >> bam) & 1)) sink.bam_count(v, bam, (int)cls);
      ++ne_sum; ++ng;
    }
  }
  // ---- the others
  for (;;) {
    u32 start = 0;
    if (lane == 0) start = atomicAdd(&s_nextm, 32u);
    start = __shfl_sync(0xFFFFFFFFu, start, 0);
    if (start >= nm) break;
    const u32 i = start + lane;
    if (i < nm) {
      const int64_t f = f0 + s_list[FRAG_RANGE - 1 - i];
      const u32 o0 = f_off[f], n = f_off[f + 1] - o0;
      info[o0] &= 0x7FFFu;              // the head mark is not part of the tuple's class | bam
      const u32 ne = process_fragment<ONE_BAM>(c, fk + o0, info + o0, n, sink, ng, np);
      info[o0] |= INFO_HEAD;
      if (ne != n) f_ne[f] = ne;
      ne_sum += ne;
    }
  }
  // counters: registers -> warp -> one reduction per warp
  for (int o = 16; o > 0; o >>= 1) {
    ne_sum += __shfl_xor_sync(0xFFFFFFFFu, ne_sum, o); ng += __shfl_xor_sync(0xFFFFFFFFu, ng, o); np += __shfl_xor_sync(0xFFFFFFFFu, np, o);
  }
  if (lane == 0) {
    if (ne_sum) { atomicAdd(&cnt3[0], (unsigned long long)ne_sum); atomicAdd(&cnt3[1], (unsigned long long)ng); }
    if (np) atomicAdd(&cnt3[2], (unsigned long long)np);
  }
  __syncthreads();
  const u32 base = s_base;
  for (int i = tid; i < FRAG_W; i += FRAG_CTA) {
    const int64_t v = (int64_t)base + i;
    for (int x = 0; x < 3; ++x) if (s_sz[i * 3 + x]) atomicAdd(&sz[v * 3 + x], s_sz[i * 3 + x]);
    if (nb <= VB_BAMS) for (int a = 0; a < nb * 2; ++a) if (s_vb[i * nb * 2 + a]) atomicAdd(&vbc[v * nb * 2 + a], s_vb[i * nb * 2 + a]);
  }
  for (int i = tid; i < FRAG_HS; i += FRAG_CTA) {
    const unsigned long long key = h_keys[i];
    if (key == PAIR_EMPTY) continue;
    const int s = pair_slot(pt, key);
    if (s < 0) continue;
    for (int cidx = 0; cidx < PAIR_CELLS; ++cidx) {
      const u32 cv = (h_vals[i * 5 + (cidx >> 1)] >> (16 * (cidx & 1))) & 0xFFFFu;
      if (cv) atomicAdd(&pt.vals[(int64_t)s * PAIR_CELLS + cidx], cv);
    }
  }
}
#endif

#ifdef __CUDACC__
// ----------------------------------------------------------------------------- slot-chunk form (frag_stage 1, the default)
// The slots of consecutive fragments are consecutive (f_off is a scan), so a CTA can own a CHUNK OF SLOTS instead of a
// range of fragment ids: it copies its chunk of keys / info words into shared memory with coalesced loads (plus
// FRAG_EXTRA slots for the tail of its last fragment), finds the fragments from the head marks the scatter pass left in
// f_info (no look-up in the fragment table at all, and fragments without tuples cost nothing), runs the same
// process_fragment on shared memory, and writes the chunk back in one coalesced sweep.  The per-fragment dependent
// global loads (table -> key -> info) of the range form, which left it waiting on the long scoreboard, are gone; work
// per CTA is even (slots, not fragment ids).  A fragment whose tail does not fit the staged slots (more than
// FRAG_EXTRA tuples and badly placed) goes to a side list and is processed from global memory by a second, tiny pass
// once no CTA is staging any more (`deferred`: pairs of head slot and fragment id; count at deferred_n).
__device__ __forceinline__ int next_head(const u32* hb, int i, int n_st) {      // first head position > i, n_st if none
  int w = (i + 1) >> 5;
  const int nw = (n_st + 31) >> 5;
  if (w >= nw) return n_st;
  u32 m = hb[w] & (0xFFFFFFFFu << ((i + 1) & 31));
  while (!m) { if (++w >= nw) return n_st; m = hb[w]; }
  return (w << 5) + __ffs(m) - 1;
}

template <bool ONE_BAM>
__global__ void __launch_bounds__(FRAG_CTA, PHZ_FRAG_SLOT_CTAS) fragment_slots_kernel(FragCtx c, int64_t n_slots, u64* __restrict__ fk,
                                                                     uint16_t* __restrict__ info, const u32* __restrict__ gf,
                                                                     const u32* __restrict__ f_off, u32* __restrict__ f_ne, int nb,
                                                                     u32* __restrict__ sz, u32* __restrict__ vbc, PairTable pt,
                                                                     unsigned long long* __restrict__ cnt3, u32* __restrict__ deferred,
                                                                     u32* __restrict__ deferred_n) {
  constexpr int VB_BAMS = ONE_BAM ? 1 : 4;
  constexpr int SLOTS = ONE_BAM ? FRAG_SLOTS : FRAG_SLOTS * 3 / 4;
  constexpr int CAP = SLOTS + FRAG_EXTRA;
  static_assert(CAP % 32 == 0 && CAP < 32768, "staged slots: whole mark words, 15-bit positions");
  __shared__ u32 s_sz[FRAG_W * 3];
  __shared__ u32 s_vb[FRAG_W * 2 * VB_BAMS];
  __shared__ unsigned long long h_keys[FRAG_HS];
  __shared__ u32 h_vals[FRAG_HS * 5];
  __shared__ __align__(16) u64 s_k[CAP];
  __shared__ uint16_t s_i[CAP];
  __shared__ u32 s_hb[CAP / 32];           // head marks of the staged slots
  __shared__ uint16_t s_list[SLOTS];       // head positions: single-tuple fragments from the front, the others from the back
  __shared__ u32 s_base, s_n1, s_nm, s_next1, s_nextm;
  __shared__ int s_def;                    // head position of the fragment left to the second pass, -1 if none
  const int tid = threadIdx.x, lane = tid & 31;
  const int64_t s0 = (int64_t)blockIdx.x * SLOTS;
  const int n_own = (int)((n_slots - s0) < SLOTS ? (n_slots - s0) : SLOTS);
  const int n_st = (int)((n_slots - s0) < CAP ? (n_slots - s0) : CAP);
  if (*c.abort & 8u) return;        // the host takes the sort-based stage instead
  for (int i = tid; i < FRAG_W * 3; i += FRAG_CTA) s_sz[i] = 0;
  for (int i = tid; i < FRAG_W * 2 * VB_BAMS; i += FRAG_CTA) s_vb[i] = 0;
  for (int i = tid; i < FRAG_HS; i += FRAG_CTA) h_keys[i] = PAIR_EMPTY;
  for (int i = tid; i < FRAG_HS * 5; i += FRAG_CTA) h_vals[i] = 0;
  // ---- stage the chunk (whole warps, so that every mark word is one ballot)
  for (int b = tid & ~31; b < n_st; b += FRAG_CTA) {
    const int i = b + lane;
    u32 w = 0;
    if (i < n_st) { s_k[i] = fk[s0 + i]; w = info[s0 + i]; s_i[i] = (uint16_t)(w & 0x7FFFu); }
    const u32 marks = __ballot_sync(0xFFFFFFFFu, (w & INFO_HEAD) != 0);
    if (lane == 0) s_hb[b >> 5] = marks;
  }
  if (tid == 0) { s_n1 = 0; s_nm = 0; s_next1 = 0; s_nextm = 0; s_def = -1; }
  __syncthreads();
  if (tid == 0) {
    const u32 v0 = (u32)(s_k[0] >> 32);
    s_base = v0 > FRAG_W / 4 ? v0 - FRAG_W / 4 : 0;
  }
  // ---- fragments whose head lies in the chunk: two work lists (one shared-memory reduction per warp and list)
  for (int b = tid & ~31; b < n_own; b += FRAG_CTA) {
    const int i = b + lane;
    u32 n = 0;
    if (i < n_own && ((s_hb[i >> 5] >> (i & 31)) & 1u)) {
      const int e = next_head(s_hb, i, n_st);
      if (e == n_st && s0 + n_st < n_slots) {
        // the tail is not staged: side list (at most one fragment per CTA: the chunk's last)
        const u32 d = atomicAdd(deferred_n, 1u);
        deferred[2 * d] = (u32)(s0 + i); deferred[2 * d + 1] = gf[(u32)s_k[i]];
        s_def = i;                  // slots from here on stay as they are
        n = 0xFFFFFFFFu;
      } else n = (u32)(e - i);
    }
    const bool one = n == 1, many = n > 1 && n != 0xFFFFFFFFu;
    const u32 m1 = __ballot_sync(0xFFFFFFFFu, one), mm = __ballot_sync(0xFFFFFFFFu, many);
    u32 b1 = 0, bm = 0;
    if (lane == 0) { if (m1) b1 = atomicAdd(&s_n1, (u32)__popc(m1)); if (mm) bm = atomicAdd(&s_nm, (u32)__popc(mm)); }
    b1 = __shfl_sync(0xFFFFFFFFu, b1, 0); bm = __shfl_sync(0xFFFFFFFFu, bm, 0);
    const u32 lt = (1u << lane) - 1u;
    if (one) s_list[b1 + __popc(m1 & lt)] = (uint16_t)i;
    else if (many) s_list[SLOTS - 1 - (bm + __popc(mm & lt))] = (uint16_t)i;
  }
  __syncthreads();
  CtaSink<VB_BAMS> sink{s_sz, s_vb, s_base, nb, sz, vbc, h_keys, h_vals, pt};
  u32 ne_sum = 0, ng = 0, np = 0;
  const u32 n1 = s_n1, nm = s_nm;
  // ---- single-tuple fragments: one entry, one group, no pair
  for (;;) {
    u32 start = 0;
    if (lane == 0) start = atomicAdd(&s_next1, 32u);
    start = __shfl_sync(0xFFFFFFFFu, start, 0);
    if (start >= n1) break;
    const u32 i = start + lane;
    if (i < n1) {
      const int p = s_list[i];
      const u64 key = s_k[p]; const u32 v = (u32)(key >> 32);
      const u32 cb = s_i[p]; const u32 cls = cb & 3, bam = cb >> 2;
      if (cls >= 2) s_k[p] = key | 0xFFFFFFFFull;
      s_i[p] = (uint16_t)((bam << 3) | (1u << cls));
      sink.set_size(v, (int)cls);
      if (cls < 2 && !((c.excl_mask >> bam) & 1)) sink.bam_count(v, bam, (int)cls);
      ++ne_sum; ++ng;
    }
  }
  // ---- the others
  for (;;) {
    u32 start = 0;
    if (lane == 0) start = atomicAdd(&s_nextm, 32u);
    start = __shfl_sync(0xFFFFFFFFu, start, 0);
    if (start >= nm) break;
    const u32 i = start + lane;
    if (i < nm) {
      const int p = s_list[SLOTS - 1 - i];
      const u32 n = (u32)(next_head(s_hb, p, n_st) - p);
      const u32 t0 = (u32)s_k[p];                 // any tuple of the fragment names it
      const u32 ne = process_fragment<ONE_BAM>(c, s_k + p, s_i + p, n, sink, ng, np);
      if (ne != n) f_ne[gf[t0]] = ne;
      ne_sum += ne;
    }
  }
  // counters: registers -> warp -> one reduction per warp
  for (int o = 16; o > 0; o >>= 1) {
    ne_sum += __shfl_xor_sync(0xFFFFFFFFu, ne_sum, o); ng += __shfl_xor_sync(0xFFFFFFFFu, ng, o); np += __shfl_xor_sync(0xFFFFFFFFu, np, o);
  }
  if (lane == 0) {
    if (ne_sum) { atomicAdd(&cnt3[0], (unsigned long long)ne_sum); atomicAdd(&cnt3[1], (unsigned long long)ng); }
    if (np) atomicAdd(&cnt3[2], (unsigned long long)np);
  }
  __syncthreads();
  // ---- write the chunk back: from its first head to the end of its last fragment (a deferred fragment stays untouched)
  {
    int first = n_st;
    for (int w = 0; w < (n_own + 31) >> 5; ++w) if (s_hb[w]) { first = (w << 5) + __ffs(s_hb[w]) - 1; break; }
    int last_head = -1;
    for (int w = ((n_own + 31) >> 5) - 1; w >= 0; --w) {
      u32 m = s_hb[w];
      if (w == (n_own >> 5) && (n_own & 31)) m &= (1u << (n_own & 31)) - 1u;
      if (m) { last_head = (w << 5) + 31 - __clz(m); break; }
    }
    int end = 0;
    if (last_head >= 0) end = s_def >= 0 ? s_def : next_head(s_hb, last_head, n_st);
    for (int i = first + tid; i < end; i += FRAG_CTA) {
      fk[s0 + i] = s_k[i];
      info[s0 + i] = (uint16_t)(s_i[i] | (((s_hb[i >> 5] >> (i & 31)) & 1u) ? INFO_HEAD : 0));
    }
  }
  const u32 base = s_base;
  for (int i = tid; i < FRAG_W; i += FRAG_CTA) {
    const int64_t v = (int64_t)base + i;
    for (int x = 0; x < 3; ++x) if (s_sz[i * 3 + x]) atomicAdd(&sz[v * 3 + x], s_sz[i * 3 + x]);
    if (nb <= VB_BAMS) for (int a = 0; a < nb * 2; ++a) if (s_vb[i * nb * 2 + a]) atomicAdd(&vbc[v * nb * 2 + a], s_vb[i * nb * 2 + a]);
  }
  for (int i = tid; i < FRAG_HS; i += FRAG_CTA) {
    const unsigned long long key = h_keys[i];
    if (key == PAIR_EMPTY) continue;
    const int s = pair_slot(pt, key);
    if (s < 0) continue;
    for (int cidx = 0; cidx < PAIR_CELLS; ++cidx) {
      const u32 cv = (h_vals[i * 5 + (cidx >> 1)] >> (16 * (cidx & 1))) & 0xFFFFu;
      if (cv) atomicAdd(&pt.vals[(int64_t)s * PAIR_CELLS + cidx], cv);
    }
  }
}
#endif

}  // namespace phz

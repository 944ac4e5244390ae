// Per-block phasing logic (one logical thread per haplotype block).
//
// Replaces phase_v3 and its helpers of the reference (phaser/phaser.py:2107-2324):
//   resolve_phase :2172-2207, sub_block_phase :2209-2258, inverse_conifg :2260-2269,
//   split_by_weak :2271-2294, split_variants :2296-2307, find_weak_points :2309-2324.
// Strings over {0,1,-} become small integer arrays; allele links become signed edges
// (0 = cis: a:0-b:0 & a:1-b:1, 1 = trans, 2 = tie: keys exist, no links, phaser.py:708-726).
// Reference behaviours kept on purpose: the seed is variants[0]:0; a conflicted component that
// happens to reach exactly n alleles "resolves" into a SHORT all-zero string; sub-block merges slice
// `variants[split_start : split_start+used]` with split_start = used after a failed merge (Q14);
// blocks whose first allele is '-' are dropped.
#pragma once
#include "phz_backend.h"

namespace phz {

// A block is phased by `n` cooperating lanes (one warp on the device, a single lane in the host
// simulation): lane 0 runs the serial control flow, all lanes share the 2^n enumeration.
struct Coop { int lane; int n; };
PHZ_HD void coop_sync(const Coop& c) {
#if defined(__CUDA_ARCH__)
  if (c.n > 1) __syncwarp();
#endif
  (void)c;
}
PHZ_HD int coop_bcast(const Coop& c, int v) {
#if defined(__CUDA_ARCH__)
  if (c.n > 1) return __shfl_sync(0xffffffffu, v, 0);
#endif
  (void)c;
  return v;
}

// any / sum over the cooperating lanes (every lane gets the result; includes the memory ordering of coop_sync)
PHZ_HD bool coop_any(const Coop& c, bool v) {
#if defined(__CUDA_ARCH__)
  if (c.n > 1) return __any_sync(0xffffffffu, v);
#endif
  (void)c;
  return v;
}
PHZ_HD int coop_sum(const Coop& c, int v) {
#if defined(__CUDA_ARCH__)
  if (c.n > 1) { for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); }
#endif
  (void)c;
  return v;
}

constexpr u8 CH_DASH = 2;
constexpr int EDGE_CIS = 0, EDGE_TRANS = 1, EDGE_TIE = 2;
constexpr int MAX_ENUM = 24;

struct BlockEdges {
  // edges of ONE block, local variant indices (position in the sorted member list)
  const u32* list;      // indices into the edge arrays, [n_edges]
  u32 n_edges;
  const u32* ed_a;      // global variant ids
  const u32* ed_b;
  const u8* ed_cfg;     // EDGE_*
  const u32* pos_in_blk;
  PHZ_HD void get(u32 k, int& i, int& j, int& s) const {
    u32 e = list[k];
    i = (int)pos_in_blk[ed_a[e]]; j = (int)pos_in_blk[ed_b[e]]; s = ed_cfg[e];
  }
};

// Reach of allele (lo:0) over allele links restricted to local variants [lo, hi).
// color[i-lo] in {0,1} for reached variants, 0xFF otherwise.  Returns the number of reached
// VARIANTS; *conflict is set when both alleles of some variant are reachable (then every reached
// variant has both alleles reachable, so the reference's reach set has 2*m alleles).
PHZ_HD int reach_range(const BlockEdges& be, int lo, int hi, u8* color, bool* conflict, const Coop& cp) {
  // The cooperating lanes share the edge list (edge k -> lane k mod n): label propagation until nothing changes.  The
  // reached SET is a closure, independent of the order of propagation; a conflict (an edge between two reached variants
  // whose colours contradict its sign) exists at the fixed point iff the component has no consistent 2-colouring, again
  // whatever the order -- so the parallel sweep returns what the serial one does.
  volatile u8* col = color;
  for (int i = lo + cp.lane; i < hi; i += cp.n) col[i - lo] = 0xFF;
  coop_sync(cp);
  if (cp.lane == 0) col[0] = 0;
  coop_sync(cp);
  bool changed = true;
  while (changed) {
    bool ch = false;
    for (u32 k = (u32)cp.lane; k < be.n_edges; k += (u32)cp.n) {
      int i, j, s; be.get(k, i, j, s);
      if (s == EDGE_TIE || i < lo || i >= hi || j < lo || j >= hi) continue;
      u8 ci = col[i - lo], cj = col[j - lo];
      if (ci != 0xFF && cj == 0xFF) { col[j - lo] = ci ^ (u8)s; ch = true; }
      else if (ci == 0xFF && cj != 0xFF) { col[i - lo] = cj ^ (u8)s; ch = true; }
    }
    coop_sync(cp);
    changed = coop_any(cp, ch);
  }
  int m = 0; bool conf = false;
  for (int i = lo + cp.lane; i < hi; i += cp.n) if (col[i - lo] != 0xFF) m++;
  for (u32 k = (u32)cp.lane; k < be.n_edges; k += (u32)cp.n) {
    int i, j, s; be.get(k, i, j, s);
    if (s == EDGE_TIE || i < lo || i >= hi || j < lo || j >= hi) continue;
    u8 ci = col[i - lo], cj = col[j - lo];
    if (ci != 0xFF && cj != 0xFF && ((ci ^ cj) != (u8)s)) conf = true;
  }
  *conflict = coop_any(cp, conf);
  return coop_sum(cp, m);
}

// resolve_phase on local range [lo, hi): writes the string into out (chars 0/1), returns its
// length, or -1 when the reference returns None.
// (called by every cooperating lane; the return value is the same on all of them)
PHZ_HD int resolve_range(const BlockEdges& be, int lo, int hi, u8* color, u8* out, const Coop& cp) {
  bool conf;
  int n = hi - lo;
  int m = reach_range(be, lo, hi, color, &conf, cp);
  int len = -1;
  if (!conf && m == n) { for (int i = cp.lane; i < n; i += cp.n) out[i] = color[i]; len = n; }
  else if (conf && 2 * m == n) { for (int i = cp.lane; i < m; i += cp.n) out[i] = 0; len = m; }   // short string quirk
  coop_sync(cp);
  return len;
}

// 2^n enumeration of sub_block_phase on local range [lo, hi), shared by the cooperating lanes.
// Returns length n (on every lane); out is the unique best configuration or all dashes (written by
// lane 0).  n > MAX_ENUM sets *err.
PHZ_HD int enumerate_range(const BlockEdges& be, int lo, int hi, u8* out, int* err, const Coop& cp) {
  int n = hi - lo;
  if (n > MAX_ENUM) {
    if (cp.lane == 0) { *err |= 1; for (int i = 0; i < n; ++i) out[i] = CH_DASH; }
    coop_sync(cp);
    return n;
  }
  u32 cis[MAX_ENUM], trans[MAX_ENUM];
  for (int i = 0; i < n; ++i) { cis[i] = 0; trans[i] = 0; }
  for (u32 k = 0; k < be.n_edges; ++k) {
    int i, j, s; be.get(k, i, j, s);
    if (s == EDGE_TIE || i < lo || i >= hi || j < lo || j >= hi) continue;
    i -= lo; j -= lo;
    if (s == EDGE_CIS) { cis[i] |= 1u << j; cis[j] |= 1u << i; }
    else { trans[i] |= 1u << j; trans[j] |= 1u << i; }
  }
  // variant i <-> bit i; only configurations with allele 0 at variant 0 are scored (the complement of
  // every other one was scored earlier in lexicographic order, phaser.py:2226-2234)
  u32 full = (1u << n) - 1;
  int best = -1; u32 best_w = 0; u32 n_best = 0;
  u32 total = 1u << (n - 1);
  for (u32 y = (u32)cp.lane; y < total; y += (u32)cp.n) {
    u32 w = y << 1;
    int s = 0;
    for (int i = 0; i < n; ++i) {
      u32 eq = ((w >> i) & 1u) ? w : ~w;       // bit j set <=> allele_j == allele_i
#if defined(__CUDA_ARCH__)
      s += __popc(cis[i] & eq & full) + __popc(trans[i] & ~eq & full);
#else
      s += __builtin_popcount(cis[i] & eq & full) + __builtin_popcount(trans[i] & ~eq & full);
#endif
    }
    if (s > best) { best = s; best_w = w; n_best = 1; }
    else if (s == best) n_best++;
  }
#if defined(__CUDA_ARCH__)
  if (cp.n > 1) {
    int gbest = best;
    for (int o = 16; o > 0; o >>= 1) { int t = __shfl_xor_sync(0xffffffffu, gbest, o); gbest = t > gbest ? t : gbest; }
    u32 mine = (best == gbest) ? n_best : 0u;
    u32 cnt = mine;
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    unsigned holders = __ballot_sync(0xffffffffu, best == gbest);
    u32 w0 = __shfl_sync(0xffffffffu, best_w, __ffs(holders) - 1);
    best = gbest; n_best = cnt; best_w = w0;
  }
#endif
  if (cp.lane == 0) {
    if (n_best == 1) { for (int i = 0; i < n; ++i) out[i] = (best_w >> i) & 1u; }
    else { for (int i = 0; i < n; ++i) out[i] = CH_DASH; }
  }
  coop_sync(cp);
  return n;
}

// support of a configuration string laid over variants[start : start+len) (zip truncation at n)
// (cooperative: the lanes share the edge list, every lane gets the sum)
PHZ_HD int score_config(const BlockEdges& be, int n, int start, const u8* cfg, int len, const Coop& cp) {
  int hi = start + len; if (hi > n) hi = n;
  int s = 0;
  for (u32 k = (u32)cp.lane; k < be.n_edges; k += (u32)cp.n) {
    int i, j, sg; be.get(k, i, j, sg);
    if (sg == EDGE_TIE || i < start || i >= hi || j < start || j >= hi) continue;
    u8 ci = cfg[i - start], cj = cfg[j - start];
    if (ci == CH_DASH || cj == CH_DASH) continue;
    if ((ci ^ cj) == (u8)sg) s += 2;
  }
  return coop_sum(cp, s);
}

PHZ_HD void close_final_block(const u8* str, int len, int n, int* consumed, int* n_runs,
                              u32* run_start, u32* run_len, u8* hap, u32* fin_local) {
  // phaser.py:2160-2168: strings are laid over the sorted variants one after another; a block whose
  // first allele is '-' is dropped (its variants fall back to singletons)
  if (len > 0 && str[0] != CH_DASH) {
    run_start[*n_runs] = (u32)(*consumed); run_len[*n_runs] = (u32)len;
    for (int i = 0; i < len && *consumed + i < n; ++i) { hap[*consumed + i] = str[i]; fin_local[*consumed + i] = (u32)(*n_runs); }
    (*n_runs)++;
  }
  *consumed += len;
}

// 6n+16 words for phase_block_hard itself, n more for the caller's fin_local view
PHZ_HD size_t hard_scratch_words(size_t n) { return 7 * n + 16; }

// Whole phase_v3 for a block of n variants that failed the fast path.  Scratch `w` has
// hard_scratch_words(n) u32 words.  Output: runs (run_start[k], run_len[k]) of consecutive local
// variants that form the final blocks (k < returned count), hap[i] = allele of haplotype A for local
// variant i (meaningful inside runs), fin_local[i] = run index (caller pre-fills 0xFFFFFFFF).
// Called by every cooperating lane; the return value is the same on all of them.
PHZ_HD int phase_block_hard(const BlockEdges& be, int n, int max_block_size, u32* w,
                            u32* run_start, u32* run_len, u8* hap, u32* fin_local, int* err, const Coop& cp) {
  int* cnt = (int*)w;                       // [n+1] crossing counts
  u32* sub_off = w + (n + 1);               // [n+1] variant offsets of sub-blocks
  u32* sub_len = sub_off + (n + 1);         // [n]   string length of each sub-block phase
  u32* sub_pos = sub_len + n;               // [n]   offset of each string inside `sub`
  u8* chosen = (u8*)(sub_pos + n);          // [n+1]
  u8* color = chosen + (n + 1);             // [n]
  u8* sub = color + n;                      // [n] concatenated sub-block phase strings
  u8* fin = sub + n;                        // [n] current final_phase string
  u8* cand = fin + n;                       // [n] candidate
  int n_sub = 0;
  if (cp.lane == 0) {
    // ---- find_weak_points (phaser.py:2309-2324): edge (i<j) crosses cut p iff i < p <= j
    for (int p = 0; p <= n; ++p) { cnt[p] = 0; chosen[p] = 0; }
    for (u32 k = 0; k < be.n_edges; ++k) {
      int i, j, s; be.get(k, i, j, s);
      if (i > j) { int t = i; i = j; j = t; }
      if (i == j) continue;
      cnt[i + 1] += 1; cnt[j + 1] -= 1;      // j <= n-1
    }
    for (int p = 1; p <= n; ++p) cnt[p] += cnt[p - 1];
    // ---- split_by_weak (phaser.py:2271-2294)
    int xmax = (max_block_size == 0) ? n : max_block_size;
    int split_at = 1;
    while (true) {
      for (int p = 2; p <= n - 2; ++p)
        if (cnt[p] == split_at && !chosen[p + 1] && !chosen[p - 1]) chosen[p] = 1;
      int last = 0, max_frag = 0;
      for (int p = 1; p <= n; ++p)
        if (p == n || chosen[p]) { if (p - last > max_frag) max_frag = p - last; last = p; }
      split_at++;
      if (!(max_frag > xmax)) break;
      int next = 0x7FFFFFFF;       // levels without positions change nothing: jump to the next one
      for (int p = 2; p <= n - 2; ++p) if (cnt[p] >= split_at && cnt[p] < next) next = cnt[p];
      if (next == 0x7FFFFFFF) { *err |= 2; break; }     // the reference would loop forever here
      split_at = next;
    }
    int last = 0;
    for (int p = 1; p <= n; ++p) if (p == n || chosen[p]) { sub_off[n_sub++] = (u32)last; last = p; }
    sub_off[n_sub] = (u32)n;
  }
  n_sub = coop_bcast(cp, n_sub);
  coop_sync(cp);
  // ---- phase every sub-block (phaser.py:2133-2136)
  u32 used_chars = 0;
  for (int s = 0; s < n_sub; ++s) {
    int lo = (int)sub_off[s], hi = (int)sub_off[s + 1];
    int len = -1;
    if (n_sub > 1) len = resolve_range(be, lo, hi, color, sub + used_chars, cp);
    if (len < 0) len = enumerate_range(be, lo, hi, sub + used_chars, err, cp);
    if (cp.lane == 0) { sub_pos[s] = used_chars; sub_len[s] = (u32)len; }
    used_chars += (u32)len;
  }
  coop_sync(cp);
  // ---- merge left to right (phaser.py:2140-2157).  A and B are each all digits or all dashes, so
  // of the four concatenations only A+B and A+B' are scored (the other two are their complements,
  // phaser.py:2234) and the merge succeeds iff both are digits and the two supports differ.
  int n_runs = 0;
  {
    // every lane runs the control flow (its scalars depend only on data all lanes read after a sync); lane 0 alone
    // writes the strings and the runs; the two supports of a merge are summed by all lanes over the shared edge list
    int consumed = 0;
    int fin_len = (int)sub_len[0];
    if (cp.lane == 0) for (int i = 0; i < fin_len; ++i) fin[i] = sub[sub_pos[0] + i];
    coop_sync(cp);
    int split_start = 0;
    for (int s = 1; s < n_sub; ++s) {
      const u8* nb = sub + sub_pos[s];
      int nb_len = (int)sub_len[s];
      int used = fin_len + nb_len;
      bool a_dash = (fin_len > 0 && fin[0] == CH_DASH), b_dash = (nb_len > 0 && nb[0] == CH_DASH);
      bool ok = false, flip_b = false;
      if (!a_dash && !b_dash) {
        if (cp.lane == 0) {
          for (int i = 0; i < fin_len; ++i) cand[i] = fin[i];
          for (int i = 0; i < nb_len; ++i) cand[fin_len + i] = nb[i];
        }
        coop_sync(cp);
        int s0 = score_config(be, n, split_start, cand, used, cp);
        coop_sync(cp);
        if (cp.lane == 0) for (int i = 0; i < nb_len; ++i) cand[fin_len + i] = nb[i] ^ 1;
        coop_sync(cp);
        int s1 = score_config(be, n, split_start, cand, used, cp);
        if (s0 > s1) { ok = true; flip_b = false; } else if (s1 > s0) { ok = true; flip_b = true; }
      }
      coop_sync(cp);
      if (ok) {
        if (cp.lane == 0) for (int i = 0; i < nb_len; ++i) fin[fin_len + i] = flip_b ? (u8)(nb[i] ^ 1) : nb[i];
        fin_len = used;
      } else {
        if (cp.lane == 0) close_final_block(fin, fin_len, n, &consumed, &n_runs, run_start, run_len, hap, fin_local);
        split_start = used;                 // Q14: not an offset sum
        fin_len = nb_len;
        if (cp.lane == 0) for (int i = 0; i < nb_len; ++i) fin[i] = nb[i];
      }
      coop_sync(cp);
    }
    if (cp.lane == 0) close_final_block(fin, fin_len, n, &consumed, &n_runs, run_start, run_len, hap, fin_local);
  }
  n_runs = coop_bcast(cp, n_runs);
  coop_sync(cp);
  return n_runs;
}

}  // namespace phz

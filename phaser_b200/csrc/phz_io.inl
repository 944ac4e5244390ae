// Native host ingest: BAM (BGZF) or SAM text -> the packed SoA layout of include/phz.h.
//
// Replaces the two `samtools view` stages of the reference pipeline and the per-line parsing of its
// mapper (phaser/phaser.py:1346, read_variant_map.py:25-64) -- SURVEY.md section 8f row N2.  BGZF
// blocks are inflated in parallel (zlib, one raw-deflate stream per block), records are decoded
// straight into the device layout (BAM's 4-bit bases / len<<4|op CIGAR words / phred bytes ARE that
// layout), filters as the reference's samtools arguments (phaser.py:505-513): contig named by the
// VCF, -F 0x400 iff remove_dups, -f 2 iff paired_end, -q MAPQ.  Fragment ids: one per distinct QNAME,
// shared by all BAMs of a run (exact: names are compared, not just hashed).
#include <zlib.h>
#include <array>
#include <thread>
#include <mutex>
#include <memory>
#include <exception>
#include <atomic>
#include <fstream>
#include <chrono>
#include <cstdlib>
#include <sys/mman.h>

namespace phzio {

using phz::u64; using phz::u32; using phz::u8; using phz::PhzError;

template <class F>
static void parallel_for(size_t n, int n_threads, F f) {      // f(i) for i in [0, n), dynamic schedule
  if (n_threads < 1) n_threads = 1;
  if ((size_t)n_threads > n) n_threads = (int)(n ? n : 1);
  std::atomic<size_t> next(0);
  std::exception_ptr err; std::mutex em;
  auto work = [&]() {
    try { while (true) { size_t i = next.fetch_add(1); if (i >= n) break; f(i); } }
    catch (...) { std::lock_guard<std::mutex> g(em); if (!err) err = std::current_exception(); next = n; }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < n_threads; ++t) th.emplace_back(work);
  work();
  for (auto& t : th) t.join();
  if (err) std::rethrow_exception(err);
}

// vector whose resize() leaves new elements uninitialised: a multi-hundred-megabyte buffer that is about to be filled by
// parallel threads must not be zero-filled (and its pages first-touched) by ONE thread beforehand
template <class T>
struct NoInit : std::allocator<T> {
  template <class U> struct rebind { typedef NoInit<U> other; };
  NoInit() = default;
  template <class U> NoInit(const NoInit<U>&) {}
  template <class U, class... A> void construct(U* p, A&&... a) {
    if constexpr (sizeof...(A) == 0) ::new ((void*)p) U; else ::new ((void*)p) U(std::forward<A>(a)...);
  }
  // big arrays (inflated BAM, the SoA columns) are first touched by many threads at once: 2 MB pages cut the number of
  // page faults -- and the time the threads spend queueing on the address-space lock -- by 512
  static constexpr size_t HUGE = (size_t)2 << 20;
  T* allocate(size_t n) {
    const size_t bytes = n * sizeof(T);
    if (bytes < 4 * HUGE) return std::allocator<T>::allocate(n);
    const size_t len = (bytes + HUGE - 1) / HUGE * HUGE;
    void* p = std::aligned_alloc(HUGE, len);
    if (!p) throw std::bad_alloc();
    madvise(p, len, MADV_HUGEPAGE);
    return (T*)p;
  }
  void deallocate(T* p, size_t n) {
    if (n * sizeof(T) < 4 * HUGE) std::allocator<T>::deallocate(p, n); else std::free(p);
  }
};
typedef std::vector<u8, NoInit<u8>> Bytes;

// read-only view of a whole file: mapped, not copied (SAM text is parsed in place, BGZF blocks are inflated from it)
static bool read_file(const char* path, std::vector<u8>& out);
struct FileMap {
  const u8* p = nullptr; size_t n = 0; bool mapped = false; std::vector<u8> fallback;
  const u8* data() const { return p; }
  size_t size() const { return n; }
  u8 operator[](size_t i) const { return p[i]; }
  bool open(const char* path);
  ~FileMap();
};

static u64 name_hash(const char* s, size_t n) {
  u64 h = 1469598103934665603ull;
  for (size_t i = 0; i < n; ++i) { h ^= (u8)s[i]; h *= 1099511628211ull; }
  h ^= h >> 29; h *= 0xbf58476d1ce4e5b9ull; h ^= h >> 32;
  return h;
}

// QNAME -> dense fragment id.  64 independent shards (by hash) so that one BAM's names are inserted by
// many threads at once; ids are handed out in order of first appearance in the file, which keeps the
// fragments of one locus numerically close (the device code aggregates on that).
struct FragDict {
  static constexpr int NS = 64;
  struct Shard {
    std::vector<u64> slot_hash; std::vector<u32> slot_ent;       // open addressing -> entry index
    std::vector<u64> name_off; std::string arena;                // entry -> name
    std::vector<u32> gid; std::vector<u64> first_rec;            // entry -> global id / first record of the current file
    size_t mask = 0;
    Shard() { name_off.push_back(0); resize(1 << 12); }
    // room for `extra` more names (an upper bound: every record of the bucket could be a new name)
    void reserve_for(size_t extra) {
      const size_t want = gid.size() + extra;
      size_t slots = mask + 1; while (slots < want * 2 + 2) slots <<= 1;
      if (slots != mask + 1) resize(slots);
      gid.reserve(want); first_rec.reserve(want); name_off.reserve(want + 1);
      const size_t per = gid.empty() ? 24 : arena.size() / gid.size() + 1;
      arena.reserve(arena.size() + extra * per / 2);
    }
    void resize(size_t n) {
      std::vector<u64> oh(n, 0); std::vector<u32> oe(n, 0xFFFFFFFFu);
      for (size_t i = 0; i < slot_hash.size(); ++i)
        if (slot_ent[i] != 0xFFFFFFFFu) { size_t p = slot_hash[i] & (n - 1); while (oe[p] != 0xFFFFFFFFu) p = (p + 1) & (n - 1); oh[p] = slot_hash[i]; oe[p] = slot_ent[i]; }
      slot_hash.swap(oh); slot_ent.swap(oe); mask = n - 1;
    }
    // entry index of the name, inserting it (gid unknown yet, first seen at record `rec`) when new
    u32 find_or_add(u64 h, const char* s, size_t n, u64 rec) {
      size_t p = h & mask;
      while (slot_ent[p] != 0xFFFFFFFFu) {
        if (slot_hash[p] == h) {
          u32 e = slot_ent[p]; u64 o = name_off[e];
          if ((size_t)(name_off[e + 1] - o) == n && std::memcmp(arena.data() + o, s, n) == 0) return e;
        }
        p = (p + 1) & mask;
      }
      u32 e = (u32)gid.size();
      arena.append(s, n); name_off.push_back(arena.size());
      gid.push_back(0xFFFFFFFFu); first_rec.push_back(rec);
      slot_hash[p] = h; slot_ent[p] = e;
      if (gid.size() * 2 > mask + 1) resize((mask + 1) * 2);
      return e;
    }
  };
  Shard sh[NS];
  std::vector<u8> id_shard; std::vector<u32> id_ent;             // global id -> (shard, entry)
  size_t count = 0;
  static int shard_of(u64 h) { return (int)(h >> 58); }

  // ids for the kept records of one file (names[i], lens[i], hashes[i] in file order) -> out[i]
  void assign(const char* const* names, const u32* lens, const u64* hashes, size_t n, u32* out, int n_threads) {
    std::vector<u32, NoInit<u32>> ent(n);
    // 1. record indices bucketed by shard (counting sort over fixed chunks: file order is kept inside a bucket)
    const size_t CH = 1 << 16, nch = (n + CH - 1) / CH;
    std::vector<u32> cnt(nch * NS + 1, 0);
    parallel_for(nch, n_threads, [&](size_t c) {
      u32* k = cnt.data() + c * NS;
      const size_t i1 = std::min(n, (c + 1) * CH);
      for (size_t i = c * CH; i < i1; ++i) k[shard_of(hashes[i])]++;
    });
    std::vector<size_t> start(nch * NS + 1, 0), sh_begin(NS + 1, 0);
    { size_t acc = 0;
      for (int s = 0; s < NS; ++s) { sh_begin[s] = acc; for (size_t c = 0; c < nch; ++c) { start[c * NS + s] = acc; acc += cnt[c * NS + s]; } }
      sh_begin[NS] = acc; }
    std::vector<u32, NoInit<u32>> order(n);
    parallel_for(nch, n_threads, [&](size_t c) {
      size_t pos[NS];
      for (int s = 0; s < NS; ++s) pos[s] = start[c * NS + s];
      const size_t i1 = std::min(n, (c + 1) * CH);
      for (size_t i = c * CH; i < i1; ++i) order[pos[shard_of(hashes[i])]++] = (u32)i;
    });
    // 2. every shard looks its own records up (in file order: first_rec is the first record of a new name)
    parallel_for(NS, n_threads, [&](size_t s) {
      Shard& S = sh[s];
      S.reserve_for(sh_begin[s + 1] - sh_begin[s]);          // no rehash / regrowth in the middle of the file
      for (size_t k = sh_begin[s]; k < sh_begin[s + 1]; ++k) { const u32 i = order[k]; ent[i] = S.find_or_add(hashes[i], names[i], lens[i], (u64)i); }
    });
    // 3. new names get ids in order of first appearance: flag the first record of every new name, scan, hand out
    std::vector<u8, NoInit<u8>> pre(n + 1);
    std::vector<size_t> chunk_new(nch + 1, 0);
    parallel_for(nch, n_threads, [&](size_t c) {
      const size_t i1 = std::min(n, (c + 1) * CH); size_t k = 0;
      for (size_t i = c * CH; i < i1; ++i) {
        const Shard& S = sh[shard_of(hashes[i])]; const u32 e = ent[i];
        const bool fresh = S.gid[e] == 0xFFFFFFFFu && S.first_rec[e] == (u64)i;
        pre[i] = fresh ? 1 : 0; k += fresh;
      }
      chunk_new[c + 1] = k;
    });
    for (size_t c = 0; c < nch; ++c) chunk_new[c + 1] += chunk_new[c];
    const size_t base = count, fresh_total = chunk_new[nch];
    id_shard.resize(base + fresh_total); id_ent.resize(base + fresh_total);
    parallel_for(nch, n_threads, [&](size_t c) {
      const size_t i1 = std::min(n, (c + 1) * CH); size_t g = base + chunk_new[c];
      for (size_t i = c * CH; i < i1; ++i)
        if (pre[i]) {
          const int s = shard_of(hashes[i]); const u32 e = ent[i];
          sh[s].gid[e] = (u32)g; id_shard[g] = (u8)s; id_ent[g] = e; ++g;      // one writer per entry: its first record
        }
    });
    count = base + fresh_total;
    parallel_for(nch, n_threads, [&](size_t c) {
      const size_t i1 = std::min(n, (c + 1) * CH);
      for (size_t i = c * CH; i < i1; ++i) out[i] = sh[shard_of(hashes[i])].gid[ent[i]];
    });
  }
  u32 get(const char* s, size_t n) {      // single lookup (tests / small inputs)
    const char* a = s; u32 l = (u32)n; u64 h = name_hash(s, n); u32 o = 0;
    assign(&a, &l, &h, 1, &o, 1);
    return o;
  }
};

struct HostReads {
  int n_contigs = 0;
  std::vector<int64_t> contig_rec_off;
  std::vector<int32_t, NoInit<int32_t>> pos, tlen; std::vector<int16_t, NoInit<int16_t>> aln; std::vector<u32, NoInit<u32>> frag, cigar;
  std::vector<u32, NoInit<u32>> cigar_off; std::vector<u64, NoInit<u64>> seq_off; Bytes seq; Bytes qual;
  int sorted = 1;
  std::string error;
};

static bool read_file(const char* path, std::vector<u8>& out) {
  std::ifstream f(path, std::ios::binary | std::ios::ate);
  if (!f) return false;
  std::streamsize n = f.tellg(); f.seekg(0);
  out.resize((size_t)n);
  return n == 0 || (bool)f.read((char*)out.data(), n);
}

}  // namespace phzio
#include <sys/mman.h>
#include <sys/stat.h>
#include <fcntl.h>
#include <unistd.h>
namespace phzio {
inline bool FileMap::open(const char* path) {
  int fd = ::open(path, O_RDONLY);
  if (fd < 0) return false;
  struct stat st;
  if (fstat(fd, &st) != 0) { ::close(fd); return false; }
  n = (size_t)st.st_size;
  if (n == 0) { ::close(fd); p = (const u8*)""; return true; }
  void* m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
  ::close(fd);
  if (m == MAP_FAILED) { if (!read_file(path, fallback)) return false; p = fallback.data(); n = fallback.size(); return true; }
  p = (const u8*)m; mapped = true;
  return true;
}
inline FileMap::~FileMap() { if (mapped && p) munmap((void*)p, n); }

// ---- BGZF: index the blocks, inflate them in parallel into one buffer
template <class In, class Out>
static void inflate_bgzf(const In& in, Out& out, int n_threads) {
  struct Blk { size_t in_off, in_len, out_off, out_len; };
  std::vector<Blk> blks;
  size_t p = 0, total = 0;
  while (p + 18 <= in.size()) {
    if (in[p] != 0x1f || in[p + 1] != 0x8b) throw PhzError("corrupt BGZF block header");
    u32 xlen = in[p + 10] | (in[p + 11] << 8);
    size_t q = p + 12, xe = q + xlen; u32 bsize = 0; bool found = false;
    while (q + 4 <= xe) {
      u32 slen = in[q + 2] | (in[q + 3] << 8);
      if (in[q] == 'B' && in[q + 1] == 'C' && slen == 2) { bsize = (in[q + 4] | (in[q + 5] << 8)) + 1; found = true; }
      q += 4 + slen;
    }
    if (!found || p + bsize > in.size()) throw PhzError("BGZF block without BC field or truncated file");
    u32 isize = in[p + bsize - 4] | (in[p + bsize - 3] << 8) | (in[p + bsize - 2] << 16) | ((u32)in[p + bsize - 1] << 24);
    blks.push_back(Blk{xe, p + bsize - 8 - xe, total, isize});
    total += isize; p += bsize;
  }
  out.resize(total);
  std::atomic<size_t> next(0); std::atomic<int> bad(0);
  auto work = [&]() {
    z_stream zs;
    while (true) {
      size_t i = next.fetch_add(1);
      if (i >= blks.size()) break;
      const Blk& b = blks[i];
      if (b.out_len == 0) continue;
      std::memset(&zs, 0, sizeof(zs));
      if (inflateInit2(&zs, -15) != Z_OK) { bad = 1; break; }
      zs.next_in = (Bytef*)(in.data() + b.in_off); zs.avail_in = (uInt)b.in_len;
      zs.next_out = out.data() + b.out_off; zs.avail_out = (uInt)b.out_len;
      int rc = inflate(&zs, Z_FINISH);
      inflateEnd(&zs);
      if (rc != Z_STREAM_END) { bad = 1; break; }
    }
  };
  if (n_threads < 1) n_threads = 1;
  std::vector<std::thread> th;
  for (int t = 1; t < n_threads; ++t) th.emplace_back(work);
  work();
  for (auto& t : th) t.join();
  if (bad) throw PhzError("BGZF inflate failed");
}

// The same inflate as a background job: worker threads take the blocks in file order and publish how far the output is
// complete, so that the caller can start walking the BAM records behind the inflate front instead of after it.
struct InflateJob {
  struct Blk { size_t in_off, in_len, out_off, out_len; };
  std::vector<Blk> blks; std::vector<std::thread> th;
  std::unique_ptr<std::atomic<u8>[]> done;
  std::atomic<size_t> next{0}; std::atomic<int> bad{0};
  size_t frontier = 0, total = 0;          // frontier: only the consumer thread touches it
  template <class In, class Out>
  void start(const In& in, Out& out, int n_threads) {
    size_t p = 0;
    while (p + 18 <= in.size()) {
      if (in[p] != 0x1f || in[p + 1] != 0x8b) throw PhzError("corrupt BGZF block header");
      u32 xlen = in[p + 10] | (in[p + 11] << 8);
      size_t q = p + 12, xe = q + xlen; u32 bsize = 0; bool found = false;
      while (q + 4 <= xe) {
        u32 slen = in[q + 2] | (in[q + 3] << 8);
        if (in[q] == 'B' && in[q + 1] == 'C' && slen == 2) { bsize = (in[q + 4] | (in[q + 5] << 8)) + 1; found = true; }
        q += 4 + slen;
      }
      if (!found || p + bsize > in.size()) throw PhzError("BGZF block without BC field or truncated file");
      u32 isize = in[p + bsize - 4] | (in[p + bsize - 3] << 8) | (in[p + bsize - 2] << 16) | ((u32)in[p + bsize - 1] << 24);
      blks.push_back(Blk{xe, p + bsize - 8 - xe, total, isize});
      total += isize; p += bsize;
    }
    out.resize(total);
    done.reset(new std::atomic<u8>[blks.size() + 1]);
    for (size_t i = 0; i <= blks.size(); ++i) done[i] = 0;
    const u8* src = in.data(); u8* dst = out.data();
    auto work = [this, src, dst]() {
      z_stream zs;
      while (!bad) {
        size_t i = next.fetch_add(1);
        if (i >= blks.size()) break;
        const Blk& b = blks[i];
        if (b.out_len) {
          std::memset(&zs, 0, sizeof(zs));
          if (inflateInit2(&zs, -15) != Z_OK) { bad = 1; break; }
          zs.next_in = (Bytef*)(src + b.in_off); zs.avail_in = (uInt)b.in_len;
          zs.next_out = dst + b.out_off; zs.avail_out = (uInt)b.out_len;
          int rc = inflate(&zs, Z_FINISH);
          inflateEnd(&zs);
          if (rc != Z_STREAM_END) { bad = 1; break; }
        }
        done[i].store(1, std::memory_order_release);
      }
    };
    int nw = n_threads > 1 ? n_threads - 1 : 1;        // the caller's thread walks the records meanwhile
    for (int t = 0; t < nw; ++t) th.emplace_back(work);
  }
  // blocks until output bytes [0, off) are complete
  void need(size_t off) {
    if (off > total) off = total;
    while (true) {
      while (frontier < blks.size() && done[frontier].load(std::memory_order_acquire)) ++frontier;
      size_t ready = frontier < blks.size() ? blks[frontier].out_off : total;
      if (ready >= off) return;
      if (bad) { join(); throw PhzError("BGZF inflate failed"); }
      std::this_thread::yield();
    }
  }
  void join() { for (auto& t : th) if (t.joinable()) t.join(); th.clear(); if (bad) throw PhzError("BGZF inflate failed"); }
  ~InflateJob() { for (auto& t : th) if (t.joinable()) t.join(); }
};

template <class In, class Out>
static void inflate_gzip_stream(const In& in, Out& out) {
  z_stream zs; std::memset(&zs, 0, sizeof(zs));
  if (inflateInit2(&zs, 31) != Z_OK) throw PhzError("zlib init failed");
  out.clear(); std::vector<u8> buf(1 << 20);
  zs.next_in = (Bytef*)in.data(); zs.avail_in = (uInt)in.size();
  while (true) {
    zs.next_out = buf.data(); zs.avail_out = (uInt)buf.size();
    int rc = inflate(&zs, Z_NO_FLUSH);
    out.insert(out.end(), buf.data(), buf.data() + (buf.size() - zs.avail_out));
    if (rc == Z_STREAM_END) { if (zs.avail_in == 0) break; if (inflateReset(&zs) != Z_OK) break; continue; }
    if (rc != Z_OK) { inflateEnd(&zs); throw PhzError("gzip inflate failed"); }
  }
  inflateEnd(&zs);
}

struct Rec {              // one alignment that passed the filters, as offsets into the raw buffer
  int contig; int32_t pos, tlen; int16_t aln; u32 frag;
  const char* qname; u32 l_qname;
  const u8* cig; u32 n_cig;        // BAM: packed words; SAM: text
  const u8* seq; const u8* qual; u32 l_seq;
  bool text;
};

struct BaseTable {
  u8 t[256];
  BaseTable() {
    const char* alpha = "=ACMGRSVTWYHKDBN";      // BAM 4-bit codes; anything else encodes as N, like samtools
    for (int i = 0; i < 256; ++i) t[i] = 15;
    for (int i = 0; i < 16; ++i) t[(u8)alpha[i]] = (u8)i;
  }
};
static const BaseTable BASE_OF_TABLE;

static int16_t as_from_aux(const u8* p, const u8* end, std::string& err) {
  while (p + 3 <= end) {
    char t0 = p[0], t1 = p[1], ty = p[2]; p += 3;
    bool is_as = (t0 == 'A' && t1 == 'S');
    int64_t v = 0; bool have = false;
    switch (ty) {
      case 'A': p += 1; break;
      case 'c': v = (int8_t)p[0]; have = true; p += 1; break;
      case 'C': v = p[0]; have = true; p += 1; break;
      case 's': v = (int16_t)(p[0] | (p[1] << 8)); have = true; p += 2; break;
      case 'S': v = (uint16_t)(p[0] | (p[1] << 8)); have = true; p += 2; break;
      case 'i': v = (int32_t)(p[0] | (p[1] << 8) | (p[2] << 16) | ((u32)p[3] << 24)); have = true; p += 4; break;
      case 'I': v = (u32)(p[0] | (p[1] << 8) | (p[2] << 16) | ((u32)p[3] << 24)); have = true; p += 4; break;
      case 'f': p += 4; break;
      case 'Z': case 'H': while (p < end && *p) ++p; ++p; break;
      case 'B': { char st = p[0]; u32 n = p[1] | (p[2] << 8) | (p[3] << 16) | ((u32)p[4] << 24); p += 5;
                  int w = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : 4; p += (size_t)n * w; break; }
      default: return -32768;
    }
    if (is_as && have) {
      if (v < -32767 || v > 32767) { err = "AS:i value outside the int16 range of the packed layout"; return -32768; }
      return (int16_t)v;
    }
  }
  return -32768;
}

template <class RV>
static void assign_fragments(RV& recs, FragDict* fd, int n_threads) {
  size_t n = recs.size();
  std::vector<const char*, NoInit<const char*>> names(n); std::vector<u32, NoInit<u32>> lens(n), ids(n);
  std::vector<u64, NoInit<u64>> hashes(n);
  parallel_for((n + 65535) / 65536, n_threads, [&](size_t c) {
    size_t i1 = std::min(n, (c + 1) * 65536);
    for (size_t i = c * 65536; i < i1; ++i) { names[i] = recs[i].qname; lens[i] = recs[i].l_qname; hashes[i] = name_hash(recs[i].qname, recs[i].l_qname); }
  });
  fd->assign(names.data(), lens.data(), hashes.data(), n, ids.data(), n_threads);
  parallel_for((n + 65535) / 65536, n_threads, [&](size_t c) {
    size_t i1 = std::min(n, (c + 1) * 65536);
    for (size_t i = c * 65536; i < i1; ++i) recs[i].frag = ids[i];
  });
}

static u32 count_cigar_ops(const Rec& r) {
  if (!r.text) return r.n_cig;
  if (r.n_cig == 1 && r.cig[0] == '*') return 0;
  u32 nops = 0;
  for (u32 i = 0; i < r.n_cig; ++i) if (r.cig[i] < '0' || r.cig[i] > '9') nops++;
  return nops;
}

template <class RV>
static HostReads* build(RV& recs, int nc, int n_threads) {
  HostReads* H = new HostReads();
  H->n_contigs = nc;
  size_t R = recs.size();
  const size_t CH = 16384, n_chunks = (R + CH - 1) / CH;
  // records grouped by contig (VCF order), file order kept inside a contig: a counting sort over fixed chunks
  std::vector<int64_t> ccnt(n_chunks * (size_t)nc + 1, 0);
  parallel_for(n_chunks, n_threads, [&](size_t c) {
    int64_t* k = ccnt.data() + c * nc;
    const size_t i1 = std::min(R, (c + 1) * CH);
    for (size_t i = c * CH; i < i1; ++i) k[recs[i].contig]++;
  });
  H->contig_rec_off.assign(nc + 1, 0);
  { int64_t acc = 0;
    for (int ci = 0; ci < nc; ++ci) {
      H->contig_rec_off[ci] = acc;
      for (size_t c = 0; c < n_chunks; ++c) { const int64_t k = ccnt[c * nc + ci]; ccnt[c * nc + ci] = acc; acc += k; }
    }
    H->contig_rec_off[nc] = acc; }
  std::vector<u32, NoInit<u32>> order(R);
  parallel_for(n_chunks, n_threads, [&](size_t c) {
    int64_t* at = ccnt.data() + c * nc;
    const size_t i1 = std::min(R, (c + 1) * CH);
    for (size_t i = c * CH; i < i1; ++i) order[at[recs[i].contig]++] = (u32)i;
  });
  H->pos.resize(R); H->tlen.resize(R); H->aln.resize(R); H->frag.resize(R);
  H->cigar_off.resize(R + 1); H->seq_off.resize(R + 1);
  H->cigar_off[0] = 0; H->seq_off[0] = 0;
  // sizes: per-record counts and per-chunk sums in parallel, the running sum over the chunks serially, then every chunk
  // adds its base
  std::vector<u64> csum(n_chunks + 1, 0), ssum(n_chunks + 1, 0);
  parallel_for(n_chunks, n_threads, [&](size_t c) {
    size_t k1 = std::min(R, (c + 1) * CH); u64 a = 0, b = 0;
    for (size_t k = c * CH; k < k1; ++k) {
      const Rec& r = recs[order[k]];
      a += count_cigar_ops(r); b += r.l_seq;
      H->cigar_off[k + 1] = (u32)a; H->seq_off[k + 1] = b;
    }
    csum[c + 1] = a; ssum[c + 1] = b;
  });
  for (size_t c = 0; c < n_chunks; ++c) { csum[c + 1] += csum[c]; ssum[c + 1] += ssum[c]; }
  if (csum[n_chunks] >= 0xFFFFFFFFull) throw PhzError("more than 2^32 CIGAR operations in one file");
  parallel_for(n_chunks, n_threads, [&](size_t c) {
    if (c == 0) return;
    size_t k1 = std::min(R, (c + 1) * CH); const u32 a = (u32)csum[c]; const u64 b = ssum[c];
    for (size_t k = c * CH; k < k1; ++k) { H->cigar_off[k + 1] += a; H->seq_off[k + 1] += b; }
  });
  H->cigar.resize(H->cigar_off[R]);
  u64 nb = H->seq_off[R];
  H->qual.resize(nb); H->seq.resize((nb + 1) / 2);
  { const size_t ZB = (size_t)1 << 22, nz = (H->seq.size() + ZB - 1) / ZB;      // zeroed by all threads (first touch as well)
    parallel_for(nz, n_threads, [&](size_t z) { std::memset(H->seq.data() + z * ZB, 0, std::min(ZB, H->seq.size() - z * ZB)); }); }
  static const char* OPS = "MIDNSHP=X";
  parallel_for(n_chunks, n_threads, [&](size_t c) {
    size_t k1 = std::min(R, (c + 1) * CH);
    for (size_t k = c * CH; k < k1; ++k) {
      const Rec& r = recs[order[k]];
      H->pos[k] = r.pos; H->tlen[k] = r.tlen; H->aln[k] = r.aln; H->frag[k] = r.frag;
      u32* co = H->cigar.data() + H->cigar_off[k];
      if (!r.text) {
        std::memcpy(co, r.cig, (size_t)r.n_cig * 4);
      } else if (!(r.n_cig == 1 && r.cig[0] == '*')) {
        u32 n = 0, w = 0;
        for (u32 i = 0; i < r.n_cig; ++i) {
          u8 ch = r.cig[i];
          if (ch >= '0' && ch <= '9') n = n * 10 + (ch - '0');
          else { const char* q = std::strchr(OPS, ch); co[w++] = (n << 4) | (u32)(q ? q - OPS : 0); n = 0; }
        }
      }
      u64 b0 = H->seq_off[k];
      const u32 L = r.l_seq;
      auto code_at = [&](u32 j) -> u8 {
        return r.text ? BASE_OF_TABLE.t[r.seq[j]] : ((j & 1) ? (u8)(r.seq[j >> 1] & 15) : (u8)(r.seq[j >> 1] >> 4));
      };
      // bases: the first / last byte of a record can be shared with a neighbour handled by another thread (atomic
      // or of one nibble); every byte in between belongs to this record alone and is written whole
      u32 j = 0;
      if (L > 0 && (b0 & 1)) { __atomic_fetch_or(&H->seq[b0 >> 1], code_at(0), __ATOMIC_RELAXED); j = 1; }
      u8* sq = H->seq.data() + ((b0 + j) >> 1);
      for (; j + 1 < L; j += 2) *sq++ = (u8)((code_at(j) << 4) | code_at(j + 1));
      if (j < L) __atomic_fetch_or(&H->seq[(b0 + j) >> 1], (u8)(code_at(j) << 4), __ATOMIC_RELAXED);
      u8* qo = H->qual.data() + b0;
      if (r.text) for (u32 t = 0; t < L; ++t) { int q = (int)r.qual[t] - 33; qo[t] = (u8)(q < 0 ? 0 : q); }
      else std::memcpy(qo, r.qual, L);
    }
  });
  std::atomic<int> unsorted{0};
  parallel_for(n_chunks, n_threads, [&](size_t c) {
    if (unsorted.load(std::memory_order_relaxed)) return;
    const size_t k0 = c * CH, k1 = std::min(R, (c + 1) * CH);
    int ci = (int)(std::upper_bound(H->contig_rec_off.begin(), H->contig_rec_off.end(), (int64_t)k0) - H->contig_rec_off.begin()) - 1;
    for (size_t k = k0; k < k1; ++k) {
      while (ci + 1 <= nc && (int64_t)k >= H->contig_rec_off[ci + 1]) ++ci;
      if ((int64_t)k > H->contig_rec_off[ci] && H->pos[k] < H->pos[k - 1]) { unsorted = 1; return; }
    }
  });
  if (unsorted) H->sorted = 0;
  return H;
}

template <class D>
static HostReads* parse_bam(const D& d, const char* const* contigs, int nc, FragDict* fd, int remove_dups,
                            int proper_pair, int min_mapq, int n_threads, InflateJob* job = nullptr) {
  auto rd32 = [&](size_t o) { return (int32_t)(d[o] | (d[o + 1] << 8) | (d[o + 2] << 16) | ((u32)d[o + 3] << 24)); };
  auto need = [&](size_t off) { if (job) job->need(off); };         // the inflate front (BGZF input inflated in the background)
  need(12);
  if (d.size() < 12 || std::memcmp(d.data(), "BAM\1", 4) != 0) throw PhzError("not a BAM file");
  need(8 + (size_t)rd32(4) + 4);
  size_t p = 8 + (size_t)rd32(4);
  int n_ref = rd32(p); p += 4;
  std::vector<int> ref_to_contig(n_ref, -1);
  for (int i = 0; i < n_ref; ++i) {
    need(p + 4); need(p + 4 + (size_t)rd32(p) + 4);
    int ln = rd32(p); p += 4;
    std::string name((const char*)d.data() + p, ln > 0 ? ln - 1 : 0); p += ln + 4;
    for (int c = 0; c < nc; ++c) if (name == contigs[c]) ref_to_contig[i] = c;
  }
  // record boundaries (serial pointer chase), then filter + decode headers in parallel chunks
  std::vector<size_t> starts;
  starts.reserve(d.size() / 96 + 16);
  size_t have = 0;                       // bytes known to be inflated
  while (p + 4 <= d.size()) {
    if (job && p + 4 > have) { need(p + (1 << 20)); have = p + (1 << 20); }        // wait in 1 MB strides, not per record
    int32_t bs = rd32(p);
    if (bs < 32 || p + 4 + (size_t)bs > d.size()) throw PhzError("truncated or corrupt BAM record");
    starts.push_back(p + 4); p += 4 + (size_t)bs;
  }
  if (job) job->join();                  // the decode below reads every byte
  const size_t CH = 16384, n_chunks = (starts.size() + CH - 1) / CH;
  std::vector<std::vector<Rec>> parts(n_chunks);
  double tp0 = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
  parallel_for(n_chunks, n_threads, [&](size_t c) {
    size_t i1 = std::min(starts.size(), (c + 1) * CH);
    std::vector<Rec>& out = parts[c];
    std::string err;
    for (size_t i = c * CH; i < i1; ++i) {
      size_t s = starts[i]; size_t e = s + (size_t)rd32(s - 4);
      int ref = rd32(s); if (ref < 0 || ref >= n_ref) continue;
      int ci = ref_to_contig[ref]; if (ci < 0) continue;
      u32 l_rn = d[s + 8], mapq = d[s + 9], n_cig = d[s + 12] | (d[s + 13] << 8), flag = d[s + 14] | (d[s + 15] << 8);
      int32_t l_seq = rd32(s + 16);
      if (remove_dups && (flag & 0x400)) continue;
      if (proper_pair && !(flag & 2)) continue;
      if ((int)mapq < min_mapq) continue;
      size_t o = s + 32;
      Rec r; r.text = false; r.contig = ci; r.pos = rd32(s + 4) + 1; r.tlen = rd32(s + 28); r.frag = 0;
      r.qname = (const char*)d.data() + o; r.l_qname = l_rn ? l_rn - 1 : 0; o += l_rn;
      r.cig = d.data() + o; r.n_cig = n_cig; o += (size_t)n_cig * 4;
      r.seq = d.data() + o; o += ((size_t)l_seq + 1) / 2; r.qual = d.data() + o; r.l_seq = (u32)l_seq; o += l_seq;
      if (o > e) throw PhzError("corrupt BAM record (fields exceed block_size)");
      if (l_seq > 0 && r.qual[0] == 0xFF) throw PhzError("record without QUAL (unsupported)");
      r.aln = as_from_aux(d.data() + o, d.data() + e, err);
      if (!err.empty()) throw PhzError(err);
      out.push_back(r);
    }
  });
  const bool timing = std::getenv("PHZ_IO_TIMING") != nullptr;
  auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double t0 = now();
  if (timing) std::fprintf(stderr, "[phz_io] records: decode %.3fs\n", t0 - tp0);
  std::vector<Rec, NoInit<Rec>> recs;
  size_t total = 0; for (auto& v : parts) total += v.size();
  recs.resize(total);
  { std::vector<size_t> at(parts.size() + 1, 0);
    for (size_t c = 0; c < parts.size(); ++c) at[c + 1] = at[c] + parts[c].size();
    parallel_for(parts.size(), n_threads, [&](size_t c) {
      std::copy(parts[c].begin(), parts[c].end(), recs.begin() + at[c]); std::vector<Rec>().swap(parts[c]);
    }); }
  double t1 = now();
  assign_fragments(recs, fd, n_threads);
  double t2 = now();
  HostReads* H = build(recs, nc, n_threads);
  if (timing) std::fprintf(stderr, "[phz_io] records: merge %.3fs fragment ids %.3fs build %.3fs\n", t1 - t0, t2 - t1, now() - t2);
  return H;
}

static void parse_sam_line(const u8* p, const u8* le, const char* const* contigs, int nc, int remove_dups, int proper_pair,
                           int min_mapq, std::vector<Rec>& out) {
  const u8* f[12]; int nf = 0; const u8* q = p; f[nf++] = q;
  while (q < le && nf < 12) { if (*q == '\t') f[nf++] = q + 1; ++q; }
  if (nf < 11) return;
  auto len = [&](int i) { return (size_t)((i + 1 < nf ? f[i + 1] - 1 : le) - f[i]); };
  auto toint = [&](int i) { char b[24]; size_t n = len(i); if (n > 23) n = 23; std::memcpy(b, f[i], n); b[n] = 0; return std::strtol(b, nullptr, 10); };
  int ci = -1;
  for (int c = 0; c < nc; ++c) if (std::strlen(contigs[c]) == len(2) && std::memcmp(contigs[c], f[2], len(2)) == 0) { ci = c; break; }
  if (ci < 0) return;
  long flag = toint(1), mapq = toint(4);
  if ((remove_dups && (flag & 0x400)) || (proper_pair && !(flag & 2)) || mapq < min_mapq) return;
  Rec r; r.text = true; r.contig = ci; r.pos = (int32_t)toint(3); r.tlen = (int32_t)toint(8); r.frag = 0;
  r.qname = (const char*)f[0]; r.l_qname = (u32)len(0);
  r.cig = f[5]; r.n_cig = (u32)len(5);
  r.seq = f[9]; r.l_seq = (u32)len(9);
  const u8* qe = le; if (nf > 11) qe = f[11] - 1;
  r.qual = f[10]; size_t lq = (size_t)(qe - f[10]);
  if (r.l_seq == 1 && r.seq[0] == '*') r.l_seq = 0;
  if (lq != r.l_seq) throw PhzError("SAM record with QUAL missing or not the length of SEQ (unsupported)");
  r.aln = -32768;
  if (nf > 11) {          // first AS: tag from column 12 on (read_variant_map.py:56-59)
    const u8* t = f[11];
    while (t < le) {
      const u8* te = (const u8*)std::memchr(t, '\t', le - t); if (!te) te = le;
      if (te - t > 5 && t[0] == 'A' && t[1] == 'S' && t[2] == ':') {
        const u8* c2 = (const u8*)std::memchr(t + 3, ':', te - t - 3);
        if (c2) {
          char b[24]; size_t n = (size_t)(te - c2 - 1); if (n > 23) n = 23; std::memcpy(b, c2 + 1, n); b[n] = 0;
          long v = std::strtol(b, nullptr, 10);
          if (v < -32767 || v > 32767) throw PhzError("AS:i value outside the int16 range of the packed layout");
          r.aln = (int16_t)v; break;
        }
      }
      t = te + 1;
    }
  }
  out.push_back(r);
}

template <class D>
static HostReads* parse_sam(const D& d, const char* const* contigs, int nc, FragDict* fd, int remove_dups,
                            int proper_pair, int min_mapq, int n_threads) {
  // chunk boundaries at line starts
  const u8* base = d.data(); const u8* end = base + d.size();
  size_t want = (size_t)(n_threads < 1 ? 1 : n_threads) * 8;
  size_t step = d.size() / want + 1;
  std::vector<const u8*> cuts{base};
  for (size_t k = 1; k < want; ++k) {
    const u8* q = base + std::min(d.size(), k * step);
    if (q <= cuts.back()) continue;
    const u8* nl = (const u8*)std::memchr(q, '\n', end - q);
    if (!nl) break;
    if (nl + 1 > cuts.back()) cuts.push_back(nl + 1);
  }
  cuts.push_back(end);
  std::vector<std::vector<Rec>> parts(cuts.size() - 1);
  double tp0 = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
  parallel_for(parts.size(), n_threads, [&](size_t c) {
    const u8* p = cuts[c]; const u8* e = cuts[c + 1];
    while (p < e) {
      const u8* eol = (const u8*)std::memchr(p, '\n', e - p); if (!eol) eol = e;
      const u8* le = eol; if (le > p && le[-1] == '\r') --le;
      if (p < le && *p != '@') parse_sam_line(p, le, contigs, nc, remove_dups, proper_pair, min_mapq, parts[c]);
      p = eol + 1;
    }
  });
  const bool timing = std::getenv("PHZ_IO_TIMING") != nullptr;
  auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double t0 = now();
  if (timing) std::fprintf(stderr, "[phz_io] records: decode %.3fs\n", t0 - tp0);
  std::vector<Rec, NoInit<Rec>> recs;
  size_t total = 0; for (auto& v : parts) total += v.size();
  recs.resize(total);
  { std::vector<size_t> at(parts.size() + 1, 0);
    for (size_t c = 0; c < parts.size(); ++c) at[c + 1] = at[c] + parts[c].size();
    parallel_for(parts.size(), n_threads, [&](size_t c) {
      std::copy(parts[c].begin(), parts[c].end(), recs.begin() + at[c]); std::vector<Rec>().swap(parts[c]);
    }); }
  double t1 = now();
  assign_fragments(recs, fd, n_threads);
  double t2 = now();
  HostReads* H = build(recs, nc, n_threads);
  if (timing) std::fprintf(stderr, "[phz_io] records: merge %.3fs fragment ids %.3fs build %.3fs\n", t1 - t0, t2 - t1, now() - t2);
  return H;
}

}  // namespace phzio

struct phz_fragdict { phzio::FragDict d; };
struct phz_host_reads { phzio::HostReads* h; };

extern "C" {

phz_fragdict* phz_fragdict_create(void) { return new phz_fragdict(); }
void phz_fragdict_destroy(phz_fragdict* d) { delete d; }
int64_t phz_fragdict_size(phz_fragdict* d) { return (int64_t)d->d.count; }
int64_t phz_fragdict_name(phz_fragdict* d, int64_t id, char* buf, int64_t buflen) {
  if (id < 0 || (size_t)id >= d->d.count) return -1;
  const phzio::FragDict::Shard& S = d->d.sh[d->d.id_shard[id]]; u32 e = d->d.id_ent[id];
  u64 o = S.name_off[e]; int64_t n = (int64_t)(S.name_off[e + 1] - o);
  if (n + 1 > buflen) return -(n + 1);
  std::memcpy(buf, S.arena.data() + o, n); buf[n] = 0;
  return n;
}

// Bulk export / import of the QNAME dictionary (ids in order): what the SoA cache file stores beside the arrays so that a
// later run can skip the ingest ("parse once") and still name its reads or meet the same QNAME in
// another BAM.
int64_t phz_fragdict_blob_bytes(phz_fragdict* d) {
  int64_t n = 0;
  for (size_t g = 0; g < d->d.count; ++g) { auto& S = d->d.sh[d->d.id_shard[g]]; u32 e = d->d.id_ent[g]; n += (int64_t)(S.name_off[e + 1] - S.name_off[e]); }
  return n;
}
int phz_fragdict_export(phz_fragdict* d, char* blob, int64_t* off) {
  PHZ_TRY
  int64_t at = 0;
  for (size_t g = 0; g < d->d.count; ++g) {
    auto& S = d->d.sh[d->d.id_shard[g]]; u32 e = d->d.id_ent[g];
    const u64 o = S.name_off[e]; const int64_t n = (int64_t)(S.name_off[e + 1] - o);
    off[g] = at; std::memcpy(blob + at, S.arena.data() + o, (size_t)n); at += n;
  }
  off[d->d.count] = at;
  PHZ_CATCH
}
int phz_fragdict_import(phz_fragdict* d, const char* blob, const int64_t* off, int64_t n, int n_threads) {
  PHZ_TRY
  if (d->d.count != 0) throw PhzError("phz_fragdict_import: the dictionary must be empty");
  std::vector<const char*> names((size_t)n); std::vector<u32> lens((size_t)n); std::vector<u64> hashes((size_t)n); std::vector<u32> ids((size_t)n);
  phzio::parallel_for((size_t)((n + 65535) / 65536), n_threads, [&](size_t c) {
    const int64_t i1 = std::min<int64_t>(n, (int64_t)(c + 1) * 65536);
    for (int64_t i = (int64_t)c * 65536; i < i1; ++i) {
      names[i] = blob + off[i]; lens[i] = (u32)(off[i + 1] - off[i]); hashes[i] = phzio::name_hash(names[i], lens[i]);
    }
  });
  d->d.assign(names.data(), lens.data(), hashes.data(), (size_t)n, ids.data(), n_threads);
  for (int64_t i = 0; i < n; ++i) if (ids[i] != (u32)i) throw PhzError("phz_fragdict_import: duplicate names in the cache");
  PHZ_CATCH
}

phz_host_reads* phz_read_alignments(const char* path, const char* const* contigs, int n_contigs, phz_fragdict* fd,
                                    int remove_dups, int proper_pair, int min_mapq, int n_threads) {
  try {
    phzio::FileMap raw;
    const bool timing = std::getenv("PHZ_IO_TIMING") != nullptr;
    auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t0 = now();
    if (!raw.open(path)) throw PhzError(std::string("cannot read ") + path);
    double t1 = now();
    phzio::Bytes data;
    bool gz = raw.size() >= 18 && raw[0] == 0x1f && raw[1] == 0x8b;
    bool bgzf = gz && (raw[3] & 4) && raw[12] == 'B' && raw[13] == 'C';
    phzio::InflateJob job;
    if (bgzf) { job.start(raw, data, n_threads); job.need(4); }       // the blocks inflate in the background from here on
    else if (gz) phzio::inflate_gzip_stream(raw, data);
    double t2 = now();
    std::unique_ptr<phz_host_reads> out(new phz_host_reads());
    auto parse = [&](const auto& d, phzio::InflateJob* j) {
      if (d.size() >= 4 && std::memcmp(d.data(), "BAM\1", 4) == 0)
        return phzio::parse_bam(d, contigs, n_contigs, &fd->d, remove_dups, proper_pair, min_mapq, n_threads, j);
      if (j) j->join();
      return phzio::parse_sam(d, contigs, n_contigs, &fd->d, remove_dups, proper_pair, min_mapq, n_threads);
    };
    out->h = gz ? parse(data, bgzf ? &job : nullptr) : parse(raw, nullptr);
    if (timing) std::fprintf(stderr, "[phz_io] read %.3fs inflate %.3fs parse+build %.3fs (%lld records)\n", t1 - t0, t2 - t1,
                             now() - t2, (long long)out->h->pos.size());
    return out.release();
  } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}

int phz_host_reads_view(phz_host_reads* r, phz_reads* out, int* sorted) {
  PHZ_TRY
  phzio::HostReads* h = r->h;
  out->n_records = (int64_t)h->pos.size(); out->n_cigar_ops = (int64_t)h->cigar.size(); out->n_bases = (int64_t)h->qual.size();
  out->h_contig_rec_off = h->contig_rec_off.data();
  out->pos = h->pos.data(); out->tlen = h->tlen.data(); out->aln_score = h->aln.data(); out->frag = h->frag.data();
  out->cigar_off = h->cigar_off.data(); out->cigar = h->cigar.data(); out->seq_off = (const uint64_t*)h->seq_off.data();
  out->seq = h->seq.data(); out->qual = h->qual.data();
  *sorted = h->sorted;
  PHZ_CATCH
}

void phz_host_reads_free(phz_host_reads* r) { if (r) { delete r->h; delete r; } }

// SAM text twin of generated records (test / baseline input for the reference, which reads SAM text
// through the harness).  ops/opl: n x n_ops padded CIGAR table (len 0 = unused); bases: 4-bit codes.
int phz_write_sam(const char* path, const char* const* contig_names, const int64_t* contig_lengths, int n_contigs, int64_t n,
                  const int64_t* contig, const int64_t* pos, const int64_t* tlen, const int64_t* flag, const int64_t* mapq,
                  const int64_t* aln, const int64_t* frag, const int64_t* ops, const int64_t* opl, int n_ops,
                  const uint8_t* bases, const uint8_t* qual, int read_len, const char* bam_name) {
  PHZ_TRY
  FILE* f = std::fopen(path, "w");
  if (!f) throw PhzError(std::string("cannot write ") + path);
  static const char* ALPHA = "=ACMGRSVTWYHKDBN"; static const char* OPS = "MIDNSHP=X";
  std::fprintf(f, "@HD\tVN:1.6\tSO:coordinate\n");
  for (int c = 0; c < n_contigs; ++c) std::fprintf(f, "@SQ\tSN:%s\tLN:%lld\n", contig_names[c], (long long)contig_lengths[c]);
  std::string line; std::vector<char> sq(read_len + 1), ql(read_len + 1);
  for (int64_t i = 0; i < n; ++i) {
    char cig[512]; int w = 0;
    for (int j = 0; j < n_ops; ++j) if (opl[i * n_ops + j] > 0) w += std::snprintf(cig + w, sizeof(cig) - w, "%lld%c", (long long)opl[i * n_ops + j], OPS[ops[i * n_ops + j]]);
    cig[w] = 0;
    for (int j = 0; j < read_len; ++j) { sq[j] = ALPHA[bases[i * read_len + j] & 15]; ql[j] = (char)(qual[i * read_len + j] + 33); }
    sq[read_len] = 0; ql[read_len] = 0;
    long long pn = tlen[i] > 0 ? (pos[i] + tlen[i] > 1 ? pos[i] + tlen[i] : 1) : pos[i];
    std::fprintf(f, "%s.%lld\t%lld\t%s\t%lld\t%lld\t%s\t=\t%lld\t%lld\t%s\t%s\tNH:i:1\tAS:i:%lld\n", bam_name, (long long)frag[i],
                 (long long)flag[i], contig_names[contig[i]], (long long)pos[i], (long long)mapq[i], cig, pn, (long long)tlen[i],
                 sq.data(), ql.data(), (long long)aln[i]);
  }
  std::fclose(f);
  PHZ_CATCH
}

}  // extern "C"

// =============================================================================== packed transport (include/phz.h)
// Host side of phz_packed_reads: offsets -> per-record counts, 4-bit bases -> 2 bits + exception list, phred bytes ->
// indices into the table of distinct values.  Lossless; expanded again on the device by phz_map_reads_packed.
// Host -> device copy of a pageable host array at PCIe speed: the bytes go through two page-locked staging buffers that
// host threads fill (parallel memcpy) while the previous chunk is on the bus.  A plain cudaMemcpy from pageable memory
// stages through one driver thread at ~5 GB/s; the command line uploads 150 bytes per record this way.
int phz_upload(phz_ctx* ctx, void* d_dst, const void* h_src, int64_t bytes, int n_threads) {
  PHZ_TRY
  auto& be = ctx->p.be;
  if (bytes <= 0) return 0;
#ifdef __CUDACC__
  static const size_t CHUNK = 32u << 20;
  if (!be.up_stage[0]) {
    for (int b = 0; b < 2; ++b) {
      PHZ_CUDA(cudaHostAlloc(&be.up_stage[b], CHUNK, cudaHostAllocDefault));
      PHZ_CUDA(cudaEventCreateWithFlags(&be.up_ev[b], cudaEventDisableTiming));
    }
  }
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 8) n_threads = 8;
  size_t off = 0; int b = 0; bool used[2] = {false, false};
  while (off < (size_t)bytes) {
    const size_t n = std::min(CHUNK, (size_t)bytes - off);
    if (used[b]) PHZ_CUDA(cudaEventSynchronize(be.up_ev[b]));          // the copy that last read this buffer is done
    const u8* src = (const u8*)h_src + off; u8* st = (u8*)be.up_stage[b];
    const size_t piece = (n + n_threads - 1) / n_threads;
    phzio::parallel_for((size_t)n_threads, n_threads, [&](size_t t) {
      const size_t a = t * piece; if (a >= n) return;
      std::memcpy(st + a, src + a, std::min(piece, n - a));
    });
    PHZ_CUDA(cudaMemcpyAsync((u8*)d_dst + off, st, n, cudaMemcpyHostToDevice, be.stream));
    PHZ_CUDA(cudaEventRecord(be.up_ev[b], be.stream));
    used[b] = true; off += n; b ^= 1;
  }
  for (int k = 0; k < 2; ++k) if (used[k]) PHZ_CUDA(cudaEventSynchronize(be.up_ev[k]));      // the staging buffers are free again
#else
  (void)n_threads;
  std::memcpy(d_dst, h_src, (size_t)bytes);
#endif
  PHZ_CATCH
}

// BAM (BGZF) twin of generated records: what the product's command line reads in the files-to-files benchmark while the
// reference baseline gets the SAM text twin of the same records (test / bench infrastructure).  Records are encoded on
// n_threads threads, the byte stream is cut into 0xff00-byte BGZF blocks that are deflated in parallel.
int phz_write_bam(const char* path, const char* const* contig_names, const int64_t* contig_lengths, int n_contigs, int64_t n,
                  const int64_t* contig, const int64_t* pos, const int64_t* tlen, const int64_t* flag, const int64_t* mapq,
                  const int64_t* aln, const int64_t* frag, const int64_t* ops, const int64_t* opl, int n_ops,
                  const uint8_t* bases, const uint8_t* qual, int read_len, const char* bam_name, int n_threads) {
  PHZ_TRY
  auto put32 = [](std::vector<u8>& o, u32 v) { o.push_back((u8)v); o.push_back((u8)(v >> 8)); o.push_back((u8)(v >> 16)); o.push_back((u8)(v >> 24)); };
  std::vector<u8> head;
  { std::string text = "@HD\tVN:1.6\tSO:coordinate\n";
    for (int c = 0; c < n_contigs; ++c) text += std::string("@SQ\tSN:") + contig_names[c] + "\tLN:" + std::to_string((long long)contig_lengths[c]) + "\n";
    head.insert(head.end(), {'B', 'A', 'M', 1});
    put32(head, (u32)text.size()); head.insert(head.end(), text.begin(), text.end());
    put32(head, (u32)n_contigs);
    for (int c = 0; c < n_contigs; ++c) {
      u32 l = (u32)std::strlen(contig_names[c]) + 1;
      put32(head, l); head.insert(head.end(), contig_names[c], contig_names[c] + l); put32(head, (u32)contig_lengths[c]);
    } }
  const size_t CH = 1 << 15, nch = ((size_t)n + CH - 1) / CH;
  std::vector<std::vector<u8>> parts(nch);
  const size_t name_len = std::strlen(bam_name);
  phzio::parallel_for(nch, n_threads, [&](size_t c) {
    std::vector<u8>& o = parts[c];
    o.reserve(CH * (size_t)(64 + read_len * 3 / 2));
    const int64_t i1 = std::min<int64_t>(n, (int64_t)(c + 1) * (int64_t)CH);
    char nm[96];
    for (int64_t i = (int64_t)c * (int64_t)CH; i < i1; ++i) {
      const int l_name = std::snprintf(nm, sizeof(nm), "%.*s.%lld", (int)(name_len > 60 ? 60 : name_len), bam_name, (long long)frag[i]) + 1;
      u32 cig[64]; int nc = 0; int64_t ref_span = 0;
      for (int j = 0; j < n_ops && nc < 64; ++j) {
        const int64_t l = opl[i * n_ops + j]; const int op = (int)ops[i * n_ops + j];
        if (l > 0) { cig[nc++] = ((u32)l << 4) | (u32)op; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) ref_span += l; }
      }
      const int64_t beg = pos[i] - 1, end = beg + (ref_span > 0 ? ref_span : 1);
      auto reg2bin = [](int64_t b, int64_t e) -> u32 {
        --e;
        if (b >> 14 == e >> 14) return (u32)(((1 << 15) - 1) / 7 + (b >> 14));
        if (b >> 17 == e >> 17) return (u32)(((1 << 12) - 1) / 7 + (b >> 17));
        if (b >> 20 == e >> 20) return (u32)(((1 << 9) - 1) / 7 + (b >> 20));
        if (b >> 23 == e >> 23) return (u32)(((1 << 6) - 1) / 7 + (b >> 23));
        if (b >> 26 == e >> 26) return (u32)(((1 << 3) - 1) / 7 + (b >> 26));
        return 0;
      };
      const int64_t pn = tlen[i] > 0 ? (pos[i] + tlen[i] > 1 ? pos[i] + tlen[i] : 1) : pos[i];
      const bool as8 = aln[i] >= 0 && aln[i] <= 255;
      const u32 block = 32 + (u32)l_name + 4 * (u32)nc + (u32)(read_len + 1) / 2 + (u32)read_len + 4 + (as8 ? 4 : 5);
      auto p32 = [&](u32 v) { o.push_back((u8)v); o.push_back((u8)(v >> 8)); o.push_back((u8)(v >> 16)); o.push_back((u8)(v >> 24)); };
      p32(block); p32((u32)contig[i]); p32((u32)beg);
      o.push_back((u8)l_name); o.push_back((u8)mapq[i]); const u32 bin = reg2bin(beg, end); o.push_back((u8)bin); o.push_back((u8)(bin >> 8));
      o.push_back((u8)nc); o.push_back((u8)(nc >> 8)); o.push_back((u8)flag[i]); o.push_back((u8)(flag[i] >> 8));
      p32((u32)read_len); p32((u32)contig[i]); p32((u32)(pn - 1)); p32((u32)(int32_t)tlen[i]);
      o.insert(o.end(), nm, nm + l_name);
      for (int j = 0; j < nc; ++j) p32(cig[j]);
      const u8* b = bases + i * read_len;
      for (int j = 0; j + 1 < read_len; j += 2) o.push_back((u8)(((b[j] & 15) << 4) | (b[j + 1] & 15)));
      if (read_len & 1) o.push_back((u8)((b[read_len - 1] & 15) << 4));
      o.insert(o.end(), qual + i * read_len, qual + (i + 1) * read_len);
      o.push_back('N'); o.push_back('H'); o.push_back('C'); o.push_back(1);
      o.push_back('A'); o.push_back('S');
      if (as8) { o.push_back('C'); o.push_back((u8)aln[i]); }
      else { o.push_back('s'); o.push_back((u8)aln[i]); o.push_back((u8)((uint16_t)(int16_t)aln[i] >> 8)); }
    }
  });
  // one logical byte stream: header + parts; cut into BGZF blocks
  std::vector<size_t> part_off(nch + 2, 0);
  part_off[1] = head.size();
  for (size_t c = 0; c < nch; ++c) part_off[c + 2] = part_off[c + 1] + parts[c].size();
  const size_t total = part_off[nch + 1];
  auto copy_out = [&](size_t from, size_t len, u8* dst) {
    size_t k = std::upper_bound(part_off.begin(), part_off.end(), from) - part_off.begin() - 1;      // 0 = header, 1.. = parts
    while (len) {
      const std::vector<u8>& src = k == 0 ? head : parts[k - 1];
      const size_t o = from - part_off[k]; const size_t take = std::min(len, src.size() - o);
      std::memcpy(dst, src.data() + o, take); dst += take; from += take; len -= take; ++k;
    }
  };
  const size_t BS = 0xff00, nblk = (total + BS - 1) / BS;
  std::vector<std::vector<u8>> blocks(nblk);
  phzio::parallel_for(nblk, n_threads, [&](size_t b) {
    const size_t from = b * BS, len = std::min(BS, total - from);
    std::vector<u8> raw(len); copy_out(from, len, raw.data());
    std::vector<u8>& o = blocks[b];
    o.resize(len + 1024);
    z_stream zs; std::memset(&zs, 0, sizeof(zs));
    if (deflateInit2(&zs, 1, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) throw PhzError("zlib deflate init failed");
    zs.next_in = raw.data(); zs.avail_in = (uInt)len; zs.next_out = o.data() + 18; zs.avail_out = (uInt)(o.size() - 18 - 8);
    if (deflate(&zs, Z_FINISH) != Z_STREAM_END) { deflateEnd(&zs); throw PhzError("BGZF deflate failed"); }
    const size_t clen = zs.total_out; deflateEnd(&zs);
    const u8 hdr[18] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0, 0};
    std::memcpy(o.data(), hdr, 18);
    const size_t bsize = 18 + clen + 8 - 1;
    o[16] = (u8)bsize; o[17] = (u8)(bsize >> 8);
    const u32 crc = (u32)crc32(crc32(0L, Z_NULL, 0), raw.data(), (uInt)len);
    u8* t = o.data() + 18 + clen;
    t[0] = (u8)crc; t[1] = (u8)(crc >> 8); t[2] = (u8)(crc >> 16); t[3] = (u8)(crc >> 24);
    t[4] = (u8)len; t[5] = (u8)(len >> 8); t[6] = (u8)(len >> 16); t[7] = (u8)(len >> 24);
    o.resize(18 + clen + 8);
  });
  FILE* f = std::fopen(path, "wb");
  if (!f) throw PhzError(std::string("cannot write ") + path);
  for (auto& b : blocks) std::fwrite(b.data(), 1, b.size(), f);
  static const u8 eof[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  std::fwrite(eof, 1, 28, f);
  std::fclose(f);
  PHZ_CATCH
}

struct phz_packed_host {
  phz_packed_reads v;
  struct Owned { void* p; bool pinned; int64_t bytes; };
  std::vector<Owned> bufs;
  std::vector<int64_t> contig_off;
  int64_t bytes = 0;
  bool want_pinned = false;
  template <class T> T* alloc(size_t n) {
    bool pinned = false;
    void* p = want_pinned ? PHZ_BACKEND::host_alloc((n ? n : 1) * sizeof(T), &pinned) : std::malloc((n ? n : 1) * sizeof(T));
    if (!p) throw PhzError("out of host memory for the packed transport buffers");
    bufs.push_back(Owned{p, pinned, (int64_t)(n * sizeof(T))});
    bytes += (int64_t)(n * sizeof(T));
    return (T*)p;
  }
  static void free_one(const Owned& b) { if (b.pinned) PHZ_BACKEND::host_free(b.p, true); else std::free(b.p); }
  void release(const void* p) {          // a buffer the packer allocated for a coding it then did not take
    for (size_t i = 0; i < bufs.size(); ++i)
      if (bufs[i].p == p) { bytes -= bufs[i].bytes; free_one(bufs[i]); bufs.erase(bufs.begin() + (long)i); return; }
  }
  ~phz_packed_host() { for (auto& b : bufs) free_one(b); }
};

extern "C" {

phz_packed_host* phz_pack_reads(const phz_reads* h, int n_contigs, int n_threads, int page_locked) {
  phz_packed_host* P = nullptr;
  try {
    if (n_threads < 1) n_threads = 1;
    const int64_t R = h->n_records, NB = h->n_bases, NCG = h->n_cigar_ops;
    for (int64_t r = 0; r < R; ++r) {
      if (h->cigar_off[r + 1] - h->cigar_off[r] > 65535u || h->seq_off[r + 1] - h->seq_off[r] > 65535ull)
        throw PhzError("record " + std::to_string(r) + " has more than 65535 CIGAR operations or bases: not packable");
    }
    P = new phz_packed_host();
    P->want_pinned = page_locked != 0;
    phz_packed_reads& v = P->v;
    std::memset(&v, 0, sizeof(v));
    v.n_records = R; v.n_cigar_ops = NCG; v.n_bases = NB;
    P->contig_off.assign(h->h_contig_rec_off, h->h_contig_rec_off + n_contigs + 1);
    v.h_contig_rec_off = P->contig_off.data();
    const size_t CH = 1 << 20;
    const size_t nrc = (size_t)((R + (int64_t)CH - 1) / (int64_t)CH);
    // ---- pos: u16 difference to the previous record, tlen: i16, each with an exception list
    uint16_t* pd = P->alloc<uint16_t>(R); int16_t* t16 = P->alloc<int16_t>(R);
    std::vector<std::vector<std::pair<u32, int32_t>>> pex(nrc), tex(nrc);
    std::vector<std::vector<u32>> as_seen(nrc, std::vector<u32>(65536 / 32, 0));
    std::vector<u32> max_ncg(nrc, 0), lmin(nrc, 0xFFFFFFFFu), lmax(nrc, 0);
    phzio::parallel_for(nrc, n_threads, [&](size_t c) {
      int64_t r0 = (int64_t)(c * CH), r1 = std::min<int64_t>(R, r0 + (int64_t)CH);
      auto& seen = as_seen[c];
      for (int64_t r = r0; r < r1; ++r) {
        int64_t d = (int64_t)h->pos[r] - (r > 0 ? (int64_t)h->pos[r - 1] : 0);
        if (d >= 0 && d < 65535) pd[r] = (uint16_t)d; else { pd[r] = 65535; pex[c].emplace_back((u32)r, (int32_t)d); }
        int32_t t = h->tlen[r];
        if (t > -32768 && t <= 32767) t16[r] = (int16_t)t; else { t16[r] = (int16_t)-32768; tex[c].emplace_back((u32)r, t); }
        u32 a = (u32)(uint16_t)h->aln_score[r]; seen[a >> 5] |= 1u << (a & 31);
        u32 nc = h->cigar_off[r + 1] - h->cigar_off[r]; if (nc > max_ncg[c]) max_ncg[c] = nc;
        u32 ls = (u32)(h->seq_off[r + 1] - h->seq_off[r]); if (ls < lmin[c]) lmin[c] = ls; if (ls > lmax[c]) lmax[c] = ls;
      }
    });
    auto flatten = [&](std::vector<std::vector<std::pair<u32, int32_t>>>& parts, const uint32_t*& oi, const int32_t*& ov) {
      int64_t n = 0; for (auto& e : parts) n += (int64_t)e.size();
      uint32_t* xi = P->alloc<uint32_t>(n); int32_t* xv = P->alloc<int32_t>(n);
      int64_t o = 0; for (auto& e : parts) for (auto& x : e) { xi[o] = x.first; xv[o] = x.second; ++o; }
      oi = xi; ov = xv; return n;
    };
    v.pos_delta = pd; v.n_pos_exc = flatten(pex, v.pos_exc_index, v.pos_exc_delta);
    v.tlen16 = t16; v.n_tlen_exc = flatten(tex, v.tlen_exc_index, v.tlen_exc_value);
    // ---- fragment ids numbered by first appearance are implicit (phz.h: frag_first / frag_back); taken when smaller.
    // A record "opens" an id iff its id exceeds every id before it (a prefix maximum: chunk maxima first, then the chunks
    // in parallel); whatever the implicit rule then gets wrong -- ids that are not dense, references further back than
    // 16 bits -- is listed as an exception, so the coding is lossless for any input.
    v.frag_bits = 32;
    if (R > 0) {
      const u32 base = h->frag[0];
      std::vector<u32> cmax(nrc, 0); std::vector<int64_t> cfirst(nrc + 1, 0);
      phzio::parallel_for(nrc, n_threads, [&](size_t c) {
        int64_t r0 = (int64_t)(c * CH), r1 = std::min<int64_t>(R, r0 + (int64_t)CH);
        u32 m = 0; for (int64_t r = r0; r < r1; ++r) m = std::max(m, h->frag[r]);
        cmax[c] = m;
      });
      // exclusive prefix maximum over the chunks; "nothing seen yet" = base - 1 (64-bit: base may be 0)
      std::vector<int64_t> pmax(nrc, (int64_t)base - 1);
      for (size_t c = 1; c < nrc; ++c) pmax[c] = std::max(pmax[c - 1], (int64_t)cmax[c - 1]);
      phzio::parallel_for(nrc, n_threads, [&](size_t c) {
        int64_t r0 = (int64_t)(c * CH), r1 = std::min<int64_t>(R, r0 + (int64_t)CH);
        int64_t m = pmax[c], k = 0;
        for (int64_t r = r0; r < r1; ++r) if ((int64_t)h->frag[r] > m) { m = h->frag[r]; ++k; }
        cfirst[c + 1] = k;
      });
      for (size_t c = 0; c < nrc; ++c) cfirst[c + 1] += cfirst[c];
      const int64_t n_first = cfirst[nrc], n_back = R - n_first;
      std::vector<std::vector<std::pair<u32, u32>>> fex(nrc);
      uint32_t* fb = P->alloc<uint32_t>((R + 31) / 32); uint16_t* bk = P->alloc<uint16_t>(n_back);
      static_assert(CH % 32 == 0, "chunks own whole words of the first-appearance bitmap");
      phzio::parallel_for(nrc, n_threads, [&](size_t c) {
        int64_t r0 = (int64_t)(c * CH), r1 = std::min<int64_t>(R, r0 + (int64_t)CH);
        for (int64_t w = r0 >> 5; w < (r1 + 31) >> 5; ++w) fb[w] = 0;
        int64_t m = pmax[c], nf = cfirst[c];          // nf: set bits before r
        for (int64_t r = r0; r < r1; ++r) {
          const u32 id = h->frag[r];
          if ((int64_t)id > m) {
            m = id; fb[r >> 5] |= 1u << (r & 31);
            if ((int64_t)id != (int64_t)base + nf) fex[c].emplace_back((u32)r, id);
            ++nf;
          } else {
            const int64_t d = (int64_t)base + nf - (int64_t)id;
            if (d >= 1 && d <= 65535) bk[r - nf] = (uint16_t)d;
            else { bk[r - nf] = 0; fex[c].emplace_back((u32)r, id); }
          }
        }
      });
      int64_t n_exc = 0; for (auto& e : fex) n_exc += (int64_t)e.size();
      if (((R + 31) / 32) * 4 + n_back * 2 + n_exc * 8 < R * 4) {
        uint32_t* xi = P->alloc<uint32_t>(n_exc); uint32_t* xv = P->alloc<uint32_t>(n_exc);
        int64_t o = 0; for (auto& e : fex) for (auto& x : e) { xi[o] = x.first; xv[o] = x.second; ++o; }
        v.frag_bits = 16; v.frag_base = base; v.frag_first = fb; v.n_frag_back = n_back; v.frag_back = bk;
        v.n_frag_exc = n_exc; v.frag_exc_index = xi; v.frag_exc_value = xv;
      } else { P->release(fb); P->release(bk); }
    }
    if (v.frag_bits == 32) {
      uint32_t* frag = P->alloc<uint32_t>(R);
      phzio::parallel_for(nrc, n_threads, [&](size_t c) {
        int64_t r0 = (int64_t)(c * CH), r1 = std::min<int64_t>(R, r0 + (int64_t)CH);
        std::memcpy(frag + r0, h->frag + r0, (r1 - r0) * 4);
      });
      v.frag = frag;
    }
    // ---- alignment scores: table of distinct values when it fits a byte
    {
      std::vector<u32> seen(65536 / 32, 0);
      for (auto& sc : as_seen) for (size_t k = 0; k < seen.size(); ++k) seen[k] |= sc[k];
      std::vector<uint16_t> distinct;
      for (u32 a = 0; a < 65536; ++a) if (seen[a >> 5] >> (a & 31) & 1) distinct.push_back((uint16_t)a);
      if (distinct.size() <= 256) {
        std::vector<u8> index_of(65536, 0);
        for (size_t k = 0; k < distinct.size(); ++k) { index_of[distinct[k]] = (u8)k; v.as_table[k] = (int16_t)distinct[k]; }
        uint8_t* a8 = P->alloc<uint8_t>(R);
        phzio::parallel_for(nrc, n_threads, [&](size_t c) {
          int64_t r0 = (int64_t)(c * CH), r1 = std::min<int64_t>(R, r0 + (int64_t)CH);
          for (int64_t r = r0; r < r1; ++r) a8[r] = index_of[(uint16_t)h->aln_score[r]];
        });
        v.as_bits = 8; v.as_data = a8;
      } else {
        int16_t* a16 = P->alloc<int16_t>(R);
        phzio::parallel_for(nrc, n_threads, [&](size_t c) {
          int64_t r0 = (int64_t)(c * CH), r1 = std::min<int64_t>(R, r0 + (int64_t)CH);
          std::memcpy(a16 + r0, h->aln_score + r0, (r1 - r0) * 2);
        });
        v.as_bits = 16; v.as_data = a16;
      }
    }
    // ---- per-record counts
    {
      u32 mx = 0; for (u32 x : max_ncg) mx = std::max(mx, x);
      u32 lo = 0xFFFFFFFFu, hi = 0; for (size_t c = 0; c < nrc; ++c) { lo = std::min(lo, lmin[c]); hi = std::max(hi, lmax[c]); }
      v.n_cigar_bits = mx <= 255 ? 8 : 16;
      void* ncg = v.n_cigar_bits == 8 ? (void*)P->alloc<uint8_t>(R) : (void*)P->alloc<uint16_t>(R);
      const bool lconst = R > 0 && lo == hi;
      uint16_t* lsq = lconst ? nullptr : P->alloc<uint16_t>(R);
      v.l_seq_const = lconst ? (int32_t)lo : -1; v.l_seq = lsq;
      const int nb = v.n_cigar_bits;
      phzio::parallel_for(nrc, n_threads, [&](size_t c) {
        int64_t r0 = (int64_t)(c * CH), r1 = std::min<int64_t>(R, r0 + (int64_t)CH);
        for (int64_t r = r0; r < r1; ++r) {
          u32 nc = h->cigar_off[r + 1] - h->cigar_off[r];
          if (nb == 8) ((uint8_t*)ncg)[r] = (uint8_t)nc; else ((uint16_t*)ncg)[r] = (uint16_t)nc;
          if (lsq) lsq[r] = (uint16_t)(h->seq_off[r + 1] - h->seq_off[r]);
        }
      });
      v.n_cigar = ncg;
    }
    // ---- CIGAR words: table of distinct words when it fits 16 bits
    {
      const size_t ncc = (size_t)((NCG + (int64_t)CH - 1) / (int64_t)CH);
      std::vector<std::vector<u32>> part(ncc);
      phzio::parallel_for(ncc, n_threads, [&](size_t c) {
        int64_t i0 = (int64_t)(c * CH), i1 = std::min<int64_t>(NCG, i0 + (int64_t)CH);
        std::vector<u32> w(h->cigar + i0, h->cigar + i1);
        std::sort(w.begin(), w.end()); w.erase(std::unique(w.begin(), w.end()), w.end());
        if (w.size() > 65536) w.resize(65537);           // hopeless already: keep it short
        part[c].swap(w);
      });
      std::vector<u32> table;
      for (auto& w : part) { table.insert(table.end(), w.begin(), w.end()); if (table.size() > (1u << 22)) { std::sort(table.begin(), table.end()); table.erase(std::unique(table.begin(), table.end()), table.end()); } }
      std::sort(table.begin(), table.end()); table.erase(std::unique(table.begin(), table.end()), table.end());
      if (table.size() <= 65536) {
        uint32_t* tab = P->alloc<uint32_t>(table.size()); std::memcpy(tab, table.data(), table.size() * 4);
        uint16_t* c16 = P->alloc<uint16_t>(NCG);
        phzio::parallel_for(ncc, n_threads, [&](size_t c) {
          int64_t i0 = (int64_t)(c * CH), i1 = std::min<int64_t>(NCG, i0 + (int64_t)CH);
          u32 last = 0xFFFFFFFFu; uint16_t last_ix = 0;
          for (int64_t i = i0; i < i1; ++i) {
            u32 w = h->cigar[i];
            if (w != last) { last = w; last_ix = (uint16_t)(std::lower_bound(table.begin(), table.end(), w) - table.begin()); }
            c16[i] = last_ix;
          }
        });
        v.cigar_bits = 16; v.cigar = c16; v.n_cigar_table = (int32_t)table.size(); v.cigar_table = tab;
      } else {
        uint32_t* cig = P->alloc<uint32_t>(NCG);
        phzio::parallel_for(ncc, n_threads, [&](size_t c) {
          int64_t i0 = (int64_t)(c * CH), i1 = std::min<int64_t>(NCG, i0 + (int64_t)CH);
          std::memcpy(cig + i0, h->cigar + i0, (i1 - i0) * 4);
        });
        v.cigar_bits = 32; v.cigar = cig; v.n_cigar_table = 0; v.cigar_table = nullptr;
      }
    }
    // ---- bases: 2 bits + exceptions.  Chunks are multiples of 4 bases so that no output byte is shared.
    const size_t nchunks = (size_t)((NB + (int64_t)CH - 1) / (int64_t)CH);
    uint8_t* s2 = P->alloc<uint8_t>((NB + 3) / 4);
    std::vector<std::vector<std::pair<u64, u8>>> exc(nchunks);
    std::vector<std::array<u64, 256>> hist(nchunks);
    phzio::parallel_for(nchunks, n_threads, [&](size_t c) {
      int64_t b0 = (int64_t)(c * CH), b1 = std::min<int64_t>(NB, b0 + (int64_t)CH);
      auto& hh = hist[c]; hh.fill(0);
      auto& ex = exc[c];
      for (int64_t b = b0; b < b1; b += 4) {
        u8 out = 0;
        for (int k = 0; k < 4 && b + k < b1; ++k) {
          int64_t i = b + k;
          u8 code = (h->seq[i >> 1] >> ((i & 1) ? 0 : 4)) & 15;
          u8 two;
          switch (code) { case 1: two = 0; break; case 2: two = 1; break; case 4: two = 2; break; case 8: two = 3; break;
                          default: two = 0; ex.emplace_back((u64)i, code); }
          out |= (u8)(two << (2 * k));
        }
        s2[b >> 2] = out;
      }
      for (int64_t b = b0; b < b1; ++b) hh[h->qual[b]]++;
    });
    int64_t nex = 0; for (auto& e : exc) nex += (int64_t)e.size();
    uint64_t* exi = P->alloc<uint64_t>(nex); uint8_t* exc_code = P->alloc<uint8_t>(nex);
    { int64_t o = 0; for (auto& e : exc) for (auto& x : e) { exi[o] = x.first; exc_code[o] = x.second; ++o; } }
    v.seq2 = s2; v.n_exceptions = nex; v.exc_index = exi; v.exc_code = exc_code;
    // ---- base qualities: table of distinct values, index width = what the data needs
    u64 tot[256]; for (int q = 0; q < 256; ++q) { tot[q] = 0; for (auto& hh : hist) tot[q] += hh[q]; }
    u8 index_of[256]; int nd = 0;
    for (int q = 0; q < 256; ++q) { index_of[q] = 0; if (tot[q]) { index_of[q] = (u8)nd; v.qual_table[nd] = (u8)q; nd++; } }
    int bits = nd <= 2 ? 1 : nd <= 4 ? 2 : nd <= 16 ? 4 : 8;
    v.qual_bits = bits;
    uint8_t* qp = P->alloc<uint8_t>((size_t)((NB * bits + 7) / 8));
    const int per = 8 / bits;                   // bases per packed byte
    phzio::parallel_for(nchunks, n_threads, [&](size_t c) {
      int64_t b0 = (int64_t)(c * CH), b1 = std::min<int64_t>(NB, b0 + (int64_t)CH);     // CH is a multiple of 8
      for (int64_t b = b0; b < b1; b += per) {
        u8 out = 0;
        for (int k = 0; k < per && b + k < b1; ++k) out |= (u8)(index_of[h->qual[b + k]] << (k * bits));
        qp[b / per] = out;
      }
    });
    v.qualp = qp;
    return P;
  } catch (const std::exception& e) { g_err = e.what(); delete P; return nullptr; }
}

int phz_packed_view(phz_packed_host* p, phz_packed_reads* out) { PHZ_TRY *out = p->v; PHZ_CATCH }
int64_t phz_packed_bytes(phz_packed_host* p) { return p->bytes; }
void phz_packed_free(phz_packed_host* p) { delete p; }

}  // extern "C"

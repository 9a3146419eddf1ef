// ingest.cpp -- native target preparation: parallel gz decode + FASTQ/FASTA parse, then the exact
// tagging / splitting / exact-duplicate merging / chunk ordering of the reference's prepare_target()
// (/root/reference/lib/aln.py:274-433, read_fastx lib/library.py:157-176, check_file_num :262-271).
//
// The reference does this in one Python thread (gzip text decode, dict of strings).  Here every input
// file is inflated and split into records on its own host thread (phase 1); tagging and the
// order-dependent duplicate merge run afterwards in one pass over the records in file order (phase 2),
// so all.raw.fasta, total_counts.tsv and merged_redundance.txt are byte-identical to the reference's,
// quirks included:
//   * FASTQ id = whole header line without the first '@', spaces -> '_'; FASTA id = first token after '>'
//   * tag  {id}_PhAlTag_{species}_fastx{n}[_split_{a}_{b}]
//   * duplicates: exact string, case-sensitive, only sequences of length <= merge_len, across all
//     species/files; representative = first occurrence; the last FASTA record of a file is appended to
//     its duplicate group by its UNTAGGED id (reference bug, lib/aln.py:407)
//   * the FASTA fast path (merge_len == 0, no split) copies sequence lines verbatim and never counts the
//     last record of a file
//   * chunks of >= 10 MB of sequence; past 1000 chunks, chunks 100..999 fold into chunk (i % 100) and the
//     chunk size grows tenfold, which permutes the final record order exactly as the reference does
// Host-only code (zlib); lives in libphyloaln_b200.so behind the pa_ingest_* C ABI.
#include <sys/mman.h>
#include <sys/stat.h>
#include <zlib.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <string>
#include <string_view>
#include <thread>
#include <vector>

#include "../../include/phyloaln_b200.h"

namespace {

struct FileJob {
  std::string path, species;
  int file_index = 1;  // j + 1
  int format = 0;      // 0 guess, 1 fastq, 2 fasta
  // phase 1 output: the inflated text and its line table (rstripped)
  std::string text;
  std::vector<std::pair<uint64_t, uint64_t>> lines;  // (offset, length after rstrip)
  std::string error;
  // unsplit FASTQ (preformat): every record the pass can emit, already written as '>' tag '\n' seq '\n' into one
  // buffer that outlives the file text, with the hashes of its sequence and tag; the serial pass then only decides
  struct PreRec { uint64_t off, seq_len, hseq, hid; uint32_t tag_len; };
  std::vector<PreRec> pre;
  char *fmt = nullptr;
  size_t fmt_bytes = 0;
  bool preformatted = false;
};

bool inflate_file(FileJob &fj) {
  gzFile g = gzopen(fj.path.c_str(), "rb");  // transparently reads plain files too
  if (!g) { fj.error = "cannot open " + fj.path; return false; }
  gzbuffer(g, 1 << 20);
  std::string &t = fj.text;
  size_t cap = 1 << 22;
  t.resize(cap);
  size_t n = 0;
  for (;;) {
    if (cap - n < (1 << 20)) { cap *= 2; t.resize(cap); }
    const int got = gzread(g, &t[n], (unsigned)std::min<size_t>(cap - n, 1u << 30));
    if (got < 0) { int e; fj.error = std::string("gzread: ") + gzerror(g, &e); gzclose(g); return false; }
    if (got == 0) break;
    n += (size_t)got;
  }
  gzclose(g);
  t.resize(n);
  return true;
}

inline bool py_space(unsigned char c) {  // what str.rstrip() / str.split() strip for ASCII text
  return c == ' ' || (c >= 9 && c <= 13) || (c >= 0x1c && c <= 0x1f);
}

void split_lines(FileJob &fj) {
  const std::string &t = fj.text;
  size_t i = 0, n = t.size();
  fj.lines.reserve(n / 64 + 16);
  while (i < n) {  // universal newlines, like Python's text mode: \n, \r\n and a lone \r all end a line
    const void *nl = memchr(t.data() + i, '\n', n - i);
    size_t e = nl ? (size_t)((const char *)nl - t.data()) : n;
    size_t next = e + 1;
    const void *cr = memchr(t.data() + i, '\r', e - i);
    if (cr) {
      const size_t c = (size_t)((const char *)cr - t.data());
      if (c + 1 < e || (c + 1 == e && !nl)) { e = c; next = c + 1; }  // lone \r inside the segment
      else if (c + 1 == e) { e = c; }                                  // \r\n
    }
    size_t r = e;
    while (r > i && py_space((unsigned char)t[r - 1])) r--;
    fj.lines.emplace_back((uint64_t)i, (uint64_t)(r - i));
    i = next;
  }
  if (fj.format == 0) {  // guess from the first line (lib/library.py:160-166)
    if (!fj.lines.empty() && fj.lines[0].second > 0) {
      const char c = t[fj.lines[0].first];
      if (c == '@') fj.format = 1;
      else if (c == '>') fj.format = 2;
    }
    // an unreadable first line leaves 'guess': no branch of prepare_target matches, nothing is emitted
  }
}

// Large, long-lived buffers (record texts, hash tables, record tables) come straight from mmap with transparent huge
// pages requested: the single pass touches gigabytes of fresh memory and random table slots, and 4 KB pages make that one
// page fault per ~20 records plus a TLB miss per probe.
constexpr size_t HUGE_MIN = (size_t)2 << 20;
inline void *big_alloc(size_t bytes) {
  bytes = (bytes + HUGE_MIN - 1) & ~(HUGE_MIN - 1);
  // over-allocate by one huge page and give the unaligned head and the tail back: the region is 2 MB-aligned
  char *raw = static_cast<char *>(mmap(nullptr, bytes + HUGE_MIN, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0));
  if (raw == MAP_FAILED) throw std::bad_alloc();
  char *p = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(raw) + HUGE_MIN - 1) & ~(uintptr_t)(HUGE_MIN - 1));
  if (p > raw) munmap(raw, (size_t)(p - raw));
  const size_t tail = (size_t)((raw + bytes + HUGE_MIN) - (p + bytes));
  if (tail) munmap(p + bytes, tail);
#ifdef MADV_HUGEPAGE
  madvise(p, bytes, MADV_HUGEPAGE);
#endif
  return p;
}
inline void big_free(void *p, size_t bytes) {
  bytes = (bytes + HUGE_MIN - 1) & ~(HUGE_MIN - 1);
  munmap(p, bytes);
}
template <class T>
struct BigAlloc {   // std::vector allocator: mmap above 2 MB, malloc below
  using value_type = T;
  BigAlloc() = default;
  template <class U> BigAlloc(const BigAlloc<U> &) {}
  T *allocate(size_t n) {
    const size_t b = n * sizeof(T);
    if (b >= HUGE_MIN) return static_cast<T *>(big_alloc(b));
    void *p = malloc(b ? b : 1);
    if (!p) throw std::bad_alloc();
    return static_cast<T *>(p);
  }
  void deallocate(T *p, size_t n) {
    const size_t b = n * sizeof(T);
    if (b >= HUGE_MIN) big_free(p, b);
    else free(p);
  }
  template <class U> bool operator==(const BigAlloc<U> &) const { return true; }
  template <class U> bool operator!=(const BigAlloc<U> &) const { return false; }
};
template <class T> using BigVec = std::vector<T, BigAlloc<T>>;

// bump allocator: keys, tags and record texts are never freed individually; pointers stay valid until the ingest is destroyed
struct Arena {
  std::vector<std::pair<char *, size_t>> slabs;
  size_t left = 0;
  char *cur = nullptr;
  Arena() = default;
  Arena(const Arena &) = delete;
  Arena &operator=(const Arena &) = delete;
  ~Arena() {
    for (auto &sl : slabs) big_free(sl.first, sl.second);
  }
  char *alloc(size_t n) {
    if (n > left) {
      const size_t sz = std::max<size_t>(n, (size_t)32 << 20);
      slabs.emplace_back(static_cast<char *>(big_alloc(sz)), sz);
      cur = slabs.back().first;
      left = sz;
    }
    char *d = cur;
    cur += n;
    left -= n;
    return d;
  }
  const char *put(const char *p, size_t n) {
    char *d = alloc(n);
    memcpy(d, p, n);
    return d;
  }
};

inline uint64_t hash_bytes(const char *p, size_t n) {  // 8 bytes at a time, multiply-xorshift mixing
  uint64_t h = 0x9E3779B97F4A7C15ull ^ (n * 0xFF51AFD7ED558CCDull);
  while (n >= 8) {
    uint64_t w;
    memcpy(&w, p, 8);
    h = (h ^ w) * 0xC4CEB9FE1A85EC53ull;
    h ^= h >> 29;
    p += 8;
    n -= 8;
  }
  uint64_t w = 0;
  memcpy(&w, p, n);
  h = (h ^ w ^ (n << 56)) * 0xFF51AFD7ED558CCDull;
  return h ^ (h >> 32);
}

// FASTQ id: header line without its first '@', ' ' -> '_' (lib/aln.py:296: line.replace('@', '', 1).replace(' ', '_'))
inline void put_fastq_id(char *d, const char *id, size_t n) {
  for (size_t i = 0; i < n; i++) d[i] = id[i] == ' ' ? '_' : id[i];
}

// Unsplit FASTQ, on the file's worker thread: what the reference's loop does with such a file is fixed by the line
// numbers alone -- record r is emitted iff line 4r starts with '@' and has an id, and line 4r+1 exists (lib/aln.py:
// 293-318 with line_num % 4; the state `seqid` never survives a record) -- so every candidate record is written here in
// its final form '>' id suffix '\n' seq '\n' and hashed, and the order-dependent pass is left with the merge decisions.
void preformat(FileJob &fj, int merge_len) {
  const std::string &T = fj.text;
  const std::string suffix = "_PhAlTag_" + fj.species + "_fastx" + std::to_string(fj.file_index);
  const size_t nlines = fj.lines.size();
  size_t bytes = 0, nvalid = 0;
  for (size_t r = 0; 4 * r + 1 < nlines; r++) {
    const auto &hl = fj.lines[4 * r];
    if (hl.second > 1 && T[hl.first] == '@') {
      if (hl.second - 1 + suffix.size() >= (1ull << 32)) { fj.error = "FASTQ identifier longer than 4 GiB in " + fj.path; return; }
      bytes += (hl.second - 1) + suffix.size() + fj.lines[4 * r + 1].second + 3;
      nvalid++;
    }
  }
  fj.pre.reserve(nvalid);
  if (bytes) {
    fj.fmt = static_cast<char *>(big_alloc(bytes));
    fj.fmt_bytes = bytes;
  }
  size_t pos = 0;
  for (size_t r = 0; 4 * r + 1 < nlines; r++) {
    const auto &hl = fj.lines[4 * r];
    if (!(hl.second > 1 && T[hl.first] == '@')) continue;
    const auto &sl = fj.lines[4 * r + 1];
    const size_t idl = hl.second - 1, tl = idl + suffix.size();
    char *d = fj.fmt + pos;
    d[0] = '>';
    put_fastq_id(d + 1, T.data() + hl.first + 1, idl);
    memcpy(d + 1 + idl, suffix.data(), suffix.size());
    d[1 + tl] = '\n';
    memcpy(d + 2 + tl, T.data() + sl.first, sl.second);
    d[2 + tl + sl.second] = '\n';
    FileJob::PreRec R;
    R.off = pos;
    R.tag_len = (uint32_t)tl;
    R.seq_len = sl.second;
    R.hseq = (int64_t)sl.second <= merge_len ? hash_bytes(d + 2 + tl, sl.second) : 0;
    R.hid = hash_bytes(d + 1, tl);
    fj.pre.push_back(R);
    pos += tl + sl.second + 3;
  }
  std::string().swap(fj.text);
  std::vector<std::pair<uint64_t, uint64_t>>().swap(fj.lines);
  fj.preformatted = true;
}

// one duplicate group: the sequence, the representative's tag, then a chain of further tags
struct MergeGroup {
  const char *key;
  uint32_t key_len;
  uint32_t tag_len;
  const char *tag;
  int64_t extra_head = -1, extra_tail = -1;  // indices into Builder::extras
};
struct ExtraTag {
  const char *tag;
  uint32_t tag_len;
  int64_t next;
};

// One record of all.raw.fasta, in EMISSION order: text = '>' id '\n' [body '\n'].  The body is the sequence; on the FASTA
// fast path it keeps the line breaks of the input.  Records are appended by the single phase-2 thread and read by any number
// of consumers up to `published` (release/acquire), so the search can start while the ingest is still running.
struct Rec {
  const char *text;
  uint64_t body_len;
  uint32_t id_len;
  uint32_t has_body;
  uint64_t text_len() const { return (uint64_t)id_len + 2 + (has_body ? body_len + 1 : 0); }
};
struct RecStore {
  static constexpr uint64_t BLK = 1ull << 20;
  std::unique_ptr<Rec *[]> blocks{new Rec *[1 << 16]()};  // fixed directory: readers never see it move
  uint64_t n = 0;
  std::atomic<uint64_t> published{0};
  ~RecStore() {
    for (uint64_t b = 0; b * BLK < n; b++) big_free(blocks[b], BLK * sizeof(Rec));
  }
  bool push(const Rec &r) {
    if (n % BLK == 0) {
      if (n / BLK >= (1u << 16)) return false;
      blocks[n / BLK] = static_cast<Rec *>(big_alloc(BLK * sizeof(Rec)));
    }
    blocks[n / BLK][n % BLK] = r;
    n++;
    published.store(n, std::memory_order_release);
    return true;
  }
  const Rec &at(uint64_t i) const { return blocks[i / BLK][i % BLK]; }
};

}  // namespace

struct pa_ingest {
  int merge_len = 250, split_len = 0, split_slide = 0, nthreads = 1;
  std::vector<FileJob> files;
  std::string err;
  std::atomic<bool> finished{false};
  bool started = false;
  int rc = 0;
  std::thread runner;
  std::mutex wait_mu;
  // results
  Arena text;       // record texts
  std::vector<std::pair<char *, size_t>> fmt_bufs;   // preformatted record texts of the unsplit FASTQ files (FileJob::fmt)
  RecStore recs;    // emission order
  std::vector<uint64_t> final_order;  // emission index of every final position; empty = identity (no chunk fold happened)
  uint64_t id_dups = 0;               // records whose id repeats an earlier record's id (read_fastx keeps one entry per id)
  std::vector<std::pair<std::string, int64_t>> counts;
  std::string merged_txt, counts_txt;
  // materialised on demand (pa_ingest_views): the final all.raw.fasta text and its record table in final order
  bool have_raw = false;
  std::string raw;
  std::vector<uint64_t> id_off, id_len, seq_off, seq_len;
  ~pa_ingest() {
    if (runner.joinable()) runner.join();
    for (auto &b : fmt_bufs) big_free(b.first, b.second);
    for (auto &f : files)
      if (f.fmt) big_free(f.fmt, f.fmt_bytes);     // a file that was formatted but never reached by the pass (error exit)
  }
};

namespace {

struct Builder {
  pa_ingest &S;
  // chunk files of the reference as runs (first emission index, count) of records
  std::vector<std::vector<std::pair<uint64_t, uint64_t>>> chunks;
  bool folded = false;
  size_t cur = 0;
  uint64_t data_size = 0, each_size = 10000000;
  // open-addressing table (hash, group index + 1); groups in insertion order == Python dict order
  BigVec<std::pair<uint64_t, uint64_t>> table = BigVec<std::pair<uint64_t, uint64_t>>(1 << 16);
  BigVec<MergeGroup> groups;
  BigVec<ExtraTag> extras;
  Arena arena;
  // hashes of the ids seen so far: counts the records whose id is (very probably) not new
  BigVec<uint64_t> idtab = BigVec<uint64_t>(1 << 16, 0);
  uint64_t n_ids = 0;

  int64_t find(std::string_view seq, uint64_t h) const {
    const size_t mask = table.size() - 1;
    for (size_t i = h & mask;; i = (i + 1) & mask) {
      const auto &e = table[i];
      if (e.second == 0) return -1;
      if (e.first == h) {
        const MergeGroup &g = groups[e.second - 1];
        if (g.key_len == seq.size() && !memcmp(g.key, seq.data(), seq.size())) return (int64_t)e.second - 1;
      }
    }
  }
  void insert(uint64_t h, uint64_t gi) {
    if ((groups.size() + 1) * 2 > table.size()) {
      BigVec<std::pair<uint64_t, uint64_t>> nt(table.size() * 2);
      const size_t mask = nt.size() - 1;
      for (const auto &e : table)
        if (e.second) {
          size_t i = e.first & mask;
          while (nt[i].second) i = (i + 1) & mask;
          nt[i] = e;
        }
      table.swap(nt);
    }
    const size_t mask = table.size() - 1;
    size_t i = h & mask;
    while (table[i].second) i = (i + 1) & mask;
    table[i] = {h, gi + 1};
  }
  // Set of 64-bit id hashes (0 = empty slot).  A hit counts as "this id was seen before"; the rare false positive of a
  // hash collision only sends the caller down the exact path (dropin._dictionary_view compares the ids themselves).
  void note_id(const char *id, uint32_t len) { note_id_hash(hash_bytes(id, len)); }
  void note_id_hash(uint64_t h) {
    if (h == 0) h = 1;
    if ((n_ids + 1) * 2 > idtab.size()) {
      BigVec<uint64_t> nt(idtab.size() * 2, 0);
      const size_t mask = nt.size() - 1;
      for (const uint64_t e : idtab)
        if (e) {
          size_t i = e & mask;
          while (nt[i]) i = (i + 1) & mask;
          nt[i] = e;
        }
      idtab.swap(nt);
    }
    const size_t mask = idtab.size() - 1;
    for (size_t i = h & mask;; i = (i + 1) & mask) {
      if (idtab[i] == 0) { idtab[i] = h; n_ids++; return; }
      if (idtab[i] == h) { S.id_dups++; return; }
    }
  }
  // look-ahead of the single pass: pull the table slots of an upcoming record into the cache
  void prefetch_hashes(uint64_t hseq, uint64_t hid) const {
    __builtin_prefetch(&table[hseq & (table.size() - 1)]);
    __builtin_prefetch(&idtab[(hid ? hid : 1) & (idtab.size() - 1)]);
  }
  // size the tables for about n records before the pass starts, so that they are not rebuilt log2(n) times on the way
  void reserve(uint64_t n, bool with_groups) {
    if (!groups.empty() || n_ids) return;
    n = std::min<uint64_t>(n, 1ull << 31);
    size_t cap = idtab.size();
    while (cap < 2 * (n + 1)) cap <<= 1;
    if (cap > idtab.size()) idtab.assign(cap, 0);
    if (with_groups && cap > table.size()) {
      table.assign(cap, {0, 0});
      groups.reserve((size_t)n);
    }
  }
  // emit() for a preformatted record (FileJob::pre): same decisions in the same order; the text is already final and
  // stays where it is -- an emitted record points at it, a merged one contributes its tag from it
  bool emit_pre(char *d, uint32_t tl, uint64_t sl, uint64_t hseq, uint64_t hid) {
    const bool small = (int64_t)sl <= S.merge_len;
    if (small) {
      const int64_t gi = find(std::string_view(d + 2 + tl, sl), hseq);
      if (gi >= 0) {
        extras.push_back({d + 1, tl, -1});
        MergeGroup &g = groups[(size_t)gi];
        if (g.extra_tail >= 0) extras[(size_t)g.extra_tail].next = (int64_t)extras.size() - 1;
        else g.extra_head = (int64_t)extras.size() - 1;
        g.extra_tail = (int64_t)extras.size() - 1;
        return true;
      }
    }
    if (!add_record(cur, d, tl, sl, true, &hid)) return false;
    data_size += sl;
    rotate_if_full();
    if (small) {
      MergeGroup g;
      g.key = d + 2 + tl;
      g.key_len = (uint32_t)sl;
      g.tag = d + 1;
      g.tag_len = tl;
      insert(hseq, groups.size());
      groups.push_back(g);
    }
    return true;
  }
  explicit Builder(pa_ingest &s) : S(s) {
    chunks.emplace_back();
    // test knob: the reference's chunk size (10 MB, lib/aln.py:283) scaled down so that the >1000-chunk fold can be
    // checked against the reference on kilobytes of input (tests/golden/make_ingest_golden.py patches the same constant)
    if (const char *e = getenv("PA_INGEST_EACH_SIZE")) {
      const unsigned long long v = strtoull(e, nullptr, 10);
      if (v > 0) each_size = v;
    }
  }

  void rotate_if_full() {  // lib/aln.py:313-318 + check_file_num
    if (data_size >= each_size) {
      data_size = 0;
      cur++;
      if (cur >= 1000) {
        for (size_t i = 100; i < 1000; i++) chunks[i % 100].insert(chunks[i % 100].end(), chunks[i].begin(), chunks[i].end());
        chunks.resize(100);
        cur = 100;
        each_size *= 10;
        folded = true;
      }
      if (chunks.size() <= cur) chunks.resize(cur + 1);
    }
  }
  // append a finished record (text already in S.text) to chunk `ch`
  bool add_record(size_t ch, const char *text, uint32_t id_len, uint64_t body_len, bool has_body, const uint64_t *id_hash = nullptr) {
    const uint64_t idx = S.recs.n;
    if (id_hash) note_id_hash(*id_hash);
    else note_id(text + 1, id_len);
    auto &runs = chunks[ch];
    if (!runs.empty() && runs.back().first + runs.back().second == idx) runs.back().second++;
    else runs.emplace_back(idx, 1);
    return S.recs.push(Rec{text, body_len, id_len, has_body ? 1u : 0u});
  }
  // one sequence (or split fragment) with its tag: merge or write
  bool emit(const std::string &tagid, std::string_view seq, uint64_t size_add, bool may_register, const std::string *append_as = nullptr) {
    uint64_t h = 0;
    if ((int64_t)seq.size() <= S.merge_len) {
      h = hash_bytes(seq.data(), seq.size());
      const int64_t gi = find(seq, h);
      if (gi >= 0) {
        const std::string &t = append_as ? *append_as : tagid;
        extras.push_back({arena.put(t.data(), t.size()), (uint32_t)t.size(), -1});
        MergeGroup &g = groups[(size_t)gi];
        if (g.extra_tail >= 0) extras[(size_t)g.extra_tail].next = (int64_t)extras.size() - 1;
        else g.extra_head = (int64_t)extras.size() - 1;
        g.extra_tail = (int64_t)extras.size() - 1;
        return true;
      }
    }
    if (tagid.size() >= (1ull << 32)) return false;
    char *d = S.text.alloc(tagid.size() + seq.size() + 3);
    d[0] = '>';
    memcpy(d + 1, tagid.data(), tagid.size());
    d[1 + tagid.size()] = '\n';
    memcpy(d + 2 + tagid.size(), seq.data(), seq.size());
    d[2 + tagid.size() + seq.size()] = '\n';
    if (!add_record(cur, d, (uint32_t)tagid.size(), seq.size(), true)) return false;
    data_size += size_add;
    rotate_if_full();
    if (may_register) {
      MergeGroup g;
      g.key = d + 2 + tagid.size();   // the record text itself is the key: no second copy of the sequence
      g.key_len = (uint32_t)seq.size();
      g.tag = d + 1;
      g.tag_len = (uint32_t)tagid.size();
      insert(h, groups.size());
      groups.push_back(g);
    }
    return true;
  }
};

// the whole pipeline; runs on S->runner (pa_ingest_start) or on the caller's thread (pa_ingest_run)
int run_pipeline(pa_ingest *S) {
  if (S->split_len && S->split_slide <= 0) { S->err = "split_slide must be positive (the reference would loop forever)"; return -2; }
  const bool dbg = getenv("PA_INGEST_DEBUG") != nullptr;
  auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t0 = now();
  // ---- phase 1: inflate + line split on worker threads, at most `ahead` files beyond the phase-2 cursor (bounded memory)
  const size_t nfiles = S->files.size();
  std::vector<char> ready(nfiles, 0);
  std::mutex mu;
  std::condition_variable cv;
  size_t consumed = 0;           // files phase 2 is done with
  bool stop = false;
  const int nworkers = (int)std::min<size_t>((size_t)std::max(1, S->nthreads - 1), std::max<size_t>(1, nfiles));
  const size_t ahead = (size_t)nworkers + 1;
  std::atomic<size_t> next{0};
  auto work = [&]() {
    for (;;) {
      const size_t i = next.fetch_add(1);
      if (i >= nfiles) break;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return stop || i < consumed + ahead; });
        if (stop) break;
      }
      if (inflate_file(S->files[i])) {
        split_lines(S->files[i]);
        if (S->files[i].format == 1 && !S->split_len) preformat(S->files[i], S->merge_len);
      }
      {
        std::lock_guard<std::mutex> lk(mu);
        ready[i] = 1;
      }
      cv.notify_all();
    }
  };
  std::vector<std::thread> th;
  for (int t = 0; t < nworkers; t++) th.emplace_back(work);
  auto finish_workers = [&]() {
    {
      std::lock_guard<std::mutex> lk(mu);
      stop = true;
    }
    cv.notify_all();
    for (auto &t : th) t.join();
  };

  double t_wait = 0;
  // ---- phase 2: the reference's single pass, in file order
  Builder B(*S);
  const bool fastpath = S->merge_len == 0 && S->split_len == 0;
  const int split_len = S->split_len, slide = S->split_slide;
  std::string tagid, seqstr, prefix_tag;
  std::string fast_text;      // FASTA fast path: the record being copied
  size_t fast_chunk = 0;
  uint32_t fast_idlen = 0;
  bool fast_open = false;
  auto fast_close = [&]() -> bool {
    if (!fast_open) return true;
    fast_open = false;
    const uint64_t body = fast_text.size() - (fast_idlen + 2);
    const char *d = S->text.put(fast_text.data(), fast_text.size());
    return B.add_record(fast_chunk, d, fast_idlen, body ? body - 1 : 0, body != 0);
  };
  bool ok = true;
  for (size_t fi = 0; fi < nfiles && ok; fi++) {
    FileJob &fj = S->files[fi];
    {
      const double w0 = now();
      std::unique_lock<std::mutex> lk(mu);
      cv.wait(lk, [&] { return ready[fi] != 0; });
      t_wait += now() - w0;
    }
    if (!fj.error.empty()) { S->err = fj.error; ok = false; break; }
    if (fi == 0) {  // records expected over all files, from the first file's record density per input byte
      struct stat st;
      double first = 0, total = 0;
      for (size_t k = 0; k < nfiles; k++)
        if (stat(S->files[k].path.c_str(), &st) == 0) { total += (double)st.st_size; if (k == 0) first = (double)st.st_size; }
      const size_t nrec0 = fj.preformatted ? fj.pre.size() : fj.lines.size() / (fj.format == 1 ? 4 : 2);
      if (first > 0) B.reserve((uint64_t)((double)nrec0 * (total / first) * 1.05), S->merge_len > 0);
    }
    int64_t *count = nullptr;
    for (auto &kv : S->counts)
      if (kv.first == fj.species) count = &kv.second;
    if (!count) { S->counts.emplace_back(fj.species, 0); count = &S->counts.back().second; }
    const std::string suffix = "_PhAlTag_" + fj.species + "_fastx" + std::to_string(fj.file_index);
    const std::string &T = fj.text;
    std::string seqid;
    bool have_id = false;  // Python truthiness of seqid: not None and not ''
    seqstr.clear();
    bool have_seqstr_obj = false;  // seqstr is not None (only matters for the trailing FASTA record)
    int64_t split_start = 0;
    auto make_tag = [&](int64_t a, int64_t b) {
      tagid = seqid;
      tagid += suffix;
      if (split_len) { tagid += "_split_"; tagid += std::to_string(a); tagid += '_'; tagid += std::to_string(b); }
    };
    auto emit_windows = [&](std::string &s) {  // while len(seqstr) > split_len (lib/aln.py:338-357, 381-399)
      size_t pos = 0;
      while ((int64_t)(s.size() - pos) > split_len) {
        make_tag(split_start + 1, split_start + split_len);
        ok = ok && B.emit(tagid, std::string_view(s).substr(pos, (size_t)split_len), (uint64_t)split_len, split_len <= S->merge_len);
        split_start += slide;
        pos += (size_t)slide;
        if (pos > s.size()) pos = s.size();
      }
      if (pos) s.erase(0, pos);
    };
    if (fj.preformatted) {   // unsplit FASTQ: texts and hashes come from the worker thread (preformat), table slots prefetched
      const size_t n = fj.pre.size();
      for (size_t k = 0; k < n && ok; k++) {
        if (k + 12 < n) B.prefetch_hashes(fj.pre[k + 12].hseq, fj.pre[k + 12].hid);
        const FileJob::PreRec &R = fj.pre[k];
        ok = B.emit_pre(fj.fmt + R.off, R.tag_len, R.seq_len, R.hseq, R.hid);
        (*count)++;
      }
      if (fj.fmt) { S->fmt_bufs.emplace_back(fj.fmt, fj.fmt_bytes); fj.fmt = nullptr; }
      std::vector<FileJob::PreRec>().swap(fj.pre);
    }
    uint64_t line_num = 0;
    for (const auto &ln : fj.lines) {
      const char *lp = T.data() + ln.first;
      const uint64_t ll = ln.second;
      if (fj.format == 1 && ll > 0 && lp[0] == '@' && line_num % 4 == 0) {
        seqid.assign(lp + 1, ll - 1);
        for (char &c : seqid)
          if (c == ' ') c = '_';
        have_id = !seqid.empty();
      } else if (fj.format == 2 && ll > 0 && lp[0] == '>') {
        if (have_id) {
          if (!seqstr.empty()) {
            make_tag(split_start + 1, split_start + (int64_t)seqstr.size());
            ok = ok && B.emit(tagid, seqstr, seqstr.size(), (int64_t)seqstr.size() <= S->merge_len);
          }
          (*count)++;
          split_start = 0;
        }
        // seqid = line.lstrip('>').split()[0]
        uint64_t a = 0;
        while (a < ll && lp[a] == '>') a++;
        while (a < ll && py_space((unsigned char)lp[a])) a++;
        uint64_t b = a;
        while (b < ll && !py_space((unsigned char)lp[b])) b++;
        if (a == b) { S->err = "FASTA header without an identifier in " + fj.path + " (the reference raises IndexError)"; ok = false; break; }
        seqid.assign(lp + a, b - a);
        have_id = true;
        seqstr.clear();
        have_seqstr_obj = true;
        if (fastpath) {
          ok = ok && fast_close();
          B.rotate_if_full();
          fast_chunk = B.cur;
          fast_text.assign(1, '>');
          fast_text += seqid;
          fast_text += suffix;
          fast_idlen = (uint32_t)(fast_text.size() - 1);
          fast_text += '\n';
          fast_open = true;
        }
      } else if (have_id) {
        if (fj.format == 1) {
          seqstr.assign(lp, ll);
          if (split_len) {
            split_start = 0;
            emit_windows(seqstr);
          }
          make_tag(split_start + 1, split_start + (int64_t)seqstr.size());
          ok = ok && B.emit(tagid, seqstr, seqstr.size(), (int64_t)seqstr.size() <= S->merge_len);
          (*count)++;
          have_id = false;
          seqid.clear();
        } else if (fastpath) {
          fast_text.append(lp, ll);
          fast_text += '\n';
          B.data_size += ll;
        } else {
          seqstr.append(lp, ll);
          if (split_len) emit_windows(seqstr);
        }
      }
      line_num++;
    }
    if (ok && fj.format == 2 && have_seqstr_obj && !seqstr.empty()) {  // trailing FASTA record (lib/aln.py:401-420)
      make_tag(split_start + 1, split_start + (int64_t)seqstr.size());
      ok = ok && B.emit(tagid, seqstr, seqstr.size(), (int64_t)seqstr.size() <= S->merge_len, &seqid);
      (*count)++;
    }
    ok = ok && fast_close();
    std::string().swap(fj.text);
    std::vector<std::pair<uint64_t, uint64_t>>().swap(fj.lines);
    {
      std::lock_guard<std::mutex> lk(mu);
      consumed = fi + 1;
    }
    cv.notify_all();
  }
  finish_workers();
  if (!ok) {
    if (S->err.empty()) S->err = "ingest: record table overflow";
    return -2;
  }
  const double t2 = now();
  // ---- outputs
  for (auto &kv : S->counts) { S->counts_txt += kv.first; S->counts_txt += '\t'; S->counts_txt += std::to_string(kv.second); S->counts_txt += '\n'; }
  for (auto &g : B.groups)
    if (g.extra_head >= 0) {
      S->merged_txt.append(g.tag, g.tag_len);
      for (int64_t e = g.extra_head; e >= 0; e = B.extras[(size_t)e].next) {
        S->merged_txt += ' ';
        S->merged_txt.append(B.extras[(size_t)e].tag, B.extras[(size_t)e].tag_len);
      }
      S->merged_txt += '\n';
    }
  if (B.folded) {  // the chunk fold permuted the record order (inputs beyond 10 Gbases): final position -> emission index
    S->final_order.reserve(S->recs.n);
    for (auto &runs : B.chunks)
      for (auto &r : runs)
        for (uint64_t k = 0; k < r.second; k++) S->final_order.push_back(r.first + k);
  }
  if (dbg) fprintf(stderr, "pa_ingest: phase 2 %.3f s (of which %.3f s waiting for inflate), outputs %.3f s, %llu records, %llu repeated ids\n",
                   t2 - t0, t_wait, now() - t2, (unsigned long long)S->recs.n, (unsigned long long)S->id_dups);
  return 0;
}

inline uint64_t final_to_emitted(const pa_ingest *S, uint64_t pos) { return S->final_order.empty() ? pos : S->final_order[pos]; }

// final all.raw.fasta text + record table, built on first use (the streaming path never needs it)
void materialise_raw(pa_ingest *S) {
  if (S->have_raw) return;
  const uint64_t n = S->recs.n;
  uint64_t total = 0;
  for (uint64_t i = 0; i < n; i++) total += S->recs.at(i).text_len();
  S->raw.reserve(total);
  S->id_off.reserve(n); S->id_len.reserve(n); S->seq_off.reserve(n); S->seq_len.reserve(n);
  for (uint64_t pos = 0; pos < n; pos++) {
    const Rec &r = S->recs.at(final_to_emitted(S, pos));
    const uint64_t o = S->raw.size();
    S->raw.append(r.text, r.text_len());
    S->id_off.push_back(o + 1);
    S->id_len.push_back(r.id_len);
    S->seq_off.push_back(o + r.id_len + 2);
    S->seq_len.push_back(r.body_len);
  }
  S->have_raw = true;
}

}  // namespace

extern "C" pa_ingest *pa_ingest_create(int merge_len, int split_len, int split_slide, int nthreads) {
  pa_ingest *S = new pa_ingest();
  S->merge_len = merge_len;
  S->split_len = split_len > 0 ? split_len : 0;
  S->split_slide = S->split_len ? (split_slide > 0 ? split_slide : S->split_len / 2) : 0;
  S->nthreads = nthreads > 0 ? nthreads : 1;
  return S;
}

extern "C" void pa_ingest_destroy(pa_ingest *S) { delete S; }
extern "C" const char *pa_ingest_last_error(pa_ingest *S) { return S ? S->err.c_str() : "null ingest handle"; }

extern "C" int pa_ingest_add_file(pa_ingest *S, const char *path, const char *species, int file_index, const char *format) {
  if (!S || S->started) return -1;
  FileJob fj;
  fj.path = path;
  fj.species = species;
  fj.file_index = file_index;
  fj.format = !strcmp(format, "fastq") ? 1 : !strcmp(format, "fasta") ? 2 : !strcmp(format, "guess") ? 0 : -1;
  if (fj.format < 0) { S->err = "file format must be guess, fastq or fasta"; return -2; }
  S->files.push_back(std::move(fj));
  return 0;
}

extern "C" int pa_ingest_start(pa_ingest *S) {
  if (!S || S->started) return -1;
  S->started = true;
  S->runner = std::thread([S] {
    S->rc = run_pipeline(S);
    S->finished.store(true, std::memory_order_release);
  });
  return 0;
}

extern "C" int pa_ingest_wait(pa_ingest *S) {
  if (!S || !S->started) return -1;
  std::lock_guard<std::mutex> lk(S->wait_mu);   // any number of threads may wait; one of them joins
  if (S->runner.joinable()) S->runner.join();
  return S->rc;
}

extern "C" int pa_ingest_run(pa_ingest *S) {
  const int rc = pa_ingest_start(S);
  return rc ? rc : pa_ingest_wait(S);
}

extern "C" int pa_ingest_done(pa_ingest *S) { return S && S->finished.load(std::memory_order_acquire) ? 1 : 0; }

extern "C" int64_t pa_ingest_emitted(pa_ingest *S) { return S ? (int64_t)S->recs.published.load(std::memory_order_acquire) : -1; }

extern "C" int64_t pa_ingest_num_records(pa_ingest *S) { return S && S->finished.load(std::memory_order_acquire) && S->rc == 0 ? (int64_t)S->recs.n : -1; }

extern "C" int64_t pa_ingest_repeated_ids(pa_ingest *S) { return S && S->finished.load(std::memory_order_acquire) && S->rc == 0 ? (int64_t)S->id_dups : -1; }

extern "C" int pa_ingest_final_order(pa_ingest *S, uint64_t *emitted_index) {
  if (!S || !S->finished.load(std::memory_order_acquire) || S->rc != 0) return -1;
  if (S->final_order.empty()) return 0;
  if (emitted_index) memcpy(emitted_index, S->final_order.data(), S->final_order.size() * sizeof(uint64_t));
  return 1;
}

extern "C" int pa_ingest_views(pa_ingest *S, const char **raw, uint64_t *raw_len, const uint64_t **id_off, const uint64_t **id_len,
                               const uint64_t **seq_off, const uint64_t **seq_len, const char **counts, uint64_t *counts_len,
                               const char **merged, uint64_t *merged_len) {
  if (!S || !S->finished.load(std::memory_order_acquire) || S->rc != 0) return -1;
  if (raw || raw_len || id_off || id_len || seq_off || seq_len) materialise_raw(S);
  if (raw) *raw = S->raw.data();
  if (raw_len) *raw_len = S->raw.size();
  if (id_off) *id_off = S->id_off.data();
  if (id_len) *id_len = S->id_len.data();
  if (seq_off) *seq_off = S->seq_off.data();
  if (seq_len) *seq_len = S->seq_len.data();
  if (counts) *counts = S->counts_txt.data();
  if (counts_len) *counts_len = S->counts_txt.size();
  if (merged) *merged = S->merged_txt.data();
  if (merged_len) *merged_len = S->merged_txt.size();
  return 0;
}

extern "C" int pa_ingest_write(pa_ingest *S, const char *raw_fasta, const char *total_counts, const char *merged_redundance) {
  if (!S || !S->finished.load(std::memory_order_acquire) || S->rc != 0) return -1;
  auto dump = [&](const char *path, const std::string &s) -> bool {
    if (!path) return true;
    FILE *f = fopen(path, "wb");
    if (!f) { S->err = std::string("cannot write ") + path; return false; }
    const bool ok = fwrite(s.data(), 1, s.size(), f) == s.size();
    fclose(f);
    if (!ok) S->err = std::string("short write on ") + path;
    return ok;
  };
  if (raw_fasta) {  // straight from the record texts, contiguous spans in one write: no second copy of the file in memory
    FILE *f = fopen(raw_fasta, "wb");
    if (!f) { S->err = std::string("cannot write ") + raw_fasta; return -2; }
    setvbuf(f, nullptr, _IOFBF, 1 << 22);
    const uint64_t n = S->recs.n;
    bool ok = true;
    const char *span = nullptr;
    uint64_t span_len = 0;
    for (uint64_t pos = 0; pos < n && ok; pos++) {
      const Rec &r = S->recs.at(final_to_emitted(S, pos));
      if (span && span + span_len == r.text) span_len += r.text_len();
      else {
        if (span) ok = fwrite(span, 1, span_len, f) == span_len;
        span = r.text;
        span_len = r.text_len();
      }
    }
    if (ok && span) ok = fwrite(span, 1, span_len, f) == span_len;
    ok = (fclose(f) == 0) && ok;
    if (!ok) { S->err = std::string("short write on ") + raw_fasta; return -2; }
  }
  return dump(total_counts, S->counts_txt) && dump(merged_redundance, S->merged_txt) ? 0 : -2;
}

static int64_t pack_records(pa_ingest *S, uint64_t first, uint64_t count, bool final_order, char *nt_out, uint64_t cap, uint64_t *off_out) {
  uint64_t pos = 0;
  for (uint64_t i = 0; i < count; i++) {
    const Rec &r = S->recs.at(final_order ? final_to_emitted(S, first + i) : first + i);
    const char *body = r.text + r.id_len + 2;
    off_out[i] = pos;
    if (pos + r.body_len > cap) { S->err = "pa_ingest_pack: output buffer too small"; return -3; }
    if (!memchr(body, '\n', r.body_len)) { memcpy(nt_out + pos, body, r.body_len); pos += r.body_len; }
    else
      for (uint64_t k = 0; k < r.body_len; k++)  // the FASTA fast path keeps line breaks inside a record: drop them
        if (body[k] != '\n') nt_out[pos++] = body[k];
  }
  off_out[count] = pos;
  return (int64_t)pos;
}

// Sequences [first, first + count) of the FINAL record order (all.raw.fasta order), concatenated for pa_load_reads
// (off_out has count + 1 entries). Returns the number of bytes written or a negative error.
extern "C" int64_t pa_ingest_pack(pa_ingest *S, uint64_t first, uint64_t count, char *nt_out, uint64_t cap, uint64_t *off_out) {
  if (!S || !S->finished.load(std::memory_order_acquire) || S->rc != 0) return -1;
  if (first + count > S->recs.n) { S->err = "pa_ingest_pack: range out of bounds"; return -2; }
  return pack_records(S, first, count, true, nt_out, cap, off_out);
}

// Same in EMISSION order, valid while the ingest is still running for records below pa_ingest_emitted()
extern "C" int64_t pa_ingest_pack_emitted(pa_ingest *S, uint64_t first, uint64_t count, char *nt_out, uint64_t cap, uint64_t *off_out) {
  if (!S) return -1;
  if (first + count > S->recs.published.load(std::memory_order_acquire)) { S->err = "pa_ingest_pack_emitted: records not emitted yet"; return -2; }
  return pack_records(S, first, count, false, nt_out, cap, off_out);
}

// total sequence bytes of emitted records [first, first + count) (sizing of the pack buffer / byte-bounded chunks)
extern "C" int64_t pa_ingest_emitted_bytes(pa_ingest *S, uint64_t first, uint64_t count) {
  if (!S || first + count > S->recs.published.load(std::memory_order_acquire)) return -1;
  uint64_t tot = 0;
  for (uint64_t i = 0; i < count; i++) tot += S->recs.at(first + i).body_len;
  return (int64_t)tot;
}

// ids (what = 0) or sequences (what = 1, line breaks of the FASTA fast path removed) of selected records (emission
// indices), concatenated; out_off has n + 1 entries; out may be NULL (sizes only). Returns bytes written or <0.
static int64_t emitted_fields(pa_ingest *S, int what, const uint64_t *idx, uint64_t n, char *out, uint64_t cap, uint64_t *out_off) {
  if (!S) return -1;
  const uint64_t have = S->recs.published.load(std::memory_order_acquire);
  uint64_t pos = 0;
  for (uint64_t i = 0; i < n; i++) {
    if (idx[i] >= have) { S->err = "pa_ingest_emitted_*: record not emitted yet"; return -2; }
    const Rec &r = S->recs.at(idx[i]);
    out_off[i] = pos;
    const char *src = what ? r.text + r.id_len + 2 : r.text + 1;
    const uint64_t len = what ? r.body_len : r.id_len;
    if (out && pos + len > cap) { S->err = "pa_ingest_emitted_*: output buffer too small"; return -3; }
    if (what && memchr(src, '\n', len)) {
      for (uint64_t k = 0; k < len; k++)
        if (src[k] != '\n') { if (out) out[pos] = src[k]; pos++; }
    } else {
      if (out) memcpy(out + pos, src, len);
      pos += len;
    }
  }
  out_off[n] = pos;
  return (int64_t)pos;
}
extern "C" int64_t pa_ingest_emitted_ids(pa_ingest *S, const uint64_t *idx, uint64_t n, char *out, uint64_t cap, uint64_t *out_off) {
  return emitted_fields(S, 0, idx, n, out, cap, out_off);
}
extern "C" int64_t pa_ingest_emitted_seqs(pa_ingest *S, const uint64_t *idx, uint64_t n, char *out, uint64_t cap, uint64_t *out_off) {
  return emitted_fields(S, 1, idx, n, out, cap, out_off);
}

// ---------------------------------------------------------------------------------------------------------------------
// pa_fasta_*: all.raw.fasta -> (ids, concatenated nucleotides, offsets) for pa_load_reads, replacing the Python line loop
// of dropin.read_raw_fasta (19 s per 10 M reads).  Same semantics as the reference's read_fastx(..., 'fasta') dictionary
// (lib/library.py:186-198): the id is the header up to the first blank, multi-line records are joined, a repeated id
// keeps its FIRST position but its LAST sequence.
// ---------------------------------------------------------------------------------------------------------------------
struct pa_fasta {
  BigVec<char> ids, nt;          // untouched until the pack threads write them (no serial zero fill)
  BigVec<uint64_t> id_off, off;
  std::string err;
};

extern "C" pa_fasta *pa_fasta_open(const char *path, char *err, int errlen) {
  FileJob fj;
  fj.path = path ? path : "";
  bool plain = false;
  if (FILE *fp = fopen(fj.path.c_str(), "rb")) {   // uncompressed input: one read of the whole file, no zlib copy loop
    unsigned char magic[2] = {0, 0};
    const size_t got = fread(magic, 1, 2, fp);
    if (!(got == 2 && magic[0] == 0x1f && magic[1] == 0x8b) && fseek(fp, 0, SEEK_END) == 0) {
      const long sz = ftell(fp);
      if (sz >= 0 && fseek(fp, 0, SEEK_SET) == 0) {
        fj.text.resize((size_t)sz);
        plain = sz == 0 || fread(&fj.text[0], 1, (size_t)sz, fp) == (size_t)sz;
      }
    }
    fclose(fp);
  }
  if (!plain && !inflate_file(fj)) {
    if (err && errlen > 0) snprintf(err, (size_t)errlen, "%s", fj.error.c_str());
    return nullptr;
  }
  const bool dbg = getenv("PA_INGEST_DEBUG") != nullptr;
  auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double tt0 = now();
  const std::string &t = fj.text;
  const size_t n = t.size();
  // ---- 1. the file cut at header lines into one range per thread; every range lists its records as they stand in the
  // text (no dictionary yet): id, first sequence line, further lines joined on the side, hash of the id
  struct Raw { uint64_t id_off, seq_off, seq_len, hash; uint32_t id_len; int32_t multi; };   // multi: index into the range's `joined`, -1 = one line
  struct Range { size_t begin = 0, end = 0; std::vector<Raw> recs; std::vector<std::string> joined; };
  int nth = (int)std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
  if (n < ((size_t)1 << 20)) nth = 1;
  if (const char *e = getenv("PA_INGEST_THREADS")) nth = std::max(1, std::min(256, atoi(e)));   // tests: several ranges on small files
  std::vector<Range> ranges((size_t)nth);
  for (int k = 0; k < nth; k++) {
    size_t pos = k == 0 ? 0 : n / (size_t)nth * (size_t)k;
    if (k > 0) {   // forward to the next line that starts with '>'
      for (;;) {
        const void *nl = pos < n ? memchr(t.data() + pos, '\n', n - pos) : nullptr;
        if (!nl) { pos = n; break; }
        pos = (size_t)((const char *)nl - t.data()) + 1;
        if (pos < n && t[pos] == '>') break;
      }
    }
    ranges[(size_t)k].begin = pos;
  }
  for (int k = 0; k < nth; k++) ranges[(size_t)k].end = k + 1 < nth ? ranges[(size_t)k + 1].begin : n;
  auto parse_range = [&](Range &R) {
    R.recs.reserve((R.end - R.begin) / 160 + 16);
    size_t i = R.begin;
    long cur = -1;
    int nparts = 0;
    while (i < R.end) {
      const void *nl = memchr(t.data() + i, '\n', R.end - i);
      const size_t e = nl ? (size_t)((const char *)nl - t.data()) : R.end;
      size_t r = e;
      while (r > i && (t[r - 1] == '\r' || t[r - 1] == '\n')) r--;   // rstrip(b"\r\n") as in dropin.read_raw_fasta
      if (r > i && t[i] == '>') {
        size_t b = i + 1;
        while (b < r && t[b] != ' ') b++;
        R.recs.push_back(Raw{(uint64_t)(i + 1), 0, 0, hash_bytes(t.data() + i + 1, b - i - 1), (uint32_t)(b - i - 1), -1});
        if (b - i - 1 > 0xffffffffull) R.recs.back().id_len = 0xffffffffu;   // reported below
        cur = (long)R.recs.size() - 1;
        nparts = 0;
      } else if (cur >= 0) {
        Raw &W = R.recs[(size_t)cur];
        if (nparts == 0) { W.seq_off = i; W.seq_len = r - i; }
        else if (r > i) {
          if (W.multi < 0) { W.multi = (int32_t)R.joined.size(); R.joined.emplace_back(t.data() + W.seq_off, W.seq_len); }
          R.joined[(size_t)W.multi].append(t.data() + i, r - i);
        }
        nparts++;
      }
      i = e + 1;
    }
  };
  {
    std::vector<std::thread> th;
    for (int k = 1; k < nth; k++) th.emplace_back([&, k] { parse_range(ranges[(size_t)k]); });
    parse_range(ranges[0]);
    for (auto &x : th) x.join();
  }
  const double tt1 = now();
  // ---- 2. read_fastx's dictionary, in file order: a repeated id keeps its position and takes the later record's sequence
  size_t nraw = 0;
  for (auto &R : ranges) nraw += R.recs.size();
  if (nraw >= 0xfffffffeull) {   // 32-bit slots: say so, never wrap
    if (err && errlen > 0) snprintf(err, (size_t)errlen, "%s: more than 4,294,967,294 records", fj.path.c_str());
    return nullptr;
  }
  struct Ent { uint32_t range, idx; };      // the raw record an entry takes its id from / its sequence from
  BigVec<Ent> id_src, seq_src;
  id_src.reserve(nraw);
  seq_src.reserve(nraw);
  size_t tcap = 1024;
  while (tcap < 2 * (nraw + 16)) tcap <<= 1;
  BigVec<uint32_t> slots(tcap, 0u);        // open addressing: entry index + 1
  for (size_t k = 0; k < ranges.size(); k++) {
    const auto &recs = ranges[k].recs;
    for (size_t j = 0; j < recs.size(); j++) {
      if (j + 16 < recs.size()) __builtin_prefetch(&slots[(size_t)(recs[j + 16].hash ^ (recs[j + 16].hash >> 32)) & (tcap - 1)]);
      const Raw &W = recs[j];
      size_t p = (size_t)(W.hash ^ (W.hash >> 32)) & (tcap - 1);
      bool found = false;
      while (slots[p]) {
        const Ent &E = id_src[slots[p] - 1];
        const Raw &O = ranges[E.range].recs[E.idx];
        if (O.hash == W.hash && O.id_len == W.id_len && memcmp(t.data() + O.id_off, t.data() + W.id_off, W.id_len) == 0) { found = true; break; }
        p = (p + 1) & (tcap - 1);
      }
      if (found) seq_src[slots[p] - 1] = Ent{(uint32_t)k, (uint32_t)j};
      else {
        id_src.push_back(Ent{(uint32_t)k, (uint32_t)j});
        seq_src.push_back(Ent{(uint32_t)k, (uint32_t)j});
        slots[p] = (uint32_t)id_src.size();
      }
    }
  }
  const double tt2 = now();
  // ---- 3. pack ids and sequences: offsets by prefix sum, the copies on all threads
  pa_fasta *F = new pa_fasta();
  const size_t nent = id_src.size();
  F->id_off.resize(nent + 1);
  F->off.resize(nent + 1);
  auto seq_of = [&](const Ent &E, const char *&ptr) -> uint64_t {
    const Range &R = ranges[E.range];
    const Raw &W = R.recs[E.idx];
    if (W.multi >= 0) { ptr = R.joined[(size_t)W.multi].data(); return R.joined[(size_t)W.multi].size(); }
    ptr = t.data() + W.seq_off;
    return W.seq_len;
  };
  uint64_t idb = 0, ntb = 0;
  for (size_t e = 0; e < nent; e++) {
    F->id_off[e] = idb;
    F->off[e] = ntb;
    idb += ranges[id_src[e].range].recs[id_src[e].idx].id_len;
    const char *ptr;
    ntb += seq_of(seq_src[e], ptr);
  }
  F->id_off[nent] = idb;
  F->off[nent] = ntb;
  F->ids.reserve(idb + 1);       // capacity only: BigVec memory above 2 MB is fresh mmap, first touched by the pack threads
  F->nt.reserve(ntb + 16);
  auto pack_part = [&](size_t e0, size_t e1) {
    for (size_t e = e0; e < e1; e++) {
      const Raw &W = ranges[id_src[e].range].recs[id_src[e].idx];
      memcpy(F->ids.data() + F->id_off[e], t.data() + W.id_off, W.id_len);
      const char *ptr;
      const uint64_t len = seq_of(seq_src[e], ptr);
      if (len) memcpy(F->nt.data() + F->off[e], ptr, len);
    }
  };
  {
    std::vector<std::thread> th;
    const size_t per = (nent + (size_t)nth - 1) / (size_t)nth;
    for (int k = 1; k < nth; k++) th.emplace_back([&, k] { pack_part(std::min(nent, per * (size_t)k), std::min(nent, per * (size_t)(k + 1))); });
    pack_part(0, std::min(nent, per));
    for (auto &x : th) x.join();
  }
  if (dbg) fprintf(stderr, "pa_fasta_open: parse %.3f s (%d threads), dictionary %.3f s, pack %.3f s\n", tt1 - tt0, nth, tt2 - tt1, now() - tt2);
  return F;
}
extern "C" int64_t pa_fasta_num_records(pa_fasta *F) { return F ? (int64_t)F->off.size() - 1 : -1; }
extern "C" int pa_fasta_views(pa_fasta *F, const char **ids, const uint64_t **id_off, const char **nt, const uint64_t **off) {
  if (!F) return -1;
  if (ids) *ids = F->ids.data();
  if (id_off) *id_off = F->id_off.data();
  if (nt) *nt = F->nt.data();
  if (off) *off = F->off.data();
  return 0;
}
extern "C" void pa_fasta_close(pa_fasta *F) { delete F; }

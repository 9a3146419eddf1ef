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
#include <zlib.h>

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
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
  std::vector<std::pair<uint64_t, uint32_t>> lines;  // (offset, length after rstrip)
  std::string error;
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
    fj.lines.emplace_back((uint64_t)i, (uint32_t)(r - i));
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

// bump allocator: keys and tags are never freed individually
struct Arena {
  std::vector<std::unique_ptr<char[]>> slabs;
  size_t left = 0;
  char *cur = nullptr;
  const char *put(const char *p, size_t n) {
    if (n > left) {
      const size_t sz = std::max<size_t>(n, (size_t)16 << 20);
      slabs.emplace_back(new char[sz]);
      cur = slabs.back().get();
      left = sz;
    }
    char *d = cur;
    memcpy(d, p, n);
    cur += n;
    left -= n;
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

}  // namespace

struct pa_ingest {
  int merge_len = 250, split_len = 0, split_slide = 0, nthreads = 1;
  std::vector<FileJob> files;
  std::string err;
  bool finished = false;
  // results
  std::string raw;  // final all.raw.fasta text
  std::vector<std::pair<std::string, int64_t>> counts;
  std::string merged_txt, counts_txt;
  // record table of `raw` in final order (ids / sequences for direct GPU feeding)
  std::vector<uint64_t> id_off, id_len, seq_off, seq_len;
};

namespace {

struct Builder {
  pa_ingest &S;
  std::vector<std::string> chunks;
  size_t cur = 0;
  uint64_t data_size = 0, each_size = 10000000;
  // open-addressing table (hash, group index + 1); groups in insertion order == Python dict order
  std::vector<std::pair<uint64_t, uint64_t>> table = std::vector<std::pair<uint64_t, uint64_t>>(1 << 16);
  std::vector<MergeGroup> groups;
  std::vector<ExtraTag> extras;
  Arena arena;

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
      std::vector<std::pair<uint64_t, uint64_t>> nt(table.size() * 2);
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

  explicit Builder(pa_ingest &s) : S(s) { chunks.emplace_back(); }

  void rotate_if_full() {  // lib/aln.py:313-318 + check_file_num
    if (data_size >= each_size) {
      data_size = 0;
      cur++;
      if (cur >= 1000) {
        for (size_t i = 100; i < 1000; i++) { chunks[i % 100] += chunks[i]; }
        chunks.resize(100);
        cur = 100;
        each_size *= 10;
      }
      if (chunks.size() <= cur) chunks.resize(cur + 1);
    }
  }
  // one sequence (or split fragment) with its tag: merge or write
  void emit(const std::string &tagid, std::string_view seq, uint64_t size_add, bool may_register, const std::string *append_as = nullptr) {
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
        return;
      }
    }
    std::string &c = chunks[cur];
    c += '>'; c += tagid; c += '\n'; c.append(seq.data(), seq.size()); c += '\n';
    data_size += size_add;
    rotate_if_full();
    if (may_register) {
      MergeGroup g;
      g.key = arena.put(seq.data(), seq.size());
      g.key_len = (uint32_t)seq.size();
      g.tag = arena.put(tagid.data(), tagid.size());
      g.tag_len = (uint32_t)tagid.size();
      insert(h, groups.size());
      groups.push_back(g);
    }
  }
};

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
  if (!S || S->finished) return -1;
  FileJob fj;
  fj.path = path;
  fj.species = species;
  fj.file_index = file_index;
  fj.format = !strcmp(format, "fastq") ? 1 : !strcmp(format, "fasta") ? 2 : !strcmp(format, "guess") ? 0 : -1;
  if (fj.format < 0) { S->err = "file format must be guess, fastq or fasta"; return -2; }
  S->files.push_back(std::move(fj));
  return 0;
}

extern "C" int pa_ingest_run(pa_ingest *S) {
  if (!S || S->finished) return -1;
  if (S->split_len && S->split_slide <= 0) { S->err = "split_slide must be positive (the reference would loop forever)"; return -2; }
  const bool dbg = getenv("PA_INGEST_DEBUG") != nullptr;
  auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t0 = now();
  // ---- phase 1: inflate + line split, one file per worker
  {
    std::atomic<size_t> next{0};
    auto work = [&]() {
      for (;;) {
        const size_t i = next.fetch_add(1);
        if (i >= S->files.size()) break;
        if (inflate_file(S->files[i])) split_lines(S->files[i]);
      }
    };
    std::vector<std::thread> th;
    const int nt = (int)std::min<size_t>((size_t)S->nthreads, std::max<size_t>(1, S->files.size()));
    for (int t = 1; t < nt; t++) th.emplace_back(work);
    work();
    for (auto &t : th) t.join();
  }
  for (auto &f : S->files)
    if (!f.error.empty()) { S->err = f.error; return -2; }

  const double t1 = now();
  // ---- phase 2: the reference's single pass, in file order
  Builder B(*S);
  const bool fastpath = S->merge_len == 0 && S->split_len == 0;
  const int split_len = S->split_len, slide = S->split_slide;
  std::string tagid, seqstr, prefix_tag;
  for (FileJob &fj : S->files) {
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
        B.emit(tagid, std::string_view(s).substr(pos, (size_t)split_len), (uint64_t)split_len, split_len <= S->merge_len);
        split_start += slide;
        pos += (size_t)slide;
        if (pos > s.size()) pos = s.size();
      }
      if (pos) s.erase(0, pos);
    };
    uint64_t line_num = 0;
    for (const auto &ln : fj.lines) {
      const char *lp = T.data() + ln.first;
      const uint32_t ll = ln.second;
      if (fj.format == 1 && ll > 0 && lp[0] == '@' && line_num % 4 == 0) {
        seqid.assign(lp + 1, ll - 1);
        for (char &c : seqid)
          if (c == ' ') c = '_';
        have_id = !seqid.empty();
      } else if (fj.format == 2 && ll > 0 && lp[0] == '>') {
        if (have_id) {
          if (!seqstr.empty()) {
            make_tag(split_start + 1, split_start + (int64_t)seqstr.size());
            B.emit(tagid, seqstr, seqstr.size(), (int64_t)seqstr.size() <= S->merge_len);
          }
          (*count)++;
          split_start = 0;
        }
        // seqid = line.lstrip('>').split()[0]
        uint32_t a = 0;
        while (a < ll && lp[a] == '>') a++;
        while (a < ll && py_space((unsigned char)lp[a])) a++;
        uint32_t b = a;
        while (b < ll && !py_space((unsigned char)lp[b])) b++;
        if (a == b) { S->err = "FASTA header without an identifier in " + fj.path + " (the reference raises IndexError)"; return -2; }
        seqid.assign(lp + a, b - a);
        have_id = true;
        seqstr.clear();
        have_seqstr_obj = true;
        if (fastpath) {
          B.rotate_if_full();
          std::string &c = B.chunks[B.cur];
          c += '>'; c += seqid; c += suffix; c += '\n';
        }
      } else if (have_id) {
        if (fj.format == 1) {
          seqstr.assign(lp, ll);
          if (split_len) {
            split_start = 0;
            emit_windows(seqstr);
          }
          make_tag(split_start + 1, split_start + (int64_t)seqstr.size());
          B.emit(tagid, seqstr, seqstr.size(), (int64_t)seqstr.size() <= S->merge_len);
          (*count)++;
          have_id = false;
          seqid.clear();
        } else if (fastpath) {
          std::string &c = B.chunks[B.cur];
          c.append(lp, ll);
          c += '\n';
          B.data_size += ll;
        } else {
          seqstr.append(lp, ll);
          if (split_len) emit_windows(seqstr);
        }
      }
      line_num++;
    }
    if (fj.format == 2 && have_seqstr_obj && !seqstr.empty()) {  // trailing FASTA record (lib/aln.py:401-420)
      make_tag(split_start + 1, split_start + (int64_t)seqstr.size());
      B.emit(tagid, seqstr, seqstr.size(), (int64_t)seqstr.size() <= S->merge_len, &seqid);
      (*count)++;
    }
    std::string().swap(fj.text);
    std::vector<std::pair<uint64_t, uint32_t>>().swap(fj.lines);
  }
  const double t2 = now();
  // ---- outputs
  size_t total = 0;
  for (auto &c : B.chunks) total += c.size();
  S->raw.reserve(total);
  for (auto &c : B.chunks) { S->raw += c; std::string().swap(c); }
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
  // record table (the fast path may hold multi-line records: sequence = first line..last line, newlines inside)
  {
    const std::string &R = S->raw;
    size_t i = 0, n = R.size();
    while (i < n) {
      if (R[i] != '>') { S->err = "internal: malformed raw FASTA"; return -1; }
      const char *nl = (const char *)memchr(R.data() + i, '\n', n - i);
      const size_t e = nl ? (size_t)(nl - R.data()) : n;
      S->id_off.push_back(i + 1);
      S->id_len.push_back(e - i - 1);
      size_t s = e + 1, j = s;
      // sequence lines until the next header at line start
      while (j < n && R[j] != '>') {
        const char *nl2 = (const char *)memchr(R.data() + j, '\n', n - j);
        j = nl2 ? (size_t)(nl2 - R.data()) + 1 : n;
      }
      S->seq_off.push_back(s);
      S->seq_len.push_back(j > s ? j - s - 1 : 0);
      i = j;
    }
  }
  if (dbg) fprintf(stderr, "pa_ingest_run: inflate+split %.3f s, tag+merge %.3f s, outputs %.3f s\n", t1 - t0, t2 - t1, now() - t2);
  S->finished = true;
  return 0;
}

extern "C" int64_t pa_ingest_num_records(pa_ingest *S) { return S && S->finished ? (int64_t)S->id_off.size() : -1; }

extern "C" int pa_ingest_views(pa_ingest *S, const char **raw, uint64_t *raw_len, const uint64_t **id_off, const uint64_t **id_len,
                               const uint64_t **seq_off, const uint64_t **seq_len, const char **counts, uint64_t *counts_len,
                               const char **merged, uint64_t *merged_len) {
  if (!S || !S->finished) return -1;
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
  if (!S || !S->finished) return -1;
  auto dump = [&](const char *path, const std::string &s) -> bool {
    if (!path) return true;
    FILE *f = fopen(path, "wb");
    if (!f) { S->err = std::string("cannot write ") + path; return false; }
    const bool ok = fwrite(s.data(), 1, s.size(), f) == s.size();
    fclose(f);
    if (!ok) S->err = std::string("short write on ") + path;
    return ok;
  };
  return dump(raw_fasta, S->raw) && dump(total_counts, S->counts_txt) && dump(merged_redundance, S->merged_txt) ? 0 : -2;
}

// Sequences [first, first + count) of the final record order, concatenated for pa_load_reads
// (off_out has count + 1 entries). Returns the number of bytes written or a negative error.
extern "C" int64_t pa_ingest_pack(pa_ingest *S, uint64_t first, uint64_t count, char *nt_out, uint64_t cap, uint64_t *off_out) {
  if (!S || !S->finished) return -1;
  if (first + count > S->id_off.size()) { S->err = "pa_ingest_pack: range out of bounds"; return -2; }
  uint64_t pos = 0;
  for (uint64_t i = 0; i < count; i++) {
    const uint64_t o = S->seq_off[first + i], l = S->seq_len[first + i];
    off_out[i] = pos;
    if (pos + l > cap) { S->err = "pa_ingest_pack: output buffer too small"; return -3; }
    for (uint64_t k = 0; k < l; k++) {  // the FASTA fast path keeps line breaks inside a record: drop them
      const char c = S->raw[o + k];
      if (c != '\n') nt_out[pos++] = c;
    }
  }
  off_out[count] = pos;
  return (int64_t)pos;
}

// ---------------------------------------------------------------------------------------------------------------------
// pa_fasta_*: all.raw.fasta -> (ids, concatenated nucleotides, offsets) for pa_load_reads, replacing the Python line loop
// of dropin.read_raw_fasta (19 s per 10 M reads).  Same semantics as the reference's read_fastx(..., 'fasta') dictionary
// (lib/library.py:186-198): the id is the header up to the first blank, multi-line records are joined, a repeated id
// keeps its FIRST position but its LAST sequence.
// ---------------------------------------------------------------------------------------------------------------------
struct pa_fasta {
  std::string ids, nt;
  std::vector<uint64_t> id_off, off;
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
  struct Rec { uint64_t id_off, seq_off, seq_len; uint32_t id_len; int32_t multi; };   // multi: index into `joined`, -1 = one line
  std::vector<std::string> joined;
  std::vector<Rec> recs;
  recs.reserve(t.size() / 160 + 16);
  // open-addressing id table (FNV-1a): slot = record index + 1
  size_t tcap = 1024;
  while (tcap < 2 * (t.size() / 96 + 16)) tcap <<= 1;
  std::vector<uint32_t> slots(tcap, 0u);
  auto find_or_add = [&](std::string_view id, size_t next_index, bool &found) -> size_t {
    if (2 * (next_index + 1) > tcap) {   // grow: re-insert every record
      tcap <<= 1;
      std::vector<uint32_t> bigger(tcap, 0u);
      for (size_t k = 0; k < recs.size(); k++) {
        uint64_t h = 1469598103934665603ull;
        for (uint32_t z = 0; z < recs[k].id_len; z++) { h ^= (unsigned char)t[recs[k].id_off + z]; h *= 1099511628211ull; }
        size_t p = (size_t)(h ^ (h >> 32)) & (tcap - 1);
        while (bigger[p]) p = (p + 1) & (tcap - 1);
        bigger[p] = (uint32_t)k + 1;
      }
      slots.swap(bigger);
    }
    uint64_t h = 1469598103934665603ull;
    for (unsigned char c : id) { h ^= c; h *= 1099511628211ull; }
    size_t p = (size_t)(h ^ (h >> 32)) & (tcap - 1);
    while (slots[p]) {
      const Rec &R = recs[slots[p] - 1];
      if (R.id_len == id.size() && memcmp(t.data() + R.id_off, id.data(), id.size()) == 0) { found = true; return slots[p] - 1; }
      p = (p + 1) & (tcap - 1);
    }
    slots[p] = (uint32_t)next_index + 1;
    found = false;
    return next_index;
  };
  size_t i = 0;
  const size_t n = t.size();
  long cur = -1;
  int nparts = 0;
  while (i < n) {
    const void *nl = memchr(t.data() + i, '\n', n - i);
    const size_t e = nl ? (size_t)((const char *)nl - t.data()) : n;
    size_t r = e;
    while (r > i && (t[r - 1] == '\r' || t[r - 1] == '\n')) r--;   // rstrip(b"\r\n") as in dropin.read_raw_fasta
    if (r > i && t[i] == '>') {
      size_t b = i + 1;
      while (b < r && t[b] != ' ') b++;
      const std::string_view id(t.data() + i + 1, b - i - 1);
      bool found = false;
      const size_t at = find_or_add(id, recs.size(), found);
      if (!found) {
        recs.push_back(Rec{(uint64_t)(i + 1), 0, 0, (uint32_t)(b - i - 1), -1});
        cur = (long)recs.size() - 1;
      } else {
        cur = (long)at;                    // the later record replaces the sequence, the position stays
        recs[(size_t)cur].seq_len = 0;
        recs[(size_t)cur].multi = -1;
      }
      nparts = 0;
    } else if (cur >= 0) {
      Rec &R = recs[(size_t)cur];
      if (nparts == 0) { R.seq_off = i; R.seq_len = r - i; }
      else if (r > i) {
        if (R.multi < 0) { R.multi = (int32_t)joined.size(); joined.emplace_back(t.data() + R.seq_off, R.seq_len); }
        joined[(size_t)R.multi].append(t.data() + i, r - i);
      }
      nparts++;
    }
    i = e + 1;
  }
  const double tt1 = now();
  pa_fasta *F = new pa_fasta();
  size_t idb = 0, ntb = 0;
  for (auto &R : recs) { idb += R.id_len; ntb += R.multi >= 0 ? joined[(size_t)R.multi].size() : R.seq_len; }
  F->ids.reserve(idb);
  F->nt.reserve(ntb + 16);
  F->id_off.reserve(recs.size() + 1);
  F->off.reserve(recs.size() + 1);
  F->id_off.push_back(0);
  F->off.push_back(0);
  for (auto &R : recs) {
    F->ids.append(t.data() + R.id_off, R.id_len);
    if (R.multi >= 0) F->nt.append(joined[(size_t)R.multi]);
    else F->nt.append(t.data() + R.seq_off, R.seq_len);
    F->id_off.push_back(F->ids.size());
    F->off.push_back(F->nt.size());
  }
  if (dbg) fprintf(stderr, "pa_fasta_open: parse %.3f s, pack %.3f s\n", tt1 - tt0, now() - tt1);
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

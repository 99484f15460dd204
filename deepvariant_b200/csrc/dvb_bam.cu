// dvb_bam.cu — host-only BAM -> Structure-of-Arrays read table (SURVEY.md §8(f) "next" row #1).
//
// Replaces, for the pileup path, what the reference does per read through htslib + protobuf:
//   SamReader::Iterate / Query          third_party/nucleus/io/sam_reader.cc:1065-1135
//   ConvertToPb (bam1_t -> Read proto)  third_party/nucleus/io/sam_reader.cc:760-975
//   ReadSatisfiesRequirements           third_party/nucleus/io/sam_reader.cc:217-245,
//                                       third_party/nucleus/util/utils.cc:242-266 (IsReadProperlyPlaced)
// Here a BAM file is inflated block-parallel (BGZF members are independent gzip streams; htslib is not in this
// image, zlib is) and every record that passes the ReadRequirements filter is decoded ONCE straight into the
// flat arrays the DvbBatch packer consumes (positions, flags, CIGAR words, ASCII bases, qualities, names) —
// no per-read heap objects, no protos.  Field conversions follow ConvertToPb:
//   aligned_sequence  = "=ACMGRSVTWYHKDBN"[4-bit code]           (sam_reader.cc:815-825)
//   read_number       = 0 if (FREAD1 or unpaired) else 1         (sam_reader.cc:786-793)
//   number_reads      = 2 if paired else 1
//   fragment_length   = isize
//   alignment end     = pos + sum of M/D/N/=/X lengths            (utils.cc:222-240 ReadEnd)
//   HP aux tag        -> info["HP"] integer (parse_sam_aux_fields)
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <climits>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>

#include "dvb_common.h"

namespace {

constexpr uint16_t FPAIRED = 0x1, FPROPER = 0x2, FUNMAP = 0x4, FMUNMAP = 0x8, FREAD1 = 0x40, FSECONDARY = 0x100, FQCFAIL = 0x200,
                   FDUP = 0x400, FSUPP = 0x800;

inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline int32_t rdi32(const uint8_t* p) { int32_t v; memcpy(&v, p, 4); return v; }
inline uint16_t rd16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }

struct BgzfBlock { size_t in_off, in_len, out_off, out_len; };

// HP aux tag (integer types only), INT32_MIN when absent.
int32_t ParseHp(const uint8_t* aux, size_t n) {
  size_t i = 0;
  while (i + 3 <= n) {
    const uint8_t t0 = aux[i], t1 = aux[i + 1], typ = aux[i + 2];
    i += 3;
    size_t sz = 0;
    switch (typ) {
      case 'A': case 'c': case 'C': sz = 1; break;
      case 's': case 'S': sz = 2; break;
      case 'i': case 'I': case 'f': sz = 4; break;
      case 'Z': case 'H': {
        while (i < n && aux[i]) ++i;
        ++i;
        continue;
      }
      case 'B': {
        if (i + 5 > n) return INT32_MIN;
        const uint8_t sub = aux[i];
        const uint32_t cnt = rd32(aux + i + 1);
        const size_t ssz = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
        i += 5 + (size_t)cnt * ssz;
        continue;
      }
      default: return INT32_MIN;
    }
    if (i + sz > n) return INT32_MIN;
    if (t0 == 'H' && t1 == 'P' && typ != 'f' && typ != 'A') {
      switch (typ) {
        case 'c': return (int8_t)aux[i];
        case 'C': return aux[i];
        case 's': return (int16_t)rd16(aux + i);
        case 'S': return rd16(aux + i);
        case 'i': return rdi32(aux + i);
        case 'I': return (int32_t)rd32(aux + i);
      }
    }
    i += sz;
  }
  return INT32_MIN;
}

}  // namespace

struct DvbBam {
  std::vector<std::string> refs;
  std::vector<int32_t> ref_len;
  std::vector<int32_t> ref_id, pos, end, fragment_length, hp;
  std::vector<uint8_t> mapq, read_number, number_reads;
  std::vector<uint16_t> flag;
  std::vector<int64_t> seq_begin, cigar_begin, name_begin;
  std::vector<uint8_t> bases, quals;
  std::vector<uint32_t> cigar;
  std::vector<char> names;
  std::vector<int64_t> aux_begin;         // parse_hp & 2: the raw aux bytes of every kept record (MM / ML / MN, tp, t0 ... are parsed by the caller)
  std::vector<uint8_t> aux;
  int64_t n_records_seen = 0;
  bool seeked = false;                    // opened through the .bai linear index (dvb_bam_open_regions): only part of the file was read
  // region-packer index, built on first use (EnsureIndex)
  mutable std::once_flag index_once;
  mutable bool sorted = false;            // rows ordered by (ref_id, pos): region queries can binary-search
  mutable int32_t max_span = 0;           // longest end - pos
  mutable std::unordered_map<uint64_t, int32_t> first_by_key;   // hash(fragment_name, read_number) -> first row with that hash
  mutable std::vector<int32_t> next_same_hash;                   // chain to the next row with the same hash (file order), -1 ends
};

namespace {

struct Region { int32_t ref_id; int64_t start, end; };

// BGZF members of `buf` (complete ones only): appends to `blocks` with out_off relative to `out_base`; returns the number of
// compressed bytes consumed, or (size_t)-1 on a malformed member.  A member is: gzip header with the 'BC' extra subfield
// (BSIZE = member size - 1), raw deflate data, CRC32, ISIZE.
size_t IndexBgzf(const uint8_t* buf, size_t n, size_t in_base, size_t* out_total, std::vector<BgzfBlock>* blocks) {
  size_t off = 0;
  while (off + 18 <= n) {
    const uint8_t* h = buf + off;
    if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) return (size_t)-1;
    const size_t xlen = rd16(h + 10);
    if (off + 12 + xlen > n) break;                       // extra field not complete in this buffer
    size_t x = 12, bsize = 0;
    while (x + 4 <= 12 + xlen) {
      const size_t slen = rd16(h + x + 2);
      if (h[x] == 'B' && h[x + 1] == 'C' && slen == 2 && x + 6 <= 12 + xlen) bsize = (size_t)rd16(h + x + 4) + 1;
      x += 4 + slen;
    }
    if (!bsize || bsize < 12 + xlen + 8) return (size_t)-1;   // no BC subfield / a size that cannot hold header + trailer (would underflow below)
    if (off + bsize > n) break;                           // member not complete in this buffer
    const size_t isize = rd32(h + bsize - 4);
    blocks->push_back({in_base + off + 12 + xlen, bsize - 12 - xlen - 8, *out_total, isize});
    *out_total += isize;
    off += bsize;
  }
  return off;
}

// Inflates blocks[b0 ..) from `comp` (in_off relative to comp_base) into out (out_off relative to out_base), in parallel.
bool InflateBlocks(const std::vector<BgzfBlock>& blocks, size_t b0, const uint8_t* comp, size_t comp_base, uint8_t* out, size_t out_base, int threads) {
  int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
  nt = std::max(1, std::min(nt, (int)std::max<size_t>(1, (blocks.size() - b0) / 4)));
  std::atomic<size_t> next{b0};
  std::atomic<int> bad{0};
  auto work = [&]() {
    for (;;) {
      const size_t b = next.fetch_add(1);
      if (b >= blocks.size()) return;
      const BgzfBlock& k = blocks[b];
      if (!k.out_len) continue;
      z_stream zs;
      memset(&zs, 0, sizeof(zs));
      if (inflateInit2(&zs, -15) != Z_OK) { bad = 1; return; }
      zs.next_in = const_cast<uint8_t*>(comp + (k.in_off - comp_base)); zs.avail_in = (uInt)k.in_len;
      zs.next_out = out + (k.out_off - out_base); zs.avail_out = (uInt)k.out_len;
      const int rc = inflate(&zs, Z_FINISH);
      inflateEnd(&zs);
      if (rc != Z_STREAM_END || zs.avail_out != 0) { bad = 1; return; }
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < nt; ++t) pool.emplace_back(work);
  work();
  for (auto& t : pool) t.join();
  return !bad;
}

// Smallest virtual file offset at which a read overlapping [start, ...) of reference `ref_id` can begin, from the .bai linear index
// (SAM spec 5.2: ioffset[w] = offset of the first alignment overlapping the 16-kb window w); 0 = not known (scan from the beginning).
uint64_t BaiLinearOffset(const std::string& bai_path, int32_t ref_id, int64_t start) {
  FILE* f = fopen(bai_path.c_str(), "rb");
  if (!f) return 0;
  std::vector<uint8_t> b;
  fseek(f, 0, SEEK_END);
  const long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  b.resize((size_t)std::max(0L, n));
  const bool ok = n > 8 && fread(b.data(), 1, b.size(), f) == b.size();
  fclose(f);
  if (!ok || memcmp(b.data(), "BAI\1", 4) != 0) return 0;
  size_t p = 8;
  const int32_t n_ref = rdi32(b.data() + 4);
  for (int32_t r = 0; r < n_ref; ++r) {
    if (p + 4 > b.size()) return 0;
    const int32_t n_bin = rdi32(b.data() + p); p += 4;
    for (int32_t k = 0; k < n_bin; ++k) {
      if (p + 8 > b.size()) return 0;
      const int32_t n_chunk = rdi32(b.data() + p + 4);
      p += 8 + 16 * (size_t)std::max(0, n_chunk);
    }
    if (p + 4 > b.size()) return 0;
    const int32_t n_intv = rdi32(b.data() + p); p += 4;
    if (p + 8 * (size_t)std::max(0, n_intv) > b.size()) return 0;
    if (r == ref_id) {
      int64_t w = start >> 14;
      if (n_intv <= 0) return 0;
      if (w >= n_intv) w = n_intv - 1;
      for (; w >= 0; --w) {                                   // an empty window has offset 0: an earlier one is a valid (earlier) start
        uint64_t v;
        memcpy(&v, b.data() + p + 8 * (size_t)w, 8);
        if (v) return v;
      }
      return 0;
    }
    p += 8 * (size_t)std::max(0, n_intv);
  }
  return 0;
}

int OpenImpl(const char* path, const DvbReadRequirements* req_in, int parse_hp, int threads, const char* const* contigs, const int64_t* starts,
             const int64_t* ends, int32_t n_regions, DvbBam** out) {
  if (!path || !out || (n_regions > 0 && (!contigs || !starts || !ends)) || n_regions < 0) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_bam_open: bad arguments");
  *out = nullptr;
  DvbReadRequirements req;
  if (req_in) req = *req_in; else dvb_read_requirements_default(&req);
  FILE* f = fopen(path, "rb");
  if (!f) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "cannot open %s", path);
  struct Closer { FILE* f; ~Closer() { if (f) fclose(f); } } closer{f};
  fseek(f, 0, SEEK_END);
  const size_t fsz = (size_t)std::max(0L, ftell(f));
  fseek(f, 0, SEEK_SET);
  constexpr size_t kChunk = 32u << 20;      // compressed bytes read, indexed and inflated per round

  std::unique_ptr<DvbBam> bam(new DvbBam());
  std::vector<Region> regions;
  std::vector<uint8_t> comp, data;          // compressed chunk; inflated bytes not yet consumed
  size_t comp_off = 0;                      // file offset of the next byte to read
  size_t skip_first = 0;                    // uncompressed bytes of the first inflated block to drop (after a seek by the index)
  bool header_done = false, stop = false;
  int32_t span_ref = -1;
  int64_t span_lo = 0, span_hi = 0;
  static const char kSeq[] = "=ACMGRSVTWYHKDBN";
  bam->seq_begin.push_back(0); bam->cigar_begin.push_back(0); bam->name_begin.push_back(0); bam->aux_begin.push_back(0);

  // one alignment record (after its 4-byte block_size); returns false on a malformed record
  auto take_record = [&](const uint8_t* r, int32_t block_size) -> bool {
    bam->n_records_seen++;
    const int32_t ref_id = rdi32(r), pos = rdi32(r + 4);
    const uint8_t l_read_name = r[8], mapq = r[9];
    const uint16_t n_cigar = rd16(r + 12), flag = rd16(r + 14);
    const int32_t l_seq = rdi32(r + 16), next_ref = rdi32(r + 20), tlen = rdi32(r + 28);
    const size_t need = 32 + (size_t)l_read_name + 4 * (size_t)n_cigar + ((size_t)l_seq + 1) / 2 + (size_t)l_seq;
    if (l_seq < 0 || need > (size_t)block_size) return false;
    if (span_ref >= 0 && ((uint32_t)ref_id > (uint32_t)span_ref || (ref_id == span_ref && pos >= span_hi))) { stop = true; return true; }   // sorted file: nothing later can overlap
    // ReadSatisfiesRequirements (sam_reader.cc:217-245)
    if (((flag & FDUP) && !req.keep_duplicates) || ((flag & FQCFAIL) && !req.keep_failed_vendor_quality_checks) ||
        ((flag & FSECONDARY) && !req.keep_secondary_alignments) || ((flag & FSUPP) && !req.keep_supplementary_alignments))
      return true;
    const bool mapped = !(flag & FUNMAP);
    if (!mapped && !req.keep_unaligned) return true;
    const bool paired = flag & FPAIRED;
    if (!req.keep_improperly_placed && mapped) {   // IsReadProperlyPlaced (utils.cc:261-266)
      const bool mate_has_contig = paired && !(flag & FMUNMAP) && next_ref >= 0;
      if (!(!paired || (flag & FPROPER) || !mate_has_contig || next_ref == ref_id)) return true;
    }
    if (mapped && (int)mapq < req.min_mapping_quality) return true;
    const uint8_t* name = r + 32;
    const uint8_t* cig = name + l_read_name;
    const uint8_t* seq = cig + 4 * (size_t)n_cigar;
    const uint8_t* qual = seq + ((size_t)l_seq + 1) / 2;
    const uint8_t* aux = qual + l_seq;
    int32_t e = pos;
    if (mapped)
      for (uint16_t k = 0; k < n_cigar; ++k) {
        const uint32_t c = rd32(cig + 4 * (size_t)k);
        const uint32_t op = c & 0xF;
        if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) e += (int32_t)(c >> 4);
      }
    if (!regions.empty()) {                      // --regions: only reads that overlap one of them are kept (ReadOverlapsRegion, utils.cc:172-188)
      // regions are sorted and merged: the first one that ends behind the read's start decides (binary search; records of a
      // coordinate-sorted file move the answer forward monotonically, but nothing here depends on the order)
      const int32_t rend = std::max(e, pos + 1);
      size_t lo = 0, hi = regions.size();
      while (lo < hi) {
        const size_t mid = (lo + hi) >> 1;
        const Region& g = regions[mid];
        if (g.ref_id < ref_id || (g.ref_id == ref_id && g.end <= pos)) lo = mid + 1; else hi = mid;
      }
      if (lo == regions.size() || regions[lo].ref_id != ref_id || regions[lo].start >= rend) return true;
    }
    // ---- decode
    bam->ref_id.push_back(ref_id); bam->pos.push_back(pos); bam->mapq.push_back(mapq); bam->flag.push_back(flag);
    bam->fragment_length.push_back(tlen);
    bam->read_number.push_back(((flag & FREAD1) || !paired) ? 0 : 1);
    bam->number_reads.push_back(paired ? 2 : 1);
    if (mapped)
      for (uint16_t k = 0; k < n_cigar; ++k) bam->cigar.push_back(rd32(cig + 4 * (size_t)k));
    bam->end.push_back(e);
    bam->cigar_begin.push_back((int64_t)bam->cigar.size());
    const size_t b0 = bam->bases.size();
    bam->bases.resize(b0 + (size_t)l_seq);
    for (int32_t i = 0; i < l_seq; ++i) bam->bases[b0 + i] = (uint8_t)kSeq[(seq[i >> 1] >> ((~i & 1) << 2)) & 0xF];
    bam->quals.insert(bam->quals.end(), qual, qual + l_seq);
    bam->seq_begin.push_back((int64_t)bam->bases.size());
    bam->names.insert(bam->names.end(), name, name + (l_read_name ? l_read_name - 1 : 0));
    bam->name_begin.push_back((int64_t)bam->names.size());
    bam->hp.push_back((parse_hp & 1) ? ParseHp(aux, (size_t)block_size - need) : INT32_MIN);
    if (parse_hp & 2) bam->aux.insert(bam->aux.end(), aux, aux + ((size_t)block_size - need));
    bam->aux_begin.push_back((int64_t)bam->aux.size());
    return true;
  };

  // header + reference list from the front of `data`; returns bytes consumed, 0 = need more data, (size_t)-1 = malformed
  auto take_header = [&]() -> size_t {
    if (data.size() < 12) return 0;
    if (memcmp(data.data(), "BAM\1", 4) != 0) return (size_t)-1;
    size_t p = 8 + (size_t)rdi32(data.data() + 4);
    if (p + 4 > data.size()) return 0;
    const int32_t n_ref = rdi32(data.data() + p);
    p += 4;
    std::vector<std::string> refs;
    std::vector<int32_t> lens;
    for (int32_t i = 0; i < n_ref; ++i) {
      if (p + 4 > data.size()) return 0;
      const int32_t l_name = rdi32(data.data() + p);
      if (l_name < 1) return (size_t)-1;
      if (p + 8 + (size_t)l_name > data.size()) return 0;
      refs.emplace_back(reinterpret_cast<const char*>(data.data() + p + 4), (size_t)l_name - 1);
      lens.push_back(rdi32(data.data() + p + 4 + l_name));
      p += 8 + (size_t)l_name;
    }
    bam->refs = refs;
    bam->ref_len = lens;
    return p;
  };

  while (!stop) {
    // ---- read the next chunk of compressed bytes, index its complete BGZF members, inflate them behind the unconsumed bytes
    if (comp_off >= fsz && comp.empty()) break;
    const size_t have = comp.size();
    // (the first rounds are small: the header is at the front and an index seek should not be preceded by 32 MB of inflating)
    const size_t want = std::min(header_done ? kChunk : (size_t)(256u << 10), fsz - std::min(fsz, comp_off));
    comp.resize(have + want);
    if (want) {
      if (fseek(f, (long)comp_off, SEEK_SET) != 0 || fread(comp.data() + have, 1, want, f) != want) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "short read on %s", path);
      comp_off += want;
    }
    std::vector<BgzfBlock> blocks;
    size_t out_total = data.size();
    const size_t used = IndexBgzf(comp.data(), comp.size(), 0, &out_total, &blocks);
    if (used == (size_t)-1) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "%s: not a BGZF block (or a malformed one) near offset %zu", path, comp_off - comp.size());
    if (blocks.empty()) {
      if (!want) {
        if (!comp.empty()) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "%s: truncated BGZF block at the end of the file", path);
        break;
      }
      continue;                                  // a member larger than what is buffered so far: read on
    }
    const size_t data0 = data.size();
    data.resize(out_total);
    if (!InflateBlocks(blocks, 0, comp.data(), 0, data.data(), 0, threads)) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "%s: BGZF inflate failed", path);
    comp.erase(comp.begin(), comp.begin() + (long)used);
    if (skip_first) {                            // first block after a seek: drop the bytes before the indexed record
      const size_t drop = std::min(skip_first, data.size() - data0);
      data.erase(data.begin() + (long)data0, data.begin() + (long)(data0 + drop));
      skip_first = 0;
    }
    // ---- consume
    size_t p = 0;
    if (!header_done) {
      const size_t h = take_header();
      if (h == (size_t)-1) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "%s is not a BAM file", path);
      if (h == 0) { if (!want && comp.empty()) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "%s: truncated header", path); continue; }
      p = h;
      header_done = true;
      // regions -> reference ids; when they all lie on one contig and a .bai sits beside the file, jump to the first block that can
      // hold an overlapping read and stop at the first read behind the last region (coordinate-sorted files only - the index implies it)
      for (int32_t i = 0; i < n_regions; ++i) {
        int32_t id = -1;
        for (size_t k = 0; k < bam->refs.size(); ++k)
          if (bam->refs[k] == contigs[i]) { id = (int32_t)k; break; }
        if (id < 0) continue;                    // a contig the file does not know: no reads
        regions.push_back({id, std::max<int64_t>(0, starts[i]), ends[i]});
      }
      if (n_regions > 0 && regions.empty()) { stop = true; break; }
      if (!regions.empty()) {                    // sort by (contig, start) and merge what overlaps or touches
        std::sort(regions.begin(), regions.end(), [](const Region& a, const Region& b) { return a.ref_id != b.ref_id ? a.ref_id < b.ref_id : a.start < b.start; });
        std::vector<Region> merged;
        for (const Region& g : regions) {
          if (g.end <= g.start) continue;
          if (!merged.empty() && merged.back().ref_id == g.ref_id && g.start <= merged.back().end) merged.back().end = std::max(merged.back().end, g.end);
          else merged.push_back(g);
        }
        regions.swap(merged);
        if (regions.empty()) { stop = true; break; }
      }
      if (!regions.empty()) {
        bool one = true;
        span_lo = regions[0].start; span_hi = regions[0].end;
        for (const Region& g : regions) { one = one && g.ref_id == regions[0].ref_id; span_lo = std::min(span_lo, g.start); span_hi = std::max(span_hi, g.end); }
        std::string bai = std::string(path) + ".bai";
        FILE* t = fopen(bai.c_str(), "rb");
        if (!t) {
          const std::string ps(path);
          const size_t dot = ps.rfind('.');
          if (dot != std::string::npos) { bai = ps.substr(0, dot) + ".bai"; t = fopen(bai.c_str(), "rb"); }
        }
        if (t) fclose(t);
        if (one && t) {
          span_ref = regions[0].ref_id;
          const uint64_t v = BaiLinearOffset(bai, span_ref, span_lo);
          const size_t coff = (size_t)(v >> 16);
          if (v && coff < fsz && coff >= comp_off - comp.size()) {   // seek only forward of what has been read and inflated
            comp.clear();
            data.clear();
            comp_off = coff;
            skip_first = (size_t)(v & 0xFFFF);
            bam->seeked = true;
            continue;
          }
        }
      }
    }
    while (p + 4 <= data.size() && !stop) {
      const int32_t block_size = rdi32(data.data() + p);
      if (block_size < 32) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "%s: truncated alignment record", path);
      if (p + 4 + (size_t)block_size > data.size()) break;      // the record continues in the next chunk
      if (!take_record(data.data() + p + 4, block_size)) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "%s: malformed alignment record", path);
      p += 4 + (size_t)block_size;
    }
    data.erase(data.begin(), data.begin() + (long)p);
    if (!want && comp.empty()) {
      if (!data.empty() && !stop) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "%s: truncated alignment record", path);
      break;
    }
  }
  if (!header_done) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "%s is not a BAM file", path);
  *out = bam.release();
  return DVB_OK;
}

}  // namespace

extern "C" {

void dvb_read_requirements_default(DvbReadRequirements* r) {
  memset(r, 0, sizeof(*r));
  r->min_mapping_quality = 5;   // make_examples_options.py:957-964
}

int dvb_bam_open(const char* path, const DvbReadRequirements* req_in, int parse_hp, int threads, DvbBam** out) {
  return OpenImpl(path, req_in, parse_hp, threads, nullptr, nullptr, nullptr, 0, out);
}

int dvb_bam_open_regions(const char* path, const DvbReadRequirements* req_in, int parse_hp, int threads, const char* const* contigs,
                         const int64_t* starts, const int64_t* ends, int32_t n_regions, DvbBam** out) {
  return OpenImpl(path, req_in, parse_hp, threads, contigs, starts, ends, n_regions, out);
}

int dvb_bam_table(const DvbBam* bam, DvbReadTable* t) {
  if (!bam || !t) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_bam_table: null argument");
  memset(t, 0, sizeof(*t));
  t->n_reads = (int32_t)bam->pos.size();
  t->n_refs = (int32_t)bam->refs.size();
  t->n_bases = (int64_t)bam->bases.size();
  t->n_cigar = (int64_t)bam->cigar.size();
  t->n_name_bytes = (int64_t)bam->names.size();
  t->n_records_seen = bam->n_records_seen;
  t->ref_id = bam->ref_id.data(); t->pos = bam->pos.data(); t->end = bam->end.data(); t->mapq = bam->mapq.data();
  t->flag = bam->flag.data(); t->fragment_length = bam->fragment_length.data(); t->hp = bam->hp.data();
  t->read_number = bam->read_number.data(); t->number_reads = bam->number_reads.data();
  t->seq_begin = bam->seq_begin.data(); t->cigar_begin = bam->cigar_begin.data(); t->name_begin = bam->name_begin.data();
  t->bases = bam->bases.data(); t->quals = bam->quals.data(); t->cigar = bam->cigar.data(); t->names = bam->names.data();
  t->n_aux_bytes = (int64_t)bam->aux.size(); t->aux_begin = bam->aux_begin.data(); t->aux = bam->aux.data();
  return DVB_OK;
}

const char* dvb_bam_ref_name(const DvbBam* bam, int32_t i) {
  if (!bam || i < 0 || i >= (int32_t)bam->refs.size()) return nullptr;
  return bam->refs[(size_t)i].c_str();
}

int32_t dvb_bam_ref_length(const DvbBam* bam, int32_t i) {
  if (!bam || i < 0 || i >= (int32_t)bam->ref_len.size()) return -1;
  return bam->ref_len[(size_t)i];
}

void dvb_bam_close(DvbBam* bam) { delete bam; }

// A table of `n_rows` reads of `src` in the given order, row i with the alignment (new_pos[i], new_cigar[new_cigar_begin[i] ..
// new_cigar_begin[i + 1])) when that range is not empty and with its own otherwise; everything else of the record (name, flag,
// bases, qualities, mate fields, HP, raw aux bytes) is copied.  This is how realigned / normalised reads reach the candidate
// generator and the region packer: what in_memory_sam_reader.replace_reads does in the reference (make_examples_core.py:2290-2300),
// without a BAM file in between.
int dvb_bam_derive(const DvbBam* src, const int64_t* rows, int64_t n_rows, const int32_t* new_pos, const int64_t* new_cigar_begin,
                   const uint32_t* new_cigar, DvbBam** out) {
  if (!src || !out || n_rows < 0 || (n_rows > 0 && !rows)) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_bam_derive: null argument");
  *out = nullptr;
  const bool replace = new_cigar_begin != nullptr;
  if (replace && (!new_pos || (new_cigar_begin[n_rows] > new_cigar_begin[0] && !new_cigar)))
    return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_bam_derive: new_cigar_begin without new_pos / new_cigar");
  const int64_t n_src = (int64_t)src->pos.size();
  std::unique_ptr<DvbBam> d(new DvbBam());
  d->refs = src->refs;
  d->ref_len = src->ref_len;
  d->n_records_seen = n_rows;
  const bool has_aux = (int64_t)src->aux_begin.size() == n_src + 1 && n_src > 0;      // the source kept its records' raw aux bytes (parse_hp & 2)
  d->seq_begin.push_back(0); d->cigar_begin.push_back(0); d->name_begin.push_back(0); d->aux_begin.push_back(0);
  for (int64_t i = 0; i < n_rows; ++i) {
    const int64_t r = rows[i];
    if (r < 0 || r >= n_src) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_bam_derive: row %lld out of range (%lld reads)", (long long)r, (long long)n_src);
    d->ref_id.push_back(src->ref_id[r]); d->mapq.push_back(src->mapq[r]); d->flag.push_back(src->flag[r]);
    d->fragment_length.push_back(src->fragment_length[r]); d->hp.push_back(src->hp[r]);
    d->read_number.push_back(src->read_number[r]); d->number_reads.push_back(src->number_reads[r]);
    const bool swap = replace && new_cigar_begin[i + 1] > new_cigar_begin[i];
    if (replace && new_cigar_begin[i + 1] < new_cigar_begin[i]) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_bam_derive: new_cigar_begin is not ascending at %lld", (long long)i);
    if (swap) {
      int64_t query = 0, e = new_pos[i];
      for (int64_t k = new_cigar_begin[i]; k < new_cigar_begin[i + 1]; ++k) {
        const uint32_t c = new_cigar[k], op = c & 0xF;
        if (op > 8) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_bam_derive: CIGAR operation %u of read %lld", op, (long long)i);
        if (op == 0 || op == 1 || op == 4 || op == 7 || op == 8) query += (int64_t)(c >> 4);
        if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) e += (int64_t)(c >> 4);
        d->cigar.push_back(c);
      }
      if (query != src->seq_begin[r + 1] - src->seq_begin[r] || new_pos[i] < 0)
        return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_bam_derive: the new CIGAR of read %lld consumes %lld bases, the read has %lld", (long long)i,
                         (long long)query, (long long)(src->seq_begin[r + 1] - src->seq_begin[r]));
      d->pos.push_back(new_pos[i]);
      d->end.push_back((int32_t)e);
    } else {
      d->cigar.insert(d->cigar.end(), src->cigar.begin() + src->cigar_begin[r], src->cigar.begin() + src->cigar_begin[r + 1]);
      d->pos.push_back(src->pos[r]);
      d->end.push_back(src->end[r]);
    }
    d->cigar_begin.push_back((int64_t)d->cigar.size());
    d->bases.insert(d->bases.end(), src->bases.begin() + src->seq_begin[r], src->bases.begin() + src->seq_begin[r + 1]);
    d->quals.insert(d->quals.end(), src->quals.begin() + src->seq_begin[r], src->quals.begin() + src->seq_begin[r + 1]);
    d->seq_begin.push_back((int64_t)d->bases.size());
    d->names.insert(d->names.end(), src->names.begin() + src->name_begin[r], src->names.begin() + src->name_begin[r + 1]);
    d->name_begin.push_back((int64_t)d->names.size());
    if (has_aux) {
      d->aux.insert(d->aux.end(), src->aux.begin() + src->aux_begin[r], src->aux.begin() + src->aux_begin[r + 1]);
      d->aux_begin.push_back((int64_t)d->aux.size());
    }
  }
  *out = d.release();
  return DVB_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// Region packer: candidates + BAM table rows -> DvbBatch arrays, on the host, without per-read objects
// ---------------------------------------------------------------------------------------------
// Restates, over the flat read table, what ExamplesGenerator::CreateAndWriteExamplesForCandidate does per candidate before
// the pixels (deepvariant/make_examples_native.cc:632-736): the InMemoryReader::Query scan (:802-810 with ReadOverlapsRegion,
// third_party/nucleus/util/utils.cc:172-188) as an index filter on pos / end, ReadSupportsAlt's read-name search
// (deepvariant/channels/read_supports_variant_channel.cc:75-104) as hash lookups of "fragment_name/read_number" keys, and the
// gather of the per-read arrays.  Output = exactly the arrays packing.pack_images_from_table builds (tests compare them).
namespace {

inline std::string_view NameOf(const DvbBam& bam, int32_t r) {
  return std::string_view(bam.names.data() + bam.name_begin[(size_t)r], (size_t)(bam.name_begin[(size_t)r + 1] - bam.name_begin[(size_t)r]));
}

inline uint64_t KeyHash(std::string_view name, uint8_t read_number) {
  uint64_t h = 0xcbf29ce484222325ull ^ read_number;   // FNV-1a over the name bytes
  for (const char ch : name) h = (h ^ (uint8_t)ch) * 0x100000001b3ull;
  return h;
}

void EnsureIndex(const DvbBam& bam) {
  std::call_once(bam.index_once, [&]() {
    const int32_t n = (int32_t)bam.pos.size();
    bool sorted = true;
    int32_t span = 0;
    for (int32_t r = 0; r < n; ++r) {
      span = std::max(span, bam.end[r] - bam.pos[r]);
      if (r > 0) {   // unmapped reads (ref_id -1) sort last, as in coordinate-sorted BAMs
        const uint32_t a = (uint32_t)bam.ref_id[r - 1], b = (uint32_t)bam.ref_id[r];
        if (a > b || (a == b && bam.pos[r - 1] > bam.pos[r])) sorted = false;
      }
    }
    bam.sorted = sorted;
    bam.max_span = span;
    bam.next_same_hash.assign((size_t)n, -1);
    bam.first_by_key.reserve((size_t)n * 2);
    std::unordered_map<uint64_t, int32_t> last;
    last.reserve((size_t)n * 2);
    for (int32_t r = 0; r < n; ++r) {
      const uint64_t h = KeyHash(NameOf(bam, r), bam.read_number[r]);
      auto ins = last.emplace(h, r);
      if (ins.second) bam.first_by_key.emplace(h, r);
      else { bam.next_same_hash[(size_t)ins.first->second] = r; ins.first->second = r; }
    }
  });
}

}  // namespace

struct DvbPackedRegion {
  std::vector<uint8_t> ref_bases, pair_support, pair_allele_group, read_flags, bases, quals;
  std::vector<int32_t> image_start_pos, variant_start, pair_read, read_pos, read_sort_pos, read_mapq, read_fragment_length, read_hp;
  std::vector<uint32_t> read_name_rank, cigar;
  std::vector<int64_t> pair_begin, read_seq_begin, read_cigar_begin;
  int32_t n_images = 0, n_reads = 0, ref_stride = 0;
};

extern "C" {

int dvb_pack_region_from_bam(const DvbBam* bam, const DvbRegionCandidates* c, int32_t region_ref_id, int32_t region_start,
                             int32_t region_end, int32_t read_overlap_buffer_bp, int32_t width, DvbPackedRegion** out) {
  if (!bam || !c || !out || c->n_images < 0 || width < 1 || c->ref_stride < width)
    return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_pack_region_from_bam: bad arguments");
  *out = nullptr;
  EnsureIndex(*bam);
  const int32_t NR = (int32_t)bam->pos.size();
  // reads of the region (what the reference hands WriteExamplesInRegion), in file order; rows lie in [row_lo, row_hi)
  std::vector<int32_t> region_rows;
  int32_t row_lo = 0, row_hi = NR;
  if (bam->sorted) {
    auto before = [&](int32_t r, int32_t rid, int64_t p) { return (uint32_t)bam->ref_id[r] < (uint32_t)rid || (bam->ref_id[r] == rid && bam->pos[r] < p); };
    int32_t a = 0, b = NR;
    const int64_t p_lo = (int64_t)region_start - bam->max_span;
    while (a < b) { const int32_t m = a + (b - a) / 2; if (before(m, region_ref_id, p_lo)) a = m + 1; else b = m; }
    row_lo = a;
    b = NR;
    while (a < b) { const int32_t m = a + (b - a) / 2; if (before(m, region_ref_id, region_end)) a = m + 1; else b = m; }
    row_hi = a;
  }
  for (int32_t r = row_lo; r < row_hi; ++r)
    if (bam->ref_id[r] == region_ref_id && bam->pos[r] < region_end && bam->end[r] > region_start) region_rows.push_back(r);

  DvbPackedRegion* pr = new DvbPackedRegion();
  pr->n_images = c->n_images;
  pr->ref_stride = (width + 15) / 16 * 16;
  pr->ref_bases.assign((size_t)c->n_images * pr->ref_stride, 0);
  pr->pair_begin.push_back(0);
  std::vector<int32_t> table_rows_of_pairs;             // table row per pair, remapped to batch read indices below
  std::vector<int32_t> local((size_t)(row_hi - row_lo), -1);   // table row - row_lo -> position in the image's read list
  for (int32_t i = 0; i < c->n_images; ++i) {
    memcpy(pr->ref_bases.data() + (size_t)i * pr->ref_stride, c->ref_bases + (size_t)i * c->ref_stride, (size_t)width);
    pr->image_start_pos.push_back(c->image_start_pos[i]);
    pr->variant_start.push_back(c->variant_start[i]);
    const int32_t q_start = c->variant_start[i] - read_overlap_buffer_bp, q_end = c->variant_end[i] + read_overlap_buffer_bp;
    const size_t first_pair = table_rows_of_pairs.size();
    std::vector<int32_t> members;
    if (c->ref_id[i] == region_ref_id) {
      for (const int32_t r : region_rows) {
        if (bam->pos[r] < q_end && bam->end[r] > q_start) {
          local[(size_t)(r - row_lo)] = (int32_t)members.size();
          members.push_back(r);
          table_rows_of_pairs.push_back(r);
        }
      }
    }
    const size_t n = members.size();
    pr->pair_support.resize(first_pair + n, 0);
    pr->pair_allele_group.resize(first_pair + n, c->group_default ? (uint8_t)c->group_default[i] : (uint8_t)0);
    // ReadSupportsAlt: entries come in alt order; the first entry that names a read decides its class
    for (int64_t e = c->support_begin[i]; e < c->support_begin[i + 1]; ++e) {
      const std::string_view key(c->support_names + c->support_name_begin[e], (size_t)(c->support_name_begin[e + 1] - c->support_name_begin[e]));
      // table keys are fragment_name + "/" + ("0" | "1"): anything else cannot equal one
      if (key.size() < 2 || key[key.size() - 2] != '/' || (key.back() != '0' && key.back() != '1')) continue;
      const std::string_view name = key.substr(0, key.size() - 2);
      const uint8_t rn = (uint8_t)(key.back() - '0');
      auto it = bam->first_by_key.find(KeyHash(name, rn));
      if (it == bam->first_by_key.end()) continue;
      for (int32_t r = it->second; r >= 0; r = bam->next_same_hash[(size_t)r]) {
        if (r < row_lo || r >= row_hi) continue;
        const int32_t j = local[(size_t)(r - row_lo)];
        if (j < 0 || bam->read_number[r] != rn || NameOf(*bam, r) != name) continue;
        if (pr->pair_support[first_pair + j] == 0) pr->pair_support[first_pair + j] = c->support_class[e];
        if (c->support_group) pr->pair_allele_group[first_pair + j] = c->support_group[e];   // later alts overwrite (pileup_image_native.cc:345-360)
      }
    }
    for (const int32_t r : members) local[(size_t)(r - row_lo)] = -1;
    pr->pair_begin.push_back((int64_t)table_rows_of_pairs.size());
  }
  // reads are stored once, in order of first use
  std::unordered_map<int32_t, int32_t> batch_index;
  std::vector<int32_t> rows;
  pr->pair_read.reserve(table_rows_of_pairs.size());
  for (int32_t r : table_rows_of_pairs) {
    auto ins = batch_index.emplace(r, (int32_t)rows.size());
    if (ins.second) rows.push_back(r);
    pr->pair_read.push_back(ins.first->second);
  }
  pr->n_reads = (int32_t)rows.size();
  pr->read_seq_begin.push_back(0);
  pr->read_cigar_begin.push_back(0);
  for (int32_t r : rows) {
    pr->read_pos.push_back(bam->pos[r]);
    pr->read_sort_pos.push_back(bam->pos[r]);
    pr->read_mapq.push_back(bam->mapq[r]);
    const bool has_hp = bam->hp[r] != INT32_MIN;
    pr->read_flags.push_back((uint8_t)(((bam->flag[r] & 0x10) ? DVB_READ_REVERSE_STRAND : 0) | ((bam->flag[r] & FSUPP) ? DVB_READ_SUPPLEMENTARY : 0) |
                                       (has_hp ? DVB_READ_HAS_HP : 0)));
    pr->read_fragment_length.push_back(bam->fragment_length[r]);
    pr->read_hp.push_back(has_hp ? bam->hp[r] : 0);
    pr->bases.insert(pr->bases.end(), bam->bases.begin() + bam->seq_begin[r], bam->bases.begin() + bam->seq_begin[r + 1]);
    pr->quals.insert(pr->quals.end(), bam->quals.begin() + bam->seq_begin[r], bam->quals.begin() + bam->seq_begin[r + 1]);
    pr->cigar.insert(pr->cigar.end(), bam->cigar.begin() + bam->cigar_begin[r], bam->cigar.begin() + bam->cigar_begin[r + 1]);
    pr->read_seq_begin.push_back((int64_t)pr->bases.size());
    pr->read_cigar_begin.push_back((int64_t)pr->cigar.size());
  }
  // dense rank of (fragment_name bytes, read_number): std::tuple<std::string, int> order (pileup_image_native.cc:98-101)
  {
    std::vector<int32_t> order(rows.size());
    for (size_t k = 0; k < rows.size(); ++k) order[k] = (int32_t)k;
    auto name_of = [&](int32_t r) { return std::string_view(bam->names.data() + bam->name_begin[r], (size_t)(bam->name_begin[r + 1] - bam->name_begin[r])); };
    auto less = [&](int32_t a, int32_t b) {
      const std::string_view na = name_of(rows[a]), nb = name_of(rows[b]);
      if (na != nb) return na < nb;
      return bam->read_number[rows[a]] < bam->read_number[rows[b]];
    };
    std::sort(order.begin(), order.end(), less);
    pr->read_name_rank.assign(rows.size(), 0);
    uint32_t rank = 0;
    for (size_t k = 0; k < order.size(); ++k) {
      if (k > 0 && less(order[k - 1], order[k])) ++rank;
      pr->read_name_rank[order[k]] = rank;
    }
  }
  *out = pr;
  return DVB_OK;
}

int dvb_packed_region_batch(const DvbPackedRegion* pr, DvbBatch* b) {
  if (!pr || !b) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_packed_region_batch: null argument");
  memset(b, 0, sizeof(*b));
  b->n_images = pr->n_images; b->n_reads = pr->n_reads; b->n_pairs = (int64_t)pr->pair_read.size();
  b->n_bases = (int64_t)pr->bases.size(); b->n_cigar = (int64_t)pr->cigar.size(); b->ref_stride = pr->ref_stride;
  b->ref_bases = pr->ref_bases.data(); b->image_start_pos = pr->image_start_pos.data(); b->variant_start = pr->variant_start.data();
  b->pair_begin = pr->pair_begin.data(); b->pair_read = pr->pair_read.data(); b->pair_support = pr->pair_support.data();
  b->pair_allele_group = pr->pair_allele_group.data(); b->read_pos = pr->read_pos.data(); b->read_sort_pos = pr->read_sort_pos.data();
  b->read_mapq = pr->read_mapq.data(); b->read_flags = pr->read_flags.data(); b->read_fragment_length = pr->read_fragment_length.data();
  b->read_hp = pr->read_hp.data(); b->read_name_rank = pr->read_name_rank.data(); b->read_seq_begin = pr->read_seq_begin.data();
  b->read_cigar_begin = pr->read_cigar_begin.data(); b->bases = pr->bases.data(); b->quals = pr->quals.data(); b->cigar = pr->cigar.data();
  return DVB_OK;
}

void dvb_packed_region_free(DvbPackedRegion* pr) { delete pr; }

}  // extern "C"

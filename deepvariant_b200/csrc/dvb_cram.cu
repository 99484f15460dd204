// dvb_cram.cu — CRAM 3.0 input (host code): containers -> alignment records -> an uncompressed BAM the native BAM decoder takes.
//
// The reference reads CRAM through htslib (third_party/nucleus/io/sam_reader.cc:325-399 opens any of SAM / BAM / CRAM with hts_open and
// needs the FASTA for CRAM, :347-372); htslib is an un-vendored dependency (third_party/htslib.BUILD, 1.18).  What is restated here is
// the published CRAM 3.0 format (samtools/hts-specs CRAMv3) and the three decisions of htslib's decoder a reader of its output can see:
//   * cram_decode_seq: sequence, qualities and CIGAR from the read features over the reference ('X' substitutions and 'B' bases are
//     part of an M run; adjacent equal operations merge);
//   * cram_decode_slice_xref: mates inside a slice are linked through NF; both get FPAIRED, the mate's strand / unmapped bits, the
//     mate's position, and TLEN = rightmost end - leftmost start of the chain (sign: + for the leftmost record, READ1 breaking a tie);
//     detached records carry MF / NS / NP / TS verbatim;
//   * aux: the tag dictionary line TL names the tags, values are BAM-encoded; Z / H values end with NUL.
// Block compression: raw, gzip, rANS 4x8 order 0 / 1 (bzip2 / lzma are reported as unsupported).  Encodings: EXTERNAL, HUFFMAN,
// BYTE_ARRAY_LEN, BYTE_ARRAY_STOP, BETA (the ones htslib and htsjdk write for 3.0).
// Pinned by the reference's own test: make_examples over testdata/input/NA12878_S1.chr20.10_10p1mb.cram must give the goldens of the
// BAM (make_examples_test.py:330-372) - tests/test_cram.py compares the decoded table with the BAM's field by field.
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <climits>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "dvb_common.h"

namespace {

struct Rd {                          // bounds-checked cursor
  const uint8_t* p; const uint8_t* e; bool bad = false;
  uint8_t u8() { if (p >= e) { bad = true; return 0; } return *p++; }
  int32_t itf8() {
    const uint8_t v = u8();
    if (v < 0x80) return v;
    if (v < 0xC0) return ((v & 0x3f) << 8) | u8();
    if (v < 0xE0) { const uint32_t a = u8(), b = u8(); return ((v & 0x1f) << 16) | (a << 8) | b; }
    if (v < 0xF0) { const uint32_t a = u8(), b = u8(), c = u8(); return ((v & 0x0f) << 24) | (a << 16) | (b << 8) | c; }
    const uint32_t a = u8(), b = u8(), c = u8(), d = u8();
    return (int32_t)(((uint32_t)(v & 0x0f) << 28) | (a << 20) | (b << 12) | (c << 4) | (d & 0x0f));
  }
  int64_t ltf8() {
    const uint8_t v = u8();
    int n = 0;
    while (n < 8 && (v & (0x80 >> n))) ++n;
    uint64_t val = n >= 7 ? 0 : (uint64_t)(v & (0xff >> (n + 1)));
    for (int i = 0; i < n; ++i) val = (val << 8) | u8();
    return (int64_t)val;
  }
  void skip(size_t n) { if ((size_t)(e - p) < n) { bad = true; p = e; } else p += n; }
};

struct Block { int method = 0, ctype = 0, id = 0; std::vector<uint8_t> data; size_t pos = 0; int bit = 7; };

inline uint32_t Le32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }

// rANS 4x8 (CRAM 3.0 section 13): 12-bit frequencies, four interleaved 32-bit states, byte-wise renormalisation at 2^23.
bool ReadFreqs(const uint8_t*& cp, const uint8_t* end, uint16_t* freq, uint16_t* cum, uint8_t* lookup) {
  int rle = 0, x = 0;
  if (cp >= end) return false;
  int j = *cp++;
  do {
    if (cp >= end) return false;
    int f = *cp++;
    if (f >= 128) { if (cp >= end) return false; f = ((f & 127) << 8) | *cp++; }
    if (x + f > 4096) return false;
    freq[j] = (uint16_t)f; cum[j] = (uint16_t)x;
    memset(lookup + x, j, (size_t)f);
    x += f;
    if (cp >= end) return false;
    if (!rle && j + 1 == *cp) { j = *cp++; if (cp >= end) return false; rle = *cp++; }
    else if (rle) { --rle; ++j; if (j > 255) return false; }
    else j = *cp++;
  } while (j);
  return true;
}

bool RansDecode(const uint8_t* in, size_t n, size_t raw_size, std::vector<uint8_t>* out) {
  if (n < 9) return false;
  const int order = in[0];
  const size_t out_sz = Le32(in + 5);
  if (out_sz != raw_size || Le32(in + 1) + 9 != n) return false;
  out->assign(out_sz, 0);
  if (!out_sz) return true;
  const uint8_t* cp = in + 9;
  const uint8_t* end = in + n;
  uint32_t R[4];
  auto renorm = [&](uint32_t& r) { while (r < (1u << 23) && cp < end) r = (r << 8) | *cp++; };
  if (order == 0) {
    std::vector<uint16_t> freq(256, 0), cum(256, 0);
    std::vector<uint8_t> lookup(4096, 0);
    if (!ReadFreqs(cp, end, freq.data(), cum.data(), lookup.data()) || end - cp < 16) return false;
    for (int k = 0; k < 4; ++k, cp += 4) R[k] = Le32(cp);
    for (size_t i = 0; i < out_sz; ++i) {
      uint32_t& r = R[i & 3];
      const uint32_t m = r & 0xfff;
      const uint8_t s = lookup[m];
      (*out)[i] = s;
      r = freq[s] * (r >> 12) + m - cum[s];
      renorm(r);
    }
    return true;
  }
  if (order != 1) return false;
  std::vector<uint16_t> freq(256 * 256, 0), cum(256 * 256, 0);
  std::vector<uint8_t> lookup(256 * 4096, 0);
  {
    int rle = 0;
    if (cp >= end) return false;
    int i = *cp++;
    do {
      if (!ReadFreqs(cp, end, &freq[(size_t)i * 256], &cum[(size_t)i * 256], &lookup[(size_t)i * 4096])) return false;
      if (cp >= end) return false;
      if (!rle && i + 1 == *cp) { i = *cp++; if (cp >= end) return false; rle = *cp++; }
      else if (rle) { --rle; ++i; if (i > 255) return false; }
      else i = *cp++;
    } while (i);
  }
  if (end - cp < 16) return false;
  for (int k = 0; k < 4; ++k, cp += 4) R[k] = Le32(cp);
  const size_t q = out_sz >> 2;
  size_t idx[4] = {0, q, 2 * q, 3 * q};
  int last[4] = {0, 0, 0, 0};
  auto step = [&](int k) {
    const uint32_t m = R[k] & 0xfff;
    const int l = last[k];
    const uint8_t c = lookup[(size_t)l * 4096 + m];
    (*out)[idx[k]++] = c;
    R[k] = freq[(size_t)l * 256 + c] * (R[k] >> 12) + m - cum[(size_t)l * 256 + c];
    renorm(R[k]);
    last[k] = c;
  };
  for (size_t i = 0; i < q; ++i) { step(0); step(1); step(2); step(3); }
  while (idx[3] < out_sz) step(3);
  return true;
}

int ReadBlock(Rd& r, Block* b) {
  b->method = r.u8(); b->ctype = r.u8(); b->id = r.itf8();
  const int32_t csz = r.itf8(), rsz = r.itf8();
  if (r.bad || csz < 0 || rsz < 0 || (size_t)(r.e - r.p) < (size_t)csz + 4) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "CRAM: truncated block");
  const uint8_t* d = r.p;
  r.skip((size_t)csz + 4);          // data + CRC32
  b->pos = 0; b->bit = 7;
  if (b->method == 0) { b->data.assign(d, d + csz); return DVB_OK; }
  if (b->method == 1) {
    b->data.assign((size_t)rsz, 0);
    z_stream z; memset(&z, 0, sizeof z);
    if (inflateInit2(&z, 15 + 32) != Z_OK) return dvb::fail(DVB_ERR_INTERNAL, "CRAM: inflateInit2");
    z.next_in = const_cast<Bytef*>(d); z.avail_in = (uInt)csz; z.next_out = b->data.data(); z.avail_out = (uInt)rsz;
    const int rc = inflate(&z, Z_FINISH);
    const bool ok = (rc == Z_STREAM_END || (rc == Z_OK && z.avail_out == 0) || (rc == Z_BUF_ERROR && rsz == 0)) && z.total_out == (uLong)rsz;
    inflateEnd(&z);
    return ok ? DVB_OK : dvb::fail(DVB_ERR_INVALID_ARGUMENT, "CRAM: gzip block does not inflate to its raw size");
  }
  if (b->method == 4) return RansDecode(d, (size_t)csz, (size_t)rsz, &b->data) ? DVB_OK : dvb::fail(DVB_ERR_INVALID_ARGUMENT, "CRAM: malformed rANS block");
  return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "CRAM: block compression method %d (bzip2 = 2, lzma = 3) is not supported", b->method);
}

struct Enc {
  int codec = 0, ext_id = -1, stop = 0, off = 0, bits = 0;
  std::vector<int32_t> syms, lens;                 // HUFFMAN
  std::vector<std::pair<uint32_t, int>> codes;     // canonical (code, index into syms) sorted by (len, symbol)
  std::unique_ptr<Enc> len_enc, val_enc;           // BYTE_ARRAY_LEN
  bool present = false;
};

int ParseEnc(Rd& r, Enc* e) {
  e->present = true;
  e->codec = r.itf8();
  const int32_t n = r.itf8();
  if (r.bad || n < 0 || (r.e - r.p) < n) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "CRAM: truncated encoding");
  Rd p{r.p, r.p + n};
  r.skip((size_t)n);
  switch (e->codec) {
    case 0: break;
    case 1: e->ext_id = p.itf8(); break;
    case 3: {
      const int32_t ns = p.itf8();
      for (int32_t i = 0; i < ns && !p.bad; ++i) e->syms.push_back(p.itf8());
      const int32_t nl = p.itf8();
      for (int32_t i = 0; i < nl && !p.bad; ++i) e->lens.push_back(p.itf8());
      if (e->syms.size() != e->lens.size() || e->syms.empty()) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "CRAM: malformed HUFFMAN encoding");
      std::vector<int> order(e->syms.size());
      for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
      std::sort(order.begin(), order.end(), [&](int a, int b) { return e->lens[a] != e->lens[b] ? e->lens[a] < e->lens[b] : e->syms[a] < e->syms[b]; });
      uint32_t code = 0; int prev_len = e->lens[order[0]];
      for (size_t k = 0; k < order.size(); ++k) {
        const int len = e->lens[order[k]];
        if (len > 31) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "CRAM: HUFFMAN code of %d bits", len);
        code <<= (len - prev_len); prev_len = len;
        e->codes.emplace_back(code, order[k]);
        ++code;
      }
      break;
    }
    case 4: {
      e->len_enc.reset(new Enc()); e->val_enc.reset(new Enc());
      int st = ParseEnc(p, e->len_enc.get());
      if (st) return st;
      if ((st = ParseEnc(p, e->val_enc.get()))) return st;
      break;
    }
    case 5: e->stop = p.u8(); e->ext_id = p.itf8(); break;
    case 6: e->off = p.itf8(); e->bits = p.itf8(); break;
    default: return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "CRAM: encoding %d (GOLOMB / SUBEXP / GOLOMB_RICE / GAMMA) is not supported", e->codec);
  }
  return p.bad ? dvb::fail(DVB_ERR_INVALID_ARGUMENT, "CRAM: malformed encoding parameters") : DVB_OK;
}

struct Slice {
  Block* core = nullptr;
  std::map<int, Block*> ext;
  bool bad = false;
  Block* Ext(int id) { auto it = ext.find(id); if (it == ext.end()) { bad = true; return nullptr; } return it->second; }
  int Bit() {
    Block* b = core;
    if (!b || b->pos >= b->data.size()) { bad = true; return 0; }
    const int v = (b->data[b->pos] >> b->bit) & 1;
    if (--b->bit < 0) { b->bit = 7; ++b->pos; }
    return v;
  }
  int32_t Int(const Enc& e) {
    switch (e.codec) {
      case 1: { Block* b = Ext(e.ext_id); if (!b) return 0; Rd r{b->data.data() + b->pos, b->data.data() + b->data.size()}; const int32_t v = r.itf8(); bad |= r.bad; b->pos = (size_t)(r.p - b->data.data()); return v; }
      case 3: {
        if (e.syms.size() == 1 && e.lens[0] == 0) return e.syms[0];
        uint32_t code = 0; int len = 0; size_t k = 0;
        while (k < e.codes.size()) {
          const int want = e.lens[e.codes[k].second];
          while (len < want) { code = (code << 1) | (uint32_t)Bit(); ++len; }
          for (; k < e.codes.size() && e.lens[e.codes[k].second] == want; ++k)
            if (e.codes[k].first == code) return e.syms[e.codes[k].second];
          if (bad) return 0;
        }
        bad = true; return 0;
      }
      case 6: { uint32_t v = 0; for (int i = 0; i < e.bits; ++i) v = (v << 1) | (uint32_t)Bit(); return (int32_t)v - e.off; }
      default: bad = true; return 0;
    }
  }
  int Byte(const Enc& e) {
    if (e.codec == 1) { Block* b = Ext(e.ext_id); if (!b || b->pos >= b->data.size()) { bad = true; return 0; } return b->data[b->pos++]; }
    return Int(e) & 0xff;
  }
  void Bytes(const Enc& e, std::string* out) {
    if (e.codec == 5) {
      Block* b = Ext(e.ext_id); if (!b) return;
      const uint8_t* s = b->data.data() + b->pos; const uint8_t* end = b->data.data() + b->data.size();
      const uint8_t* q = (const uint8_t*)memchr(s, e.stop, (size_t)(end - s));
      if (!q) { bad = true; return; }
      out->append((const char*)s, (size_t)(q - s));
      b->pos = (size_t)(q - b->data.data()) + 1;
    } else if (e.codec == 4) {
      const int32_t n = Int(*e.len_enc);
      if (n < 0) { bad = true; return; }
      if (e.val_enc->codec == 1) {
        Block* b = Ext(e.val_enc->ext_id); if (!b || b->data.size() - b->pos < (size_t)n) { bad = true; return; }
        out->append((const char*)b->data.data() + b->pos, (size_t)n); b->pos += (size_t)n;
      } else {
        for (int32_t i = 0; i < n && !bad; ++i) out->push_back((char)Byte(*e.val_enc));
      }
    } else {
      bad = true;
    }
  }
};

struct Feature { char code; int32_t pos; int32_t n; std::string bytes; int base, qual; };

struct Rec {
  int32_t bf = 0, cf = 0, ref = -1, rl = 0, ap = 0, mq = 0, mf = 0, ns = -1, np = 0, tlen = INT_MIN, mate_line = -1;
  int64_t end = 0;              // 1-based inclusive, as htslib's aend
  std::string name, aux, seq, qual;
  std::vector<uint32_t> cigar;
};

inline void PushCigar(std::vector<uint32_t>* c, int op, int64_t n) {
  if (n <= 0) return;
  if (!c->empty() && (int)(c->back() & 0xF) == op) c->back() += (uint32_t)n << 4; else c->push_back(((uint32_t)n << 4) | (uint32_t)op);
}

struct RefSeq { const uint8_t* bases = nullptr; int64_t len = 0; int64_t origin = 0; };   // bases[i] = position origin + i (0-based)

struct Writer {                  // BAM bytes -> BGZF stored blocks
  FILE* f = nullptr; std::vector<uint8_t> buf; bool failed = false;
  void Flush(bool all) {
    size_t off = 0;
    while (buf.size() - off >= 0xff00 || (all && off < buf.size())) {
      const size_t n = std::min<size_t>(0xff00, buf.size() - off);
      uint8_t h[18 + 5] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
      const uint16_t bsize = (uint16_t)(n + 5 + 26 - 1);
      memcpy(h + 16, &bsize, 2);
      h[18] = 1;                                         // one final stored deflate block
      const uint16_t len = (uint16_t)n, nlen = (uint16_t)~len;
      memcpy(h + 19, &len, 2); memcpy(h + 21, &nlen, 2);
      const uint32_t crc = (uint32_t)crc32(crc32(0, nullptr, 0), buf.data() + off, (uInt)n), isize = (uint32_t)n;
      failed |= fwrite(h, 1, sizeof h, f) != sizeof h || fwrite(buf.data() + off, 1, n, f) != n || fwrite(&crc, 4, 1, f) != 1 || fwrite(&isize, 4, 1, f) != 1;
      off += n;
    }
    buf.erase(buf.begin(), buf.begin() + (long)off);
  }
  void Put(const void* p, size_t n) { buf.insert(buf.end(), (const uint8_t*)p, (const uint8_t*)p + n); if (buf.size() > (4u << 20)) Flush(false); }
  void I32(int32_t v) { Put(&v, 4); }
  void Eof() {
    Flush(true);
    static const uint8_t eof[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    failed |= fwrite(eof, 1, 28, f) != 28;
  }
};

int Reg2Bin(int64_t beg, int64_t end) {
  --end;
  if (beg >> 14 == end >> 14) return (int)(((1 << 15) - 1) / 7 + (beg >> 14));
  if (beg >> 17 == end >> 17) return (int)(((1 << 12) - 1) / 7 + (beg >> 17));
  if (beg >> 20 == end >> 20) return (int)(((1 << 9) - 1) / 7 + (beg >> 20));
  if (beg >> 23 == end >> 23) return (int)(((1 << 6) - 1) / 7 + (beg >> 23));
  if (beg >> 26 == end >> 26) return (int)(((1 << 3) - 1) / 7 + (beg >> 26));
  return 0;
}

struct CompressionHeader {
  bool rn = true, ap_delta = true, rr = true;
  uint8_t sub[5][4];                                  // [reference base A C G T N][code] -> read base
  std::vector<std::vector<std::string>> td;           // tag lines: 3-byte (tag, tag, type) entries
  std::map<std::string, Enc> ds;
  std::map<int32_t, Enc> tags;
};

int ParseCompressionHeader(const Block& b, CompressionHeader* h) {
  Rd r{b.data.data(), b.data.data() + b.data.size()};
  static const char kBases[] = "ACGTN";
  for (int i = 0; i < 5; ++i) for (int k = 0; k < 4; ++k) h->sub[i][k] = 'N';
  r.itf8();
  int32_t n = r.itf8();
  for (int32_t i = 0; i < n && !r.bad; ++i) {
    const char k0 = (char)r.u8(), k1 = (char)r.u8();
    if (k0 == 'R' && k1 == 'N') h->rn = r.u8() != 0;
    else if (k0 == 'A' && k1 == 'P') h->ap_delta = r.u8() != 0;
    else if (k0 == 'R' && k1 == 'R') h->rr = r.u8() != 0;
    else if (k0 == 'S' && k1 == 'M') {
      for (int ref = 0; ref < 5; ++ref) {
        const uint8_t m = r.u8();
        int k = 0;
        for (int alt = 0; alt < 5; ++alt) {
          if (alt == ref) continue;
          h->sub[ref][(m >> (6 - 2 * k)) & 3] = (uint8_t)kBases[alt];
          ++k;
        }
      }
    } else if (k0 == 'T' && k1 == 'D') {
      const int32_t len = r.itf8();
      if (r.bad || len < 0 || (r.e - r.p) < len) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "CRAM: truncated tag dictionary");
      std::vector<std::string> line;
      for (int32_t q = 0; q < len;) {
        if (r.p[q] == 0) { h->td.push_back(line); line.clear(); ++q; continue; }
        if (q + 3 > len) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "CRAM: malformed tag dictionary");
        line.emplace_back((const char*)r.p + q, 3); q += 3;
      }
      r.skip((size_t)len);
    } else {
      return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "CRAM: unknown preservation key %c%c", k0, k1);
    }
  }
  r.itf8(); n = r.itf8();
  for (int32_t i = 0; i < n && !r.bad; ++i) {
    std::string key; key.push_back((char)r.u8()); key.push_back((char)r.u8());
    const int st = ParseEnc(r, &h->ds[key]);
    if (st) return st;
  }
  r.itf8(); n = r.itf8();
  for (int32_t i = 0; i < n && !r.bad; ++i) {
    const int32_t key = r.itf8();
    const int st = ParseEnc(r, &h->tags[key]);
    if (st) return st;
  }
  return r.bad ? dvb::fail(DVB_ERR_INVALID_ARGUMENT, "CRAM: truncated compression header") : DVB_OK;
}

// One data container -> BAM records (appended to `w`).  Containers are independent of each other: dvb_cram_to_bam decodes a batch of
// them on several threads and writes the results in file order.
struct Out {
  std::vector<uint8_t> buf;
  void Put(const void* p, size_t n) { buf.insert(buf.end(), (const uint8_t*)p, (const uint8_t*)p + n); }
  void I32(int32_t v) { Put(&v, 4); }
};

struct Ctx {
  const char* cram_path;
  const std::vector<std::string>* sq_names;
  const std::vector<int>* sq_to_given;
  const uint8_t* const* ref_bases;
  const int64_t* ref_lens;
};

int DecodeContainer(const Ctx& cx, const std::vector<uint8_t>& cbuf, const std::vector<int32_t>& landmarks, const std::string& name_prefix, Out& w,
                    int64_t* n_written) {
  const char* cram_path = cx.cram_path;
  const std::vector<std::string>& sq_names = *cx.sq_names;
  const std::vector<int>& sq_to_given = *cx.sq_to_given;
  const uint8_t* const* ref_bases = cx.ref_bases;
  const int64_t* ref_lens = cx.ref_lens;
  int64_t name_count = 0;
  int64_t* name_counter = &name_count;
  Rd cr{cbuf.data(), cbuf.data() + cbuf.size()};
  // ---- a data container: compression header, then slices at the landmarks
  Block chb;
  int st = ReadBlock(cr, &chb);
  if (st) return st;
  if (chb.ctype != 1) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "%s: container without a compression header", cram_path);
  CompressionHeader ch;
  if ((st = ParseCompressionHeader(chb, &ch))) return st;
  auto enc = [&](const char* key) -> const Enc& { static const Enc none; auto it = ch.ds.find(key); return it == ch.ds.end() ? none : it->second; };
  for (size_t li = 0; li < landmarks.size(); ++li) {
    if (landmarks[li] < 0 || (size_t)landmarks[li] >= cbuf.size()) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "%s: landmark outside its container", cram_path);
    Rd sr{cbuf.data() + landmarks[li], cbuf.data() + cbuf.size()};
    Block shb;
    if ((st = ReadBlock(sr, &shb))) return st;
    if (shb.ctype != 2) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "%s: landmark does not point at a slice header", cram_path);
    Rd hs{shb.data.data(), shb.data.data() + shb.data.size()};
    const int32_t s_ref = hs.itf8(), s_start = hs.itf8();
    hs.itf8();
    const int32_t s_nrec = hs.itf8();
    hs.ltf8();
    const int32_t s_nblocks = hs.itf8(), n_ids = hs.itf8();
    for (int32_t i = 0; i < n_ids; ++i) hs.itf8();
    const int32_t embedded = hs.itf8();
    if (hs.bad || s_nrec < 0 || s_nblocks < 0) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "%s: malformed slice header", cram_path);
    std::vector<std::unique_ptr<Block>> blocks;
    Slice sl;
    for (int32_t i = 0; i < s_nblocks; ++i) {
      blocks.emplace_back(new Block());
      if ((st = ReadBlock(sr, blocks.back().get()))) return st;
      if (blocks.back()->ctype == 5) sl.core = blocks.back().get(); else sl.ext[blocks.back()->id] = blocks.back().get();
    }
    std::vector<Rec> recs((size_t)s_nrec);
    int32_t prev_ap = s_start;
    for (int32_t ri = 0; ri < s_nrec; ++ri) {
      Rec& c = recs[(size_t)ri];
      c.bf = sl.Int(enc("BF")); c.cf = sl.Int(enc("CF"));
      c.ref = s_ref == -2 ? sl.Int(enc("RI")) : s_ref;
      c.rl = sl.Int(enc("RL"));
      c.ap = sl.Int(enc("AP"));
      if (ch.ap_delta) { c.ap += prev_ap; prev_ap = c.ap; }
      sl.Int(enc("RG"));
      if (ch.rn) sl.Bytes(enc("RN"), &c.name);
      if (c.cf & 2) {                                    // detached: the mate fields are stored
        c.mf = sl.Int(enc("MF"));
        if (!ch.rn) sl.Bytes(enc("RN"), &c.name);
        c.ns = sl.Int(enc("NS")); c.np = sl.Int(enc("NP")); c.tlen = sl.Int(enc("TS"));
      } else if (c.cf & 4) {
        c.mate_line = ri + sl.Int(enc("NF")) + 1;
      }
      const int32_t tl = sl.Int(enc("TL"));
      if (sl.bad || c.rl < 0 || tl < 0 || (size_t)tl >= ch.td.size()) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "%s: malformed record %d of a slice", cram_path, ri);
      for (const std::string& t : ch.td[(size_t)tl]) {
        const int32_t key = ((uint8_t)t[0] << 16) | ((uint8_t)t[1] << 8) | (uint8_t)t[2];
        auto it = ch.tags.find(key);
        if (it == ch.tags.end()) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "%s: tag %c%c without an encoding", cram_path, t[0], t[1]);
        c.aux.append(t);
        const size_t v0 = c.aux.size();
        sl.Bytes(it->second, &c.aux);
        if ((t[2] == 'Z' || t[2] == 'H') && (c.aux.size() == v0 || c.aux.back() != 0)) c.aux.push_back(0);
      }
      c.seq.assign((size_t)c.rl, 'N');
      c.qual.assign((size_t)c.rl, (char)0xff);
      c.end = c.ap;
      if (!(c.bf & 4)) {
        RefSeq ref;
        if (embedded >= 0) {
          Block* eb = sl.Ext(embedded);
          if (eb) { ref.bases = eb->data.data(); ref.len = (int64_t)eb->data.size(); ref.origin = (int64_t)s_start - 1; }
        } else if (c.ref >= 0 && (size_t)c.ref < sq_to_given.size() && sq_to_given[(size_t)c.ref] >= 0) {
          ref.bases = ref_bases[sq_to_given[(size_t)c.ref]]; ref.len = ref_lens[sq_to_given[(size_t)c.ref]];
        }
        bool ref_missing = false;
        auto ref_at = [&](int64_t pos0) -> char {
          const int64_t i = pos0 - ref.origin;
          if (!ref.bases || i < 0 || i >= ref.len) { ref_missing = true; return 'N'; }
          const char b = (char)ref.bases[i];
          return (b >= 'a' && b <= 'z') ? (char)(b - 32) : b;
        };
        const int32_t fn = sl.Int(enc("FN"));
        int64_t read_pos = 0, ref_pos = (int64_t)c.ap - 1;
        int32_t fpos = 0;
        auto fill = [&](int64_t upto) {                   // reference matches up to read index `upto` (exclusive)
          const int64_t gap = std::min<int64_t>(upto, c.rl) - read_pos;
          if (gap <= 0) return;
          for (int64_t k = 0; k < gap; ++k) c.seq[(size_t)(read_pos + k)] = ref_at(ref_pos + k);
          PushCigar(&c.cigar, 0, gap); read_pos += gap; ref_pos += gap;
        };
        for (int32_t f = 0; f < fn && !sl.bad; ++f) {
          const char code = (char)sl.Byte(enc("FC"));
          fpos += sl.Int(enc("FP"));
          fill((int64_t)fpos - 1);
          std::string bytes;
          switch (code) {
            case 'X': {
              const int bs = sl.Byte(enc("BS")) & 3;
              const char rb = ref_at(ref_pos);
              const int ri5 = rb == 'A' ? 0 : rb == 'C' ? 1 : rb == 'G' ? 2 : rb == 'T' ? 3 : 4;
              if (read_pos < c.rl) c.seq[(size_t)read_pos] = (char)ch.sub[ri5][bs];
              PushCigar(&c.cigar, 0, 1); ++read_pos; ++ref_pos; break;
            }
            case 'B': {
              const int b = sl.Byte(enc("BA")), q = sl.Byte(enc("QS"));
              if (read_pos < c.rl) { c.seq[(size_t)read_pos] = (char)b; c.qual[(size_t)read_pos] = (char)q; }
              PushCigar(&c.cigar, 0, 1); ++read_pos; ++ref_pos; break;
            }
            case 'b':
              sl.Bytes(enc("BB"), &bytes);
              for (size_t k = 0; k < bytes.size() && read_pos + (int64_t)k < c.rl; ++k) c.seq[(size_t)read_pos + k] = bytes[k];
              PushCigar(&c.cigar, 0, (int64_t)bytes.size()); read_pos += (int64_t)bytes.size(); ref_pos += (int64_t)bytes.size(); break;
            case 'I':
              sl.Bytes(enc("IN"), &bytes);
              for (size_t k = 0; k < bytes.size() && read_pos + (int64_t)k < c.rl; ++k) c.seq[(size_t)read_pos + k] = bytes[k];
              PushCigar(&c.cigar, 1, (int64_t)bytes.size()); read_pos += (int64_t)bytes.size(); break;
            case 'i': {
              const int b = sl.Byte(enc("BA"));
              if (read_pos < c.rl) c.seq[(size_t)read_pos] = (char)b;
              PushCigar(&c.cigar, 1, 1); ++read_pos; break;
            }
            case 'S':
              sl.Bytes(enc("SC"), &bytes);
              for (size_t k = 0; k < bytes.size() && read_pos + (int64_t)k < c.rl; ++k) c.seq[(size_t)read_pos + k] = bytes[k];
              PushCigar(&c.cigar, 4, (int64_t)bytes.size()); read_pos += (int64_t)bytes.size(); break;
            case 'D': { const int32_t n = sl.Int(enc("DL")); PushCigar(&c.cigar, 2, n); ref_pos += n; break; }
            case 'N': { const int32_t n = sl.Int(enc("RS")); PushCigar(&c.cigar, 3, n); ref_pos += n; break; }
            case 'H': PushCigar(&c.cigar, 5, sl.Int(enc("HC"))); break;
            case 'P': PushCigar(&c.cigar, 6, sl.Int(enc("PD"))); break;
            case 'Q': { const int q = sl.Byte(enc("QS")); if (fpos >= 1 && fpos <= c.rl) c.qual[(size_t)fpos - 1] = (char)q; break; }
            case 'q':
              sl.Bytes(enc("QQ"), &bytes);
              for (size_t k = 0; k < bytes.size() && (int64_t)fpos - 1 + (int64_t)k < c.rl; ++k) c.qual[(size_t)fpos - 1 + k] = bytes[k];
              break;
            default: return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "%s: read feature '%c'", cram_path, code);
          }
        }
        fill(c.rl);
        if (ref_missing && ch.rr)
          return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "%s: the reference bases of %s are needed to decode its reads (pass the FASTA the file was written against)", cram_path,
                           c.ref >= 0 && (size_t)c.ref < sq_names.size() ? sq_names[(size_t)c.ref].c_str() : "?");
        c.end = ref_pos;                                  // 1-based inclusive end = 0-based exclusive end
        c.mq = sl.Int(enc("MQ"));
        if (c.cf & 1) for (int32_t k = 0; k < c.rl; ++k) c.qual[(size_t)k] = (char)sl.Byte(enc("QS"));
      } else {
        for (int32_t k = 0; k < c.rl; ++k) c.seq[(size_t)k] = (char)sl.Byte(enc("BA"));
        if (c.cf & 1) for (int32_t k = 0; k < c.rl; ++k) c.qual[(size_t)k] = (char)sl.Byte(enc("QS"));
      }
      if (c.cf & 8) { c.seq.clear(); c.qual.clear(); }      // sequence unknown ('*')
      if (sl.bad) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "%s: a data series ran out inside record %d of a slice", cram_path, ri);
    }
    // ---- mates (cram_decode_slice_xref)
    std::vector<int32_t> mate_ref((size_t)s_nrec, -1), mate_pos((size_t)s_nrec, 0);
    for (int32_t ri = 0; ri < s_nrec; ++ri) {
      Rec& c = recs[(size_t)ri];
      if (c.mate_line >= 0) {
        if (c.mate_line >= s_nrec) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "%s: mate line outside its slice", cram_path);
        if (c.tlen == INT_MIN) {
          int id2 = ri; int64_t aleft = c.ap, aright = c.end; int ref = c.ref, left_cnt = 0;
          do {
            Rec& m = recs[(size_t)id2];
            if (aleft > m.ap) { aleft = m.ap; left_cnt = 1; } else if (aleft == m.ap) ++left_cnt;
            if (aright < m.end) aright = m.end;
            if (m.mate_line == -1) { m.mate_line = ri; break; }
            if (m.mate_line <= id2 || m.mate_line >= s_nrec) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "%s: mate chain does not move forward", cram_path);
            id2 = m.mate_line;
            if (recs[(size_t)id2].ref != ref) ref = -1;
          } while (id2 != ri);
          const int64_t tlen = aright - aleft + 1;
          id2 = ri;
          do {
            Rec& m = recs[(size_t)id2];
            m.tlen = ref == -1 ? 0 : (m.ap == aleft && (left_cnt == 1 || (m.bf & 0x40))) ? (int32_t)tlen : (int32_t)-tlen;
            id2 = m.mate_line;
          } while (id2 != ri && id2 >= 0);
        }
        const Rec& m = recs[(size_t)c.mate_line];
        mate_pos[(size_t)ri] = m.ap; mate_ref[(size_t)ri] = m.ref;
        c.bf |= 1;
        if (m.bf & 4) { c.bf |= 8; c.tlen = 0; }
        if (c.bf & 4) c.tlen = 0;
        if (m.bf & 0x10) c.bf |= 0x20;
      } else {
        if (c.mf & 1) c.bf |= 1 | 0x20;
        if (c.mf & 2) c.bf |= 8;
        mate_ref[(size_t)ri] = (c.bf & 1) ? c.ns : -1;
        mate_pos[(size_t)ri] = c.np;
      }
      if (c.tlen == INT_MIN) c.tlen = 0;
    }
    // ---- BAM records
    for (int32_t ri = 0; ri < s_nrec; ++ri) {
      Rec& c = recs[(size_t)ri];
      if (c.name.empty()) {                                  // names not preserved: mates share a generated one
        int head = ri;
        for (int32_t k = 0; k < ri; ++k) if (recs[(size_t)k].mate_line == ri && k < head) head = k;
        c.name = head < ri ? recs[(size_t)head].name : name_prefix + std::to_string((*name_counter)++);
      }
      if (c.name.size() > 254) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "%s: read name of %zu bytes", cram_path, c.name.size());
      if (c.cigar.size() > 65535) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "%s: a read with %zu CIGAR operations (BAM holds 65535; the CG-tag form is not written)", cram_path, c.cigar.size());
      const int32_t pos0 = (c.ref >= 0 || c.ap > 0) ? c.ap - 1 : -1;
      const int64_t end0 = (c.bf & 4) || c.end <= pos0 ? (int64_t)pos0 + 1 : c.end;
      const int32_t l_seq = (int32_t)c.seq.size();
      const size_t body = 32 + c.name.size() + 1 + 4 * c.cigar.size() + ((size_t)l_seq + 1) / 2 + (size_t)l_seq + c.aux.size();
      w.I32((int32_t)body);
      w.I32(c.ref); w.I32(pos0);
      const uint8_t lname = (uint8_t)(c.name.size() + 1), mapq = (uint8_t)c.mq;
      w.Put(&lname, 1); w.Put(&mapq, 1);
      const uint16_t bin = (uint16_t)Reg2Bin(std::max(pos0, 0), std::max<int64_t>(end0, 1)), ncig = (uint16_t)c.cigar.size(), flag = (uint16_t)c.bf;
      w.Put(&bin, 2); w.Put(&ncig, 2); w.Put(&flag, 2);
      w.I32(l_seq); w.I32(mate_ref[(size_t)ri]); w.I32(mate_pos[(size_t)ri] - 1); w.I32(c.tlen);
      w.Put(c.name.c_str(), c.name.size() + 1);
      if (!c.cigar.empty()) w.Put(c.cigar.data(), 4 * c.cigar.size());
      std::vector<uint8_t> packed(((size_t)l_seq + 1) / 2, 0);
      for (int32_t k = 0; k < l_seq; ++k) {
        const char* tbl = "=ACMGRSVTWYHKDBN";
        const char b = c.seq[(size_t)k] >= 'a' && c.seq[(size_t)k] <= 'z' ? (char)(c.seq[(size_t)k] - 32) : c.seq[(size_t)k];
        const char* at = strchr(tbl, b);
        const uint8_t code = (at && b) ? (uint8_t)(at - tbl) : 15;
        packed[(size_t)k >> 1] |= (k & 1) ? code : (uint8_t)(code << 4);
      }
      if (!packed.empty()) w.Put(packed.data(), packed.size());
      if (l_seq) w.Put(c.qual.data(), (size_t)l_seq);
      if (!c.aux.empty()) w.Put(c.aux.data(), c.aux.size());
      ++*n_written;
    }
  }
  return DVB_OK;
}

}  // namespace

extern "C" {

int dvb_cram_to_bam(const char* cram_path, const char* bam_path, const char* const* ref_names, const uint8_t* const* ref_bases,
                    const int64_t* ref_lens, int32_t n_refs, const char* const* region_contigs, const int64_t* region_starts,
                    const int64_t* region_ends, int32_t n_regions, int64_t* n_records_out) {
  if (!cram_path || !bam_path || n_refs < 0 || (n_refs > 0 && (!ref_names || !ref_bases || !ref_lens)) || n_regions < 0 ||
      (n_regions > 0 && (!region_contigs || !region_starts || !region_ends)))
    return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_cram_to_bam: bad arguments");
  if (n_records_out) *n_records_out = 0;
  FILE* in = fopen(cram_path, "rb");
  if (!in) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "cannot open %s", cram_path);
  struct Closer { FILE* f; ~Closer() { if (f) fclose(f); } } in_closer{in};
  uint8_t def[26];
  if (fread(def, 1, 26, in) != 26 || memcmp(def, "CRAM", 4) != 0) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "%s is not a CRAM file", cram_path);
  if (def[4] != 3) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "%s: CRAM %d.%d (3.0 / 3.1 containers with the 3.0 codecs are supported)", cram_path, def[4], def[5]);
  FILE* out = fopen(bam_path, "wb");
  if (!out) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "cannot write %s", bam_path);
  Closer out_closer{out};
  Writer w; w.f = out;

  std::vector<std::string> sq_names;          // the CRAM header's @SQ order: what ref ids index
  std::vector<int64_t> sq_lens;
  std::vector<int> sq_to_given;               // -> index into ref_names, -1 when the caller did not supply that contig
  std::vector<std::pair<int, std::pair<int64_t, int64_t>>> regions;   // by CRAM ref id
  bool header_done = false;
  int64_t n_written = 0;
  std::vector<uint8_t> cbuf;
  struct Job { std::vector<uint8_t> cbuf; std::vector<int32_t> landmarks; std::string name_prefix, error; Out out; int64_t n = 0; int status = DVB_OK; };
  std::vector<Job> jobs;
  int64_t container_no = 0;
  const unsigned hw = std::thread::hardware_concurrency();
  const int n_threads = (int)std::max(1u, std::min(16u, hw ? hw : 1u));
  const int batch = 4 * n_threads;
  const Ctx cx{cram_path, &sq_names, &sq_to_given, ref_bases, ref_lens};
  auto run_jobs = [&]() -> int {
    if (jobs.empty()) return DVB_OK;
    std::atomic<size_t> next{0};
    auto work = [&]() {
      for (size_t k; (k = next.fetch_add(1)) < jobs.size();) {
        Job& j = jobs[k];
        j.status = DecodeContainer(cx, j.cbuf, j.landmarks, j.name_prefix, j.out, &j.n);
        if (j.status) j.error = dvb::last_error();        // the message is thread-local: carry it to the caller's thread
        std::vector<uint8_t>().swap(j.cbuf);
      }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads && (size_t)t < jobs.size(); ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    for (Job& j : jobs) {
      if (j.status) return dvb::fail(j.status, "%s", j.error.c_str());
      if (!j.out.buf.empty()) w.Put(j.out.buf.data(), j.out.buf.size());
      n_written += j.n;
    }
    jobs.clear();
    return DVB_OK;
  };
  for (;;) {
    uint8_t lenb[4];
    if (fread(lenb, 1, 4, in) != 4) break;              // no EOF container: accept the end of the file
    const int32_t clen = (int32_t)Le32(lenb);
    uint8_t hdr[64];
    const size_t got = fread(hdr, 1, sizeof hdr, in);
    Rd hr{hdr, hdr + got};
    const int32_t c_ref = hr.itf8(), c_start = hr.itf8(), c_span = hr.itf8(), c_nrec = hr.itf8();
    hr.ltf8(); hr.ltf8();
    const int32_t c_nblocks = hr.itf8(), n_land = hr.itf8();
    if (hr.bad || clen < 0 || n_land < 0) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "%s: malformed container header", cram_path);
    // landmarks may run past the 64 bytes read: re-read exactly
    long hdr_used = (long)(hr.p - hdr);
    fseek(in, (long)hdr_used - (long)got, SEEK_CUR);
    std::vector<int32_t> landmarks;
    {
      std::vector<uint8_t> lb((size_t)n_land * 5 + 4);
      const size_t g2 = fread(lb.data(), 1, lb.size(), in);
      Rd lr{lb.data(), lb.data() + g2};
      for (int32_t i = 0; i < n_land; ++i) landmarks.push_back(lr.itf8());
      lr.skip(4);                                   // CRC32 of the container header
      if (lr.bad) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "%s: truncated container header", cram_path);
      fseek(in, (long)(lr.p - lb.data()) - (long)g2, SEEK_CUR);
    }
    if (c_nrec == 0 && c_ref == -1 && c_start == 4542278 && header_done) break;       // the EOF container
    bool wanted = true;
    if (header_done && !regions.empty()) {
      wanted = c_ref == -2;
      for (const auto& g : regions)
        if (g.first == c_ref && g.second.first < (int64_t)c_start - 1 + c_span && g.second.second > (int64_t)c_start - 1) wanted = true;
    }
    if (header_done && (!wanted || c_nblocks == 0)) { fseek(in, clen, SEEK_CUR); continue; }
    cbuf.resize((size_t)clen);
    if (fread(cbuf.data(), 1, cbuf.size(), in) != cbuf.size()) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "%s: truncated container", cram_path);
    Rd cr{cbuf.data(), cbuf.data() + cbuf.size()};
    if (!header_done) {                                   // the SAM header container
      Block hb;
      int st = ReadBlock(cr, &hb);
      if (st) return st;
      if (hb.data.size() < 4) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "%s: empty header block", cram_path);
      const int32_t tl = (int32_t)Le32(hb.data.data());
      if (tl < 0 || (size_t)tl + 4 > hb.data.size()) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "%s: malformed header block", cram_path);
      const std::string text((const char*)hb.data.data() + 4, (size_t)tl);
      for (size_t p = 0; p < text.size();) {
        size_t e = text.find('\n', p); if (e == std::string::npos) e = text.size();
        const std::string line = text.substr(p, e - p); p = e + 1;
        if (line.compare(0, 3, "@SQ") != 0) continue;
        std::string name; int64_t len = 0;
        for (size_t q = 3; q < line.size();) {
          size_t t = line.find('\t', q + 1); if (t == std::string::npos) t = line.size();
          const std::string field = line.substr(q + 1, t - q - 1);
          if (field.compare(0, 3, "SN:") == 0) name = field.substr(3);
          if (field.compare(0, 3, "LN:") == 0) len = atoll(field.c_str() + 3);
          q = t;
        }
        sq_names.push_back(name); sq_lens.push_back(len);
        int given = -1;
        for (int32_t i = 0; i < n_refs; ++i) if (name == ref_names[i]) given = i;
        sq_to_given.push_back(given);
      }
      for (int32_t i = 0; i < n_regions; ++i)
        for (size_t k = 0; k < sq_names.size(); ++k)
          if (sq_names[k] == region_contigs[i]) regions.push_back({(int)k, {region_starts[i], region_ends[i]}});
      if (n_regions > 0 && regions.empty()) regions.push_back({-3, {0, 0}});        // none of the regions' contigs is in the file: nothing is wanted
      w.Put("BAM\1", 4); w.I32((int32_t)text.size()); w.Put(text.data(), text.size()); w.I32((int32_t)sq_names.size());
      for (size_t k = 0; k < sq_names.size(); ++k) {
        w.I32((int32_t)sq_names[k].size() + 1); w.Put(sq_names[k].c_str(), sq_names[k].size() + 1); w.I32((int32_t)sq_lens[k]);
      }
      header_done = true;
      continue;
    }
    // ---- a data container: queued; a batch is decoded on several threads and written in file order
    jobs.emplace_back();
    jobs.back().cbuf.swap(cbuf);
    jobs.back().landmarks = landmarks;
    jobs.back().name_prefix = "cram." + std::to_string(container_no++) + ".";
    if (jobs.size() >= (size_t)batch) { const int st = run_jobs(); if (st) return st; }
  }
  { const int st = run_jobs(); if (st) return st; }
  if (!header_done) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "%s: no header container", cram_path);
  w.Eof();
  if (w.failed) return dvb::fail(DVB_ERR_INTERNAL, "writing %s failed", bam_path);
  if (n_records_out) *n_records_out = n_written;
  return DVB_OK;
}

}  // extern "C"

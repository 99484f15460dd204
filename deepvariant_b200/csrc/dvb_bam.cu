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
#include <string>
#include <thread>
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
  int64_t n_records_seen = 0;
};

extern "C" {

void dvb_read_requirements_default(DvbReadRequirements* r) {
  memset(r, 0, sizeof(*r));
  r->min_mapping_quality = 5;   // make_examples_options.py:957-964
}

int dvb_bam_open(const char* path, const DvbReadRequirements* req_in, int parse_hp, int threads, DvbBam** out) {
  if (!path || !out) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_bam_open: null argument");
  *out = nullptr;
  DvbReadRequirements req;
  if (req_in) req = *req_in; else dvb_read_requirements_default(&req);
  FILE* f = fopen(path, "rb");
  if (!f) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "cannot open %s", path);
  fseek(f, 0, SEEK_END);
  const long fsz = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::vector<uint8_t> file((size_t)std::max(0L, fsz));
  if (fsz > 0 && fread(file.data(), 1, file.size(), f) != file.size()) { fclose(f); return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "short read on %s", path); }
  fclose(f);

  // ---- BGZF block index: gzip member with the 'BC' extra subfield (BSIZE), ISIZE in the last 4 bytes
  std::vector<BgzfBlock> blocks;
  size_t off = 0, total_out = 0;
  while (off + 18 <= file.size()) {
    const uint8_t* h = file.data() + off;
    if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "%s: not a BGZF block at offset %zu", path, off);
    const uint16_t xlen = rd16(h + 10);
    size_t x = 12, bsize = 0;
    while (x + 4 <= 12 + (size_t)xlen) {
      const uint16_t slen = rd16(h + x + 2);
      if (h[x] == 'B' && h[x + 1] == 'C' && slen == 2) bsize = (size_t)rd16(h + x + 4) + 1;
      x += 4 + slen;
    }
    if (!bsize || off + bsize > file.size()) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "%s: truncated BGZF block at offset %zu", path, off);
    const size_t isize = rd32(h + bsize - 4);
    blocks.push_back({off + 12 + xlen, bsize - 12 - xlen - 8, total_out, isize});
    total_out += isize;
    off += bsize;
  }
  // ---- block-parallel inflate
  std::vector<uint8_t> data(total_out);
  {
    int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    nt = std::max(1, std::min(nt, (int)std::max<size_t>(1, blocks.size() / 4)));
    std::atomic<size_t> next{0};
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
        zs.next_in = file.data() + k.in_off; zs.avail_in = (uInt)k.in_len;
        zs.next_out = data.data() + k.out_off; zs.avail_out = (uInt)k.out_len;
        const int rc = inflate(&zs, Z_FINISH);
        inflateEnd(&zs);
        if (rc != Z_STREAM_END || zs.avail_out != 0) { bad = 1; return; }
      }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    if (bad) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "%s: BGZF inflate failed", path);
  }
  std::vector<uint8_t>().swap(file);

  // ---- header
  if (data.size() < 12 || memcmp(data.data(), "BAM\1", 4) != 0) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "%s is not a BAM file", path);
  DvbBam* bam = new DvbBam();
  size_t p = 8 + (size_t)rdi32(data.data() + 4);
  if (p + 4 > data.size()) { delete bam; return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "%s: truncated header", path); }
  const int32_t n_ref = rdi32(data.data() + p);
  p += 4;
  for (int32_t i = 0; i < n_ref; ++i) {
    if (p + 4 > data.size()) { delete bam; return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "%s: truncated reference list", path); }
    const int32_t l_name = rdi32(data.data() + p);
    if (l_name < 1 || p + 8 + (size_t)l_name > data.size()) { delete bam; return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "%s: bad reference record", path); }
    bam->refs.emplace_back(reinterpret_cast<const char*>(data.data() + p + 4), (size_t)l_name - 1);
    bam->ref_len.push_back(rdi32(data.data() + p + 4 + l_name));
    p += 8 + (size_t)l_name;
  }
  // ---- records
  static const char kSeq[] = "=ACMGRSVTWYHKDBN";
  bam->seq_begin.push_back(0); bam->cigar_begin.push_back(0); bam->name_begin.push_back(0);
  while (p + 4 <= data.size()) {
    const int32_t block_size = rdi32(data.data() + p);
    if (block_size < 32 || p + 4 + (size_t)block_size > data.size()) { delete bam; return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "%s: truncated alignment record", path); }
    const uint8_t* r = data.data() + p + 4;
    p += 4 + (size_t)block_size;
    bam->n_records_seen++;
    const int32_t ref_id = rdi32(r), pos = rdi32(r + 4);
    const uint8_t l_read_name = r[8], mapq = r[9];
    const uint16_t n_cigar = rd16(r + 12), flag = rd16(r + 14);
    const int32_t l_seq = rdi32(r + 16), next_ref = rdi32(r + 20), tlen = rdi32(r + 28);
    const size_t need = 32 + (size_t)l_read_name + 4 * (size_t)n_cigar + ((size_t)l_seq + 1) / 2 + (size_t)l_seq;
    if (l_seq < 0 || need > (size_t)block_size) { delete bam; return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "%s: malformed alignment record", path); }
    // ReadSatisfiesRequirements (sam_reader.cc:217-245)
    if (((flag & FDUP) && !req.keep_duplicates) || ((flag & FQCFAIL) && !req.keep_failed_vendor_quality_checks) ||
        ((flag & FSECONDARY) && !req.keep_secondary_alignments) || ((flag & FSUPP) && !req.keep_supplementary_alignments))
      continue;
    const bool mapped = !(flag & FUNMAP);
    if (!mapped && !req.keep_unaligned) continue;
    const bool paired = flag & FPAIRED;
    if (!req.keep_improperly_placed && mapped) {   // IsReadProperlyPlaced (utils.cc:261-266)
      const bool mate_has_contig = paired && !(flag & FMUNMAP) && next_ref >= 0;
      if (!(!paired || (flag & FPROPER) || !mate_has_contig || next_ref == ref_id)) continue;
    }
    if (mapped && (int)mapq < req.min_mapping_quality) continue;
    // ---- decode
    const uint8_t* name = r + 32;
    const uint8_t* cig = name + l_read_name;
    const uint8_t* seq = cig + 4 * (size_t)n_cigar;
    const uint8_t* qual = seq + ((size_t)l_seq + 1) / 2;
    const uint8_t* aux = qual + l_seq;
    bam->ref_id.push_back(ref_id); bam->pos.push_back(pos); bam->mapq.push_back(mapq); bam->flag.push_back(flag);
    bam->fragment_length.push_back(tlen);
    bam->read_number.push_back(((flag & FREAD1) || !paired) ? 0 : 1);
    bam->number_reads.push_back(paired ? 2 : 1);
    int32_t e = pos;
    if (mapped) {
      for (uint16_t k = 0; k < n_cigar; ++k) {
        const uint32_t c = rd32(cig + 4 * (size_t)k);
        bam->cigar.push_back(c);
        const uint32_t op = c & 0xF;
        if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) e += (int32_t)(c >> 4);
      }
    }
    bam->end.push_back(e);
    bam->cigar_begin.push_back((int64_t)bam->cigar.size());
    const size_t b0 = bam->bases.size();
    bam->bases.resize(b0 + (size_t)l_seq);
    for (int32_t i = 0; i < l_seq; ++i) bam->bases[b0 + i] = (uint8_t)kSeq[(seq[i >> 1] >> ((~i & 1) << 2)) & 0xF];
    bam->quals.insert(bam->quals.end(), qual, qual + l_seq);
    bam->seq_begin.push_back((int64_t)bam->bases.size());
    bam->names.insert(bam->names.end(), name, name + (l_read_name ? l_read_name - 1 : 0));
    bam->name_begin.push_back((int64_t)bam->names.size());
    bam->hp.push_back(parse_hp ? ParseHp(aux, (size_t)block_size - need) : INT32_MIN);
  }
  *out = bam;
  return DVB_OK;
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
  return DVB_OK;
}

const char* dvb_bam_ref_name(const DvbBam* bam, int32_t i) {
  if (!bam || i < 0 || i >= (int32_t)bam->refs.size()) return nullptr;
  return bam->refs[(size_t)i].c_str();
}

void dvb_bam_close(DvbBam* bam) { delete bam; }

}  // extern "C"

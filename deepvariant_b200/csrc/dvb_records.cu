// dvb_records.cu — host-only record I/O of the call_variants stage.
//
// Replaces, around the classifier, what the reference does through tf.data and Python:
//   input   call_variants.get_dataset              deepvariant/call_variants.py:449-538
//             TFRecordDataset.list_files(shuffle=False).interleave(TFRecordDataset(GZIP), cycle_length=32)
//             .map(parse_single_example{image/encoded, variant/encoded, alt_allele_indices/encoded}).batch(B)
//   output  round_gls / _create_cvo_proto / write_variant_call / post_processing
//                                                  deepvariant/call_variants.py:248-399, 541-602
//           variantcall_utils.set_model_id         third_party/nucleus/util/variantcall_utils.py:235
// Reader: worker threads inflate the shards (zlib) and pick the three features out of the tf.Example wire bytes in place;
// the consumer takes records in tf.data's deterministic interleave order (cycle_length slots, block_length 1) and copies
// each image straight into the caller's (pinned) batch buffer — no per-record Python objects, no float32 expansion on the
// host.  Writer: one thread per output shard rounds the probabilities (round-half-even to 10 decimals exactly as Python's
// round()), re-derives the smallest one, splices call.info["MID"] into the serialized Variant and appends the framed
// CallVariantsOutput to a gzip stream.
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <string_view>
#include <thread>
#include <vector>

#include "dvb_common.h"

namespace {

inline uint64_t rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }

// ---- protobuf wire helpers ---------------------------------------------------------------------------------------------
struct Span { const uint8_t* p = nullptr; size_t n = 0; };

inline bool ReadVarint(const uint8_t*& p, const uint8_t* end, uint64_t* v) {
  uint64_t r = 0;
  for (int shift = 0; shift < 70 && p < end; shift += 7) {
    const uint8_t b = *p++;
    r |= (uint64_t)(b & 0x7F) << (shift & 63);
    if (!(b & 0x80)) { *v = r; return true; }
  }
  return false;
}

// Next field of a message: number, wire type, payload (LEN) or value (VARINT / fixed).  false at end or on malformed input
// (*bad set).
inline bool NextField(const uint8_t*& p, const uint8_t* end, uint32_t* fn, uint32_t* wt, Span* payload, uint64_t* value, bool* bad) {
  if (p >= end) return false;
  uint64_t key;
  if (!ReadVarint(p, end, &key)) { *bad = true; return false; }
  *fn = (uint32_t)(key >> 3);
  *wt = (uint32_t)(key & 7);
  switch (*wt) {
    case 0: if (!ReadVarint(p, end, value)) { *bad = true; return false; } return true;
    case 1: if (end - p < 8) { *bad = true; return false; } *value = rd64(p); p += 8; return true;
    case 5: if (end - p < 4) { *bad = true; return false; } *value = rd32(p); p += 4; return true;
    case 2: {
      uint64_t len;
      if (!ReadVarint(p, end, &len) || len > (uint64_t)(end - p)) { *bad = true; return false; }
      payload->p = p; payload->n = (size_t)len;
      p += len;
      return true;
    }
    default: *bad = true; return false;
  }
}

inline void PutVarint(std::string& out, uint64_t v) {
  while (v >= 0x80) { out.push_back((char)(v | 0x80)); v >>= 7; }
  out.push_back((char)v);
}
inline void PutLen(std::string& out, uint32_t fn, const void* data, size_t n) {
  PutVarint(out, ((uint64_t)fn << 3) | 2);
  PutVarint(out, n);
  out.append((const char*)data, n);
}

// ---- one parsed tf.Example ---------------------------------------------------------------------------------------------
struct Example {
  std::vector<uint8_t> buf;     // the record payload; the spans below point into it
  Span image, variant, alt;
  int64_t shape[3] = {0, 0, 0};
  int n_shape = 0;
};

// tf.Example { Features features = 1 { map<string, Feature> feature = 1 } }, Feature { BytesList = 1 | FloatList = 2 |
// Int64List = 3 }.  parse_single_example with FixedLenFeature((), string): exactly one bytes value per requested key.
bool ParseExample(Example* ex, std::string* err) {
  const uint8_t* p = ex->buf.data();
  const uint8_t* end = p + ex->buf.size();
  bool bad = false;
  uint32_t fn, wt;
  Span s;
  uint64_t v;
  int n_image = 0, n_variant = 0, n_alt = 0;
  while (NextField(p, end, &fn, &wt, &s, &v, &bad)) {
    if (fn != 1 || wt != 2) continue;
    const uint8_t* fp = s.p;
    const uint8_t* fend = s.p + s.n;
    Span entry;
    while (NextField(fp, fend, &fn, &wt, &entry, &v, &bad)) {
      if (fn != 1 || wt != 2) continue;
      const uint8_t* ep = entry.p;
      const uint8_t* eend = entry.p + entry.n;
      Span key, feat, t;
      while (NextField(ep, eend, &fn, &wt, &t, &v, &bad)) {
        if (wt != 2) continue;
        if (fn == 1) key = t; else if (fn == 2) feat = t;
      }
      const std::string_view name((const char*)key.p, key.n);
      const bool is_image = name == "image/encoded", is_variant = name == "variant/encoded", is_alt = name == "alt_allele_indices/encoded",
                 is_shape = name == "image/shape";
      if (!is_image && !is_variant && !is_alt && !is_shape) continue;
      const uint8_t* qp = feat.p;
      const uint8_t* qend = feat.p + feat.n;
      Span list;
      while (NextField(qp, qend, &fn, &wt, &list, &v, &bad)) {
        if (wt != 2) continue;
        const uint8_t* lp = list.p;
        const uint8_t* lend = list.p + list.n;
        if (fn == 1 && !is_shape) {          // BytesList: the map keeps the last entry of a key, so reset the count
          int count = 0;
          Span val, last;
          while (NextField(lp, lend, &fn, &wt, &val, &v, &bad))
            if (fn == 1 && wt == 2) { last = val; ++count; }
          if (is_image) { ex->image = last; n_image = count; }
          else if (is_variant) { ex->variant = last; n_variant = count; }
          else { ex->alt = last; n_alt = count; }
          fn = 0;
        } else if (fn == 3 && is_shape) {    // Int64List, packed or not
          ex->n_shape = 0;
          Span packed;
          while (NextField(lp, lend, &fn, &wt, &packed, &v, &bad)) {
            if (fn != 1) continue;
            if (wt == 2) {
              const uint8_t* pp = packed.p;
              const uint8_t* pend = packed.p + packed.n;
              while (pp < pend && ReadVarint(pp, pend, &v)) { if (ex->n_shape < 3) ex->shape[ex->n_shape] = (int64_t)v; ++ex->n_shape; }
            } else if (wt == 0) { if (ex->n_shape < 3) ex->shape[ex->n_shape] = (int64_t)v; ++ex->n_shape; }
          }
          fn = 0;
        }
      }
    }
  }
  if (bad) { *err = "malformed tf.Example"; return false; }
  if (n_image != 1 || n_variant != 1 || n_alt != 1) {
    *err = std::string("tf.Example needs exactly one value of ") + (n_image != 1 ? "image/encoded" : n_variant != 1 ? "variant/encoded" : "alt_allele_indices/encoded");
    return false;
  }
  return true;
}

// ---- one input shard: file -> (gunzip) -> TFRecord frames -> parsed examples -----------------------------------------
struct Shard {
  std::string path;
  FILE* f = nullptr;
  bool opened = false, gz = false, file_eof = false, z_init = false, z_member_done = true;
  z_stream zs;
  std::vector<uint8_t> in;        // compressed bytes from the file
  std::vector<uint8_t> out;       // decompressed bytes not yet framed
  size_t out_pos = 0;
  // guarded by the reader mutex:
  std::deque<std::unique_ptr<Example>> q;
  bool busy = false, done = false;
  std::string error;

  ~Shard() {
    if (z_init) inflateEnd(&zs);
    if (f) fclose(f);
  }

  bool Open(std::string* err) {
    f = fopen(path.c_str(), "rb");
    if (!f) { *err = "cannot open " + path; return false; }
    uint8_t magic[2] = {0, 0};
    const size_t got = fread(magic, 1, 2, f);
    fseek(f, 0, SEEK_SET);
    gz = got == 2 && magic[0] == 0x1f && magic[1] == 0x8b;   // the reference decides by extension (tfrecord.py:88-93); sniffing accepts both
    if (gz) {
      memset(&zs, 0, sizeof(zs));
      if (inflateInit2(&zs, 15 + 16) != Z_OK) { *err = path + ": inflateInit2 failed"; return false; }
      z_init = true;
      in.resize(1 << 20);
    }
    opened = true;
    return true;
  }

  // Makes at least `need` bytes available at out[out_pos..]; false at a clean end of data (fewer bytes left) or on error
  // (*err set).
  bool Fill(size_t need, std::string* err) {
    while (out.size() - out_pos < need) {
      if (out_pos > 0 && out_pos == out.size()) { out.clear(); out_pos = 0; }
      else if (out_pos > (8u << 20)) { out.erase(out.begin(), out.begin() + (ptrdiff_t)out_pos); out_pos = 0; }
      const size_t old = out.size();
      const size_t chunk = std::max<size_t>(need, 1 << 20);
      if (!gz) {
        if (file_eof) return false;
        out.resize(old + chunk);
        const size_t got = fread(out.data() + old, 1, chunk, f);
        out.resize(old + got);
        if (got == 0) file_eof = true;
        continue;
      }
      if (zs.avail_in == 0) {
        if (file_eof) {
          if (!z_member_done) *err = path + ": truncated gzip stream";
          return false;
        }
        const size_t got = fread(in.data(), 1, in.size(), f);
        if (got == 0) { file_eof = true; continue; }
        zs.next_in = in.data();
        zs.avail_in = (uInt)got;
      }
      if (z_member_done) {   // first member, or the next of several concatenated gzip members
        inflateReset(&zs);
        z_member_done = false;
      }
      out.resize(old + chunk);
      zs.next_out = out.data() + old;
      zs.avail_out = (uInt)chunk;
      const int rc = inflate(&zs, Z_NO_FLUSH);
      out.resize(old + (chunk - zs.avail_out));
      if (rc == Z_STREAM_END) z_member_done = true;
      else if (rc != Z_OK && rc != Z_BUF_ERROR) { *err = path + ": corrupt gzip stream"; return false; }
    }
    return true;
  }

  // Reads the next record; 1 = got one, 0 = clean end of file, -1 = error.
  int Next(bool verify_crc, std::unique_ptr<Example>* out_ex, std::string* err) {
    if (!opened && !Open(err)) return -1;
    if (!Fill(12, err)) {
      if (!err->empty()) return -1;
      if (out.size() - out_pos != 0) { *err = path + ": truncated TFRecord header"; return -1; }
      return 0;
    }
    const uint8_t* h = out.data() + out_pos;
    const uint64_t len = rd64(h);
    if (verify_crc && dvb_masked_crc32c(h, 8) != rd32(h + 8)) { *err = path + ": corrupted record length"; return -1; }
    if (len > (1ull << 31)) { *err = path + ": implausible record length"; return -1; }
    out_pos += 12;
    if (!Fill((size_t)len + 4, err)) {
      if (err->empty()) *err = path + ": truncated TFRecord";
      return -1;
    }
    const uint8_t* d = out.data() + out_pos;
    if (verify_crc && dvb_masked_crc32c(d, (size_t)len) != rd32(d + len)) { *err = path + ": corrupted record data"; return -1; }
    auto ex = std::make_unique<Example>();
    ex->buf.assign(d, d + len);
    out_pos += (size_t)len + 4;
    std::string perr;
    if (!ParseExample(ex.get(), &perr)) { *err = path + ": " + perr; return -1; }
    *out_ex = std::move(ex);
    return 1;
  }
};

}  // namespace

struct DvbExamplesReader {
  std::vector<std::unique_ptr<Shard>> shards;
  int cycle_length = 1, prefetch = 32;
  bool verify_crc = true;
  std::mutex mu;
  std::condition_variable cv_work, cv_data;
  std::vector<std::thread> workers;
  bool stop = false;
  // consumer state: tf.data InterleaveDataset (sequential, block_length 1)
  std::vector<int> slots;       // shard index or -1
  int cycle_index = 0, next_file = 0, num_open = 0;
  int lookahead_limit = 0;      // workers may read shards below this index
  std::string error;
  bool finished = false;
  // batch meta handed out by dvb_examples_reader_next, valid until the next call
  std::vector<uint8_t> variant_blob, alt_blob;
  std::vector<int64_t> variant_begin, alt_begin;
  std::unique_ptr<Example> peeked;   // first record, held back by dvb_examples_reader_shape

  void WorkerLoop() {
    std::unique_lock<std::mutex> lk(mu);
    for (;;) {
      if (stop) return;
      Shard* pick = nullptr;
      size_t best = (size_t)prefetch;
      const int limit = std::min<int>(lookahead_limit, (int)shards.size());
      for (int i = 0; i < limit; ++i) {   // the emptiest queue first: the consumer will block on it soonest
        Shard* s = shards[(size_t)i].get();
        if (s->busy || s->done) continue;
        if (s->q.size() < best) { best = s->q.size(); pick = s; }
      }
      if (!pick) { cv_work.wait(lk); continue; }
      pick->busy = true;
      lk.unlock();
      std::vector<std::unique_ptr<Example>> got;
      std::string err;
      bool end = false;
      for (int k = 0; k < 4; ++k) {
        std::unique_ptr<Example> ex;
        const int rc = pick->Next(verify_crc, &ex, &err);
        if (rc == 1) got.push_back(std::move(ex)); else { end = true; break; }
      }
      lk.lock();
      for (auto& e : got) pick->q.push_back(std::move(e));
      if (end) { pick->done = true; pick->error = err; }
      pick->busy = false;
      cv_data.notify_all();
    }
  }

  void AdvanceCycle() { cycle_index = (cycle_index + 1) % cycle_length; }

  // Next record in interleave order; nullptr at the end or on error (error set).  Called with mu held.
  std::unique_ptr<Example> Pop(std::unique_lock<std::mutex>& lk) {
    if (peeked) return std::move(peeked);
    while (!finished && error.empty()) {
      const bool end_of_input = next_file >= (int)shards.size();
      if (end_of_input && num_open == 0) { finished = true; break; }
      int& slot = slots[(size_t)cycle_index];
      if (slot >= 0) {
        Shard* s = shards[(size_t)slot].get();
        while (s->q.empty() && !s->done) { cv_work.notify_all(); cv_data.wait(lk); }
        if (!s->q.empty()) {
          std::unique_ptr<Example> ex = std::move(s->q.front());
          s->q.pop_front();
          cv_work.notify_one();
          AdvanceCycle();
          return ex;
        }
        if (!s->error.empty()) { error = s->error; break; }
        slot = -1;              // element exhausted: the slot is refilled when the cycle comes round again
        --num_open;
        AdvanceCycle();
      } else if (!end_of_input) {
        slot = next_file++;
        ++num_open;
        lookahead_limit = next_file + cycle_length;
        cv_work.notify_all();
      } else {
        AdvanceCycle();
      }
    }
    return nullptr;
  }
};

// ---- CallVariantsOutput writer -----------------------------------------------------------------------------------------
namespace {

// Python's round(x, ndigits) for a double: the correctly rounded decimal with ndigits places (round-half-even on the exact
// binary value), converted back with a correctly rounded strtod (CPython Objects/floatobject.c double_round -> _Py_dg_dtoa
// mode 3).  glibc's printf("%.*f") and strtod are both exact, so this is the same function.
inline double PyRound(double x, int ndigits) {
  char buf[400];
  snprintf(buf, sizeof(buf), "%.*f", ndigits, x);
  return strtod(buf, nullptr);
}

// round_gls (deepvariant/call_variants.py:248-285).  false when the likelihoods do not sum to one.
bool RoundGls(const double gls[3], int precision, double out[3]) {
  const double sum = (gls[0] + gls[1]) + gls[2];   // Python sum(): left to right from 0
  if (std::fabs(sum - 1) > 1e-6) return false;
  out[0] = gls[0]; out[1] = gls[1]; out[2] = gls[2];
  if (precision < 0) return true;
  int min_ix = 0;
  for (int i = 1; i < 3; ++i) if (gls[i] < gls[min_ix]) min_ix = i;   // strict <: the first minimum
  for (int i = 0; i < 3; ++i) out[i] = PyRound(gls[i], precision);
  double others = 0.0;
  for (int i = 0; i < 3; ++i) if (i != min_ix) others += out[i];
  out[min_ix] = std::max(0.0, PyRound(1 - others, precision));
  return true;
}

// variantcall_utils.set_model_id on a serialized Variant: calls[0].info["MID"] = [string_value model_id]; an existing MID
// entry of calls[0] is dropped and the new one appended.  Variant.calls = 11, VariantCall.info = 2 (map<string, ListValue>),
// ListValue.values = 1, Value.string_value = 3.  false when the variant has no calls (the reference raises IndexError).
bool SetModelId(Span variant, const std::string& model_id, std::string* out, bool* bad) {
  std::string value, list_value, entry;
  PutLen(value, 3, model_id.data(), model_id.size());
  PutLen(list_value, 1, value.data(), value.size());
  PutLen(entry, 1, "MID", 3);
  PutLen(entry, 2, list_value.data(), list_value.size());
  const uint8_t* p = variant.p;
  const uint8_t* end = variant.p + variant.n;
  bool done = false;
  for (;;) {
    const uint8_t* field_start = p;
    uint32_t fn, wt;
    Span s;
    uint64_t v;
    if (!NextField(p, end, &fn, &wt, &s, &v, bad)) break;
    if (fn == 11 && wt == 2 && !done) {
      std::string call;
      const uint8_t* cp = s.p;
      const uint8_t* cend = s.p + s.n;
      for (;;) {
        const uint8_t* cstart = cp;
        Span cs;
        if (!NextField(cp, cend, &fn, &wt, &cs, &v, bad)) break;
        if (fn == 2 && wt == 2) {
          std::string_view key;
          const uint8_t* ip = cs.p;
          const uint8_t* iend = cs.p + cs.n;
          Span is;
          uint32_t f3, w3;
          while (NextField(ip, iend, &f3, &w3, &is, &v, bad))
            if (f3 == 1 && w3 == 2) key = std::string_view((const char*)is.p, is.n);
          if (key == "MID") continue;
        }
        call.append((const char*)cstart, (size_t)(cp - cstart));
      }
      PutLen(call, 2, entry.data(), entry.size());
      PutLen(*out, 11, call.data(), call.size());
      done = true;
    } else {
      out->append((const char*)field_start, (size_t)(p - field_start));
    }
  }
  return done && !*bad;
}

struct CvoBatch {
  int32_t n = 0;
  std::vector<uint8_t> variant_blob, alt_blob;
  std::vector<int64_t> variant_begin, alt_begin;
  std::vector<float> probs;
};

}  // namespace

struct DvbCvoWriter {
  std::string path;
  bool gz = false;
  int precision = 10;
  std::string model_id = "deepvariant";
  FILE* f = nullptr;
  z_stream zs;
  bool z_init = false;
  std::vector<uint8_t> zbuf;
  std::mutex mu;
  std::condition_variable cv;
  std::deque<std::unique_ptr<CvoBatch>> q;
  bool closing = false;
  std::string error;
  int64_t n_written = 0;
  std::thread thread;

  bool Emit(const void* data, size_t n, int flush) {
    if (!gz) return n == 0 || fwrite(data, 1, n, f) == n;
    zs.next_in = (Bytef*)data;
    zs.avail_in = (uInt)n;
    for (;;) {
      zs.next_out = zbuf.data();
      zs.avail_out = (uInt)zbuf.size();
      const int rc = deflate(&zs, flush);
      const size_t have = zbuf.size() - zs.avail_out;
      if (have && fwrite(zbuf.data(), 1, have, f) != have) return false;
      if (rc == Z_STREAM_END) return true;
      if (rc != Z_OK && rc != Z_BUF_ERROR) return false;
      if (zs.avail_in == 0 && zs.avail_out != 0 && flush != Z_FINISH) return true;
    }
  }

  bool WriteBatch(const CvoBatch& b, std::string* err) {
    std::string rec, variant, framed;
    for (int32_t i = 0; i < b.n; ++i) {
      const double gls[3] = {(double)b.probs[(size_t)i * 3], (double)b.probs[(size_t)i * 3 + 1], (double)b.probs[(size_t)i * 3 + 2]};
      double r[3];
      if (!RoundGls(gls, precision, r)) {
        char buf[200];
        snprintf(buf, sizeof(buf), "Invalid genotype likelihoods do not sum to one: sum([%.17g, %.17g, %.17g]) = %.17g", gls[0], gls[1], gls[2],
                 (gls[0] + gls[1]) + gls[2]);
        *err = buf;
        return false;
      }
      variant.clear();
      bool bad = false;
      const Span v{b.variant_blob.data() + b.variant_begin[(size_t)i], (size_t)(b.variant_begin[(size_t)i + 1] - b.variant_begin[(size_t)i])};
      if (!SetModelId(v, model_id, &variant, &bad)) { *err = bad ? "malformed variant/encoded" : "variant has no calls"; return false; }
      rec.clear();
      PutLen(rec, 1, variant.data(), variant.size());
      PutLen(rec, 2, b.alt_blob.data() + b.alt_begin[(size_t)i], (size_t)(b.alt_begin[(size_t)i + 1] - b.alt_begin[(size_t)i]));
      PutLen(rec, 3, r, sizeof(r));   // repeated double genotype_probabilities = 3, packed, little-endian
      framed.clear();
      const uint64_t len = rec.size();
      framed.append((const char*)&len, 8);
      const uint32_t hcrc = dvb_masked_crc32c(&len, 8);
      framed.append((const char*)&hcrc, 4);
      framed += rec;
      const uint32_t dcrc = dvb_masked_crc32c(rec.data(), rec.size());
      framed.append((const char*)&dcrc, 4);
      if (!Emit(framed.data(), framed.size(), Z_NO_FLUSH)) { *err = "write to " + path + " failed"; return false; }
      ++n_written;
    }
    return true;
  }

  void Loop() {
    for (;;) {
      std::unique_ptr<CvoBatch> b;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return !q.empty() || closing; });
        if (q.empty()) return;
        b = std::move(q.front());
        q.pop_front();
        cv.notify_all();
      }
      std::string err;
      if (error.empty() && !WriteBatch(*b, &err)) {
        std::lock_guard<std::mutex> lk(mu);
        error = err;
      }
    }
  }
};

extern "C" {

int dvb_examples_reader_open(const char* const* paths, int32_t n_paths, int32_t threads, int32_t cycle_length, int32_t verify_crc,
                             DvbExamplesReader** out) {
  if (!out || n_paths < 0 || (n_paths > 0 && !paths)) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_examples_reader_open: bad arguments");
  *out = nullptr;
  for (int32_t i = 0; i < n_paths; ++i) {
    FILE* f = paths[i] ? fopen(paths[i], "rb") : nullptr;
    if (!f) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "cannot open %s", paths[i] ? paths[i] : "(null)");
    fclose(f);
  }
  auto* r = new DvbExamplesReader();
  for (int32_t i = 0; i < n_paths; ++i) {
    r->shards.emplace_back(new Shard());
    r->shards.back()->path = paths[i];
  }
  r->cycle_length = std::max(1, cycle_length > 0 ? cycle_length : 32);   // _DEFAULT_INPUT_READ_THREADS (call_variants.py:83)
  r->slots.assign((size_t)r->cycle_length, -1);
  r->lookahead_limit = r->cycle_length;
  r->verify_crc = verify_crc != 0;
  int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
  nt = std::max(1, std::min(nt, std::max(1, n_paths)));
  for (int t = 0; t < nt; ++t) r->workers.emplace_back([r] { r->WorkerLoop(); });
  *out = r;
  return DVB_OK;
}

int dvb_examples_reader_shape(DvbExamplesReader* r, int64_t shape[3], int64_t* image_bytes) {
  if (!r || !shape) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_examples_reader_shape: null argument");
  std::unique_lock<std::mutex> lk(r->mu);
  if (!r->peeked) r->peeked = r->Pop(lk);
  if (!r->error.empty()) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "%s", r->error.c_str());
  shape[0] = shape[1] = shape[2] = 0;
  if (image_bytes) *image_bytes = 0;
  if (!r->peeked) return DVB_OK;   // no records at all
  if (r->peeked->n_shape != 3) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "first example has no image/shape of 3 values");
  for (int k = 0; k < 3; ++k) shape[k] = r->peeked->shape[k];
  if (image_bytes) *image_bytes = (int64_t)r->peeked->image.n;
  return DVB_OK;
}

int dvb_examples_reader_next(DvbExamplesReader* r, int32_t max_n, uint8_t* images_host, int64_t image_bytes, int32_t* n_out,
                             DvbExampleBatchMeta* meta) {
  if (!r || !n_out || max_n < 0 || (max_n > 0 && !images_host) || image_bytes < 0)
    return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_examples_reader_next: bad arguments");
  *n_out = 0;
  std::vector<std::unique_ptr<Example>> batch;
  {
    std::unique_lock<std::mutex> lk(r->mu);
    while ((int32_t)batch.size() < max_n) {
      std::unique_ptr<Example> ex = r->Pop(lk);
      if (!ex) break;
      batch.push_back(std::move(ex));
    }
    if (!r->error.empty()) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "%s", r->error.c_str());
  }
  for (size_t i = 0; i < batch.size(); ++i)
    if ((int64_t)batch[i]->image.n != image_bytes)
      return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "image/encoded has %zu bytes, expected %lld", batch[i]->image.n, (long long)image_bytes);
  // images straight into the caller's batch buffer (a few threads: one memcpy stream does not saturate host memory)
  {
    const size_t n = batch.size();
    const int nt = n >= 64 ? 4 : 1;
    auto copy = [&](size_t a, size_t b) {
      for (size_t i = a; i < b; ++i) memcpy(images_host + i * (size_t)image_bytes, batch[i]->image.p, (size_t)image_bytes);
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(copy, n * t / nt, n * (t + 1) / nt);
    copy(0, n / nt);
    for (auto& t : pool) t.join();
  }
  r->variant_blob.clear(); r->alt_blob.clear();
  r->variant_begin.assign(1, 0); r->alt_begin.assign(1, 0);
  for (auto& ex : batch) {
    r->variant_blob.insert(r->variant_blob.end(), ex->variant.p, ex->variant.p + ex->variant.n);
    r->alt_blob.insert(r->alt_blob.end(), ex->alt.p, ex->alt.p + ex->alt.n);
    r->variant_begin.push_back((int64_t)r->variant_blob.size());
    r->alt_begin.push_back((int64_t)r->alt_blob.size());
  }
  if (meta) {
    meta->variant_blob = r->variant_blob.data(); meta->variant_begin = r->variant_begin.data();
    meta->alt_blob = r->alt_blob.data(); meta->alt_begin = r->alt_begin.data();
  }
  *n_out = (int32_t)batch.size();
  return DVB_OK;
}

void dvb_examples_reader_close(DvbExamplesReader* r) {
  if (!r) return;
  {
    std::lock_guard<std::mutex> lk(r->mu);
    r->stop = true;
  }
  r->cv_work.notify_all();
  for (auto& t : r->workers) t.join();
  delete r;
}

int dvb_debug_round_gls(const double gls[3], int32_t precision, double out[3]) {
  if (!gls || !out) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_debug_round_gls: null argument");
  if (!RoundGls(gls, precision, out)) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "Invalid genotype likelihoods do not sum to one");
  return DVB_OK;
}

int dvb_cvo_writer_open(const char* path, int32_t gl_precision, DvbCvoWriter** out) {
  if (!path || !out) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_cvo_writer_open: null argument");
  *out = nullptr;
  auto* w = new DvbCvoWriter();
  w->path = path;
  w->precision = gl_precision;
  const size_t n = w->path.size();
  w->gz = n >= 3 && w->path.compare(n - 3, 3, ".gz") == 0;   // third_party/nucleus/io/tfrecord.py:88-93
  w->f = fopen(path, "wb");
  if (!w->f) { delete w; return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "cannot create %s", path); }
  if (w->gz) {
    memset(&w->zs, 0, sizeof(w->zs));
    if (deflateInit2(&w->zs, Z_DEFAULT_COMPRESSION, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) {
      fclose(w->f);
      delete w;
      return dvb::fail(DVB_ERR_INTERNAL, "deflateInit2 failed");
    }
    w->z_init = true;
    w->zbuf.resize(1 << 18);
  }
  w->thread = std::thread([w] { w->Loop(); });
  *out = w;
  return DVB_OK;
}

int dvb_cvo_writer_write_batch(DvbCvoWriter* w, int32_t n, const DvbExampleBatchMeta* meta, const float* probs) {
  if (!w || n < 0 || (n > 0 && (!meta || !probs))) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_cvo_writer_write_batch: bad arguments");
  auto b = std::make_unique<CvoBatch>();
  b->n = n;
  if (n > 0) {
    b->variant_begin.assign(meta->variant_begin, meta->variant_begin + n + 1);
    b->alt_begin.assign(meta->alt_begin, meta->alt_begin + n + 1);
    b->variant_blob.assign(meta->variant_blob + b->variant_begin[0], meta->variant_blob + b->variant_begin[(size_t)n]);
    b->alt_blob.assign(meta->alt_blob + b->alt_begin[0], meta->alt_blob + b->alt_begin[(size_t)n]);
    const int64_t v0 = b->variant_begin[0], a0 = b->alt_begin[0];
    for (auto& x : b->variant_begin) x -= v0;
    for (auto& x : b->alt_begin) x -= a0;
    b->probs.assign(probs, probs + (size_t)n * 3);
  }
  std::unique_lock<std::mutex> lk(w->mu);
  if (!w->error.empty()) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "%s", w->error.c_str());
  w->cv.wait(lk, [&] { return w->q.size() < 8; });   // bounded, as the reference's writer queues are drained by their processes
  w->q.push_back(std::move(b));
  w->cv.notify_all();
  return DVB_OK;
}

int dvb_cvo_writer_close(DvbCvoWriter* w, int64_t* n_written) {
  if (!w) return DVB_OK;
  {
    std::lock_guard<std::mutex> lk(w->mu);
    w->closing = true;
  }
  w->cv.notify_all();
  if (w->thread.joinable()) w->thread.join();
  bool ok = true;
  if (w->gz && w->z_init) {
    ok = w->Emit(nullptr, 0, Z_FINISH);
    deflateEnd(&w->zs);
  }
  if (w->f && fclose(w->f) != 0) ok = false;
  if (n_written) *n_written = w->n_written;
  const std::string err = !w->error.empty() ? w->error : (ok ? std::string() : "write to " + w->path + " failed");
  delete w;
  if (!err.empty()) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "%s", err.c_str());
  return DVB_OK;
}

}  // extern "C"

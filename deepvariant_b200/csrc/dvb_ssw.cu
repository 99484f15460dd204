// Smith-Waterman alignment with libssw's tie-breaking (host code) - what deepvariant/realigner/ssw.h wraps.
//
// The reference links the third-party Complete-Striped-Smith-Waterman-Library (libssw 1.2.x, third_party/libssw.BUILD; its
// sources are not vendored in the reference).  Its result is a pure function of the Gotoh recurrences plus four published
// tie-breaking rules, restated here on plain scalar arrays:
//   1. score / end: H(i,j) = max(0, H(i-1,j-1) + s, E, F) with E / F opening at -gap_open and extending at -gap_extend; the
//      reference end is the FIRST column whose column maximum reaches the global maximum, the query end the SMALLEST query index
//      holding that maximum in that column (sw_sse2_byte / sw_sse2_word: "if (temp > max)", "if (temp < end_read)");
//   2. begin: the same recurrences on the reversed query prefix, scanning the reference backwards from the end column and stopping
//      at the first column whose maximum equals the score (ssw_align, the "terminate" argument);
//   3. CIGAR: banded_sw over [begin, end] x [begin, end] (band |ref_len - query_len| + 1, doubled until the score is reached) with
//      its direction codes: a diagonal step wins ties against a gap, between gaps "e1 > f1 ? E : F", gap extension vs opening by
//      "temp1 > temp2 ? open : extend"; traceback from the end cell while the query index is positive;
//   4. ssw_cpp's ConvertAlignment / CalculateNumberMismatch: leading / trailing soft clips, M runs split into '=' / 'X';
//   5. the score matrix of ssw_cpp's BuildSwScoreMatrix (1.2.5): N against anything, N included, costs a mismatch (older releases
//      and ssw.c's own example score it 0).  Pinned by the reference's importer golden: the one read of 223 images that tells the two
//      apart (59 N-ridden bases in front of 42 clean ones) keeps its 59S42M there and becomes 4S97M with N scored 0
//      (tests/test_vcf_candidate_importer.py::test_libssw_penalises_n_like_a_mismatch).
// Pinned by the known answers of the reference's ssw tests (deepvariant/realigner/ssw_test.cc:47-58,
// deepvariant/realigner/python/ssw_misc_test.py:44-84, ssw_wrap_test.py:37-72) and of fast_pass_aligner_test.cc
// (tests/test_ssw.py, tests/test_fast_pass_aligner.py).
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "dvb_common.h"

namespace {

inline int8_t Code(char c) {
  switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 4;
  }
}

struct Best { int score, ref, read; };

// One scan of sw_sse2_*: columns = reference positions from `begin` towards `end` (exclusive) in steps of `step`, rows = the
// query.  Returns the score, the column where the running maximum last strictly increased and the smallest row holding it there.
Best Scan(const int8_t* ref, int begin, int end, int step, const int8_t* read, int read_len, const int8_t* mat, int go, int ge,
          int terminate) {
  std::vector<int> h((size_t)read_len + 1, 0), e((size_t)read_len + 1, 0), best_col;
  int max = 0, end_ref = -1;
  for (int j = begin; j != end; j += step) {
    int diag = 0, f = 0, col_max = 0;     // diag = H(i-1, j-1)
    const int8_t* row = mat + 5 * ref[j];
    for (int i = 1; i <= read_len; ++i) {
      // e[i] holds E(i, j): computed from column j-1 at the end of the previous iteration of j
      int hv = diag + row[read[i - 1]];
      if (hv < e[i]) hv = e[i];
      if (hv < f) hv = f;
      if (hv < 0) hv = 0;
      diag = h[i];
      h[i] = hv;
      if (hv > col_max) col_max = hv;
      int e_next = std::max(hv - go, e[i] - ge);
      e[i] = e_next < 0 ? 0 : e_next;
      int f_next = std::max(hv - go, f - ge);
      f = f_next < 0 ? 0 : f_next;
    }
    if (col_max > max) {
      max = col_max;
      end_ref = j;
      best_col.assign(h.begin() + 1, h.end());
    }
    if (terminate >= 0 && col_max == terminate) break;
  }
  Best b{max, end_ref, read_len - 1};
  for (int i = 0; i < (int)best_col.size(); ++i)
    if (best_col[(size_t)i] == max) { b.read = i; break; }
  return b;
}

inline void set_u(int& u, int w, int i, int j) { int x = i - w; x = x > 0 ? x : 0; u = j - x + 1; }
inline void set_d(int& u, int w, int i, int j, int p) { int x = i - w; x = x > 0 ? x : 0; x = j - x; u = x * 3 + p; }

// banded_sw: CIGAR (length, op) pairs in alignment order; false when the traceback leaves the band.
bool BandedSw(const int8_t* ref, const int8_t* read, int ref_len, int read_len, int score, int go, int ge, int band_width,
              const int8_t* mat, std::vector<std::pair<int, char>>* cigar) {
  std::vector<int> h_b, e_b, h_c;
  std::vector<int8_t> direction;
  int width_d = 0, max = 0;
  do {
    const int width = band_width * 2 + 3;
    width_d = band_width * 2 + 1;
    h_b.assign((size_t)width + 1, 0);
    e_b.assign((size_t)width + 1, 0);
    h_c.assign((size_t)width + 1, 0);
    direction.assign((size_t)width_d * read_len * 3 + 3, 0);
    max = 0;
    for (int i = 0; i < read_len; ++i) {
      int beg = 0, end = ref_len - 1, u = 0;
      int j = i - band_width;
      beg = beg > j ? beg : j;
      j = i + band_width;
      end = end < j ? end : j;
      const int edge = end + 1 < width - 1 ? end + 1 : width - 1;
      int f = 0;
      h_b[0] = e_b[0] = h_b[(size_t)edge] = e_b[(size_t)edge] = h_c[0] = 0;
      int8_t* direction_line = direction.data() + (size_t)width_d * i * 3;
      for (j = beg; j <= end; ++j) {
        int b, e, d, de, df, dh;
        set_u(u, band_width, i, j);
        set_u(e, band_width, i - 1, j);
        set_u(b, band_width, i, j - 1);
        set_u(d, band_width, i - 1, j - 1);
        set_d(de, band_width, i, j, 0);
        set_d(df, band_width, i, j, 1);
        set_d(dh, band_width, i, j, 2);
        int temp1 = i == 0 ? -go : h_b[(size_t)e] - go;
        int temp2 = i == 0 ? -ge : e_b[(size_t)e] - ge;
        e_b[(size_t)u] = temp1 > temp2 ? temp1 : temp2;
        direction_line[de] = temp1 > temp2 ? 3 : 2;
        temp1 = h_c[(size_t)b] - go;
        temp2 = f - ge;
        f = temp1 > temp2 ? temp1 : temp2;
        direction_line[df] = temp1 > temp2 ? 5 : 4;
        const int e1 = e_b[(size_t)u] > 0 ? e_b[(size_t)u] : 0;
        const int f1 = f > 0 ? f : 0;
        temp1 = e1 > f1 ? e1 : f1;
        temp2 = h_b[(size_t)d] + mat[ref[j] * 5 + read[i]];
        h_c[(size_t)u] = temp1 > temp2 ? temp1 : temp2;
        if (h_c[(size_t)u] > max) max = h_c[(size_t)u];
        if (temp1 <= temp2) direction_line[dh] = 1;
        else direction_line[dh] = e1 > f1 ? direction_line[de] : direction_line[df];
      }
      for (j = 1; j <= u; ++j) h_b[(size_t)j] = h_c[(size_t)j];
    }
    band_width *= 2;
  } while (max < score && band_width < 2 * (ref_len + read_len + 2));
  if (max < score) return false;
  band_width /= 2;
  // trace back
  int i = read_len - 1, j = ref_len - 1, e = 0, temp2 = 2;
  char op = 'M', prev_op = 'M';
  const int8_t* direction_line = direction.data() + (size_t)width_d * i * 3;
  std::vector<std::pair<int, char>> rev;
  while (i > 0) {
    int temp1;
    set_d(temp1, band_width, i, j, temp2);
    if (temp1 < 0 || temp1 >= width_d * 3) return false;
    switch (direction_line[temp1]) {
      case 1: --i; --j; temp2 = 2; direction_line -= width_d * 3; op = 'M'; break;
      case 2: --i; temp2 = 0; direction_line -= width_d * 3; op = 'I'; break;
      case 3: --i; temp2 = 2; direction_line -= width_d * 3; op = 'I'; break;
      case 4: --j; temp2 = 1; op = 'D'; break;
      case 5: --j; temp2 = 2; op = 'D'; break;
      default: return false;
    }
    if (j < -1) return false;
    if (op == prev_op) ++e;
    else {
      rev.emplace_back(e, prev_op);
      prev_op = op;
      e = 1;
    }
  }
  if (op == 'M') rev.emplace_back(e + 1, op);
  else { rev.emplace_back(e, op); rev.emplace_back(1, 'M'); }
  cigar->assign(rev.rbegin(), rev.rend());
  return true;
}

}  // namespace


namespace dvb_ssw_internal {

// Second half of the alignment: out->{sw_score, ref_begin, ref_end, query_begin, query_end} are known (from the two scans - host
// Scan() above or the CUDA wavefront kernel in dvb_ssw_gpu.cu); banded traceback inside that window, then ssw_cpp's
// ConvertAlignment + CalculateNumberMismatch (soft clips, '=' / 'X' runs).  r / q = base codes 0..4.
int FinishFromEnds(const int8_t* r, const int8_t* q, int query_len, const int8_t* mat, int gap_open, int gap_extend, DvbSswAlignment* out,
                   char* cigar_out, int64_t cigar_cap) {
  const int sub_ref = out->ref_end - out->ref_begin + 1, sub_read = out->query_end - out->query_begin + 1;
  std::vector<std::pair<int, char>> cigar;
  if (!BandedSw(r + out->ref_begin, q + out->query_begin, sub_ref, sub_read, out->sw_score, gap_open, gap_extend,
                std::abs(sub_ref - sub_read) + 1, mat, &cigar))
    return dvb::fail(DVB_ERR_INTERNAL, "dvb_ssw_align: banded traceback failed");
  std::string s;
  if (out->query_begin > 0) s += std::to_string(out->query_begin) + "S";
  const int8_t* rp = r + out->ref_begin;
  const int8_t* qp = q + out->query_begin;
  int mism = 0, len_m = 0, len_x = 0;
  auto flush = [&]() {
    if (len_m) s += std::to_string(len_m) + "=";
    if (len_x) s += std::to_string(len_x) + "X";
    len_m = len_x = 0;
  };
  for (const auto& c : cigar) {
    if (c.second == 'M') {
      for (int k = 0; k < c.first; ++k, ++rp, ++qp) {
        if (*rp != *qp) { ++mism; if (len_m) flush(); ++len_x; }
        else { if (len_x) flush(); ++len_m; }
      }
    } else if (c.second == 'I') {
      qp += c.first; mism += c.first; flush(); s += std::to_string(c.first) + "I";
    } else {
      rp += c.first; mism += c.first; flush(); s += std::to_string(c.first) + "D";
    }
  }
  flush();
  const int tail = query_len - out->query_end - 1;
  if (tail > 0) s += std::to_string(tail) + "S";
  out->mismatches = mism;
  out->cigar_len = (int32_t)s.size();
  if (cigar_cap > (int64_t)s.size()) memcpy(cigar_out, s.c_str(), s.size() + 1);
  return DVB_OK;
}

int8_t BaseCode(char c) { return Code(c); }

}  // namespace dvb_ssw_internal

extern "C" {

int dvb_ssw_align(const char* ref, int64_t ref_len, const char* query, int64_t query_len, int32_t match, int32_t mismatch, int32_t gap_open,
                  int32_t gap_extend, DvbSswAlignment* out, char* cigar_out, int64_t cigar_cap) {
  if (!ref || !query || !out || (cigar_cap > 0 && !cigar_out)) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_ssw_align: null argument");
  memset(out, 0, sizeof(*out));
  if (cigar_cap > 0) cigar_out[0] = 0;
  if (ref_len <= 0 || query_len <= 0) return DVB_OK;           // ssw_cpp's Align returns false: empty alignment, score 0
  int8_t mat[25];
  for (int i = 0; i < 5; ++i)
    for (int j = 0; j < 5; ++j) mat[i * 5 + j] = (i == 4 || j == 4) ? (int8_t)-mismatch : (i == j ? (int8_t)match : (int8_t)-mismatch);
  std::vector<int8_t> r((size_t)ref_len), q((size_t)query_len);
  for (int64_t i = 0; i < ref_len; ++i) r[(size_t)i] = Code(ref[i]);
  for (int64_t i = 0; i < query_len; ++i) q[(size_t)i] = Code(query[i]);
  const Best fwd = Scan(r.data(), 0, (int)ref_len, 1, q.data(), (int)query_len, mat, gap_open, gap_extend, -1);
  out->sw_score = fwd.score;
  if (fwd.score <= 0 || fwd.ref < 0) return DVB_OK;
  out->ref_end = fwd.ref;
  out->query_end = fwd.read;
  std::vector<int8_t> rq(q.begin(), q.begin() + fwd.read + 1);
  std::reverse(rq.begin(), rq.end());
  const Best rev = Scan(r.data(), fwd.ref, -1, -1, rq.data(), fwd.read + 1, mat, gap_open, gap_extend, fwd.score);
  out->ref_begin = rev.ref;
  out->query_begin = fwd.read - rev.read;
  return dvb_ssw_internal::FinishFromEnds(r.data(), q.data(), (int)query_len, mat, gap_open, gap_extend, out, cigar_out, cigar_cap);
}

}  // extern "C"

// ---- FastPassAligner::FastAlignReadsToHaplotypes (deepvariant/realigner/fast_pass_aligner.cc:165-279): the exact k-mer pass -------------
// For every haplotype: each k-mer of the haplotype is looked up in the index of read k-mers; a read whose k-mer matches is compared
// base by base at the implied offset and accepted with at most max_mismatches mismatches (N matches anything); the haplotype's score
// is the sum of its reads' best scores, or 0 as soon as an indexed k-mer position inside [prefix, len - suffix) is covered by no
// accepted read (never for the reference haplotype).  Outputs, per haplotype h and read r: position[h * n_reads + r] (65535 = not
// aligned) and score[h * n_reads + r]; hap_score[h].  The CIGAR of an accepted read is always "<len>=".
#include <unordered_map>

extern "C" int dvb_fast_pass_scores(const char* reference, int64_t ref_len, const char* const* haplotypes, const int64_t* hap_lens, int32_t n_haps,
                                    const char* const* reads, const int64_t* read_lens, int32_t n_reads, int32_t kmer_size,
                                    int32_t max_mismatches, int32_t match, int32_t mismatch, int32_t ref_prefix_len, int32_t ref_suffix_len,
                                    int32_t* hap_score, int32_t* position, int32_t* score) {
  if (!reference || !haplotypes || !hap_lens || !reads || !read_lens || !hap_score || !position || !score || kmer_size < 1)
    return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_fast_pass_scores: bad argument");
  const size_t k = (size_t)kmer_size;
  std::unordered_map<std::string, std::vector<std::pair<int32_t, int32_t>>> index;     // k-mer -> (read, offset), in insertion order
  for (int32_t r = 0; r < n_reads; ++r) {
    const size_t n = (size_t)read_lens[r];
    if (n <= k) continue;
    for (size_t i = 0; i + k <= n; ++i) index[std::string(reads[r] + i, k)].emplace_back(r, (int32_t)i);
  }
  std::string kmer;
  for (int32_t h = 0; h < n_haps; ++h) {
    const char* hap = haplotypes[h];
    const int64_t hl = hap_lens[h];
    int32_t* pos = position + (size_t)h * n_reads;
    int32_t* sc = score + (size_t)h * n_reads;
    for (int32_t r = 0; r < n_reads; ++r) { pos[r] = 65535; sc[r] = 0; }
    const bool is_ref = hl == ref_len && memcmp(hap, reference, (size_t)hl) == 0;
    std::vector<int32_t> coverage((size_t)std::max<int64_t>(hl, 1), 0);
    int64_t total = 0;
    bool zeroed = false;
    for (int64_t i = 0; i + (int64_t)k <= hl && !zeroed; ++i) {
      kmer.assign(hap + i, k);
      auto it = index.find(kmer);
      if (it == index.end()) continue;
      for (const auto& occ : it->second) {
        const int32_t r = occ.first;
        const int64_t start = std::max<int64_t>(0, i - occ.second), n = read_lens[r];
        if (start + n > hl) continue;
        if (pos[r] != 65535 && pos[r] == start) continue;
        int matches = 0, mism = 0;
        bool rejected = false;
        const char* a = hap + start;
        const char* b = reads[r];
        for (int64_t j = 0; j < n; ++j) {
          if (a[j] != b[j] && a[j] != 'N' && b[j] != 'N') {
            if (++mism == max_mismatches + 1) { rejected = true; break; }
          } else {
            ++matches;
          }
        }
        if (rejected || mism > max_mismatches) continue;
        const int32_t new_score = matches * match - mism * mismatch;
        for (int64_t p = start; p < start + n; ++p) ++coverage[(size_t)p];
        if (sc[r] < new_score) {
          total += new_score - sc[r];
          sc[r] = new_score;
          pos[r] = (int32_t)start;
        }
      }
      if (coverage[(size_t)i] == 0 && i >= ref_prefix_len && i < hl - ref_suffix_len && !is_ref) zeroed = true;
    }
    if (zeroed) total = 0;
    hap_score[h] = (int32_t)total;
    if (total == 0)
      for (int32_t r = 0; r < n_reads; ++r) { pos[r] = 65535; sc[r] = 0; }
  }
  return DVB_OK;
}

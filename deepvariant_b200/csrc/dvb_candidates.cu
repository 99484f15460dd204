// Candidate generation on the host (SURVEY.md 8(f) "next" row #2, first half): allele counting over the rows of the native
// BAM table + the very-sensitive variant caller, producing the DeepVariantCall protos the pileup path consumes.
//
// Restates (without copying; the reference keeps protobuf maps of strings, this keeps flat per-position entry lists keyed by
// table row):
//   AlleleCounter::Add / MakeIndelReadAllele / AddReadAlleles / CanBasesBeUsed / GetAvgBaseQuality / GetPrevBase / RefBases
//       deepvariant/allelecounter.cc:195-243, 360-519, 880-978
//   SumAlleleCounts / TotalAlleleCounts                         deepvariant/allelecounter.cc:78-193
//   multi_sample::VariantCaller::SelectAltAlleles / AlleleFilter / IsGoodAltAlleleWithReason / CalcRefBases / BuildAlleleMap /
//       AddReadDepths / CallVariant / CallVariantPosition / AddSupportingReads / AddAdjacentAlleleFractionsAtPosition
//       deepvariant/variant_calling_multisample.cc:95-290, 610-760, 1000-1330 (single sample, create_complex_alleles = false,
//       use_rejected_alleles = false, no methylation: the make_examples defaults, make_examples_options.py:597-922)
// Read normalisation (--normalize_reads) happens before this counter, on the reads (deepvariant_b200/normalize_reads.py).
// Not restated here: complex alleles, rejected alleles, methylation, multi-sample
// filters, gVCF summaries.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "dvb_allele_walk.h"
#include "dvb_common.h"

namespace {

using dvb_allele::Canonical;
using dvb_allele::Element;
using dvb_allele::kDeletion;
using dvb_allele::kInsertion;
using dvb_allele::kReference;
using dvb_allele::kSoftClip;
using dvb_allele::kSubstitution;

struct Entry {              // one value of AlleleCount.read_alleles (the key is the read: key_id)
  int32_t row;              // BAM table row
  int32_t key_id;           // id of "fragment_name/read_number" among the region's reads (equal ids = equal map keys)
  uint8_t type, low_quality, mapq, reverse;
  int32_t avg_base_quality;
  uint32_t bases_off, bases_len;   // into Counter::arena
};

struct Site {
  int32_t ref_supporting_read_count = 0;
  std::vector<Entry> entries;        // insertion order; a later entry with the same key replaces the earlier one in place
};

struct Counter {
  const DvbReadTable* t;
  const uint8_t* contig;         // upper-case bases of the whole contig
  int64_t contig_len;
  int64_t start, end;            // interval (== reads interval: no normalisation)
  DvbCandidateOptions opt;
  std::vector<Site> sites;
  std::string arena;             // allele bases
  std::vector<int32_t> candidate_positions;   // relative to start, sorted (track_ref_reads second pass)

  // AddReadAlleles (allelecounter.cc:475-546) for one surviving element of dvb_allele::WalkRead.
  struct Sink {
    Counter* c;
    int32_t row, key_id;
    const uint8_t* seq;
    void Commit(const Element& a) {
      Site& site = c->sites[(size_t)a.position];
      if (a.type == kReference && !a.low_quality) ++site.ref_supporting_read_count;
      if (a.type == kReference &&
          !(c->opt.track_ref_reads && std::binary_search(c->candidate_positions.begin(), c->candidate_positions.end(), a.position)))
        return;
      Entry e;
      e.row = row;
      e.key_id = key_id;
      e.type = a.type;
      e.low_quality = a.low_quality;
      e.mapq = c->t->mapq[row];
      e.reverse = (c->t->flag[row] & 0x10) != 0;
      e.avg_base_quality = a.avg_base_quality;
      e.bases_off = (uint32_t)c->arena.size();
      if (a.len == 0) {
        c->arena.push_back((char)seq[a.read_offset]);
        e.bases_len = 1;
      } else {
        c->arena.push_back((char)a.prev);
        if (a.type == kDeletion) c->arena.append((const char*)c->contig + a.ref_abs, (size_t)a.len);
        else c->arena.append((const char*)seq + a.read_offset, (size_t)a.len);
        e.bases_len = (uint32_t)a.len + 1;
      }
      for (Entry& old : site.entries)
        if (old.key_id == key_id) { old = e; return; }     // (*read_alleles)[key] = allele
      site.entries.push_back(e);
    }
  };

  dvb_allele::WalkParams Params() const {
    dvb_allele::WalkParams p;
    p.start = start;
    p.end = end;
    p.contig = contig;
    p.contig_origin = 0;
    p.contig_avail = contig_len;
    p.contig_len = contig_len;
    p.min_base_quality = opt.min_base_quality;
    p.keep_legacy_behavior = opt.keep_legacy_behavior;
    return p;
  }

  // AlleleCounter::Add (allelecounter.cc:880-978).
  void Add(int32_t row, int32_t key_id) {
    if (t->mapq[row] < opt.min_mapping_quality) return;
    const int64_t s0 = t->seq_begin[row];
    dvb_allele::ReadView r;
    r.seq = t->bases + s0;
    r.qual = t->quals + s0;
    r.seq_len = (int)(t->seq_begin[row + 1] - s0);
    r.cigar = t->cigar + t->cigar_begin[row];
    r.n_cigar = (int)(t->cigar_begin[row + 1] - t->cigar_begin[row]);
    r.pos = t->pos[row];
    Sink sink{this, row, key_id, r.seq};
    dvb_allele::WalkRead(r, Params(), sink);
  }
};

struct SummedAllele { uint32_t off, len; uint8_t type; int count; };

struct Caller {
  const Counter* c;
  const DvbCandidateOptions* opt;

  std::string Bases(uint32_t off, uint32_t len) const { return c->arena.substr(off, len); }

  // SumAlleleCounts(allele_count) (allelecounter.cc:78-116): non-low-quality entries grouped by (bases, type), in the
  // order of std::map<pair<string_view, AlleleType>>; the synthetic REFERENCE allele is irrelevant to selection.
  std::vector<SummedAllele> Sum(const Site& s) const {
    std::vector<SummedAllele> out;        // a handful of distinct alleles per site: linear grouping, then one sort
    const char* a = c->arena.data();
    for (const Entry& e : s.entries) {
      if (e.low_quality) continue;
      bool found = false;
      for (SummedAllele& o : out)
        if (o.type == e.type && o.len == e.bases_len && memcmp(a + o.off, a + e.bases_off, e.bases_len) == 0) { ++o.count; found = true; break; }
      if (!found) out.push_back(SummedAllele{e.bases_off, e.bases_len, e.type, 1});
    }
    std::sort(out.begin(), out.end(), [a](const SummedAllele& x, const SummedAllele& y) {   // (bases, type) as std::map orders them
      const int c = memcmp(a + x.off, a + y.off, std::min(x.len, y.len));
      if (c != 0) return c < 0;
      if (x.len != y.len) return x.len < y.len;
      return x.type < y.type;
    });
    return out;
  }

  // TotalAlleleCounts (allelecounter.cc:157-169).
  int Total(const Site& s) const {
    int n = s.ref_supporting_read_count;
    for (const Entry& e : s.entries)
      if (!e.low_quality && e.type != kReference) ++n;
    return n;
  }

  int MinCount(const SummedAllele& a) const { return a.type == kSubstitution ? opt->min_count_snps : opt->min_count_indels; }
  double MinFraction(const SummedAllele& a) const {      // variant_calling_multisample.h:357-372 (proto fields are float)
    if (a.type == kSubstitution) return (double)opt->min_fraction_snps;
    if (opt->vsc_small_indel_threshold > 0 && opt->vsc_min_indel_fraction_for_small_indels > 0.0f &&
        opt->vsc_min_indel_fraction_for_large_indels > 0.0f) {
      return (int)a.len <= opt->vsc_small_indel_threshold + 1 ? (double)opt->vsc_min_indel_fraction_for_small_indels
                                                              : (double)opt->vsc_min_indel_fraction_for_large_indels;
    }
    return (double)opt->min_fraction_indels;
  }
  // IsGoodAltAlleleWithReason (variant_calling_multisample.cc:175-196): 0 accepted, 1 ref, 2 low support, 3 other, 4 low ratio
  int Reason(const SummedAllele& a, int total, bool trio) const {
    if (a.type == kReference) return 1;
    if (a.count < MinCount(a)) return 2;
    if (a.type == kSoftClip) return 3;
    if ((1.0 * a.count) / total < MinFraction(a) * (trio ? (double)opt->min_fraction_multiplier : 1.0)) return 4;
    return 0;
  }
  // AlleleFilter for one sample (:203-260): a low-ratio / low-support allele is re-tested against the all-sample counts
  // (the same counts here) with the trio coefficient.
  bool Keep(const SummedAllele& a, int total) const {
    int r = Reason(a, total, false);
    if (r == 0) return true;
    if (r == 4 || r == 2) return Reason(a, total, true) == 0;
    return false;
  }
  std::vector<SummedAllele> SelectAlts(const Site& s) const {
    std::vector<SummedAllele> alts;
    const int total = Total(s);
    for (const SummedAllele& a : Sum(s))
      if (Keep(a, total)) alts.push_back(a);
    return alts;
  }
};

// ---- protobuf wire helpers ---------------------------------------------------------------------------------------------
void PutVarint(std::string* o, uint64_t v) {
  while (v >= 0x80) { o->push_back((char)(v | 0x80)); v >>= 7; }
  o->push_back((char)v);
}
void PutTag(std::string* o, int field, int wt) { PutVarint(o, ((uint64_t)field << 3) | (uint64_t)wt); }
void PutBytes(std::string* o, int field, const std::string& b) { PutTag(o, field, 2); PutVarint(o, b.size()); o->append(b); }
void PutInt(std::string* o, int field, int64_t v) { PutTag(o, field, 0); PutVarint(o, (uint64_t)v); }
void PutDouble(std::string* o, int field, double d) { PutTag(o, field, 1); o->append((const char*)&d, 8); }

std::string InfoEntryInts(const char* key, const std::vector<int>& vals) {    // map<string, ListValue>; Value.int_value = 7
  std::string lv;
  for (int v : vals) { std::string val; PutInt(&val, 7, v); PutBytes(&lv, 1, val); }
  std::string e;
  PutBytes(&e, 1, key);
  PutBytes(&e, 2, lv);
  return e;
}
std::string InfoEntryDoubles(const char* key, const std::vector<double>& vals) {   // Value.number_value = 2
  std::string lv;
  for (double v : vals) { std::string val; PutDouble(&val, 2, v); PutBytes(&lv, 1, val); }
  std::string e;
  PutBytes(&e, 1, key);
  PutBytes(&e, 2, lv);
  return e;
}

}  // namespace

struct DvbCandidates {
  std::string protos;                 // serialized DeepVariantCall records, concatenated
  std::vector<int64_t> proto_begin;   // [n + 1]
  std::vector<int32_t> position;      // [n] variant.start
  std::vector<int32_t> positions_only;   // dvb_candidate_positions result
  std::vector<int32_t> summary;          // [2 n_sites]: ref_supporting_read_count, total_read_count (AlleleCounter::SummaryCounts)
  int64_t n_reads_counted = 0;
};

namespace {

std::string ReadKey(const DvbReadTable& t, int32_t row) {      // AlleleCounter::ReadKey (allelecounter.cc:980-983)
  std::string k(t.names + t.name_begin[row], (size_t)(t.name_begin[row + 1] - t.name_begin[row]));
  k.push_back('/');
  k += std::to_string((int)t.read_number[row]);
  return k;
}

int BuildCounter(const DvbBam* bam, DvbReadTable* table, Counter* c, const uint8_t* contig_bases, int64_t contig_n_bases,
                 int64_t start, int64_t end, const int64_t* rows, int64_t n_rows, const DvbCandidateOptions* opt,
                 const int32_t* candidate_positions, int32_t n_candidate_positions, std::vector<std::string>* keys) {
  if (!bam || !contig_bases || !opt || (n_rows && !rows))
    return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_candidates: null argument");
  if (start < 0 || end > contig_n_bases || start >= end)
    return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_candidates: interval [%lld, %lld) is not inside the contig (%lld bases)",
                     (long long)start, (long long)end, (long long)contig_n_bases);
  int st = dvb_bam_table(bam, table);
  if (st != DVB_OK) return st;
  c->t = table;
  c->contig = contig_bases;
  c->contig_len = contig_n_bases;
  c->start = start;
  c->end = end;
  c->opt = *opt;
  c->sites.assign((size_t)(end - start), Site());
  for (int32_t i = 0; i < n_candidate_positions; ++i) c->candidate_positions.push_back(candidate_positions[i] - (int32_t)start);
  std::sort(c->candidate_positions.begin(), c->candidate_positions.end());
  std::unordered_map<std::string, int32_t> key_ids;
  keys->clear();
  for (int64_t i = 0; i < n_rows; ++i) {
    const int64_t row = rows[i];
    if (row < 0 || row >= table->n_reads) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_candidates: read row %lld out of range", (long long)row);
    std::string key = ReadKey(*table, (int32_t)row);
    auto it = key_ids.find(key);
    int32_t id;
    if (it == key_ids.end()) {
      id = (int32_t)keys->size();
      key_ids.emplace(key, id);
      keys->push_back(std::move(key));
    } else {
      id = it->second;
    }
    c->Add((int32_t)row, id);
  }
  return DVB_OK;
}

}  // namespace

extern "C" {

void dvb_candidate_options_default(DvbCandidateOptions* o) {
  memset(o, 0, sizeof(*o));
  o->min_mapping_quality = 5;      // make_examples_options.py:303-310
  o->min_base_quality = 10;        // :293-302
  o->min_count_snps = 2;           // :311-326
  o->min_count_indels = 2;
  o->min_fraction_snps = 0.12f;    // :327-343
  o->min_fraction_indels = 0.06f;
  o->min_fraction_multiplier = 1.0f;
  o->sample_name = "";
}

int dvb_candidate_positions(const DvbBam* bam, const uint8_t* contig_bases, int64_t contig_n_bases, int64_t start, int64_t end,
                            const int64_t* rows, int64_t n_rows, const DvbCandidateOptions* opt, DvbCandidates** out) {
  if (!out) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_candidate_positions: null out");
  *out = nullptr;
  DvbReadTable table;
  Counter c;
  std::vector<std::string> keys;
  int st = BuildCounter(bam, &table, &c, contig_bases, contig_n_bases, start, end, rows, n_rows, opt, nullptr, 0, &keys);
  if (st != DVB_OK) return st;
  Caller caller{&c, &c.opt};
  auto* res = new DvbCandidates();
  for (size_t i = 0; i < c.sites.size(); ++i) {
    if (c.sites[i].entries.empty()) continue;                              // no read allele, no alt allele
    if (!Canonical((char)contig_bases[start + (int64_t)i])) continue;     // CallVariantPosition (:1075-1115)
    if (!caller.SelectAlts(c.sites[i]).empty()) res->positions_only.push_back((int32_t)(start + (int64_t)i));
  }
  *out = res;
  return DVB_OK;
}

int dvb_candidates_in_region(const DvbBam* bam, const char* reference_name, const uint8_t* contig_bases, int64_t contig_n_bases,
                             int64_t start, int64_t end, const int64_t* rows, int64_t n_rows, const DvbCandidateOptions* opt,
                             const int32_t* candidate_positions, int32_t n_candidate_positions, DvbCandidates** out) {
  return dvb_candidates_at_positions(bam, reference_name, contig_bases, contig_n_bases, start, end, rows, n_rows, opt, candidate_positions,
                                     n_candidate_positions, nullptr, -1, out);
}

int dvb_candidates_at_positions(const DvbBam* bam, const char* reference_name, const uint8_t* contig_bases, int64_t contig_n_bases,
                                int64_t start, int64_t end, const int64_t* rows, int64_t n_rows, const DvbCandidateOptions* opt,
                                const int32_t* candidate_positions, int32_t n_candidate_positions, const int32_t* emit_positions,
                                int32_t n_emit_positions, DvbCandidates** out) {
  if (!out || !reference_name) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_candidates_in_region: null argument");
  if (n_emit_positions > 0 && !emit_positions) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_candidates_at_positions: null emit_positions");
  *out = nullptr;
  DvbReadTable table;
  Counter c;
  std::vector<std::string> keys;
  int st = BuildCounter(bam, &table, &c, contig_bases, contig_n_bases, start, end, rows, n_rows, opt, candidate_positions,
                        n_candidate_positions, &keys);
  if (st != DVB_OK) return st;
  Caller caller{&c, &c.opt};
  auto* res = new DvbCandidates();
  res->proto_begin.push_back(0);
  const std::string sample = opt->sample_name ? opt->sample_name : "";
  const int n_sites = (int)c.sites.size();
  res->summary.resize(2 * (size_t)n_sites);          // AlleleCounter::SummaryCounts (allelecounter.cc:986-1007), for the gVCF records
  for (int i = 0; i < n_sites; ++i) {
    res->summary[2 * (size_t)i] = c.sites[(size_t)i].ref_supporting_read_count;
    res->summary[2 * (size_t)i + 1] = caller.Total(c.sites[(size_t)i]);
  }
  for (int i = 0; i < n_sites; ++i) {
    const Site& site = c.sites[(size_t)i];
    if (site.entries.empty()) continue;                                     // no read allele, no alt allele
    const char ref_base = (char)contig_bases[start + i];
    if (n_emit_positions >= 0 && !std::binary_search(emit_positions, emit_positions + n_emit_positions, (int32_t)(start + i))) continue;
    if (!Canonical(ref_base)) continue;                                      // CallVariant (:1117-1130)
    std::vector<SummedAllele> alts = caller.SelectAlts(site);
    if (alts.empty()) continue;                                              // fraction_reference_sites_to_emit = 0
    // CalcRefBases (:95-128): the longest selected deletion extends the reference allele.
    std::string ref_bases(1, ref_base);
    {
      int best = -1;
      const SummedAllele* del = nullptr;
      for (const SummedAllele& a : alts) {
        int sz = a.type == kDeletion ? (int)a.len : -1;
        if (sz > best) { best = sz; del = a.type == kDeletion ? &a : nullptr; }
      }
      if (del) ref_bases += c.arena.substr(del->off + 1, del->len - 1);
    }
    // BuildAlleleMap (:560-606), ordered like AlleleMap = std::map<Allele, string, OrderAllele> (type, then bases).
    struct MapItem { SummedAllele a; std::string bases; std::string alt; };
    std::vector<MapItem> amap;
    for (const SummedAllele& a : alts) {
      MapItem m{a, c.arena.substr(a.off, a.len), ""};
      if (a.type == kSubstitution) {
        m.alt = (m.bases.size() > 1 && ref_bases.size() > 1) ? m.bases : m.bases + ref_bases.substr(1);
      } else if (a.type == kInsertion) {
        m.alt = m.bases + ref_bases.substr(1);
      } else if (a.type == kDeletion) {
        m.alt = m.bases.substr(0, 1) + (m.bases.size() >= ref_bases.size() ? std::string() : ref_bases.substr(m.bases.size()));
      } else {
        continue;     // SOFT_CLIP never passes the filter
      }
      amap.push_back(std::move(m));
    }
    std::sort(amap.begin(), amap.end(), [](const MapItem& x, const MapItem& y) {
      return x.a.type != y.a.type ? x.a.type < y.a.type : x.bases < y.bases;
    });
    std::vector<const MapItem*> by_alt;
    for (const MapItem& m : amap) by_alt.push_back(&m);
    std::sort(by_alt.begin(), by_alt.end(), [](const MapItem* x, const MapItem* y) { return x->alt < y->alt; });
    for (size_t k = 1; k < by_alt.size(); ++k)
      if (by_alt[k]->alt == by_alt[k - 1]->alt) {
        delete res;
        return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_candidates_in_region: non-unique alternative alleles at %lld",
                         (long long)(start + i));   // the reference CHECK-fails (AddReadDepths)
      }
    // AddReadDepths (:608-640)
    const int dp = caller.Total(site);
    std::vector<int> ad{site.ref_supporting_read_count};
    std::vector<double> vaf;
    for (const MapItem* m : by_alt) { ad.push_back(m->a.count); vaf.push_back(1.0 * m->a.count / dp); }

    // ---- Variant (variants.proto:52-118) ----
    std::string call;     // VariantCall: info = 2, genotype = 7 (packed), call_set_name = 9
    PutBytes(&call, 2, InfoEntryInts("AD", ad));
    PutBytes(&call, 2, InfoEntryInts("DP", std::vector<int>{dp}));
    PutBytes(&call, 2, InfoEntryDoubles("VAF", vaf));
    {
      std::string gt;
      PutVarint(&gt, (uint64_t)(int64_t)-1);
      PutVarint(&gt, (uint64_t)(int64_t)-1);
      PutBytes(&call, 7, gt);
    }
    if (!sample.empty()) PutBytes(&call, 9, sample);
    std::string variant;
    PutBytes(&variant, 6, ref_bases);
    for (const MapItem* m : by_alt) PutBytes(&variant, 7, m->alt);
    PutBytes(&variant, 11, call);
    PutInt(&variant, 13, start + i + (int64_t)ref_bases.size());
    PutBytes(&variant, 14, reference_name);
    if (start + i) PutInt(&variant, 16, start + i);

    // ---- AddSupportingReads (:1235-1290): every read_alleles entry, low quality and soft clips included ----
    std::string dvc;
    PutBytes(&dvc, 1, variant);
    std::vector<std::string> support_key;                       // first-seen order of supported alleles
    std::vector<std::string> support_names, support_ext;        // per supported allele: SupportingReads / SupportingReadsExt
    std::string ref_names, ref_ext;
    auto read_support = [&](const Entry& e) {
      std::string rs;
      PutBytes(&rs, 1, keys[(size_t)e.key_id]);
      if (e.low_quality) PutInt(&rs, 2, 1);
      if (e.mapq) PutInt(&rs, 3, e.mapq);
      if (e.avg_base_quality) PutInt(&rs, 4, e.avg_base_quality);
      if (e.reverse) PutInt(&rs, 5, 1);
      if (!sample.empty()) PutBytes(&rs, 7, sample);
      return rs;
    };
    for (const Entry& e : site.entries) {
      if (e.type != kReference) {
        std::string supported = "UNCALLED_ALLELE";
        for (const MapItem& m : amap)
          if (m.a.type == e.type && m.a.len == e.bases_len && c.arena.compare(e.bases_off, e.bases_len, m.bases) == 0) { supported = m.alt; break; }
        size_t k = 0;
        while (k < support_key.size() && support_key[k] != supported) ++k;
        if (k == support_key.size()) { support_key.push_back(supported); support_names.emplace_back(); support_ext.emplace_back(); }
        PutBytes(&support_names[k], 1, keys[(size_t)e.key_id]);
        PutBytes(&support_ext[k], 1, read_support(e));
      } else {
        PutBytes(&ref_names, 4, keys[(size_t)e.key_id]);          // DeepVariantCall.ref_support = 4
        PutBytes(&ref_ext, 1, read_support(e));
      }
    }
    for (size_t k = 0; k < support_key.size(); ++k) {
      std::string entry;
      PutBytes(&entry, 1, support_key[k]);
      PutBytes(&entry, 2, support_names[k]);
      PutBytes(&dvc, 2, entry);                                   // allele_support
    }
    dvc += ref_names;
    for (size_t k = 0; k < support_key.size(); ++k) {
      std::string entry;
      PutBytes(&entry, 1, support_key[k]);
      PutBytes(&entry, 2, support_ext[k]);
      PutBytes(&dvc, 5, entry);                                   // allele_support_ext
    }
    if (!ref_ext.empty()) PutBytes(&dvc, 6, ref_ext);             // ref_support_ext
    // AddAdjacentAlleleFractionsAtPosition (:1330-1360)
    if (opt->small_model_vaf_context_window_size > 0) {
      const int half = opt->small_model_vaf_context_window_size / 2;
      const int lo = i - std::min(i, half), hi = i + std::min(n_sites - i, half + 1);
      for (int j = lo; j < hi; ++j) {
        const Site& s = c.sites[(size_t)j];
        const int n_alleles = (int)s.entries.size();
        const int depth = s.ref_supporting_read_count + n_alleles;
        const int v = depth > 0 ? (100 * n_alleles) / depth : 0;
        std::string entry;
        PutInt(&entry, 1, start + j);
        PutInt(&entry, 2, v);
        PutBytes(&dvc, 7, entry);
      }
    }
    res->protos += dvc;
    res->proto_begin.push_back((int64_t)res->protos.size());
    res->position.push_back((int32_t)(start + i));
  }
  *out = res;
  return DVB_OK;
}

// VcfCandidateImporter: candidates PROPOSED by a VCF, with the evidence of this region's reads attached.  Restates
//   VariantCaller::CallsFromVcf / CallsFromVariantsInRegion / ComputeVariant      deepvariant/variant_calling.cc:393-435, 493-541
//   SelectAltAlleles / IsGoodAltAllele (the single-sample rule, no trio re-test)   :226-252
//   CalcRefBases :176-203, MakeVariantConsistentWithRefAndAlts :118-143, BuildAlleleMap :254-291 (a substitution is always
//   bases + ref[1:] here), AddReadDepths :300-345 (alleles matched through SimplifyRefAlt, utils.cc:63-84),
//   AddSupportingReads :675-712 (read name + is_low_quality only; the suffix when the proposed reference allele is longer).
// The caller passes the records that START inside [start, end) (CallsFromVcf keeps variant->start() >= range.start() of the
// records that overlap the range) after the uncalled-genotype filter; allele 0 of each record is its reference allele.
int dvb_candidates_from_proposed(const DvbBam* bam, const char* reference_name, const uint8_t* contig_bases, int64_t contig_n_bases,
                                 int64_t start, int64_t end, const int64_t* rows, int64_t n_rows, const DvbCandidateOptions* opt,
                                 const int32_t* candidate_positions, int32_t n_candidate_positions, int32_t n_proposed,
                                 const int64_t* proposed_start, const int32_t* allele_first, const int64_t* allele_begin,
                                 const char* allele_chars, DvbCandidates** out) {
  if (!out || !reference_name) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_candidates_from_proposed: null argument");
  if (n_proposed < 0 || (n_proposed > 0 && (!proposed_start || !allele_first || !allele_begin || !allele_chars)))
    return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_candidates_from_proposed: null proposed-variant arrays");
  *out = nullptr;
  DvbReadTable table;
  Counter c;
  std::vector<std::string> keys;
  int st = BuildCounter(bam, &table, &c, contig_bases, contig_n_bases, start, end, rows, n_rows, opt, candidate_positions,
                        n_candidate_positions, &keys);
  if (st != DVB_OK) return st;
  Caller caller{&c, &c.opt};
  std::unique_ptr<DvbCandidates> res(new DvbCandidates());
  res->proto_begin.push_back(0);
  const std::string sample = opt->sample_name ? opt->sample_name : "";
  const int n_sites = (int)c.sites.size();
  res->summary.resize(2 * (size_t)n_sites);
  for (int i = 0; i < n_sites; ++i) {
    res->summary[2 * (size_t)i] = c.sites[(size_t)i].ref_supporting_read_count;
    res->summary[2 * (size_t)i + 1] = caller.Total(c.sites[(size_t)i]);
  }
  auto simplify = [](const std::string& ref, const std::string& alt) {          // SimplifyRefAlt
    const size_t shortest = std::min(ref.size(), alt.size());
    size_t common = 0;
    for (size_t k = 1; k < shortest; ++k) {
      if (ref[ref.size() - k] != alt[alt.size() - k]) break;
      common = k;
    }
    return ref.substr(0, ref.size() - common) + "->" + alt.substr(0, alt.size() - common);
  };
  const Site empty_site;
  for (int32_t v = 0; v < n_proposed; ++v) {
    const int32_t a0 = allele_first[v], a1 = allele_first[v + 1];
    if (a0 < 0 || a1 <= a0) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_candidates_from_proposed: variant %d has no reference allele", v);
    for (int32_t a = a0; a < a1; ++a)
      if (allele_begin[a] < 0 || allele_begin[a + 1] < allele_begin[a])
        return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_candidates_from_proposed: allele_begin is not ascending at allele %d", a);
    if (allele_begin[a0 + 1] == allele_begin[a0]) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_candidates_from_proposed: variant %d has an empty reference allele", v);
    auto allele = [&](int32_t a) { return std::string(allele_chars + allele_begin[a], (size_t)(allele_begin[a + 1] - allele_begin[a])); };
    std::string var_ref = allele(a0);
    std::vector<std::string> var_alts;
    for (int32_t a = a0 + 1; a < a1; ++a) var_alts.push_back(allele(a));
    const int64_t pos = proposed_start[v];
    // AlleleIndex: no AlleleCount at the position = a call with no evidence (missing genotype downstream)
    const bool inside = pos >= start && pos < end;
    const Site& site = inside ? c.sites[(size_t)(pos - start)] : empty_site;
    std::string ref_base;
    if (inside) {
      ref_base.assign(1, (char)contig_bases[pos]);
      if (!Canonical(ref_base[0])) continue;
    }
    // SelectAltAlleles, single-sample rule
    std::vector<SummedAllele> alts;
    const int total = caller.Total(site);
    for (const SummedAllele& a : caller.Sum(site))
      if (a.type != kReference && a.type != kSoftClip && a.count >= caller.MinCount(a) &&
          (1.0 * a.count) / total >= (a.type == kSubstitution ? (double)opt->min_fraction_snps : (double)opt->min_fraction_indels))
        alts.push_back(a);
    std::string ref_bases = ref_base;               // CalcRefBases
    {
      int best = -1;
      const SummedAllele* del = nullptr;
      for (const SummedAllele& a : alts) {
        const int sz = a.type == kDeletion ? (int)a.len : -1;
        if (sz > best) { best = sz; del = a.type == kDeletion ? &a : nullptr; }
      }
      if (del) ref_bases += c.arena.substr(del->off + 1, del->len - 1);
    }
    int64_t var_end = pos + (int64_t)var_ref.size();
    if (var_ref != ref_bases) {                     // MakeVariantConsistentWithRefAndAlts
      if (var_ref.size() == ref_bases.size() || (var_ref.size() < ref_bases.size() && ref_bases.compare(0, var_ref.size(), var_ref) != 0))
        return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_candidates_from_proposed: proposed variant at %s:%lld has incorrect ref bases (%s, the reads say %s)",
                         reference_name, (long long)(pos + 1), var_ref.c_str(), ref_bases.c_str());      // the reference QCHECK-fails
      if (var_ref.size() < ref_bases.size()) {
        const std::string suffix = ref_bases.substr(var_ref.size());
        var_ref += suffix;
        for (std::string& a : var_alts) a += suffix;
        var_end += (int64_t)suffix.size();
      }
    }
    struct MapItem { SummedAllele a; std::string bases; std::string alt; };
    std::vector<MapItem> amap;                      // BuildAlleleMap
    for (const SummedAllele& a : alts) {
      MapItem m{a, c.arena.substr(a.off, a.len), ""};
      const std::string tail1 = ref_bases.size() > 1 ? ref_bases.substr(1) : std::string();
      if (a.type == kSubstitution || a.type == kInsertion) m.alt = m.bases + tail1;
      else if (a.type == kDeletion) m.alt = m.bases.substr(0, 1) + (m.bases.size() >= ref_bases.size() ? std::string() : ref_bases.substr(m.bases.size()));
      else continue;
      amap.push_back(std::move(m));
    }
    // AddReadDepths
    const int dp = total;
    std::string call;
    const bool only_dp = var_alts.size() == 1 && (var_alts[0] == "." || var_alts[0] == "<*>");
    if (!only_dp) {
      std::vector<std::pair<std::string, const MapItem*>> by_key;
      for (const MapItem& m : amap) by_key.emplace_back(simplify(ref_bases, m.alt), &m);
      for (size_t x = 0; x < by_key.size(); ++x)
        for (size_t y = x + 1; y < by_key.size(); ++y)
          if (by_key[x].first == by_key[y].first)
            return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_candidates_from_proposed: non-unique alternative alleles at %lld", (long long)pos);
      std::vector<int> ad{site.ref_supporting_read_count};
      std::vector<double> vaf;
      for (const std::string& alt : var_alts) {
        const std::string key = simplify(var_ref, alt);
        int count = 0;
        for (const auto& kv : by_key)
          if (kv.first == key) count = kv.second->a.count;
        ad.push_back(count);
        vaf.push_back(dp > 0 ? 1.0 * count / dp : 0.0);
      }
      PutBytes(&call, 2, InfoEntryInts("AD", ad));
      PutBytes(&call, 2, InfoEntryInts("DP", std::vector<int>{dp}));
      PutBytes(&call, 2, InfoEntryDoubles("VAF", vaf));
    } else {
      PutBytes(&call, 2, InfoEntryInts("DP", std::vector<int>{dp}));
    }
    {
      std::string gt;
      PutVarint(&gt, (uint64_t)(int64_t)-1);
      PutVarint(&gt, (uint64_t)(int64_t)-1);
      PutBytes(&call, 7, gt);
    }
    if (!sample.empty()) PutBytes(&call, 9, sample);
    std::string variant;
    PutBytes(&variant, 6, var_ref);
    for (const std::string& a : var_alts) PutBytes(&variant, 7, a);
    PutBytes(&variant, 11, call);
    PutInt(&variant, 13, var_end);
    PutBytes(&variant, 14, reference_name);
    if (pos) PutInt(&variant, 16, pos);
    // AddSupportingReads
    std::string suffix;
    if (var_ref.size() > ref_bases.size()) {
      if (var_ref.compare(0, ref_bases.size(), ref_bases) != 0)
        return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_candidates_from_proposed: %s has to be a prefix of %s (%s:%lld)", ref_bases.c_str(),
                         var_ref.c_str(), reference_name, (long long)(pos + 1));
      suffix = var_ref.substr(ref_bases.size());
    }
    std::string dvc;
    PutBytes(&dvc, 1, variant);
    std::vector<std::string> support_key, support_names, support_ext;
    std::string ref_names, ref_ext;
    auto read_support = [&](const Entry& e) {
      std::string rs;
      PutBytes(&rs, 1, keys[(size_t)e.key_id]);
      if (e.low_quality) PutInt(&rs, 2, 1);
      return rs;
    };
    for (const Entry& e : site.entries) {
      if (e.type != kReference) {
        std::string supported = "UNCALLED_ALLELE";
        for (const MapItem& m : amap)
          if (m.a.type == e.type && m.a.len == e.bases_len && c.arena.compare(e.bases_off, e.bases_len, m.bases) == 0) { supported = m.alt + suffix; break; }
        size_t k = 0;
        while (k < support_key.size() && support_key[k] != supported) ++k;
        if (k == support_key.size()) { support_key.push_back(supported); support_names.emplace_back(); support_ext.emplace_back(); }
        PutBytes(&support_names[k], 1, keys[(size_t)e.key_id]);
        PutBytes(&support_ext[k], 1, read_support(e));
      } else if (opt->track_ref_reads) {
        PutBytes(&ref_names, 4, keys[(size_t)e.key_id]);
        PutBytes(&ref_ext, 1, read_support(e));
      }
    }
    for (size_t k = 0; k < support_key.size(); ++k) {
      std::string entry;
      PutBytes(&entry, 1, support_key[k]);
      PutBytes(&entry, 2, support_names[k]);
      PutBytes(&dvc, 2, entry);
    }
    dvc += ref_names;
    for (size_t k = 0; k < support_key.size(); ++k) {
      std::string entry;
      PutBytes(&entry, 1, support_key[k]);
      PutBytes(&entry, 2, support_ext[k]);
      PutBytes(&dvc, 5, entry);
    }
    if (!ref_ext.empty()) PutBytes(&dvc, 6, ref_ext);
    res->protos += dvc;
    res->proto_begin.push_back((int64_t)res->protos.size());
    res->position.push_back((int32_t)pos);
  }
  *out = res.release();
  return DVB_OK;
}

int64_t dvb_debug_allele_counts(const DvbBam* bam, const uint8_t* contig_bases, int64_t contig_n_bases, int64_t start, int64_t end,
                                const int64_t* rows, int64_t n_rows, const DvbCandidateOptions* opt,
                                const int32_t* candidate_positions, int32_t n_candidate_positions, char* out, int64_t cap) {
  DvbReadTable table;
  Counter c;
  std::vector<std::string> keys;
  int st = BuildCounter(bam, &table, &c, contig_bases, contig_n_bases, start, end, rows, n_rows, opt, candidate_positions,
                        n_candidate_positions, &keys);
  if (st != DVB_OK) return -(int64_t)st;
  std::string js = "[";
  for (size_t i = 0; i < c.sites.size(); ++i) {
    if (i) js += ",";
    js += "{\"ref\":" + std::to_string(c.sites[i].ref_supporting_read_count) + ",\"alleles\":[";
    for (size_t k = 0; k < c.sites[i].entries.size(); ++k) {
      const Entry& e = c.sites[i].entries[k];
      if (k) js += ",";
      js += "[\"" + c.arena.substr(e.bases_off, e.bases_len) + "\"," + std::to_string((int)e.type) + "," +
            std::to_string((int)e.low_quality) + ",\"" + keys[(size_t)e.key_id] + "\"," + std::to_string((int)e.mapq) + "," +
            std::to_string(e.avg_base_quality) + "," + std::to_string((int)e.reverse) + "]";
    }
    js += "]}";
  }
  js += "]";
  if (out && cap > (int64_t)js.size()) memcpy(out, js.c_str(), js.size() + 1);
  return (int64_t)js.size();
}

// Test access to dvb_allele::ElementAt (the (candidate, read) support walk of the encoder's pre-pass): the read allele of one
// read at `target` found (a) by the full WalkRead over [start, end) with the whole contig resident - the last commit at that
// position, which is what the read's map entry ends up holding - and (b) by ElementAt from the three reference facts.
// each out: {found, type, low_quality, prev, read_offset, len}.
int dvb_debug_read_allele_at(const uint8_t* seq, const uint8_t* qual, int32_t seq_len, const uint32_t* cigar, int32_t n_cigar, int64_t pos,
                             const uint8_t* contig, int64_t contig_len, int64_t start, int64_t end, int64_t target, int32_t min_base_quality,
                             int32_t keep_legacy, int32_t* walk_out, int32_t* at_out) {
  if (!seq || !qual || !cigar || !contig || !walk_out || !at_out || target < start || target >= end || start < 0 || end > contig_len)
    return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_debug_read_allele_at: bad arguments");
  dvb_allele::ReadView r{seq, qual, seq_len, cigar, n_cigar, pos};
  dvb_allele::WalkParams wp;
  wp.start = start; wp.end = end; wp.contig = contig; wp.contig_origin = 0; wp.contig_avail = contig_len; wp.contig_len = contig_len;
  wp.min_base_quality = min_base_quality; wp.keep_legacy_behavior = keep_legacy;
  struct LastAt {
    int want;
    int32_t* o;
    void Commit(const Element& e) {
      if (e.position != want) return;
      o[0] = 1; o[1] = e.type; o[2] = e.low_quality; o[3] = e.len ? e.prev : 0; o[4] = e.read_offset; o[5] = e.len;
    }
  } sink{(int)(target - start), walk_out};
  for (int k = 0; k < 6; ++k) walk_out[k] = at_out[k] = 0;
  dvb_allele::WalkRead(r, wp, sink);
  dvb_allele::TargetParams tp;
  tp.target = target;
  tp.ref_base = contig[target];
  int64_t run = 0;
  while (target + 1 + run < contig_len && Canonical(contig[target + 1 + run])) ++run;
  tp.ref_run = (int)std::min<int64_t>(run, 0x7fffffff);
  tp.min_base_quality = min_base_quality;
  tp.keep_legacy_behavior = keep_legacy;
  dvb_allele::TargetElement e;
  if (dvb_allele::ElementAt(r, tp, &e)) {
    at_out[0] = 1; at_out[1] = e.type; at_out[2] = e.low_quality; at_out[3] = e.len ? e.prev : 0; at_out[4] = e.read_offset; at_out[5] = e.len;
  }
  return DVB_OK;
}

int64_t dvb_candidates_count(const DvbCandidates* c) { return c ? (int64_t)c->position.size() : 0; }

int dvb_candidates_protos(const DvbCandidates* c, const uint8_t** data, const int64_t** begin) {
  if (!c || !data || !begin) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_candidates_protos: null argument");
  *data = (const uint8_t*)c->protos.data();
  *begin = c->proto_begin.data();
  return DVB_OK;
}

int64_t dvb_candidates_positions(const DvbCandidates* c, const int32_t** positions) {
  if (!c) return 0;
  const std::vector<int32_t>& v = c->positions_only.empty() ? c->position : c->positions_only;
  if (positions) *positions = v.data();
  return (int64_t)v.size();
}

int64_t dvb_candidates_summary_counts(const DvbCandidates* c, const int32_t** counts) {
  if (!c) return 0;
  if (counts) *counts = c->summary.data();
  return (int64_t)(c->summary.size() / 2);
}

void dvb_candidates_free(DvbCandidates* c) { delete c; }

}  // extern "C"

// ================================================================================================================================
// CUDA allele counting (SURVEY.md 8(f) "next" row #2, device half): the reads of the BAM table live in HBM; one pass with a
// thread per read runs dvb_allele::WalkRead and adds into dense per-position counters with atomics, a second pass flags the
// positions that can carry a candidate.  HBM-bound integer work: per read 2 L + 4 n_cigar + 24 bytes in, per position 25
// bytes of counters (memset + atomics + one read by the flag pass) + 1 flag byte out.
// ================================================================================================================================
struct DvbDeviceReads {
  int device = 0;
  int64_t n_reads = 0;
  int32_t* pos = nullptr;
  uint8_t* mapq = nullptr;
  int64_t* seq_begin = nullptr;
  int64_t* cigar_begin = nullptr;
  uint8_t* bases = nullptr;
  uint8_t* quals = nullptr;
  uint32_t* cigar = nullptr;
  int32_t* end = nullptr;          // alignment end (exclusive), for the tile kernel's overlap test
  int64_t max_span = 0;            // max(end - pos) over the table
  bool sorted = false;             // coordinate-sorted within every contig: the tile kernel binary-searches the rows of a call
  bool rows_unordered = false;     // set by the host entry point when a call's rows are not ascending in position
  // per-call scratch (grow-only)
  dvb::DevBuf rows, ref, counts, flags, tiles;
  int64_t launches = 0;
};

namespace {

struct DeviceTable {
  const int32_t* pos;
  const uint8_t* mapq;
  const int64_t* seq_begin;
  const int64_t* cigar_begin;
  const uint8_t* bases;
  const uint8_t* quals;
  const uint32_t* cigar;
};

__global__ void dvb_allele_count_kernel(DeviceTable t, const int64_t* __restrict__ rows, int64_t n_rows, dvb_allele::WalkParams p,
                                        int min_mapping_quality, dvb_allele::DenseCounts out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows) return;
  const int64_t row = rows[i];
  if (t.mapq[row] < min_mapping_quality) return;
  const int64_t s0 = t.seq_begin[row];
  dvb_allele::ReadView r;
  r.seq = t.bases + s0;
  r.qual = t.quals + s0;
  r.seq_len = (int)(t.seq_begin[row + 1] - s0);
  r.cigar = t.cigar + t.cigar_begin[row];
  r.n_cigar = (int)(t.cigar_begin[row + 1] - t.cigar_begin[row]);
  r.pos = t.pos[row];
  dvb_allele::DenseSink sink{out, r.seq};
  dvb_allele::WalkRead(r, p, sink);
}

__global__ void dvb_allele_flag_kernel(dvb_allele::DenseCounts c, const uint8_t* __restrict__ ref /* ref[0] = position start */,
                                       int64_t len, dvb_allele::FlagParams f, uint8_t* __restrict__ flags) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < len) flags[p] = dvb_allele::FlagPosition(c, p, ref[p], f);
}

// ---- tile kernel: positions gather instead of reads scatter -------------------------------------------------------------------
// The thread-per-read kernel above sends one L2 atomic per counted base and reads every read byte by byte from one thread
// (measured on B200, 30x / 150 bp: 1.0 ms per 4 Mb = 0.87 TB/s of algorithmic bytes, 13 % of the HBM roofline, L2-atomic bound).
// Here a CTA owns a tile of kAlleleTile consecutive positions: its counters live in shared memory, the reads that can touch
// the tile are found by binary search (coordinate-sorted table; dvb_allele_tile_ranges_kernel), a WARP walks each read - the
// bases of an alignment-match operation 32 at a time with coalesced loads of bases, qualities and reference - and adds with
// shared-memory atomics; at the end the tile's counters, the indel bytes AND the candidate flags leave with coalesced stores:
// no memsets, no global atomics, no separate flag pass.  A read that spans several tiles is walked by each of them (7 % extra
// at 150 bp and 2048-position tiles).  Same arithmetic as dvb_allele::WalkRead: the only order-dependent part of the walk is
// "an element is dropped when the NEXT generated element has the same position", and the elements of one alignment-match
// run have distinct, increasing positions - so the run's elements commit independently except the last one, which is held
// as the pending element exactly as the sequential walk holds it.
constexpr int kAlleleTile = 2048;
constexpr int kAlleleTileThreads = 512;   // 16 warps walk reads; 2 CTAs per SM (see the carveout note at the launch)

__global__ void dvb_allele_tile_ranges_kernel(const int64_t* __restrict__ rows, int64_t n_rows, const int32_t* __restrict__ pos,
                                              int64_t start, int64_t len, int64_t max_span, int2* __restrict__ ranges) {
  // one WARP per tile: a 33-ary search (32 probes per step, ~4 steps for a million rows) instead of 20 dependent binary steps
  const int lane = threadIdx.x & 31;
  const int64_t tile = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_tiles = (len + kAlleleTile - 1) / kAlleleTile;
  if (tile >= n_tiles) return;
  const int64_t p0 = start + tile * kAlleleTile;
  auto lower_bound = [&](int64_t key) {       // first i with pos[rows[i]] >= key
    int64_t lo = 0, hi = n_rows;              // invariant: every i < lo has pos < key, every i >= hi has pos >= key
    while (hi - lo > 32) {
      const int64_t step = (hi - lo + 32) / 33;
      const int64_t probe = lo + (int64_t)(lane + 1) * step - 1;              // lane 0 .. 31: increasing probes inside [lo, hi)
      const bool less = probe < hi && (int64_t)pos[rows[probe]] < key;
      const unsigned m = __ballot_sync(0xffffffffu, less);                   // monotone: the lanes with pos < key form a prefix
      const int k = __popc(m);
      const int64_t new_lo = k ? lo + (int64_t)k * step : lo;                // probe of lane k-1 is < key -> lo = that probe + 1
      const int64_t new_hi = k < 32 ? min(hi, lo + (int64_t)(k + 1) * step - 1) : hi;
      lo = new_lo; hi = new_hi;
    }
    const int64_t i = lo + lane;
    const bool less = i < hi && (int64_t)pos[rows[i]] < key;
    return lo + __popc(__ballot_sync(0xffffffffu, less));
  };
  // reads with end > p0 (end <= pos + max_span) and pos <= p0 + tile (a leading insertion / soft clip is anchored at pos - 1)
  const int64_t a = lower_bound(p0 - max_span + 1), b = lower_bound(p0 + kAlleleTile + 1);
  if (lane == 0) ranges[tile] = make_int2((int)a, (int)b);
}

struct TileShared {
  int32_t ref_count[kAlleleTile];
  int32_t subst[4 * kAlleleTile];
  int32_t other[kAlleleTile];
  uint8_t indel[kAlleleTile];
};

struct PendingElement { int position; int read_offset; uint8_t type, low_quality; };

__device__ __forceinline__ void TileCommit(TileShared& sm, int tile_lo, int tile_n, const PendingElement& e, const uint8_t* seq) {
  const int idx = e.position - tile_lo;
  if (idx < 0 || idx >= tile_n) return;
  if (e.type == dvb_allele::kReference) {
    if (!e.low_quality) atomicAdd(&sm.ref_count[idx], 1);
    return;
  }
  if (e.low_quality) return;
  if (e.type == dvb_allele::kSubstitution) {
    const uint8_t b = seq[e.read_offset];
    atomicAdd(&sm.subst[4 * idx + (b == 'A' ? 0 : b == 'C' ? 1 : b == 'G' ? 2 : 3)], 1);
  } else {
    atomicAdd(&sm.other[idx], 1);
    if (e.type != dvb_allele::kSoftClip) sm.indel[idx] = 1;
  }
}

__global__ void __launch_bounds__(kAlleleTileThreads, 2)
dvb_allele_count_tile_kernel(DeviceTable t, const int32_t* __restrict__ read_end, const int64_t* __restrict__ rows, const int2* __restrict__ ranges,
                             dvb_allele::WalkParams p, int min_mapping_quality, dvb_allele::DenseCounts out, dvb_allele::FlagParams fp,
                             uint8_t* __restrict__ flags) {
  extern __shared__ __align__(16) uint8_t tile_smem[];
  TileShared& sm = *reinterpret_cast<TileShared*>(tile_smem);
  const int64_t len = p.end - p.start;
  const int tile_lo = (int)((int64_t)blockIdx.x * kAlleleTile);
  const int tile_n = (int)min((int64_t)kAlleleTile, len - tile_lo);
  for (int i = threadIdx.x; i < 6 * kAlleleTile; i += kAlleleTileThreads) reinterpret_cast<int32_t*>(tile_smem)[i] = 0;
  for (int i = threadIdx.x; i < kAlleleTile / 4; i += kAlleleTileThreads) reinterpret_cast<int32_t*>(sm.indel)[i] = 0;
  __syncthreads();
  const int2 range = ranges[blockIdx.x];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t abs_lo = p.start + tile_lo;
  // Reads are taken 32 at a time per warp: every lane fetches the header of one read (row -> mapping quality, position, end, sequence
  // and CIGAR offsets: two dependent levels of global loads, 32 reads in flight), then the warp walks the 32 reads one after the other
  // with the header fields broadcast by shuffle - the header latency is paid once per 32 reads instead of once per read.
  constexpr int kWarps = kAlleleTileThreads / 32;
  for (int rb = range.x + warp * 32; rb < range.y; rb += kWarps * 32) {
    const int my = rb + lane;
    long long h_row = -1, h_s0 = 0, h_s1 = 0, h_c0 = 0, h_c1 = 0;
    int h_pos = 0, h_ok = 0;
    if (my < range.y) {
      h_row = rows[my];
      h_pos = t.pos[h_row];
      h_ok = t.mapq[h_row] >= min_mapping_quality && !((int64_t)read_end[h_row] <= abs_lo && (int64_t)h_pos < abs_lo);   // ends before the tile (a zero-span read is kept: its clip is anchored at pos - 1)
      if (h_ok) { h_s0 = t.seq_begin[h_row]; h_s1 = t.seq_begin[h_row + 1]; h_c0 = t.cigar_begin[h_row]; h_c1 = t.cigar_begin[h_row + 1]; }
    }
    const unsigned ok_mask = __ballot_sync(0xffffffffu, h_ok);
    for (unsigned rest = ok_mask; rest; rest &= rest - 1) {
    const int src = __ffs(rest) - 1;
    const int64_t s0 = __shfl_sync(0xffffffffu, h_s0, src);
    const int seq_len = (int)(__shfl_sync(0xffffffffu, h_s1, src) - s0);
    const int64_t c0g = __shfl_sync(0xffffffffu, h_c0, src);
    const int n_cigar = (int)(__shfl_sync(0xffffffffu, h_c1, src) - c0g);
    const int read_pos = __shfl_sync(0xffffffffu, h_pos, src);
    const uint8_t* seq = t.bases + s0;
    const uint8_t* qual = t.quals + s0;
    if (seq_len == 0) continue;
    const uint32_t* cigar = t.cigar + c0g;
    bool have = false;
    PendingElement pending{-1, 0, 0, 0};
    int read_offset = 0;
    int64_t interval_offset = (int64_t)read_pos - p.start;
    auto usable = [&](int base_offset, bool* low_quality) {        // CanBasesBeUsed(offset, 1)
      const int q = qual[base_offset];
      if (q < p.min_base_quality && p.keep_legacy_behavior) return false;
      if (!dvb_allele::Canonical(seq[base_offset])) return false;
      *low_quality = !p.keep_legacy_behavior && q < p.min_base_quality;
      return true;
    };
    for (int c = 0; c < n_cigar; ++c) {
      const uint32_t cg = cigar[c];
      const int op = (int)(cg & 0xF), op_len = (int)(cg >> 4);
      if (op == 0 || op == 7 || op == 8) {
        int64_t i0 = interval_offset < 0 ? min(-interval_offset, (int64_t)op_len) : 0;
        int64_t i1 = min((int64_t)op_len, min(len - interval_offset, (int64_t)(seq_len - read_offset)));
        if (i1 > i0) {
          // the last generated element of the run (the usable base with the largest index) becomes the pending element
          int last = -1;
          for (int64_t hi = i1; hi > i0 && last < 0; hi -= 32) {
            const int64_t i = hi - 1 - lane;
            bool lq = false;
            const bool ok = i >= i0 && usable(read_offset + (int)i, &lq);
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            if (m) last = (int)(hi - 1 - (__ffs(m) - 1));
          }
          if (last >= 0) {
            if (have && lane == 0) TileCommit(sm, tile_lo, tile_n, pending, seq);     // positions differ: every run element lies after it
            // bulk: the run's elements inside the tile, except `last`
            const int64_t b0 = max(i0, (int64_t)tile_lo - interval_offset), b1 = min((int64_t)last, (int64_t)tile_lo + tile_n - interval_offset);
            // 8 x 32 bases per trip with all loads issued before the first use (memory-level parallelism: the kernel is bound by the
            // latency of dependent global loads, not by bandwidth - measured: the scalar form of this loop ran as slowly as the scatter kernel)
            for (int64_t base = b0; base < b1; base += 8 * 32) {
              uint8_t bb[8], qq[8], rr[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const int64_t i = base + u * 32 + lane;
                if (i < b1) {
                  bb[u] = seq[read_offset + (int)i];
                  qq[u] = qual[read_offset + (int)i];
                  rr[u] = dvb_allele::RefAt(p, p.start + interval_offset + i);
                }
              }
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const int64_t i = base + u * 32 + lane;
                if (i >= b1) continue;
                const uint8_t b = bb[u];
                const int q = qq[u];
                if ((q < p.min_base_quality && p.keep_legacy_behavior) || !dvb_allele::Canonical(b)) continue;     // CanBasesBeUsed(offset, 1)
                if (!p.keep_legacy_behavior && q < p.min_base_quality) continue;                                  // low quality: counted nowhere
                const int idx = (int)(interval_offset + i) - tile_lo;
                if (rr[u] == b) atomicAdd(&sm.ref_count[idx], 1);
                else atomicAdd(&sm.subst[4 * idx + (b == 'A' ? 0 : b == 'C' ? 1 : b == 'G' ? 2 : 3)], 1);
              }
            }
            bool lq = false;
            usable(read_offset + last, &lq);
            pending.position = (int)(interval_offset + last);
            pending.read_offset = read_offset + last;
            pending.type = dvb_allele::RefAt(p, p.start + interval_offset + last) == seq[read_offset + last] ? dvb_allele::kReference : dvb_allele::kSubstitution;
            pending.low_quality = lq;
            have = true;
          }
        }
        read_offset += op_len;
        interval_offset += op_len;
      } else if (op == 4 || op == 1 || op == 2) {
        // MakeIndelReadAllele, warp-cooperative over the operation's bases; every lane ends with the same element
        PendingElement e{-1, read_offset, 0, 0};
        bool valid = true;
        uint8_t prev = 0;
        if (read_offset == 0) {
          const int64_t abs = p.start + interval_offset - 1;
          if (!dvb_allele::RefAvailable(p, abs, 1)) valid = false; else prev = dvb_allele::RefAt(p, abs);
        } else {
          prev = seq[read_offset - 1];
        }
        if (valid && !dvb_allele::Canonical(prev)) valid = false;
        bool low_quality = false;
        if (valid && op != 2) {
          if (read_offset + op_len > seq_len) valid = false;
          else {
            int sum = 0;
            bool bad = false;
            for (int i = lane; i < op_len; i += 32) {
              const int q = qual[read_offset + i];
              sum += q;
              if ((q < p.min_base_quality && p.keep_legacy_behavior) || !dvb_allele::Canonical(seq[read_offset + i])) bad = true;
            }
            sum = __reduce_add_sync(0xffffffffu, sum);
            if (__any_sync(0xffffffffu, bad)) valid = false;
            low_quality = !p.keep_legacy_behavior && sum < p.min_base_quality * op_len;
          }
        }
        if (valid && op == 2) {
          const int64_t ref_abs = p.start + interval_offset;
          if (!dvb_allele::RefAvailable(p, ref_abs, op_len)) valid = false;
          else {
            bool bad = false;
            for (int i = lane; i < op_len; i += 32)
              if (!dvb_allele::Canonical(dvb_allele::RefAt(p, ref_abs + i))) bad = true;
            if (__any_sync(0xffffffffu, bad)) valid = false;
          }
        }
        if (valid && (interval_offset - 1 > 0x7fffffff || interval_offset - 1 < -0x7fffffff)) valid = false;
        if (valid) {
          e.position = (int)(interval_offset - 1);
          e.type = op == 2 ? dvb_allele::kDeletion : op == 1 ? dvb_allele::kInsertion : dvb_allele::kSoftClip;
          e.low_quality = low_quality;
        }
        if (have && pending.position != e.position && lane == 0) TileCommit(sm, tile_lo, tile_n, pending, seq);
        pending = e;
        have = true;
        if (op == 2) interval_offset += op_len; else read_offset += op_len;
      } else if (op == 6 || op == 3) {
        interval_offset += op_len;
      }
      if (interval_offset >= (int64_t)tile_lo + tile_n + 1 && (op == 0 || op == 7 || op == 8 || op == 2 || op == 3)) {
        // Everything later in the read lies behind the tile (an indel that follows is anchored at interval_offset - 1 >= tile end,
        // and a pending element it could supersede has that same position, outside the tile): only the pending element is left.
        break;
      }
    }
    if (have && lane == 0) TileCommit(sm, tile_lo, tile_n, pending, seq);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < tile_n; i += kAlleleTileThreads) {
    const int64_t q = tile_lo + i;
    out.ref_count[q] = sm.ref_count[i];
#pragma unroll
    for (int b = 0; b < 4; ++b) out.subst[4 * q + b] = sm.subst[4 * i + b];
    out.other[q] = sm.other[i];
    out.indel[q] = sm.indel[i];
    // FlagPosition on the tile's own counters
    const uint8_t ref_base = dvb_allele::RefAt(p, p.start + q);
    uint8_t flag = 0;
    if (dvb_allele::Canonical(ref_base)) {
      const int total = sm.ref_count[i] + sm.subst[4 * i] + sm.subst[4 * i + 1] + sm.subst[4 * i + 2] + sm.subst[4 * i + 3] + sm.other[i];
      flag = sm.indel[i] ? 2 : 0;
      for (int b = 0; b < 4; ++b) {
        const int sb = sm.subst[4 * i + b];
        if (sb > 0 && sb >= fp.min_count_snps && (1.0 * sb) / total >= 0.9 * fp.min_fraction_snps) flag |= 1;
      }
    }
    flags[q] = flag;
  }
}

dvb_allele::FlagParams MakeFlagParams(const DvbCandidateOptions& o) {
  dvb_allele::FlagParams f;
  f.min_count_snps = o.min_count_snps;
  f.min_fraction_snps = (double)o.min_fraction_snps * std::min(1.0, (double)o.min_fraction_multiplier);
  return f;
}

// Window of the contig a set of reads can touch: [min(start, first read) - 1, max(end, last read end) + 1) inside the contig.
void RefWindow(const DvbReadTable& t, const int64_t* rows, int64_t n_rows, int64_t start, int64_t end, int64_t contig_len,
               int64_t* w0, int64_t* w1) {
  int64_t lo = start, hi = end;
  for (int64_t i = 0; i < n_rows; ++i) {
    lo = std::min<int64_t>(lo, t.pos[rows[i]]);
    hi = std::max<int64_t>(hi, t.end[rows[i]]);
  }
  *w0 = std::max<int64_t>(0, lo - 1);
  *w1 = std::min<int64_t>(contig_len, hi + 1);
}

int CheckCountArgs(const void* reads, const uint8_t* contig_bases, int64_t contig_n_bases, int64_t start, int64_t end,
                   const int64_t* rows, int64_t n_rows, const DvbCandidateOptions* opt) {
  if (!reads || !contig_bases || !opt || (n_rows && !rows) || n_rows < 0)
    return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_allele_count: null argument");
  if (start < 0 || end > contig_n_bases || start >= end || end - start > (int64_t)1 << 30)
    return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_allele_count: bad interval [%lld, %lld)", (long long)start, (long long)end);
  return DVB_OK;
}

}  // namespace

extern "C" {

int dvb_device_reads_create(const DvbBam* bam, int device, DvbDeviceReads** out) {
  if (!bam || !out) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_device_reads_create: null argument");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return dvb::fail(DVB_ERR_NO_DEVICE, "no CUDA device (allele counting on the device has no CPU path)");
  if (device < 0 || device >= ndev) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "device %d out of range", device);
  DVB_CUDA(cudaSetDevice(device));
  DvbReadTable t;
  int st = dvb_bam_table(bam, &t);
  if (st != DVB_OK) return st;
  auto* d = new DvbDeviceReads();
  d->device = device;
  d->n_reads = t.n_reads;
  auto up = [&](void** dst, const void* src, size_t bytes) -> cudaError_t {
    cudaError_t e = cudaMalloc(dst, std::max<size_t>(bytes, 16));
    if (e != cudaSuccess) return e;
    return bytes ? cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice) : cudaSuccess;
  };
  const size_t n = (size_t)t.n_reads;
  cudaError_t e = cudaSuccess;
  if (e == cudaSuccess) e = up((void**)&d->pos, t.pos, n * 4);
  if (e == cudaSuccess) e = up((void**)&d->mapq, t.mapq, n);
  if (e == cudaSuccess) e = up((void**)&d->seq_begin, t.seq_begin, (n + 1) * 8);
  if (e == cudaSuccess) e = up((void**)&d->cigar_begin, t.cigar_begin, (n + 1) * 8);
  if (e == cudaSuccess) e = up((void**)&d->bases, t.bases, (size_t)t.n_bases);
  if (e == cudaSuccess) e = up((void**)&d->quals, t.quals, (size_t)t.n_bases);
  if (e == cudaSuccess) e = up((void**)&d->cigar, t.cigar, (size_t)t.n_cigar * 4);
  if (e == cudaSuccess) e = up((void**)&d->end, t.end, n * 4);
  d->sorted = true;
  for (size_t i = 0; i < n; ++i) {
    d->max_span = std::max<int64_t>(d->max_span, (int64_t)t.end[i] - t.pos[i]);
    if (i && t.ref_id[i] == t.ref_id[i - 1] && t.pos[i] < t.pos[i - 1]) d->sorted = false;
  }
  if (e != cudaSuccess) {
    dvb_device_reads_destroy(d);
    return dvb::fail(DVB_ERR_CUDA, "dvb_device_reads_create: %s", cudaGetErrorString(e));
  }
  *out = d;
  return DVB_OK;
}

void dvb_device_reads_destroy(DvbDeviceReads* d) {
  if (!d) return;
  cudaSetDevice(d->device);
  cudaFree(d->pos); cudaFree(d->mapq); cudaFree(d->seq_begin); cudaFree(d->cigar_begin);
  cudaFree(d->bases); cudaFree(d->quals); cudaFree(d->cigar); cudaFree(d->end);
  d->rows.release(); d->ref.release(); d->counts.release(); d->flags.release(); d->tiles.release();
  delete d;
}

int64_t dvb_device_reads_launch_count(const DvbDeviceReads* d) { return d ? d->launches : 0; }

// Device pointers in and out; asynchronous on `stream`.  ref_dev[0] is absolute position ref_origin and must cover the window
// the rows touch (RefWindow).  counts_dev = int32[6 * len] laid out as ref_count[len], subst[4 * len], other[len];
// indel_dev / flags_dev = uint8[len].
int dvb_allele_count_device(DvbDeviceReads* d, const uint8_t* ref_dev, int64_t ref_origin, int64_t ref_avail, int64_t contig_n_bases,
                            int64_t start, int64_t end, const int64_t* rows_dev, int64_t n_rows, const DvbCandidateOptions* opt,
                            int32_t* counts_dev, uint8_t* indel_dev, uint8_t* flags_dev, void* stream) {
  if (!d || !ref_dev || !opt || !counts_dev || !indel_dev || !flags_dev || (n_rows && !rows_dev))
    return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_allele_count_device: null argument");
  if (start < ref_origin || end > ref_origin + ref_avail || start >= end)
    return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_allele_count_device: interval outside the resident reference window");
  DVB_CUDA(cudaSetDevice(d->device));
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t len = end - start;
  // Tile kernel (coordinate-sorted table; the rows of a call are ascending table rows of one contig - what NativeBamTable.query_indices
  // and region_reads return): DVB_ALLELE_TILES=0 keeps the thread-per-read scatter kernel, as does an unsorted file.
  static const bool tiles_enabled = !getenv("DVB_ALLELE_TILES") || atoi(getenv("DVB_ALLELE_TILES")) != 0;
  const bool use_tiles = d->sorted && !d->rows_unordered && tiles_enabled && n_rows > 0;
  if (!use_tiles) {
    DVB_CUDA(cudaMemsetAsync(counts_dev, 0, (size_t)len * 24, s));
    DVB_CUDA(cudaMemsetAsync(indel_dev, 0, (size_t)len, s));
  }
  dvb_allele::WalkParams p;
  p.start = start;
  p.end = end;
  p.contig = ref_dev;
  p.contig_origin = ref_origin;
  p.contig_avail = ref_avail;
  p.contig_len = contig_n_bases;
  p.min_base_quality = opt->min_base_quality;
  p.keep_legacy_behavior = opt->keep_legacy_behavior;
  dvb_allele::DenseCounts c{counts_dev, counts_dev + len, counts_dev + 5 * len, indel_dev};
  DeviceTable t{d->pos, d->mapq, d->seq_begin, d->cigar_begin, d->bases, d->quals, d->cigar};
  if (use_tiles) {
    const int64_t n_tiles = (len + kAlleleTile - 1) / kAlleleTile;
    DVB_CUDA(d->tiles.reserve((size_t)n_tiles * sizeof(int2)));
    static bool attr_set = false;
    if (!attr_set) {
      DVB_CUDA(cudaFuncSetAttribute(dvb_allele_count_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileShared)));
      // Without this the driver picks the smallest shared-memory carveout that fits ONE block: the first version ran one 8-warp CTA per
      // SM and was bound by the latency of its dependent loads (633 us per 2 Mb, no faster than the scatter kernel).
      DVB_CUDA(cudaFuncSetAttribute(dvb_allele_count_tile_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
      attr_set = true;
    }
    dvb_allele_tile_ranges_kernel<<<(unsigned)((n_tiles * 32 + 127) / 128), 128, 0, s>>>(rows_dev, n_rows, d->pos, start, len, d->max_span, (int2*)d->tiles.p);
    dvb_allele_count_tile_kernel<<<(unsigned)n_tiles, kAlleleTileThreads, sizeof(TileShared), s>>>(t, d->end, rows_dev, (const int2*)d->tiles.p, p,
                                                                                                   opt->min_mapping_quality, c, MakeFlagParams(*opt), flags_dev);
    d->launches += 2;
    DVB_CUDA(cudaGetLastError());
    return DVB_OK;
  }
  if (n_rows) {
    dvb_allele_count_kernel<<<(unsigned)((n_rows + 127) / 128), 128, 0, s>>>(t, rows_dev, n_rows, p, opt->min_mapping_quality, c);
    ++d->launches;
  }
  dvb_allele_flag_kernel<<<(unsigned)((len + 255) / 256), 256, 0, s>>>(c, ref_dev + (start - ref_origin), len, MakeFlagParams(*opt), flags_dev);
  ++d->launches;
  DVB_CUDA(cudaGetLastError());
  return DVB_OK;
}

// Host pointers in and out (synchronous): uploads the rows and the reference window, counts, flags, copies back.
// counts_host = int32[6 * len] (ref_count, subst[4 * len], other), flags_host = uint8[len].
int dvb_allele_count_host(DvbDeviceReads* d, const DvbBam* bam, const uint8_t* contig_bases, int64_t contig_n_bases, int64_t start,
                          int64_t end, const int64_t* rows, int64_t n_rows, const DvbCandidateOptions* opt, int32_t* counts_host,
                          uint8_t* flags_host) {
  int st = CheckCountArgs(d, contig_bases, contig_n_bases, start, end, rows, n_rows, opt);
  if (st != DVB_OK) return st;
  if (!bam || !counts_host || !flags_host) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_allele_count_host: null argument");
  DvbReadTable t;
  st = dvb_bam_table(bam, &t);
  if (st != DVB_OK) return st;
  if (t.n_reads != d->n_reads) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_allele_count_host: device table was made from another BAM");
  d->rows_unordered = false;
  for (int64_t i = 0; i < n_rows; ++i) {
    if (rows[i] < 0 || rows[i] >= t.n_reads) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_allele_count_host: read row out of range");
    if (i && (t.pos[rows[i]] < t.pos[rows[i - 1]] || t.ref_id[rows[i]] != t.ref_id[rows[i - 1]])) d->rows_unordered = true;
  }
  int64_t w0, w1;
  RefWindow(t, rows, n_rows, start, end, contig_n_bases, &w0, &w1);
  const int64_t len = end - start;
  DVB_CUDA(cudaSetDevice(d->device));
  DVB_CUDA(d->rows.reserve((size_t)std::max<int64_t>(n_rows, 1) * 8));
  DVB_CUDA(d->ref.reserve((size_t)(w1 - w0)));
  DVB_CUDA(d->counts.reserve((size_t)len * 25));
  DVB_CUDA(d->flags.reserve((size_t)len));
  if (n_rows) DVB_CUDA(cudaMemcpyAsync(d->rows.p, rows, (size_t)n_rows * 8, cudaMemcpyHostToDevice, 0));
  DVB_CUDA(cudaMemcpyAsync(d->ref.p, contig_bases + w0, (size_t)(w1 - w0), cudaMemcpyHostToDevice, 0));
  int32_t* counts_dev = (int32_t*)d->counts.p;
  uint8_t* indel_dev = (uint8_t*)d->counts.p + (size_t)len * 24;
  st = dvb_allele_count_device(d, (const uint8_t*)d->ref.p, w0, w1 - w0, contig_n_bases, start, end, (const int64_t*)d->rows.p, n_rows,
                               opt, counts_dev, indel_dev, (uint8_t*)d->flags.p, nullptr);
  if (st != DVB_OK) return st;
  DVB_CUDA(cudaMemcpyAsync(counts_host, counts_dev, (size_t)len * 24, cudaMemcpyDeviceToHost, 0));
  DVB_CUDA(cudaMemcpyAsync(flags_host, d->flags.p, (size_t)len, cudaMemcpyDeviceToHost, 0));
  DVB_CUDA(cudaStreamSynchronize(0));
  return DVB_OK;
}

// Test access: the same walk, sink and flag function instantiated on the host (what the kernels must reproduce).
int dvb_debug_allele_count_dense_host(const DvbBam* bam, const uint8_t* contig_bases, int64_t contig_n_bases, int64_t start, int64_t end,
                                      const int64_t* rows, int64_t n_rows, const DvbCandidateOptions* opt, int windowed,
                                      int32_t* counts_host, uint8_t* flags_host) {
  int st = CheckCountArgs(bam, contig_bases, contig_n_bases, start, end, rows, n_rows, opt);
  if (st != DVB_OK) return st;
  if (!counts_host || !flags_host) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_debug_allele_count_dense_host: null argument");
  DvbReadTable t;
  st = dvb_bam_table(bam, &t);
  if (st != DVB_OK) return st;
  const int64_t len = end - start;
  memset(counts_host, 0, (size_t)len * 24);
  std::vector<uint8_t> indel((size_t)len, 0);
  dvb_allele::WalkParams p;
  p.start = start;
  p.end = end;
  int64_t w0 = 0, w1 = contig_n_bases;
  if (windowed) RefWindow(t, rows, n_rows, start, end, contig_n_bases, &w0, &w1);   // the window the device path uploads
  p.contig = contig_bases + w0;
  p.contig_origin = w0;
  p.contig_avail = w1 - w0;
  p.contig_len = contig_n_bases;
  p.min_base_quality = opt->min_base_quality;
  p.keep_legacy_behavior = opt->keep_legacy_behavior;
  dvb_allele::DenseCounts c{counts_host, counts_host + len, counts_host + 5 * len, indel.data()};
  for (int64_t i = 0; i < n_rows; ++i) {
    const int64_t row = rows[i];
    if (row < 0 || row >= t.n_reads) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "read row out of range");
    if (t.mapq[row] < opt->min_mapping_quality) continue;
    const int64_t s0 = t.seq_begin[row];
    dvb_allele::ReadView r;
    r.seq = t.bases + s0;
    r.qual = t.quals + s0;
    r.seq_len = (int)(t.seq_begin[row + 1] - s0);
    r.cigar = t.cigar + t.cigar_begin[row];
    r.n_cigar = (int)(t.cigar_begin[row + 1] - t.cigar_begin[row]);
    r.pos = t.pos[row];
    dvb_allele::DenseSink sink{c, r.seq};
    dvb_allele::WalkRead(r, p, sink);
  }
  const dvb_allele::FlagParams f = MakeFlagParams(*opt);
  for (int64_t q = 0; q < len; ++q) flags_host[q] = dvb_allele::FlagPosition(c, q, contig_bases[start + q], f);
  return DVB_OK;
}

}  // extern "C"

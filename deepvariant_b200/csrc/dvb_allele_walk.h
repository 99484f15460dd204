// One read's contribution to the allele counts, as a streaming walk with O(1) state - shared by the host allele counter
// (dvb_candidates.cu, Counter) and the CUDA allele-count kernels, so that both run the very same arithmetic.
//
// Restates AlleleCounter::Add + MakeIndelReadAllele + AddReadAlleles (deepvariant/allelecounter.cc:880-978, 402-473, 475-546)
// without the intermediate to_add vector: the reference builds the vector of ReadAlleles of a read and then drops every element
// whose successor has the same position (a base superseded by the indel anchored on it); here the last generated element is held
// back until the next one is known.  Elements the reference never generates (unusable bases) do not separate neighbours there
// either; indel elements are always generated, invalid ones with position -1 (ReadAllele::kInvalidPosition).
#ifndef DVB_ALLELE_WALK_H_
#define DVB_ALLELE_WALK_H_

#include <cstdint>

#if defined(__CUDACC__)
#define DVB_HD __host__ __device__ __forceinline__
#else
#define DVB_HD inline
#endif

namespace dvb_allele {

enum AlleleType : uint8_t { kUnspecified = 0, kReference = 1, kSubstitution = 2, kInsertion = 3, kDeletion = 4, kSoftClip = 5 };

DVB_HD bool Canonical(uint8_t b) { return b == 'A' || b == 'C' || b == 'G' || b == 'T'; }

struct ReadView {
  const uint8_t* seq;
  const uint8_t* qual;
  int seq_len;
  const uint32_t* cigar;     // BAM packing (len << 4 | op)
  int n_cigar;
  int64_t pos;               // alignment start, 0-based
};

struct WalkParams {
  int64_t start, end;        // the allele counter's interval
  const uint8_t* contig;     // upper-case bases; contig[i] is absolute position contig_origin + i
  int64_t contig_origin;     // 0 on the host (whole contig); the device may hold a window
  int64_t contig_avail;      // number of bases behind `contig`
  int64_t contig_len;        // bases in the whole contig (RefBases validity, allelecounter.cc:360-373)
  int min_base_quality;
  int keep_legacy_behavior;
};

struct Element {             // a ReadAllele that survives: inside the interval and not superseded
  int position;              // relative to start
  uint8_t type, low_quality;
  uint8_t prev;              // indels: the anchor base
  int avg_base_quality;
  int read_offset;           // base elements: the read base; I / S: first inserted / clipped base
  int len;                   // indels: operation length (allele bases = prev + len bases); 0 for base elements
  int64_t ref_abs;           // D: absolute position of the first deleted base
};

// AlleleCounter::RefBases: true iff [abs, abs + n) lies inside the contig.  Outside the resident window the walk cannot
// answer; callers keep the window wide enough (the host passes the whole contig).
DVB_HD bool RefAvailable(const WalkParams& p, int64_t abs, int64_t n) {
  return abs >= 0 && abs + n <= p.contig_len && abs >= p.contig_origin && abs + n <= p.contig_origin + p.contig_avail;
}
DVB_HD uint8_t RefAt(const WalkParams& p, int64_t abs) { return p.contig[abs - p.contig_origin]; }

// CanBasesBeUsed (allelecounter.cc:195-224).
DVB_HD bool CanBasesBeUsed(const ReadView& r, const WalkParams& p, int offset, int len, bool* low_quality) {
  int sum = 0;
  for (int i = 0; i < len; ++i) {
    const int q = r.qual[offset + i];
    sum += q;
    if (q < p.min_base_quality && p.keep_legacy_behavior) return false;
    if (!Canonical(r.seq[offset + i])) return false;
  }
  *low_quality = !p.keep_legacy_behavior && sum < p.min_base_quality * len;
  return true;
}

// MakeIndelReadAllele (allelecounter.cc:402-473); returns an element with position -1 when the allele is not usable.
DVB_HD Element MakeIndel(const ReadView& r, const WalkParams& p, int64_t interval_offset, int read_offset, int op, int op_len) {
  Element e;
  e.position = -1;
  e.type = kUnspecified;
  e.low_quality = 0;
  e.prev = 0;
  e.avg_base_quality = 0;
  e.read_offset = read_offset;
  e.len = op_len;
  e.ref_abs = p.start + interval_offset;
  uint8_t prev;
  if (read_offset == 0) {                 // GetPrevBase: no previous read base, take the reference's
    const int64_t abs = p.start + interval_offset - 1;
    if (!RefAvailable(p, abs, 1)) return e;
    prev = RefAt(p, abs);
  } else {
    prev = r.seq[read_offset - 1];
  }
  if (!Canonical(prev)) return e;
  bool low_quality = false;
  if (op != 2) {
    if (read_offset + op_len > r.seq_len) return e;       // the reference CHECK-fails on such a record
    if (!CanBasesBeUsed(r, p, read_offset, op_len, &low_quality)) return e;
  }
  if (op == 2) {
    if (!RefAvailable(p, e.ref_abs, op_len)) return e;
    for (int i = 0; i < op_len; ++i)
      if (!Canonical(RefAt(p, e.ref_abs + i))) return e;
    e.type = kDeletion;
    e.avg_base_quality = r.qual[read_offset > 0 ? read_offset - 1 : 0];     // GetAvgBaseQuality, DELETE
  } else {
    e.type = op == 1 ? kInsertion : kSoftClip;
    int sum = 0;
    for (int i = 0; i < op_len; ++i) sum += r.qual[read_offset + i];
    e.avg_base_quality = sum / (op_len > 1 ? op_len : 1);
  }
  e.low_quality = low_quality;
  e.prev = prev;
  if (interval_offset - 1 > 0x7fffffff || interval_offset - 1 < -0x7fffffff) return e;
  e.position = (int)(interval_offset - 1);
  return e;
}

// Calls sink.Commit(const Element&) for every surviving element of the read, in order.  The caller has already applied the
// mapping-quality filter (AlleleCounter::Add's first test).
template <class Sink>
DVB_HD void WalkRead(const ReadView& r, const WalkParams& p, Sink& sink) {
  if (r.seq_len == 0) return;
  const int64_t len = p.end - p.start;
  Element pending;
  bool have = false;
  int read_offset = 0;
  int64_t interval_offset = r.pos - p.start;
  auto push = [&](const Element& e) {
    if (have && pending.position != e.position && pending.position >= 0 && pending.position < len) sink.Commit(pending);
    pending = e;
    have = true;
  };
  for (int c = 0; c < r.n_cigar; ++c) {
    const int op = (int)(r.cigar[c] & 0xF), op_len = (int)(r.cigar[c] >> 4);
    if (op == 0 || op == 7 || op == 8) {
      // bases before the interval generate nothing: jump to the first one inside
      int i0 = 0;
      if (interval_offset < 0) i0 = (int)(-interval_offset < (int64_t)op_len ? -interval_offset : op_len);
      for (int i = i0; i < op_len; ++i) {
        const int64_t ref_offset = interval_offset + i;
        if (ref_offset >= len) break;
        const int base_offset = read_offset + i;
        if (base_offset >= r.seq_len) break;
        bool low_quality = false;
        if (!CanBasesBeUsed(r, p, base_offset, 1, &low_quality)) continue;
        Element e;
        e.position = (int)ref_offset;
        e.type = RefAt(p, p.start + ref_offset) == r.seq[base_offset] ? kReference : kSubstitution;
        e.low_quality = low_quality;
        e.prev = 0;
        e.avg_base_quality = r.qual[base_offset];
        e.read_offset = base_offset;
        e.len = 0;
        e.ref_abs = p.start + ref_offset;
        push(e);
      }
      read_offset += op_len;
      interval_offset += op_len;
    } else if (op == 4 || op == 1) {
      push(MakeIndel(r, p, interval_offset, read_offset, op, op_len));
      read_offset += op_len;
    } else if (op == 2) {
      push(MakeIndel(r, p, interval_offset, read_offset, op, op_len));
      interval_offset += op_len;
    } else if (op == 6 || op == 3) {
      interval_offset += op_len;
    }
  }
  if (have && pending.position >= 0 && pending.position < len) sink.Commit(pending);
}

// ---- the read allele of ONE position (the (candidate, read) support kernel) ------------------------------------------------------
// What AlleleCount.read_alleles of position `target` holds for this read once WalkRead has run: elements come in non-decreasing
// position order, only an element followed by one of a different position is committed, and a later commit of the same read at
// the same site replaces the earlier one - so the entry is the LAST generated element whose position is `target`.  This walk visits
// the CIGAR only (O(n_cigar)), looks at the one base that can land on `target` and at the indel operations anchored on it, and
// needs three facts about the reference instead of the contig: the base at `target`, and how many canonical in-contig bases follow
// it (`ref_run`: a deletion of op_len bases anchored here is usable iff op_len <= ref_run; allelecounter.cc:449-456).
// tests/test_pair_support.py checks it against WalkRead on random reads (host instantiation, whole contig resident).
struct TargetParams {
  int64_t target;            // absolute position
  uint8_t ref_base;          // contig[target], upper case
  int ref_run;               // canonical bases in [target + 1, contig end) before the first non-canonical one (capped by the caller)
  int min_base_quality;
  int keep_legacy_behavior;
};

struct TargetElement {
  uint8_t type, low_quality, prev;
  int read_offset, len;
};

DVB_HD bool BasesUsable(const ReadView& r, int min_base_quality, int keep_legacy, int offset, int len, bool* low_quality) {
  int sum = 0;
  for (int i = 0; i < len; ++i) {
    const int q = r.qual[offset + i];
    sum += q;
    if (q < min_base_quality && keep_legacy) return false;
    if (!Canonical(r.seq[offset + i])) return false;
  }
  *low_quality = !keep_legacy && sum < min_base_quality * len;
  return true;
}

DVB_HD bool ElementAt(const ReadView& r, const TargetParams& p, TargetElement* out) {
  if (r.seq_len == 0) return false;
  bool found = false;
  int read_offset = 0;
  int64_t ref_pos = r.pos;                              // absolute position the next reference-consuming operation starts at
  for (int c = 0; c < r.n_cigar; ++c) {
    if (ref_pos - 1 > p.target) break;                  // every later element lies right of the target
    const int op = (int)(r.cigar[c] & 0xF), op_len = (int)(r.cigar[c] >> 4);
    if (op == 0 || op == 7 || op == 8) {
      const int64_t i = p.target - ref_pos;
      if (i >= 0 && i < op_len && read_offset + i < r.seq_len) {
        const int base_offset = read_offset + (int)i;
        bool low_quality = false;
        if (BasesUsable(r, p.min_base_quality, p.keep_legacy_behavior, base_offset, 1, &low_quality)) {
          out->type = r.seq[base_offset] == p.ref_base ? kReference : kSubstitution;
          out->low_quality = low_quality;
          out->prev = 0;
          out->read_offset = base_offset;
          out->len = 0;
          found = true;
        }
      }
      read_offset += op_len;
      ref_pos += op_len;
    } else if (op == 4 || op == 1 || op == 2) {
      if (ref_pos - 1 == p.target) {
        // MakeIndelReadAllele: the anchor is the previous read base, or the reference base when the read starts here
        const uint8_t prev = read_offset == 0 ? p.ref_base : r.seq[read_offset - 1];
        bool ok = Canonical(prev), low_quality = false;
        if (ok && op != 2) ok = read_offset + op_len <= r.seq_len && BasesUsable(r, p.min_base_quality, p.keep_legacy_behavior, read_offset, op_len, &low_quality);
        if (ok && op == 2) ok = op_len <= p.ref_run;
        if (ok) {
          out->type = op == 2 ? kDeletion : op == 1 ? kInsertion : kSoftClip;
          out->low_quality = low_quality;
          out->prev = prev;
          out->read_offset = read_offset;
          out->len = op_len;
          found = true;
        }
      }
      if (op == 2) ref_pos += op_len; else read_offset += op_len;
    } else if (op == 6 || op == 3) {
      ref_pos += op_len;
    }
  }
  return found;
}

// ---- dense per-position counters (the CUDA allele-count pass; also instantiated on the host for the tests) ---------------------
// What SumAlleleCounts / TotalAlleleCounts (allelecounter.cc:78-169) need for substitutions, without read identities:
//   ref_count[p]      AlleleCount.ref_supporting_read_count
//   subst[4 p + b]    non-low-quality SUBSTITUTION entries whose read base is "ACGT"[b]
//   other[p]          non-low-quality INSERTION / DELETION / SOFT_CLIP entries (they count in the total)
//   indel[p]          1 when some non-low-quality INSERTION or DELETION is anchored at p
// (a read key that occurs twice over one position is counted twice here and once in the reference's map; the flags below are
// a conservative pre-filter and the exact caller re-derives every flagged site.)
struct DenseCounts {
  int32_t* ref_count;
  int32_t* subst;
  int32_t* other;
  uint8_t* indel;
};

DVB_HD void DenseAdd(int32_t* p) {
#if defined(__CUDA_ARCH__)
  atomicAdd(p, 1);
#else
  ++*p;
#endif
}

struct DenseSink {
  DenseCounts out;
  const uint8_t* seq;
  DVB_HD void Commit(const Element& e) {
    if (e.type == kReference) {
      if (!e.low_quality) DenseAdd(out.ref_count + e.position);
      return;
    }
    if (e.low_quality) return;
    if (e.type == kSubstitution) {
      const uint8_t b = seq[e.read_offset];
      DenseAdd(out.subst + 4 * (int64_t)e.position + (b == 'A' ? 0 : b == 'C' ? 1 : b == 'G' ? 2 : 3));
    } else {
      DenseAdd(out.other + e.position);
      if (e.type != kSoftClip) out.indel[e.position] = 1;
    }
  }
};

struct FlagParams {
  int min_count_snps;
  double min_fraction_snps;       // (double)(float) of the option, times min(1, multiplier)
};

// Candidate pre-filter of one position: bit 0 = some substitution allele passes IsGoodAltAlleleWithReason's count and ratio
// tests (variant_calling_multisample.cc:175-196) with a 10 % slack on the ratio, bit 1 = an indel is anchored here.
// A superset of the positions CallVariant emits; the exact decision is taken by the caller on the flagged sites.
DVB_HD uint8_t FlagPosition(const DenseCounts& c, int64_t p, uint8_t ref_base, const FlagParams& f) {
  if (!Canonical(ref_base)) return 0;
  const int32_t* s = c.subst + 4 * p;
  const int total = c.ref_count[p] + s[0] + s[1] + s[2] + s[3] + c.other[p];
  uint8_t flag = c.indel[p] ? 2 : 0;
  for (int b = 0; b < 4; ++b)
    if (s[b] > 0 && s[b] >= f.min_count_snps && (1.0 * s[b]) / total >= 0.9 * f.min_fraction_snps) flag |= 1;
  return flag;
}

}  // namespace dvb_allele
#endif  // DVB_ALLELE_WALK_H_

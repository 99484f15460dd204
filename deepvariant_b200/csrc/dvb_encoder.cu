// dvb_encoder.cu — batched pileup-image encoder for sm_100a (B200) + its C ABI.
//
// Replaces, for a whole batch of candidate images per launch, the reference's
//   BuildPileupForOneSample            deepvariant/pileup_image_native.cc:296-447
//     DownsampleReadIndices            :153-165   (tables built on the host: libc++ or libstdc++ form, shuffle_stdlib)
//     EncodeRead / CalculateChannels   :477-510, deepvariant/pileup_channel_lib.cc:91-261
//     GetHapIndex / SortImageRows      :449-475, :75-102
//     EncodeReference                  :512-527,  pileup_channel_lib.cc:263-293
//   FillPileupArray (planar -> HWC)    deepvariant/pileup_image_native.h:214-308
//
// Formulation (integer / byte work, HBM-write bound: H*W*C bytes out per image):
//   one CTA per image (grid-stride), 8 warps.
//   phase A  thread-per-read acceptance test (mapq, low-quality base at variant.start found by a
//            scalar CIGAR walk), block scan -> the first (H - band) accepted reads in visit order;
//   phase S  rank sort of <= (H - band) keys (hap, allele group, position, name rank, visit index);
//   phase B  warp-per-row: the row is assembled in shared memory at the SAME 16-byte phase as its
//            global address (CIGAR ops in order, __syncwarp between ops => last-writer-wins is
//            preserved), then streamed out with 16-byte coalesced stores; blank rows are stored
//            straight from registers.
// Colour arithmetic that the reference does in float32 per pixel is tabulated on the host with
// the identical float expressions (256-entry LUTs) or done once per row with IEEE _rn intrinsics.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <random>
#include <string>
#include <vector>

#include "dvb_allele_walk.h"
#include "dvb_common.h"

namespace dvb {
std::string& last_error() {
  static thread_local std::string e;
  return e;
}
}  // namespace dvb

namespace {

#ifndef DVB_ENC_THREADS
#define DVB_ENC_THREADS 128
#endif
constexpr int kThreads = DVB_ENC_THREADS;   // threads per image CTA
// Resident CTAs per SM the register allocation must allow.  The kernel is bound by dependent global-load chains
// (record -> CIGAR -> bases).  Measured on B200, 16,384 windows per launch pair (profiles/r02e_encoder_variants.json; A/B of prebuilt
// libraries of the same sources through DVB_LIB_PATH): 1 -> 1.447 ms (96 registers, 2 CTAs resident), 2 -> 0.924 (62 registers, 4 CTAs),
// 3 -> 0.941 (56), 4 -> 0.967 (56: the same count, a more constrained schedule), 5 -> 0.957 (46, 5 CTAs).  The freer allocation wins at
// the same occupancy (256-thread CTAs; minimum blocks counted for that size).
// Then the CTA size (calls 36-37, same harness): 128 threads with at least 8 resident CTAs (56 registers, 9 CTAs = 36 warps) 0.846 ms
// against 0.873 for 256 threads x 2; 64 threads 0.930, 512 threads 1.100, 128 x 10 (46 registers) 0.918, 128 x 12 (40, spills) 0.874,
// two 4-pixel groups per lane and iteration 0.901.  Smaller CTAs = more images in flight per SM and cheaper block barriers.
#ifndef DVB_ENC_MIN_BLOCKS
#define DVB_ENC_MIN_BLOCKS 8
#endif
// The 4 bases / 4 qualities of a 4-pixel group as aligned word loads + a funnel shift instead of 8 byte loads: 0.924 -> 0.877 ms.
#ifndef DVB_ENC_WORD_LOADS
#define DVB_ENC_WORD_LOADS 1
#endif
// L2 prefetch of every accepted row's CIGAR and window part of its bases / qualities in phase A: 0.847 -> 0.834 ms (WGS), 0.971 -> 0.934
// (PACBIO layout) per 16,384 windows (call 43).
#ifndef DVB_ENC_PREFETCH
#define DVB_ENC_PREFETCH 1
#endif
// (measured and dropped: asking the NEXT image's pair records into L2 during phase B - 0.867 ms against 0.836; rows handed out to the
// warps one by one through a shared counter instead of round-robin - 0.854 against 0.834)
// Measured and dropped in the same series: the blank tail of an image zeroed as one contiguous span before the read rows (0.931 ms:
// the early burst of stores competes with the rows' loads), the row's record and first CIGAR word requested before its buffer is
// cleared (0.895), a grid-stride pre-pass that fills the resident slots once (0.895: fewer, longer-lived warps lose to more CTAs).
constexpr int kWarps = kThreads / 32;
constexpr float kMaxPixelValueAsFloat = 254.0f;  // channels/channel.h:78
constexpr float kMaxFragmentLength = 1000.0f;    // channels/channel.h:81

enum Kind : uint8_t { K_ZERO = 0, K_BASE = 1, K_QUAL = 2, K_DIFF = 3, K_CONST = 4, K_ISHOMO = 5, K_HOMOW = 6, K_PLANE = 7 };

struct EncDev {
  int W, H, band, C, Cout, max_rows;
  int row_bytes;
  long long image_bytes;
  int min_bq, min_mq, anchor_char;
  int sort_by_hap, polish_tag, sort_by_group;
  float mapq_cap;
  int pos_strand, neg_strand;
  uint8_t sup_color[4];     // by support class 0/1/2
  uint8_t match_color, mismatch_color;
  uint8_t supp_color[2];    // supplementary_alignment false/true
  uint8_t kind[DVB_MAX_CHANNELS];
  uint8_t chan[DVB_MAX_CHANNELS];       // channel enum
  uint8_t ref_const[DVB_MAX_CHANNELS];  // reference-row value (K_BASE uses base_lut)
  // A pixel is assembled as up to 4 little-endian words: const + base*mask_base + qual*mask_qual +
  // diff*mask_diff, where a mask has 0x01 in the byte of every channel of that kind (no carries:
  // every term is < 256 and the byte positions of different kinds are disjoint).
  unsigned mask_base[4], mask_qual[4], mask_diff[4], mask_const[4], ref_words[4];
  unsigned mask_ishomo[4], mask_homow[4];   // per-base homopolymer channels (is_homopolymer, homopolymer_weighted)
  // the same masks as the reference band uses them: channels_enum_to_blank (blank_channel_mask) clears a channel from the READ masks only
  unsigned ref_mask_base[4], ref_mask_ishomo[4], ref_mask_homow[4];
  int has_homo;
  // per-base channel planes (DvbBatch.base_channel): channel index and plane slot of each K_PLANE channel
  int n_planes;
  uint8_t plane_chan[DVB_N_BASE_PLANES], plane_slot[DVB_N_BASE_PLANES];
  int n_pair_planes;        // per-pair channel planes (DvbBatch.pair_channel), kind K_ZERO in the pixel loop: their bytes ride in tc[]
  uint8_t pair_plane_chan[DVB_N_PAIR_PLANES], pair_plane_slot[DVB_N_PAIR_PLANES];
  // mean_coverage overlay (pileup_image_native.cc:422-444): channel index or -1; rows [0, cov_rows) are painted
  int cov_chan, cov_rows;
  int n_words;
  int gc_chan;              // index of the gc_content channel (its reference-band value is per image), or -1
  int perm_cap;             // largest n with a down-sampling table
  const int* perm;          // tables for n = max_rows+1 .. perm_cap, concatenated
  uint8_t base_lut[256];    // BaseColor(char)
  uint8_t qual_lut[256];    // ScaleColor(q, base_quality_cap)
};

__device__ __forceinline__ long long perm_offset(int n, int max_rows) {
  // sum_{k=max_rows+1}^{n-1} k
  return (long long)(n - 1) * n / 2 - (long long)max_rows * (max_rows + 1) / 2;
}

// ScaleColor (channels/mapping_quality_channel.cc:60-67) with IEEE round-to-nearest ops.
__device__ __forceinline__ uint8_t scale_color_dev(int value, float max_val) {
  float v = (float)value;
  if (v > max_val) { value = (int)max_val; v = (float)value; }
  return (uint8_t)(int)__fmul_rn(kMaxPixelValueAsFloat, __fdiv_rn(v, max_val));
}

struct ReadHdr {
  int pos, mapq, fraglen, hp;
  unsigned flags;
  long long seq0, cig0;
  int n_cig;
  int len;     // aligned_sequence length
};

__device__ __forceinline__ ReadHdr load_read(const DvbBatch& B, int r) {
  ReadHdr h;
  h.pos = B.read_pos[r];
  h.mapq = B.read_mapq[r];
  h.fraglen = B.read_fragment_length[r];
  h.hp = B.read_hp[r];
  h.flags = B.read_flags[r];
  h.seq0 = B.read_seq_begin[r];
  h.cig0 = B.read_cigar_begin[r];
  h.n_cig = (int)(B.read_cigar_begin[r + 1] - h.cig0);
  h.len = (int)(B.read_seq_begin[r + 1] - h.seq0);
  return h;
}

// EncodeRead's keep/drop decision (pileup_image_native.cc:485-491 + the low-quality-base-at-call-site
// bail-out of pileup_channel_lib.cc:144-150), evaluated without drawing.  Returns 1 keep, 0 drop,
// -1 unrecognized CIGAR op.
__device__ int accept_read(const EncDev& P, const DvbBatch& B, int r, int vstart, int image_start) {
  const ReadHdr h = load_read(B, r);
  if (h.mapq < P.min_mq) return 0;
  const unsigned vcol = (unsigned)(vstart - image_start);
  const bool v_in_window = vcol < (unsigned)P.W;
  const uint8_t* bases = B.bases + h.seq0;
  const uint8_t* quals = B.quals + h.seq0;
  int ref_i = h.pos, read_i = 0;
  for (int k = 0; k < h.n_cig; ++k) {
    const unsigned cw = B.cigar[h.cig0 + k];
    const int op = cw & 0xF, len = (int)(cw >> 4);
    switch (op) {
      case 0: case 7: case 8:
        if (v_in_window && vstart >= ref_i && vstart < ref_i + len) {
          const int ri = read_i + (vstart - ref_i);
          if (bases[ri] != 0 && (int)quals[ri] < P.min_bq) return 0;
        }
        ref_i += len; read_i += len;
        break;
      case 1:
        if (ref_i > 0 && P.anchor_char && v_in_window && ref_i - 1 == vstart &&
            (int)quals[read_i] < P.min_bq) return 0;
        read_i += len;
        break;
      case 4:
        read_i += len;
        break;
      case 2:
        if (read_i > 0 && P.anchor_char && v_in_window && ref_i - 1 == vstart &&
            (int)quals[read_i - 1] < P.min_bq) return 0;
        ref_i += len;
        break;
      case 3:
        ref_i += len;
        break;
      case 5: case 6:
        break;
      default:
        return -1;
    }
  }
  return 1;
}

// GetHapIndex, pileup_image_native.cc:449-475.
__device__ __forceinline__ int hap_index(const EncDev& P, unsigned flags, int hp) {
  if (!P.sort_by_hap || !(flags & DVB_READ_HAS_HP)) return 0;
  if (P.polish_tag > 0 && hp == P.polish_tag) return -1;
  if (hp < 0) return 0;
  return hp;
}

// Per-read constant of a K_CONST channel (channels/*_channel.cc FillReadBase).
// Per-base homopolymer channels at sequence index i of seq[0, len) (channels/is_homopolymer_channel.cc:77-105,
// channels/homopolymer_weighted_channel.cc:79-115).  *ih = 254 inside a run of >= 3 equal bases else 0;
// *hw = ScaleColor(uint8(run length), 30).
__device__ __forceinline__ void homopolymer_at(const uint8_t* seq, int len, int i, bool want_weighted, unsigned* ih, unsigned* hw) {
  const unsigned b = seq[i];
  int l = i, r = i;
  const int reach = want_weighted ? len : 2;          // is_homopolymer only needs to know whether the run reaches 3
  while (l > 0 && i - l < reach && seq[l - 1] == b) --l;
  while (r + 1 < len && r - i < reach && seq[r + 1] == b) ++r;
  const int run = r - l + 1;
  *ih = run >= 3 ? 254u : 0u;
  *hw = want_weighted ? (unsigned)scale_color_dev(run & 0xFF, 30.0f) : 0u;
}

// percent = int(float(a) / float(b) * 100) with the reference's operation order (no FMA contraction)
__device__ __forceinline__ int percent_dev(int a, int b) { return (int)__fmul_rn(__fdiv_rn((float)a, (float)b), 100.0f); }

// Bytes of the per-base channel planes at base `idx` (absolute index into bases[]), placed at their channels' byte positions.
__device__ __forceinline__ void plane_words(const EncDev& P, const DvbBatch& B, long long idx, unsigned ex[4]) {
  for (int k = 0; k < P.n_planes; ++k) {
    const int c = P.plane_chan[k];
    ex[c >> 2] += (unsigned)B.base_channel[P.plane_slot[k]][idx] << (8 * (c & 3));
  }
}

// Bytes of the per-pair channel planes of pair p - values the caller computed per (image, read): allele_frequency_channel.cc:57-68,
// read_supports_variant_fuzzy_channel.cc:99-110, allele_sample_probability_channel.cc:48-79 (functions of DeepVariantCall's
// string-keyed maps) - placed like the K_CONST bytes.
__device__ __forceinline__ void pair_plane_words(const EncDev& P, const DvbBatch& B, long long p, unsigned tc[4]) {
  for (int k = 0; k < P.n_pair_planes; ++k) {
    const int c = P.pair_plane_chan[k];
    tc[c >> 2] |= (unsigned)B.pair_channel[P.pair_plane_slot[k]][p] << (8 * (c & 3));
  }
}

__device__ uint8_t read_const(const EncDev& P, const DvbBatch& B, int chan, const ReadHdr& h, int support) {
  switch (chan) {
    // ---- "Opt Channels": whole-read statistics (channels/{read_mapping_percent,identity,avg_base_quality,
    //      gap_compressed_identity,gc_content}_channel.cc), one value per read
    case DVB_CH_READ_MAPPING_PERCENT:
    case DVB_CH_IDENTITY: {
      int match_len = 0;
      for (int k = 0; k < h.n_cig; ++k) {
        const unsigned cw = B.cigar[h.cig0 + k];
        const unsigned op = cw & 0xF;
        if (op == 0 || op == 7) match_len += (int)(cw >> 4);
      }
      return scale_color_dev(percent_dev(match_len, h.len), 100.0f);
    }
    case DVB_CH_AVG_BASE_QUALITY: {
      int sum = 0;
      for (int i = 0; i < h.len; ++i) sum += B.quals[h.seq0 + i];
      return scale_color_dev((int)__fdiv_rn((float)sum, (float)h.len), 93.0f);
    }
    case DVB_CH_GAP_COMPRESSED_IDENTITY: {
      int match_len = 0, gap_len = 0;
      for (int k = 0; k < h.n_cig; ++k) {
        const unsigned cw = B.cigar[h.cig0 + k];
        const unsigned op = cw & 0xF;
        const int len = (int)(cw >> 4);
        if (op == 0 || op == 7) { match_len += len; gap_len += len; }
        else if (op == 8) gap_len += len;
        else if (op == 1 || op == 2) gap_len += 1;
      }
      return scale_color_dev(percent_dev(match_len, gap_len), 100.0f);
    }
    case DVB_CH_GC_CONTENT: {
      int gc = 0;
      for (int i = 0; i < h.len; ++i) { const unsigned b = B.bases[h.seq0 + i]; gc += (b == 'G' || b == 'C') ? 1 : 0; }
      return scale_color_dev(percent_dev(gc, h.len), 100.0f);
    }
    case DVB_CH_MAPPING_QUALITY:
      return scale_color_dev(h.mapq, P.mapq_cap);
    case DVB_CH_STRAND:
      return (uint8_t)((h.flags & DVB_READ_REVERSE_STRAND) ? P.neg_strand : P.pos_strand);
    case DVB_CH_READ_SUPPORTS_VARIANT:
      return P.sup_color[support > 2 ? 2 : support];
    case DVB_CH_HAPLOTYPE_TAG: {  // haplotype_tag_channel.cc:74-109
      int v = 0;
      if ((h.flags & DVB_READ_HAS_HP) && !(h.flags & DVB_READ_HP_MULTI)) {
        v = h.hp;
        if (P.polish_tag == 2) { if (v == 1) v = 2; else if (v == 2) v = 1; }
      }
      return scale_color_dev(v, 2.0f);
    }
    case DVB_CH_INSERT_SIZE: {  // insert_size_channel.cc:80-89
      int f = h.fraglen < 0 ? -h.fraglen : h.fraglen;
      if ((float)f > kMaxFragmentLength) f = (int)kMaxFragmentLength;
      return (uint8_t)(int)__fmul_rn(kMaxPixelValueAsFloat, __fdiv_rn((float)f, kMaxFragmentLength));
    }
    case DVB_CH_SUPPLEMENTARY_ALIGNMENT:
      return P.supp_color[(h.flags & DVB_READ_SUPPLEMENTARY) ? 1 : 0];
    default:
      return 0;
  }
}

// ---- pre-pass: everything about a (image, read) pair that does not need the image's other reads --------------------------
// One record per pair, written by a fully parallel kernel (a warp per image, a lane per pair): EncodeRead's keep/drop
// decision, the per-read constant channel words, the sort key parts and the read's addresses.  The image kernel then reads
// ONE 64-byte record per pair (phase A) and per row (phase B) instead of chasing pair -> read header (8 arrays) -> CIGAR
// -> bases through dependent global loads, and no longer runs the divergent per-channel constant code per row.
struct __align__(16) PairRec {
  long long seq0, cig0;     // offsets of the read's bases / quals and CIGAR
  int pos, n_cig;
  unsigned tc[4];           // K_CONST channel bytes of this (read, support class), packed like a pixel
  int sort_pos;
  unsigned rank;
  int hap;
  uint8_t grp, ok, pad0, pad1;
  int len, pad3;            // sequence length (per-base homopolymer channels); (tried: first two CIGAR words inline here - 54 registers, slower)
};
static_assert(sizeof(PairRec) == 64, "PairRec is one 64-byte line");

// ---- (candidate, read) support on the device ---------------------------------------------------------------------------------
// The read allele AlleleCounter::Add records for read `h` at the image's variant_start (dvb_allele::ElementAt), matched against the
// image's alt-allele keys.  Returns bit 7 = the read holds an entry in AlleleCount.read_alleles at that site, bits 0-1 = the
// ReadSupportsAlt class of that entry; *group = the matched alt's index.
constexpr unsigned kSupportHasEntry = 0x80u;

__device__ __forceinline__ unsigned pair_read_allele(const DvbBatch& B, const ReadHdr& h, int img, int vstart, uint8_t ref_base, int ref_run,
                                                     long long a0, int n_alleles, uint8_t* group) {
  if (h.mapq < B.support_min_mapping_quality) return 0u;             // AlleleCounter::Add's first test
  dvb_allele::ReadView r;
  r.seq = B.bases + h.seq0; r.qual = B.quals + h.seq0; r.seq_len = h.len;
  r.cigar = B.cigar + h.cig0; r.n_cigar = h.n_cig; r.pos = h.pos;
  dvb_allele::TargetParams tp;
  tp.target = vstart; tp.ref_base = ref_base; tp.ref_run = ref_run;
  tp.min_base_quality = B.support_min_base_quality;
  tp.keep_legacy_behavior = (B.support_flags & DVB_SUPPORT_KEEP_LEGACY) ? 1 : 0;
  dvb_allele::TargetElement e;
  if (!dvb_allele::ElementAt(r, tp, &e)) return 0u;
  if (e.type == dvb_allele::kReference) return (B.support_flags & DVB_SUPPORT_TRACK_REF_READS) ? kSupportHasEntry : 0u;
  if (e.type == dvb_allele::kSoftClip) return kSupportHasEntry;       // never an alt allele (IsGoodAltAlleleWithReason)
  const int key_len = e.len + 1;
  for (int a = 0; a < n_alleles; ++a) {
    if (B.allele_type[a0 + a] != e.type) continue;
    const long long b0 = B.allele_bases_begin[a0 + a];
    if ((int)(B.allele_bases_begin[a0 + a + 1] - b0) != key_len) continue;
    const uint8_t* key = B.allele_bases + b0;
    bool same;
    if (e.type == dvb_allele::kSubstitution) {
      same = key[0] == r.seq[e.read_offset];
    } else {
      same = key[0] == e.prev;
      if (e.type == dvb_allele::kInsertion)                           // deleted bases are the reference's: equal lengths, equal bases
        for (int k = 0; same && k < e.len; ++k) same = key[1 + k] == r.seq[e.read_offset + k];
    }
    if (same) {
      if (B.allele_group) *group = B.allele_group[a0 + a];
      return kSupportHasEntry | B.allele_class[a0 + a];
    }
  }
  return kSupportHasEntry;                                            // "UNCALLED_ALLELE"
}

// sup (when the batch carries allele keys): uint8[4][n_pairs] = raw class byte, raw group, final class, final group.
#ifndef DVB_ENC_PREPASS_THREADS
#define DVB_ENC_PREPASS_THREADS 128
#endif
constexpr int kPreThreads = DVB_ENC_PREPASS_THREADS, kPreWarps = kPreThreads / 32;   // a warp per image
template <bool kPairPlanes>   // the batch carries per-pair channel planes (a separate instantiation keeps the common layouts' registers)
__global__ void __launch_bounds__(kPreThreads) dvb_pair_prepass_kernel(const EncDev P, const DvbBatch B, PairRec* __restrict__ recs, int* __restrict__ err,
                                                                uint8_t* __restrict__ sup) {
  const int lane = threadIdx.x & 31;
  const int img = blockIdx.x * kPreWarps + (threadIdx.x >> 5);
  if (img >= B.n_images) return;
  const int image_start = B.image_start_pos[img];
  const int vstart = B.variant_start[img];
  const long long p0 = B.pair_begin[img];
  const int n = (int)(B.pair_begin[img + 1] - p0);
  const bool derive = B.allele_begin != nullptr;
  const bool repeated = derive && (B.support_flags & DVB_SUPPORT_REPEATED_KEYS);
  const uint8_t group_default = derive && B.image_group_default ? B.image_group_default[img] : (uint8_t)0;
  if (derive) {
    const long long a0 = B.allele_begin[img];
    const int n_alleles = (int)(B.allele_begin[img + 1] - a0);
    const int off = vstart - image_start;
    const uint8_t ref_base = (off >= 0 && off < P.W) ? B.ref_bases[(size_t)img * B.ref_stride + off] : (uint8_t)0;
    const int ref_run = B.image_ref_run[img];
    for (int i = lane; i < n; i += 32) {
      const long long p = p0 + i;
      const ReadHdr h = load_read(B, B.pair_read[p]);
      uint8_t g = group_default;
      const unsigned raw = pair_read_allele(B, h, img, vstart, ref_base, ref_run, a0, n_alleles, &g);
      sup[p] = (uint8_t)raw;
      sup[B.n_pairs + p] = g;
    }
    __syncwarp();
  }
  for (int i = lane; i < n; i += 32) {
    const long long p = p0 + i;
    const int r = B.pair_read[p];
    const ReadHdr h = load_read(B, r);
    int ok = accept_read(P, B, r, vstart, image_start);
    if (ok < 0) { atomicMax(err, DVB_ERR_BAD_CIGAR); ok = 0; }
    PairRec rec;
    rec.seq0 = h.seq0; rec.cig0 = h.cig0; rec.pos = h.pos; rec.n_cig = h.n_cig;
    rec.tc[0] = rec.tc[1] = rec.tc[2] = rec.tc[3] = 0u;
    int support;
    uint8_t grp;
    if (derive) {
      // AlleleCount.read_alleles is a map keyed by (fragment_name, read_number): of the image's reads that share this one's key,
      // the last one that holds an entry decides for all of them
      long long q = p;
      if (repeated) {
        const unsigned rank = B.read_name_rank[r];
        q = -1;
        for (int k = n - 1; k >= 0; --k)
          if ((sup[p0 + k] & kSupportHasEntry) && B.read_name_rank[B.pair_read[p0 + k]] == rank) { q = p0 + k; break; }
      }
      support = q >= 0 ? (int)(sup[q] & 3u) : 0;
      grp = q >= 0 ? sup[B.n_pairs + q] : group_default;
      if (q >= 0 && !(sup[q] & kSupportHasEntry)) grp = group_default;
      sup[2 * B.n_pairs + p] = (uint8_t)support;
      sup[3 * B.n_pairs + p] = grp;
    } else {
      support = B.pair_support[p];
      grp = B.pair_allele_group ? B.pair_allele_group[p] : (uint8_t)0;
    }
    for (int c = 0; c < P.C; ++c)
      if (P.kind[c] == K_CONST) rec.tc[c >> 2] |= (unsigned)read_const(P, B, P.chan[c], h, support) << (8 * (c & 3));
    if constexpr (kPairPlanes) pair_plane_words(P, B, p, rec.tc);
    rec.sort_pos = B.read_sort_pos[r];
    rec.rank = B.read_name_rank[r];
    rec.hap = hap_index(P, h.flags, h.hp);
    rec.grp = P.sort_by_group ? grp : (uint8_t)0;
    rec.ok = (uint8_t)ok; rec.pad0 = rec.pad1 = 0; rec.len = h.len; rec.pad3 = 0;
    uint4* dst = reinterpret_cast<uint4*>(recs + p);
    const uint4* srcv = reinterpret_cast<const uint4*>(&rec);
    dst[0] = srcv[0]; dst[1] = srcv[1]; dst[2] = srcv[2]; dst[3] = srcv[3];
  }
}

// Streams `row_bytes` bytes of a row to global memory with 16-byte stores.
//   buf == nullptr : blank row, zeros straight from registers.
//   otherwise      : the row lives in shared memory at px = buf + 16 + 4 * (phase >> 2), phase = dst & 15, i.e. pixel 0 is
//                    always 4-byte aligned (so that 4-pixel groups can be stored as aligned words) and the 0-3 byte residual
//                    misalignment s = phase & 3 against the 16-byte global chunks is removed here with one funnel shift per word.
__device__ __forceinline__ void flush_row(uint8_t* __restrict__ dst, const uint8_t* buf, int row_bytes, int lane) {
  const int phase = (int)((uintptr_t)dst & 15);
  int head = (16 - phase) & 15;
  if (head > row_bytes) head = row_bytes;
  const int body = (row_bytes - head) >> 4;
  const int done = head + (body << 4);
  const int tail = row_bytes - done;
  uint4* d4 = reinterpret_cast<uint4*>(dst + head);
  if (buf == nullptr) {
    const uint4 z = make_uint4(0, 0, 0, 0);
    if (lane < head) dst[lane] = 0;
    for (int i = lane; i < body; i += 32) d4[i] = z;
    if (lane < tail) dst[done + lane] = 0;
    return;
  }
  const uint8_t* px = buf + 16 + 4 * (phase >> 2);
  if (lane < head) dst[lane] = px[lane];
  const int s = phase & 3;
  const uint8_t* blk = buf + (phase == 0 ? 16 : 32);     // aligned block holding (most of) body chunk 0
  if (s == 0) {
    for (int i = lane; i < body; i += 32) d4[i] = *reinterpret_cast<const uint4*>(blk + 16 * i);
  } else {
    const int sh = 32 - 8 * s;
    for (int i = lane; i < body; i += 32) {
      const uint4 b = *reinterpret_cast<const uint4*>(blk + 16 * i);
      const uint32_t wp = *reinterpret_cast<const uint32_t*>(blk + 16 * i - 4);
      d4[i] = make_uint4(__funnelshift_r(wp, b.x, sh), __funnelshift_r(b.x, b.y, sh), __funnelshift_r(b.y, b.z, sh),
                         __funnelshift_r(b.z, b.w, sh));
    }
  }
  if (lane < tail) dst[done + lane] = px[done + lane];
}

// FAST7: 7 computed channels, 7-byte pixels (the WGS layout): runs of 4 pixels whose first column is a multiple of 4 are
// assembled in registers and stored as 7 aligned 32-bit words instead of 28 byte stores.
// HOMO: per-base extra channels are present - the homopolymer channels and / or channel planes (a separate instantiation keeps the
// common layouts at 48 registers).
template <bool FAST7, bool HOMO>
__global__ void __launch_bounds__(kThreads, DVB_ENC_MIN_BLOCKS)
dvb_encode_kernel(const EncDev P, const DvbBatch B, uint8_t* __restrict__ out, int* __restrict__ rows_kept,
                  int* __restrict__ err, const PairRec* __restrict__ recs) {
  extern __shared__ __align__(16) uint8_t smem[];
  // layout: [base_lut 256][qual_lut 256][ref Wpad][sel arrays 6 x max_rows x 4B][order max_rows x 4B]
  //         [warp totals][row buffers kWarps x rowbuf_bytes]
  const int Wpad = (P.W + 15) & ~15;
  uint8_t* s_base = smem;
  uint8_t* s_qual = smem + 256;
  uint8_t* s_ref = smem + 512;
  int* s_pair = reinterpret_cast<int*>(s_ref + Wpad);
  int* s_hap = s_pair + P.max_rows;
  int* s_grp = s_hap + P.max_rows;
  int* s_pos = s_grp + P.max_rows;
  unsigned* s_rank = reinterpret_cast<unsigned*>(s_pos + P.max_rows);
  int* s_visit = reinterpret_cast<int*>(s_rank + P.max_rows);
  int* s_order = s_visit + P.max_rows;
  int* s_wtot = s_order + P.max_rows;          // kWarps + 2 ints
  const int rowbuf_bytes = ((P.row_bytes + 28 + 15) & ~15) + 32;
  // aligned up to 16 bytes by POINTER arithmetic: an integer round trip loses the shared address space and every pixel store below
  // becomes a generic ST instead of STS
  uint8_t* s_rows_raw = reinterpret_cast<uint8_t*>(s_wtot + kWarps + 2);
  uint8_t* s_rows = s_rows_raw + ((16u - ((unsigned)__cvta_generic_to_shared(s_rows_raw) & 15u)) & 15u);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 256; i += kThreads) {
    s_base[i] = P.base_lut[i];
    s_qual[i] = P.qual_lut[i];
  }

  for (int img = blockIdx.x; img < B.n_images; img += gridDim.x) {
    __syncthreads();  // previous image fully flushed before smem is reused
    const int image_start = B.image_start_pos[img];
    const int vstart = B.variant_start[img];
    const long long p0 = B.pair_begin[img];
    const int n = (int)(B.pair_begin[img + 1] - p0);
    for (int i = tid; i < Wpad; i += kThreads)
      s_ref[i] = i < P.W ? B.ref_bases[(long long)img * B.ref_stride + i] : (uint8_t)0;
    if (tid == 0) { s_wtot[kWarps] = 0; }

    // ---- phase A: acceptance + first max_rows accepted in visit order --------------------
    const bool shuffled = n > P.max_rows;
    const int* perm = nullptr;
    if (shuffled) {
      if (n > P.perm_cap) { if (tid == 0) atomicMax(err, DVB_ERR_TOO_MANY_READS); }
      else perm = P.perm + perm_offset(n, P.max_rows);
    }
    __syncthreads();
    if (P.gc_chan >= 0 && warp == 0) {   // gc_content reference band: GC content of this image's window (gc_content_channel.cc:64-75)
      int gc = 0;
      for (int i = lane; i < P.W; i += 32) gc += (s_ref[i] == 'G' || s_ref[i] == 'C') ? 1 : 0;
      for (int o = 16; o > 0; o >>= 1) gc += __shfl_xor_sync(0xffffffffu, gc, o);
      if (lane == 0) s_wtot[kWarps + 1] = (int)scale_color_dev(percent_dev(gc, P.W), 100.0f);
    }
    int n_acc = 0;
    for (int base = 0; base < n && n_acc < P.max_rows; base += kThreads) {
      const int i = base + tid;
      int ok = 0;
      long long p = 0;
      if (i < n) {
        p = p0 + (perm ? perm[i] : i);
        if (recs) {
          ok = recs[p].ok;
        } else {
          ok = accept_read(P, B, B.pair_read[p], vstart, image_start);
          if (ok < 0) { atomicMax(err, DVB_ERR_BAD_CIGAR); ok = 0; }
        }
      }
      const unsigned m = __ballot_sync(0xffffffffu, ok);
      if (lane == 0) s_wtot[warp] = __popc(m);
      __syncthreads();
      int before = n_acc;
      for (int w = 0; w < warp; ++w) before += s_wtot[w];
      int total = 0;
      for (int w = 0; w < kWarps; ++w) total += s_wtot[w];
      const int slot = before + __popc(m & ((1u << lane) - 1u));
      if (ok && slot < P.max_rows) {
        s_pair[slot] = (int)(p - p0);
        s_visit[slot] = i;
        if (recs) {
          const PairRec& rc = recs[p];
          s_hap[slot] = rc.hap; s_grp[slot] = rc.grp; s_pos[slot] = rc.sort_pos; s_rank[slot] = rc.rank;
#if DVB_ENC_PREFETCH
          // the row's CIGAR and the part of its bases / qualities that can land in the window are asked into L2 now, a sort and several
          // rows before phase B reads them (the row's dependent chain record -> CIGAR -> bases then runs on L2 hits)
          {
            const long long off = image_start > rc.pos ? (long long)(image_start - rc.pos) : 0;
            long long a = rc.seq0 + (off > 32 ? off - 32 : 0);
            const long long a_end = a + P.W + 96 < B.n_bases ? a + P.W + 96 : B.n_bases;
            for (; a < a_end; a += 128) {
              asm volatile("prefetch.global.L2 [%0];" ::"l"(B.bases + a));
              asm volatile("prefetch.global.L2 [%0];" ::"l"(B.quals + a));
            }
            asm volatile("prefetch.global.L2 [%0];" ::"l"(B.cigar + rc.cig0));
          }
#endif
        } else {
          const int r = B.pair_read[p];
          const unsigned fl = B.read_flags[r];
          s_hap[slot] = hap_index(P, fl, B.read_hp[r]);
          s_grp[slot] = (P.sort_by_group && B.pair_allele_group) ? (int)B.pair_allele_group[p] : 0;
          s_pos[slot] = B.read_sort_pos[r];
          s_rank[slot] = B.read_name_rank[r];
        }
      }
      n_acc += total;
      __syncthreads();
    }
    const int n_rows = n_acc < P.max_rows ? n_acc : P.max_rows;
    if (tid == 0 && rows_kept) rows_kept[img] = n_rows;

    // ---- phase S: stable sort by (hap, group, pos, name rank); visit index breaks ties -----
    for (int e = tid; e < n_rows; e += kThreads) {
      const int h = s_hap[e], g = s_grp[e], ps = s_pos[e], v = s_visit[e];
      const unsigned rk = s_rank[e];
      int less = 0;
      for (int f = 0; f < n_rows; ++f) {
        const int h2 = s_hap[f], g2 = s_grp[f], p2 = s_pos[f], v2 = s_visit[f];
        const unsigned r2 = s_rank[f];
        bool lt;
        if (h2 != h) lt = h2 < h;
        else if (g2 != g) lt = g2 < g;
        else if (p2 != ps) lt = p2 < ps;
        else if (r2 != rk) lt = r2 < rk;
        else lt = v2 < v;
        less += lt ? 1 : 0;
      }
      s_order[less] = e;
    }
    __syncthreads();

    // ---- phase B: one warp per image row ---------------------------------------------------
    uint8_t* img_out = out + (long long)img * P.image_bytes;
    for (int row = warp; row < P.H; row += kWarps) {
      uint8_t* dst = img_out + (long long)row * P.row_bytes;
      const bool cov_row = P.cov_chan >= 0 && row < P.cov_rows;   // mean_coverage is painted over finished rows, blank ones included
      if (row >= P.band + n_rows && !cov_row) {  // blank tail (pileup_image_native.cc:414-421)
        flush_row(dst, nullptr, P.row_bytes, lane);
        continue;
      }
      uint8_t* buf = s_rows + warp * rowbuf_bytes;
      const int phase = (int)((uintptr_t)dst & 15);
      uint8_t* px = buf + 16 + 4 * (phase >> 2);  // pixel (col, ch) lives at px[col * Cout + ch]; px is 4-byte aligned
      {
        uint4* b4 = reinterpret_cast<uint4*>(buf);
        const uint4 z = make_uint4(0, 0, 0, 0);
        for (int i = lane; i < (rowbuf_bytes >> 4); i += 32) b4[i] = z;
      }
      __syncwarp();
      if (row < P.band) {
        // EncodeReference / CalculateRefRows
        const unsigned gc_word = P.gc_chan >= 0 ? (unsigned)s_wtot[kWarps + 1] << (8 * (P.gc_chan & 3)) : 0u;
        for (int col = lane; col < P.W; col += 32) {
          const unsigned bc = s_base[s_ref[col]];
          unsigned ih = 0u, hw = 0u;
          if constexpr (HOMO) { if (P.has_homo) homopolymer_at(s_ref, P.W, col, P.has_homo > 1, &ih, &hw); }   // the window treated as a read
          uint8_t* q = px + col * P.Cout;
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            if (w < P.n_words) {
              const unsigned word = P.ref_words[w] + bc * P.ref_mask_base[w] + ((P.gc_chan >> 2) == w ? gc_word : 0u) +
                                    (HOMO ? ih * P.ref_mask_ishomo[w] + hw * P.ref_mask_homow[w] : 0u);
#pragma unroll
              for (int b = 0; b < 4; ++b)
                if (w * 4 + b < P.C) q[w * 4 + b] = (uint8_t)(word >> (8 * b));
            }
          }
        }
      } else if (row < P.band + n_rows) {
        const int e = s_order[row - P.band];
        const long long p = p0 + s_pair[e];
        ReadHdr h;
        unsigned tc[4] = {0u, 0u, 0u, 0u};
        if (recs) {   // one 64-byte record (all lanes read the same line)
          const uint4* rv = reinterpret_cast<const uint4*>(recs + p);
          const uint4 q0 = rv[0], q1 = rv[1];
          h.seq0 = (long long)(((unsigned long long)q0.y << 32) | q0.x);
          h.cig0 = (long long)(((unsigned long long)q0.w << 32) | q0.z);
          h.pos = (int)q1.x; h.n_cig = (int)q1.y;
          tc[0] = q1.z; tc[1] = q1.w;
          if (P.n_words > 2) { const uint4 q2 = rv[2]; tc[2] = q2.x; tc[3] = q2.y; }
          h.mapq = 0; h.fraglen = 0; h.hp = 0; h.flags = 0; h.len = 0;
          if constexpr (HOMO) h.len = (int)rv[3].z;   // (seq0 above is what the channel planes are indexed from)
        } else {
          const int r = B.pair_read[p];
          const int support = B.pair_support[p];
          h = load_read(B, r);
          // per-read constants: lane c computes channel c, then every lane gathers them into words
          unsigned myc = 0;
          if (lane < P.C && P.kind[lane] == K_CONST) myc = read_const(P, B, P.chan[lane], h, support);
#pragma unroll
          for (int c = 0; c < DVB_MAX_CHANNELS; ++c) {
            const unsigned v = __shfl_sync(0xffffffffu, myc, c);
            tc[c >> 2] |= (v & 0xFFu) << (8 * (c & 3));
          }
          if (P.n_pair_planes) pair_plane_words(P, B, p, tc);
        }
        const uint8_t* bases = B.bases + h.seq0;
        const uint8_t* quals = B.quals + h.seq0;
        int ref_i = h.pos, read_i = 0;
        for (int k = 0; k < h.n_cig; ++k) {
          const unsigned cw = B.cigar[h.cig0 + k];
          const int op = cw & 0xF, len = (int)(cw >> 4);
          if (op == 0 || op == 7 || op == 8) {
            const int c0 = ref_i - image_start;
            const int lo = c0 < 0 ? -c0 : 0;
            const int hi = P.W - c0 < len ? P.W - c0 : len;
            auto put_pixel = [&](int j) {      // byte-granular path: one pixel
              const int col = c0 + j;
              const unsigned b = bases[read_i + j];
              if (b != 0) {
                const unsigned bc = s_base[b];
                const unsigned ql = s_qual[quals[read_i + j]];
                const unsigned df = (b == s_ref[col]) ? P.match_color : P.mismatch_color;
                unsigned ih = 0u, hw = 0u;
                unsigned ex[4] = {0u, 0u, 0u, 0u};
                if constexpr (HOMO) {
                  if (P.has_homo) homopolymer_at(bases, h.len, read_i + j, P.has_homo > 1, &ih, &hw);
                  plane_words(P, B, h.seq0 + read_i + j, ex);
                }
                uint8_t* q = px + col * P.Cout;
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                  if (w < P.n_words) {
                    const unsigned word = tc[w] + bc * P.mask_base[w] + ql * P.mask_qual[w] + df * P.mask_diff[w] +
                                          (HOMO ? ih * P.mask_ishomo[w] + hw * P.mask_homow[w] + ex[w] : 0u);
#pragma unroll
                    for (int bb = 0; bb < 4; ++bb)
                      if (w * 4 + bb < P.C) q[w * 4 + bb] = (uint8_t)(word >> (8 * bb));
                  }
                }
              }
            };
            if (FAST7 && hi > lo) {
              const int cA = c0 + lo, cB = c0 + hi;            // window columns [cA, cB)
              const int gA = (cA + 3) >> 2, gB = cB >> 2;      // whole 4-pixel groups [gA, gB)
              const unsigned mb0 = P.mask_base[0], mq0 = P.mask_qual[0], md0 = P.mask_diff[0];
              const unsigned mb1 = P.mask_base[1], mq1 = P.mask_qual[1], md1 = P.mask_diff[1];
              for (int g = gA + lane; g < gB; g += 32) {
                const int col = 4 * g;
                const uint8_t* bp = bases + read_i + (col - c0);
                const uint8_t* qp = quals + read_i + (col - c0);
#if DVB_ENC_WORD_LOADS
                // the 4 bases / 4 qualities of a group as two aligned words each and a funnel shift (the misalignment is the same for
                // every group of the run) instead of 8 byte loads; the last groups of the arrays keep the byte loads (no read past the end)
                unsigned bw, qw;
                if (h.seq0 + read_i + (col - c0) + 8 <= B.n_bases) {
                  const unsigned sb = (unsigned)((uintptr_t)bp & 3u), sq = (unsigned)((uintptr_t)qp & 3u);
                  const unsigned* bpw = reinterpret_cast<const unsigned*>(bp - sb);
                  const unsigned* qpw = reinterpret_cast<const unsigned*>(qp - sq);
                  bw = __funnelshift_r(bpw[0], bpw[1], 8 * sb);
                  qw = __funnelshift_r(qpw[0], qpw[1], 8 * sq);
                } else {
                  bw = bp[0] | (bp[1] << 8) | (bp[2] << 16) | ((unsigned)bp[3] << 24);
                  qw = qp[0] | (qp[1] << 8) | (qp[2] << 16) | ((unsigned)qp[3] << 24);
                }
                const unsigned b0 = bw & 0xFF, b1 = (bw >> 8) & 0xFF, b2 = (bw >> 16) & 0xFF, b3 = bw >> 24;
                if (((bw - 0x01010101u) & ~bw & 0x80808080u) != 0) {   // a zero base is not drawn: rare, take the slow path
                  for (int j = 0; j < 4; ++j) put_pixel(col - c0 + j);
                  continue;
                }
                const unsigned r4 = *reinterpret_cast<const unsigned*>(s_ref + col);
                const unsigned q0 = s_qual[qw & 0xFF], q1 = s_qual[(qw >> 8) & 0xFF], q2 = s_qual[(qw >> 16) & 0xFF], q3 = s_qual[qw >> 24];
#else
                const unsigned b0 = bp[0], b1 = bp[1], b2 = bp[2], b3 = bp[3];
                if (b0 == 0 || b1 == 0 || b2 == 0 || b3 == 0) {   // a zero base is not drawn: rare, take the slow path
                  for (int j = 0; j < 4; ++j) put_pixel(col - c0 + j);
                  continue;
                }
                const unsigned r4 = *reinterpret_cast<const unsigned*>(s_ref + col);
                const unsigned q0 = s_qual[qp[0]], q1 = s_qual[qp[1]], q2 = s_qual[qp[2]], q3 = s_qual[qp[3]];
#endif
                const unsigned m = P.match_color, mm = P.mismatch_color;
                const unsigned f0 = b0 == (r4 & 0xFF) ? m : mm, f1 = b1 == ((r4 >> 8) & 0xFF) ? m : mm;
                const unsigned f2 = b2 == ((r4 >> 16) & 0xFF) ? m : mm, f3 = b3 == (r4 >> 24) ? m : mm;
                const unsigned c0c = s_base[b0], c1c = s_base[b1], c2c = s_base[b2], c3c = s_base[b3];
                const unsigned lo0 = tc[0] + c0c * mb0 + q0 * mq0 + f0 * md0, hi0 = tc[1] + c0c * mb1 + q0 * mq1 + f0 * md1;
                const unsigned lo1 = tc[0] + c1c * mb0 + q1 * mq0 + f1 * md0, hi1 = tc[1] + c1c * mb1 + q1 * mq1 + f1 * md1;
                const unsigned lo2 = tc[0] + c2c * mb0 + q2 * mq0 + f2 * md0, hi2 = tc[1] + c2c * mb1 + q2 * mq1 + f2 * md1;
                const unsigned lo3 = tc[0] + c3c * mb0 + q3 * mq0 + f3 * md0, hi3 = tc[1] + c3c * mb1 + q3 * mq1 + f3 * md1;
                unsigned* w = reinterpret_cast<unsigned*>(px + 28 * g);   // 4 pixels x 7 bytes = 7 aligned words
                w[0] = lo0;
                w[1] = hi0 | (lo1 << 24);
                w[2] = __funnelshift_r(lo1, hi1, 8);
                w[3] = (hi1 >> 8) | (lo2 << 16);
                w[4] = __funnelshift_r(lo2, hi2, 16);
                w[5] = (hi2 >> 16) | (lo3 << 8);
                w[6] = __funnelshift_r(lo3, hi3, 24);
              }
              // ragged ends: at most 3 + 3 pixels (or the whole run when it holds no whole group)
              const int head_end = gB > gA ? 4 * gA : cB;
              const int n_head = head_end - cA;
              const int tail_begin = gB > gA ? 4 * gB : cB;
              const int n_tail = cB - tail_begin;
              for (int i = lane; i < n_head + n_tail; i += 32) put_pixel((i < n_head ? cA + i : tail_begin + (i - n_head)) - c0);
            } else {
              for (int j = lo + lane; j < hi; j += 32) put_pixel(j);
            }
            ref_i += len; read_i += len;
          } else if (op == 1 || op == 2) {
            // INSERT: anchor = ref_i - 1, quality of read_i; DELETE: anchor = ref_i - 1 (after the
            // reference's own `ref_i -= 1`), quality of read_i - 1.  pileup_channel_lib.cc:129-137,228-245
            const bool fires = (op == 1) ? (ref_i > 0) : (read_i > 0);
            const int col = ref_i - 1 - image_start;
            if (lane == 0 && fires && P.anchor_char && (unsigned)col < (unsigned)P.W) {
              const unsigned bc = s_base[P.anchor_char & 0xFF];
              const unsigned ql = s_qual[quals[op == 1 ? read_i : read_i - 1]];
              const unsigned df = ((unsigned)P.anchor_char == s_ref[col]) ? P.match_color : P.mismatch_color;
              unsigned ih = 0u, hw = 0u;
              unsigned ex[4] = {0u, 0u, 0u, 0u};
              if constexpr (HOMO) {
                if (P.has_homo) homopolymer_at(bases, h.len, op == 1 ? read_i : read_i - 1, P.has_homo > 1, &ih, &hw);
                plane_words(P, B, h.seq0 + (op == 1 ? read_i : read_i - 1), ex);
              }
              uint8_t* q = px + col * P.Cout;
#pragma unroll
              for (int w = 0; w < 4; ++w) {
                if (w < P.n_words) {
                  const unsigned word = tc[w] + bc * P.mask_base[w] + ql * P.mask_qual[w] + df * P.mask_diff[w] +
                                        (HOMO ? ih * P.mask_ishomo[w] + hw * P.mask_homow[w] + ex[w] : 0u);
#pragma unroll
                  for (int bb = 0; bb < 4; ++bb)
                    if (w * 4 + bb < P.C) q[w * 4 + bb] = (uint8_t)(word >> (8 * bb));
                }
              }
            }
            if (op == 1) read_i += len; else ref_i += len;
          } else if (op == 4) {
            read_i += len;
          } else if (op == 3) {
            ref_i += len;
          }  // 5, 6 ignored; others were flagged in phase A
          __syncwarp();  // orders this op's shared-memory writes before the next op's
        }
      }
      __syncwarp();
      if (cov_row) {   // kChannelValue255 on the reference band, kChannelValue200 below it (pileup_image_native.cc:433-441)
        const uint8_t v = row < P.band ? (uint8_t)255 : (uint8_t)200;
        for (int col = lane; col < P.W; col += 32) px[col * P.Cout + P.cov_chan] = v;
        __syncwarp();
      }
      flush_row(dst, buf, P.row_bytes, lane);
    }
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------

struct DvbEncoder {
  DvbPileupParams params;
  EncDev dev;
  int device = 0;
  int num_sms = 148;
  int smem_bytes = 0;
  int grid_cap = 0;
  bool fast7 = false;
  int* d_perm = nullptr;
  int* d_err = nullptr;
  int64_t launches = 0;
  // staging for the host entry point
  cudaStream_t copy_stream = nullptr;
  std::vector<cudaEvent_t> copy_events;
  dvb::DevBuf d_in, d_out, d_rows, d_recs, d_support;
  bool prepass = true;
  int64_t last_support_pairs = -1;   // pairs of the last batch whose support was derived on the device
  dvb::PinBuf h_in, h_out;
  cudaStream_t stream = nullptr;
};

namespace {

// ---- host restatement of the float32 colour expressions (exactly the reference's forms) ----
uint8_t HostScaleColor(int value, float max_val) {  // channels/base_quality_channel.cc:59-66
  if (static_cast<float>(value) > max_val) value = max_val;
  return static_cast<int>(kMaxPixelValueAsFloat * (static_cast<float>(value) / max_val));
}
int HostAlphaColor(float alpha) { return static_cast<int>(kMaxPixelValueAsFloat * alpha); }

int HostBaseColor(int base, const DvbPileupParams& o) {  // channels/read_base_channel.cc:56-73
  switch (base) {
    case 'A': return o.base_color_offset_a_and_g + o.base_color_stride * 3;
    case 'G': return o.base_color_offset_a_and_g + o.base_color_stride * 2;
    case 'T': return o.base_color_offset_t_and_c + o.base_color_stride * 1;
    case 'C': return o.base_color_offset_t_and_c + o.base_color_stride * 0;
    default: return 0;
  }
}

int BuildDev(const DvbPileupParams& o, EncDev* d) {
  memset(d, 0, sizeof(*d));
  d->gc_chan = -1;
  d->cov_chan = -1;
  if (o.width < 1) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "width must be >= 1");
  if (o.num_channels < 1 || o.num_channels > DVB_MAX_CHANNELS || o.num_alt_channels < 0 ||
      o.num_channels + o.num_alt_channels > DVB_MAX_CHANNELS)
    return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "bad channel count %d + %d", o.num_channels, o.num_alt_channels);
  if (o.reference_band_height < 0 || o.height <= o.reference_band_height)
    return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "height %d must exceed reference_band_height %d", o.height,
                     o.reference_band_height);
  d->W = o.width; d->H = o.height; d->band = o.reference_band_height;
  d->C = o.num_channels; d->Cout = o.num_channels + o.num_alt_channels;
  d->max_rows = o.height - o.reference_band_height;
  d->row_bytes = d->W * d->Cout;
  d->image_bytes = (long long)d->H * d->row_bytes;
  d->min_bq = o.min_base_quality; d->min_mq = o.min_mapping_quality;
  d->anchor_char = o.indel_anchoring_base_char & 0xFF;
  d->sort_by_hap = o.sort_by_haplotypes; d->polish_tag = o.hp_tag_for_assembly_polishing;
  d->sort_by_group = o.sort_by_alt_allele_support;
  d->mapq_cap = (float)o.mapping_quality_cap;
  d->pos_strand = o.positive_strand_color; d->neg_strand = o.negative_strand_color;
  d->sup_color[0] = (uint8_t)HostAlphaColor(o.allele_unsupporting_read_alpha);
  d->sup_color[1] = (uint8_t)HostAlphaColor(o.allele_supporting_read_alpha);
  d->sup_color[2] = (uint8_t)HostAlphaColor(o.other_allele_supporting_read_alpha);
  d->match_color = (uint8_t)HostAlphaColor(o.reference_matching_read_alpha);
  d->mismatch_color = (uint8_t)HostAlphaColor(o.reference_mismatching_read_alpha);
  d->supp_color[0] = static_cast<unsigned char>(kMaxPixelValueAsFloat * o.allele_unsupporting_read_alpha);
  d->supp_color[1] = static_cast<unsigned char>(kMaxPixelValueAsFloat * o.allele_supporting_read_alpha);
  for (int c = 0; c < o.num_channels; ++c) {
    const int ch = o.channels[c];
    d->chan[c] = (uint8_t)ch;
    switch (ch) {
      case DVB_CH_READ_BASE: d->kind[c] = K_BASE; d->ref_const[c] = 0; break;
      case DVB_CH_BASE_QUALITY:
        d->kind[c] = K_QUAL; d->ref_const[c] = HostScaleColor(o.reference_base_quality, o.base_quality_cap); break;
      case DVB_CH_MAPPING_QUALITY:  // mapping_quality_channel.cc:53-58 scales by base_quality_cap
        d->kind[c] = K_CONST; d->ref_const[c] = HostScaleColor(o.reference_base_quality, o.base_quality_cap); break;
      case DVB_CH_STRAND: d->kind[c] = K_CONST; d->ref_const[c] = (uint8_t)o.positive_strand_color; break;
      case DVB_CH_READ_SUPPORTS_VARIANT: d->kind[c] = K_CONST; d->ref_const[c] = d->sup_color[0]; break;
      case DVB_CH_BASE_DIFFERS_FROM_REF: d->kind[c] = K_DIFF; d->ref_const[c] = d->match_color; break;
      case DVB_CH_HAPLOTYPE_TAG: d->kind[c] = K_CONST; d->ref_const[c] = HostScaleColor(0, 2); break;
      case DVB_CH_INSERT_SIZE: d->kind[c] = K_CONST; d->ref_const[c] = static_cast<uint8_t>(kMaxPixelValueAsFloat); break;
      case DVB_CH_SUPPLEMENTARY_ALIGNMENT:  // supplementary_alignment_channel.cc:60-63: float -> uchar
        d->kind[c] = K_CONST; d->ref_const[c] = static_cast<unsigned char>(o.allele_unsupporting_read_alpha); break;
      case DVB_CH_BLANK: d->kind[c] = K_ZERO; d->ref_const[c] = 0; break;
      case DVB_CH_READ_MAPPING_PERCENT: case DVB_CH_AVG_BASE_QUALITY: case DVB_CH_IDENTITY: case DVB_CH_GAP_COMPRESSED_IDENTITY:
        d->kind[c] = K_CONST; d->ref_const[c] = static_cast<uint8_t>(kMaxPixelValueAsFloat); break;
      case DVB_CH_IS_HOMOPOLYMER: d->kind[c] = K_ISHOMO; d->ref_const[c] = 0; d->has_homo = std::max(d->has_homo, 1); break;
      case DVB_CH_HOMOPOLYMER_WEIGHTED: d->kind[c] = K_HOMOW; d->ref_const[c] = 0; d->has_homo = 2; break;
      case DVB_CH_GC_CONTENT:   // reference band = GC content of the window: filled per image by the kernel
        d->kind[c] = K_CONST; d->ref_const[c] = 0; d->gc_chan = c; break;
      case DVB_CH_ALLELE_FREQUENCY: case DVB_CH_ALLELE_SAMPLE_PROBABILITY: case DVB_CH_READ_SUPPORTS_VARIANT_FUZZY: {
        // per pair from the caller's plane; reference band 0 (AlleleFrequencyColor(0) = 0, allele_frequency_channel.cc:70-82) or
        // SupportsAltColor(0) (read_supports_variant_fuzzy_channel.cc:112-116)
        d->kind[c] = K_ZERO; d->ref_const[c] = ch == DVB_CH_READ_SUPPORTS_VARIANT_FUZZY ? d->sup_color[0] : (uint8_t)0;
        if (d->n_pair_planes >= DVB_N_PAIR_PLANES) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "a per-pair plane channel is listed twice");
        d->pair_plane_chan[d->n_pair_planes] = (uint8_t)c;
        d->pair_plane_slot[d->n_pair_planes++] = (uint8_t)(ch == DVB_CH_ALLELE_FREQUENCY ? DVB_PAIR_PLANE_ALLELE_FREQUENCY
                                                           : ch == DVB_CH_READ_SUPPORTS_VARIANT_FUZZY ? DVB_PAIR_PLANE_FUZZY_SUPPORT
                                                                                                       : DVB_PAIR_PLANE_ALLELE_SAMPLE_PROBABILITY);
        break;
      }
      case DVB_CH_BASE_METHYLATION: case DVB_CH_BASE_6MA: case DVB_CH_HOMOPOLYMER_INSERTION_QUALITY:
      case DVB_CH_HOMOPOLYMER_DELETION_QUALITY: case DVB_CH_INTER_HOMOPOLYMER_INSERTION_QUALITY: {   // per base from the caller's plane; band 0
        static const int kSlot[] = {DVB_BASE_PLANE_5MC, DVB_BASE_PLANE_6MA, -1, -1, -1, DVB_BASE_PLANE_HMER_INSERTION,
                                    DVB_BASE_PLANE_HMER_DELETION, DVB_BASE_PLANE_INTER_HMER_INSERTION};
        d->kind[c] = K_PLANE; d->ref_const[c] = 0;
        if (d->n_planes >= DVB_N_BASE_PLANES) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "a per-base plane channel is listed twice");
        d->plane_chan[d->n_planes] = (uint8_t)c; d->plane_slot[d->n_planes] = (uint8_t)kSlot[ch - DVB_CH_BASE_METHYLATION];
        ++d->n_planes;
        break;
      }
      case DVB_CH_MEAN_COVERAGE:   // a BlankChannel per read (pileup_channel_lib.cc:418-421); the overlay is painted by the kernel
        d->kind[c] = K_ZERO; d->ref_const[c] = 0;
        if (d->cov_chan < 0) {     // std::find: the first such channel
          d->cov_chan = c;
          d->cov_rows = std::max(0, std::min(static_cast<int>(o.mean_coverage) + o.reference_band_height, o.height));
        }
        break;
      default:
        return dvb::fail(DVB_ERR_UNSUPPORTED_CHANNEL, "channel enum %d is not implemented", ch);
    }
  }
  d->n_words = (d->C + 3) / 4;
  for (int c = 0; c < d->C; ++c) {
    const unsigned one = 1u << (8 * (c & 3));
    if (d->kind[c] == K_BASE) d->mask_base[c >> 2] |= one;
    else if (d->kind[c] == K_QUAL) d->mask_qual[c >> 2] |= one;
    else if (d->kind[c] == K_DIFF) d->mask_diff[c >> 2] |= one;
    else if (d->kind[c] == K_CONST) d->mask_const[c >> 2] |= one;
    else if (d->kind[c] == K_ISHOMO) d->mask_ishomo[c >> 2] |= one;
    else if (d->kind[c] == K_HOMOW) d->mask_homow[c >> 2] |= one;
    if (d->kind[c] != K_BASE) d->ref_words[c >> 2] |= (unsigned)d->ref_const[c] << (8 * (c & 3));
  }
  // channels_enum_to_blank: FillReadBase is skipped for these channels (pileup_channel_lib.cc:152-161) while CalculateRefRows still
  // draws them - keep the reference band's masks, then take the channel out of everything the read rows use
  for (int w = 0; w < 4; ++w) { d->ref_mask_base[w] = d->mask_base[w]; d->ref_mask_ishomo[w] = d->mask_ishomo[w]; d->ref_mask_homow[w] = d->mask_homow[w]; }
  for (int c = 0; c < d->C; ++c) {
    if (!(o.blank_channel_mask & (1u << c))) continue;
    const unsigned keep = ~(0xFFu << (8 * (c & 3)));
    d->mask_base[c >> 2] &= keep; d->mask_qual[c >> 2] &= keep; d->mask_diff[c >> 2] &= keep; d->mask_const[c >> 2] &= keep;
    d->mask_ishomo[c >> 2] &= keep; d->mask_homow[c >> 2] &= keep;
    d->kind[c] = K_ZERO;
    int n = 0;
    for (int k = 0; k < d->n_planes; ++k)
      if (d->plane_chan[k] != c) { d->plane_chan[n] = d->plane_chan[k]; d->plane_slot[n++] = d->plane_slot[k]; }
    d->n_planes = n;
    n = 0;
    for (int k = 0; k < d->n_pair_planes; ++k)
      if (d->pair_plane_chan[k] != c) { d->pair_plane_chan[n] = d->pair_plane_chan[k]; d->pair_plane_slot[n++] = d->pair_plane_slot[k]; }
    d->n_pair_planes = n;
  }
  for (int i = 0; i < 256; ++i) {
    d->base_lut[i] = (uint8_t)HostBaseColor(i, o);
    d->qual_lut[i] = HostScaleColor(i, (float)o.base_quality_cap);
  }
  return DVB_OK;
}

int SmemBytes(const EncDev& d) {
  const int Wpad = (d.W + 15) & ~15;
  const int rowbuf = ((d.row_bytes + 28 + 15) & ~15) + 32;
  return 512 + Wpad + 7 * d.max_rows * 4 + (kWarps + 2) * 4 + 16 + kWarps * rowbuf;
}

int Launch(DvbEncoder* enc, const DvbBatch& b, uint8_t* out, int32_t* rows_kept, cudaStream_t stream) {
  if (b.n_images <= 0) return DVB_OK;
  for (int k = 0; k < enc->dev.n_pair_planes; ++k)   // a plane-backed channel needs its plane
    if (b.n_pairs > 0 && !b.pair_channel[enc->dev.pair_plane_slot[k]])
      return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "channel enum %d needs DvbBatch.pair_channel[%d]", (int)enc->dev.chan[enc->dev.pair_plane_chan[k]],
                       (int)enc->dev.pair_plane_slot[k]);
  for (int k = 0; k < enc->dev.n_planes; ++k)
    if (b.n_bases > 0 && !b.base_channel[enc->dev.plane_slot[k]])
      return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "channel enum %d needs DvbBatch.base_channel[%d]", (int)enc->dev.chan[enc->dev.plane_chan[k]], (int)enc->dev.plane_slot[k]);
  int grid = std::min(b.n_images, enc->grid_cap);
  PairRec* recs = nullptr;
  const bool derive = b.allele_begin != nullptr;     // pair_support comes from the allele keys: the pre-pass is where it is derived
  if ((enc->prepass || derive) && b.n_pairs > 0) {
    DVB_CUDA(enc->d_recs.reserve((size_t)b.n_pairs * sizeof(PairRec)));
    recs = static_cast<PairRec*>(enc->d_recs.p);
    uint8_t* sup = nullptr;
    if (derive) {
      if (!b.allele_bases_begin || !b.image_ref_run || (b.n_alleles > 0 && (!b.allele_type || !b.allele_class || !b.allele_bases)))
        return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "allele_begin without the other allele arrays");
      DVB_CUDA(enc->d_support.reserve((size_t)b.n_pairs * 4));
      sup = static_cast<uint8_t*>(enc->d_support.p);
      enc->last_support_pairs = b.n_pairs;
    }
    const int pre_grid = (b.n_images + kPreWarps - 1) / kPreWarps;
    if (enc->dev.n_pair_planes) dvb_pair_prepass_kernel<true><<<pre_grid, kPreThreads, 0, stream>>>(enc->dev, b, recs, enc->d_err, sup);
    else dvb_pair_prepass_kernel<false><<<pre_grid, kPreThreads, 0, stream>>>(enc->dev, b, recs, enc->d_err, sup);
    enc->launches++;
  }
  if (enc->fast7)
    dvb_encode_kernel<true, false><<<grid, kThreads, enc->smem_bytes, stream>>>(enc->dev, b, out, rows_kept, enc->d_err, recs);
  else if (enc->dev.has_homo || enc->dev.n_planes)
    dvb_encode_kernel<false, true><<<grid, kThreads, enc->smem_bytes, stream>>>(enc->dev, b, out, rows_kept, enc->d_err, recs);
  else
    dvb_encode_kernel<false, false><<<grid, kThreads, enc->smem_bytes, stream>>>(enc->dev, b, out, rows_kept, enc->d_err, recs);
  enc->launches++;
  DVB_CUDA(cudaGetLastError());
  return DVB_OK;
}

// std::shuffle(iota(n), std::mt19937_64(seed)) as the two standard libraries implement it.  The engine is standardised, the
// algorithm is not:
//   libstdc++ (bits/stl_algo.h): Fisher-Yates from the front, two positions per 64-bit draw when the range allows;
//   libc++    (__algorithm/shuffle.h + __random/uniform_int_distribution.h): for i = 0 .. n-2: j = uniform(0, n-1-i) drawn as the
//             low ceil(log2(range)) bits of one engine output, rejected while >= range; swap(v[i], v[i+j]).
// The reference's golden files were produced with the libc++ form (tools/check_downsample_golden.py).
std::vector<int> ShuffledIndices(int n, uint32_t seed, int shuffle_stdlib) {
  std::vector<int> idx(std::max(n, 0));
  std::iota(idx.begin(), idx.end(), 0);
  std::mt19937_64 gen(seed);
  if (shuffle_stdlib == DVB_SHUFFLE_LIBSTDCXX) {
    std::shuffle(idx.begin(), idx.end(), gen);   // this file is compiled with g++/libstdc++
    return idx;
  }
  for (long first = 0, d = (long)n - 1; d >= 1; ++first, --d) {
    const uint64_t range = (uint64_t)d + 1;           // uniform_int_distribution<ptrdiff_t>(0, d)
    int w = 64 - __builtin_clzll(range) - 1;
    if (range & (~uint64_t(0) >> (64 - w))) ++w;      // w = ceil(log2(range)), >= 1 here
    const uint64_t mask = ~uint64_t(0) >> (64 - w);
    uint64_t u;
    do { u = gen() & mask; } while (u >= range);
    if (u) std::swap(idx[first], idx[first + (long)u]);
  }
  return idx;
}

// Validates a host batch (what the reference leaves to CHECKs / UB), packs every input array into one pinned staging
// block and issues ONE H2D copy on `s`; *db receives the same batch with device pointers.
// One phase of a chunked upload: images [i0, i1), their pairs [p0, p1) and the reads [r0, r1) not uploaded by earlier phases.
struct StagePhase { int64_t i0, i1, p0, p1, r0, r1; cudaEvent_t done; };

int StageHostBatch(DvbEncoder* enc, const DvbBatch* hb, DvbBatch* out_db, cudaStream_t s, std::vector<StagePhase>* phases = nullptr) {
  const int64_t NI = hb->n_images, NR = hb->n_reads, NP = hb->n_pairs, NB = hb->n_bases, NC = hb->n_cigar;
  // ---- validation the reference leaves to CHECKs / UB ----
  if (hb->pair_begin[0] != 0 || hb->pair_begin[NI] != NP) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "pair_begin is not a CSR over n_pairs");
  for (int64_t i = 0; i < NI; ++i)
    if (hb->pair_begin[i + 1] < hb->pair_begin[i]) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "pair_begin not monotone");
  for (int64_t p = 0; p < NP; ++p)
    if (hb->pair_read[p] < 0 || hb->pair_read[p] >= NR) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "pair_read out of range");
  if (NR > 0 && (hb->read_seq_begin[NR] != NB || hb->read_cigar_begin[NR] != NC))
    return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "read_seq_begin / read_cigar_begin do not end at n_bases / n_cigar");
  for (int64_t r = 0; r < NR; ++r) {
    int64_t consumed = 0;
    for (int64_t k = hb->read_cigar_begin[r]; k < hb->read_cigar_begin[r + 1]; ++k) {
      const unsigned op = hb->cigar[k] & 0xF;
      if (op > 8) return dvb::fail(DVB_ERR_BAD_CIGAR, "Unrecognized CIGAR op");
      if (op == 0 || op == 1 || op == 4 || op == 7 || op == 8) consumed += hb->cigar[k] >> 4;
    }
    if (consumed > hb->read_seq_begin[r + 1] - hb->read_seq_begin[r])
      return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "read %lld: CIGAR consumes %lld bases, sequence has %lld", (long long)r,
                       (long long)consumed, (long long)(hb->read_seq_begin[r + 1] - hb->read_seq_begin[r]));
  }
  // ---- pack every input array into one pinned staging block, one H2D copy ----
  enum Domain { D_IMAGE, D_IMAGE1, D_PAIR, D_READ, D_READ1, D_BASE, D_CIGAR, D_ALLELE, D_ALLELE1, D_ABASE };
  struct Seg { const void* src; size_t bytes; size_t off; int domain; size_t esz; bool pinned; };
  constexpr int kSegs = 27 + DVB_N_PAIR_PLANES + DVB_N_BASE_PLANES;
  Seg segs[kSegs];
  int ns = 0;
  size_t total = 0;
  const Domain kDomains[kSegs] = {D_IMAGE, D_IMAGE, D_IMAGE, D_IMAGE1, D_PAIR, D_PAIR, D_PAIR, D_READ, D_READ, D_READ, D_READ, D_READ, D_READ, D_READ,
                                  D_READ1, D_READ1, D_BASE, D_BASE, D_CIGAR, D_IMAGE1, D_IMAGE, D_IMAGE, D_ALLELE, D_ALLELE, D_ALLELE, D_ALLELE1, D_ABASE,
                                  D_PAIR, D_PAIR, D_PAIR, D_BASE, D_BASE, D_BASE, D_BASE, D_BASE};
  const size_t kElem[kSegs] = {(size_t)hb->ref_stride, 4, 4, 8, 4, 1, 1, 4, 4, 4, 1, 4, 4, 4, 8, 8, 1, 1, 4, 8, 4, 1, 1, 1, 1, 8, 1, 1, 1, 1, 1, 1, 1, 1, 1};
  static_assert(DVB_N_PAIR_PLANES == 3 && DVB_N_BASE_PLANES == 5, "segment tables above list 3 + 5 planes");
  const bool derive = hb->allele_begin != nullptr;
  const int64_t NA = derive ? hb->n_alleles : 0, NAB = derive ? hb->n_allele_bases : 0;
  if (derive) {
    if (!hb->allele_bases_begin || !hb->image_ref_run || (NA > 0 && (!hb->allele_type || !hb->allele_class)) || (NAB > 0 && !hb->allele_bases))
      return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "allele_begin without the other allele arrays");
    if (hb->allele_begin[0] != 0 || hb->allele_begin[NI] != NA || hb->allele_bases_begin[0] != 0 || hb->allele_bases_begin[NA] != NAB)
      return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "allele_begin / allele_bases_begin are not CSR arrays over n_alleles / n_allele_bases");
    for (int64_t i = 0; i < NI; ++i)
      if (hb->allele_begin[i + 1] < hb->allele_begin[i]) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "allele_begin not monotone");
    for (int64_t a = 0; a < NA; ++a)
      if (hb->allele_bases_begin[a + 1] <= hb->allele_bases_begin[a]) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "empty or negative allele key");
  }
  auto add = [&](const void* p, size_t bytes) {
    segs[ns].src = p; segs[ns].bytes = p ? bytes : 0; segs[ns].off = total;
    segs[ns].domain = kDomains[ns]; segs[ns].esz = kElem[ns]; segs[ns].pinned = false;
    total += (segs[ns].bytes + 255) & ~(size_t)255;
    return ns++;
  };
  const int i_ref = add(hb->ref_bases, (size_t)NI * hb->ref_stride);
  const int i_isp = add(hb->image_start_pos, NI * 4);
  const int i_vs = add(hb->variant_start, NI * 4);
  const int i_pb = add(hb->pair_begin, (NI + 1) * 8);
  const int i_pr = add(hb->pair_read, NP * 4);
  const int i_ps = add(hb->pair_support, NP);
  const int i_pg = add(hb->pair_allele_group, NP);
  const int i_rp = add(hb->read_pos, NR * 4);
  const int i_rsp = add(hb->read_sort_pos, NR * 4);
  const int i_rmq = add(hb->read_mapq, NR * 4);
  const int i_rfl = add(hb->read_flags, NR);
  const int i_rfr = add(hb->read_fragment_length, NR * 4);
  const int i_rhp = add(hb->read_hp, NR * 4);
  const int i_rnr = add(hb->read_name_rank, NR * 4);
  const int i_rsb = add(hb->read_seq_begin, (NR + 1) * 8);
  const int i_rcb = add(hb->read_cigar_begin, (NR + 1) * 8);
  const int i_ba = add(hb->bases, NB);
  const int i_qu = add(hb->quals, NB);
  const int i_ci = add(hb->cigar, NC * 4);
  const int i_ab = add(derive ? hb->allele_begin : nullptr, (NI + 1) * 8);
  const int i_irr = add(derive ? hb->image_ref_run : nullptr, NI * 4);
  const int i_igd = add(derive ? hb->image_group_default : nullptr, NI);
  const int i_at = add(derive ? hb->allele_type : nullptr, NA);
  const int i_ac = add(derive ? hb->allele_class : nullptr, NA);
  const int i_ag = add(derive ? hb->allele_group : nullptr, NA);
  const int i_abb = add(derive ? hb->allele_bases_begin : nullptr, (NA + 1) * 8);
  const int i_aba = add(derive ? hb->allele_bases : nullptr, NAB);
  int i_pch[DVB_N_PAIR_PLANES], i_bch[DVB_N_BASE_PLANES];
  for (int k = 0; k < DVB_N_PAIR_PLANES; ++k) i_pch[k] = add(hb->pair_channel[k], NP);
  for (int k = 0; k < DVB_N_BASE_PLANES; ++k) i_bch[k] = add(hb->base_channel[k], NB);
  total = std::max<size_t>(total, 256);
  DVB_CUDA(enc->h_in.reserve(total));
  DVB_CUDA(enc->d_in.reserve(total));
  if (phases && phases->size() > 1) {
    // Chunked upload on `s` (the caller passes its copy stream): phase k brings the per-image and per-pair slices of its
    // images and the reads no earlier phase brought, then records its event, so that encode + classify of phase k
    // overlap the upload of phase k + 1.  The device arrays keep the full batch's layout (absolute indices stay valid).
    for (int i = 0; i < ns; ++i) {
      cudaPointerAttributes attr;
      segs[i].pinned = segs[i].bytes > 0 && cudaPointerGetAttributes(&attr, segs[i].src) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    }
    cudaGetLastError();
    for (StagePhase& ph : *phases) {
      for (int i = 0; i < ns; ++i) {
        if (!segs[i].bytes) continue;
        int64_t lo = 0, hi = 0;
        switch (segs[i].domain) {
          case D_IMAGE: lo = ph.i0; hi = ph.i1; break;
          case D_IMAGE1: lo = ph.i0 + (ph.i0 > 0 ? 1 : 0); hi = ph.i1 + 1; break;   // element i0 came with the previous phase
          case D_PAIR: lo = ph.p0; hi = ph.p1; break;
          case D_READ: lo = ph.r0; hi = ph.r1; break;
          case D_READ1: lo = ph.r0 + (ph.r0 > 0 ? 1 : 0); hi = ph.r1 > ph.r0 ? ph.r1 + 1 : lo; break;   // element r0 came with the previous phase
          case D_BASE: lo = hb->read_seq_begin[ph.r0]; hi = hb->read_seq_begin[ph.r1]; break;
          case D_CIGAR: lo = hb->read_cigar_begin[ph.r0]; hi = hb->read_cigar_begin[ph.r1]; break;
          case D_ALLELE: lo = hb->allele_begin[ph.i0]; hi = hb->allele_begin[ph.i1]; break;
          case D_ALLELE1: lo = hb->allele_begin[ph.i0] + (ph.i0 > 0 ? 1 : 0); hi = hb->allele_begin[ph.i1] + 1; break;
          case D_ABASE: lo = hb->allele_bases_begin[hb->allele_begin[ph.i0]]; hi = hb->allele_bases_begin[hb->allele_begin[ph.i1]]; break;
        }
        if (hi <= lo) continue;
        const size_t b0 = (size_t)lo * segs[i].esz, nb = (size_t)(hi - lo) * segs[i].esz;
        const char* src = static_cast<const char*>(segs[i].src) + b0;
        char* dst = static_cast<char*>(enc->d_in.p) + segs[i].off + b0;
        if (!segs[i].pinned) {
          char* stage = static_cast<char*>(enc->h_in.p) + segs[i].off + b0;
          memcpy(stage, src, nb);
          src = stage;
        }
        DVB_CUDA(cudaMemcpyAsync(dst, src, nb, cudaMemcpyHostToDevice, s));
      }
      DVB_CUDA(cudaEventRecord(ph.done, s));
    }
  } else
  // Arrays the caller already keeps in page-locked memory go to the device straight from where they lie; pageable
  // ones are packed into the pinned staging block first (contiguous pageable runs share one copy).
  {
    size_t run_begin = 0, run_end = 0;   // pending staged byte range [run_begin, run_end)
    auto flush = [&]() -> cudaError_t {
      if (run_end <= run_begin) return cudaSuccess;
      cudaError_t e = cudaMemcpyAsync(static_cast<char*>(enc->d_in.p) + run_begin, static_cast<char*>(enc->h_in.p) + run_begin,
                                      run_end - run_begin, cudaMemcpyHostToDevice, s);
      run_begin = run_end;
      return e;
    };
    for (int i = 0; i < ns; ++i) {
      if (!segs[i].bytes) continue;
      cudaPointerAttributes attr;
      const bool pinned = segs[i].bytes >= (64u << 10) && cudaPointerGetAttributes(&attr, segs[i].src) == cudaSuccess &&
                          attr.type == cudaMemoryTypeHost;
      if (pinned) {
        DVB_CUDA(flush());
        DVB_CUDA(cudaMemcpyAsync(static_cast<char*>(enc->d_in.p) + segs[i].off, segs[i].src, segs[i].bytes, cudaMemcpyHostToDevice, s));
        run_begin = run_end = segs[i].off + ((segs[i].bytes + 255) & ~(size_t)255);
      } else {
        if (run_end <= run_begin) run_begin = segs[i].off;
        memcpy(static_cast<char*>(enc->h_in.p) + segs[i].off, segs[i].src, segs[i].bytes);
        run_end = segs[i].off + segs[i].bytes;
      }
    }
    DVB_CUDA(flush());
    cudaGetLastError();   // cudaPointerGetAttributes on plain malloc memory may leave a sticky-free error on old drivers
  }
  DvbBatch db = *hb;
  char* base = static_cast<char*>(enc->d_in.p);
  auto dp = [&](int i) -> const void* { return segs[i].bytes || segs[i].src ? base + segs[i].off : nullptr; };
  db.ref_bases = (const uint8_t*)dp(i_ref); db.image_start_pos = (const int32_t*)dp(i_isp);
  db.variant_start = (const int32_t*)dp(i_vs); db.pair_begin = (const int64_t*)dp(i_pb);
  db.pair_read = (const int32_t*)dp(i_pr); db.pair_support = (const uint8_t*)dp(i_ps);
  db.pair_allele_group = hb->pair_allele_group ? (const uint8_t*)dp(i_pg) : nullptr;
  db.read_pos = (const int32_t*)dp(i_rp); db.read_sort_pos = (const int32_t*)dp(i_rsp);
  db.read_mapq = (const int32_t*)dp(i_rmq); db.read_flags = (const uint8_t*)dp(i_rfl);
  db.read_fragment_length = (const int32_t*)dp(i_rfr); db.read_hp = (const int32_t*)dp(i_rhp);
  db.read_name_rank = (const uint32_t*)dp(i_rnr); db.read_seq_begin = (const int64_t*)dp(i_rsb);
  db.read_cigar_begin = (const int64_t*)dp(i_rcb); db.bases = (const uint8_t*)dp(i_ba);
  db.quals = (const uint8_t*)dp(i_qu); db.cigar = (const uint32_t*)dp(i_ci);
  if (derive) {
    db.allele_begin = (const int64_t*)dp(i_ab); db.image_ref_run = (const int32_t*)dp(i_irr);
    db.image_group_default = hb->image_group_default ? (const uint8_t*)dp(i_igd) : nullptr;
    db.allele_type = (const uint8_t*)dp(i_at); db.allele_class = (const uint8_t*)dp(i_ac);
    db.allele_group = hb->allele_group ? (const uint8_t*)dp(i_ag) : nullptr;
    db.allele_bases_begin = (const int64_t*)dp(i_abb); db.allele_bases = (const uint8_t*)dp(i_aba);
  }
  for (int k = 0; k < DVB_N_PAIR_PLANES; ++k) db.pair_channel[k] = hb->pair_channel[k] ? (const uint8_t*)dp(i_pch[k]) : nullptr;
  for (int k = 0; k < DVB_N_BASE_PLANES; ++k) db.base_channel[k] = hb->base_channel[k] ? (const uint8_t*)dp(i_bch[k]) : nullptr;

  *out_db = db;
  return DVB_OK;
}

}  // namespace

extern "C" {

int dvb_abi_version(void) { return DVB_ABI_VERSION; }
const char* dvb_last_error(void) { return dvb::last_error().c_str(); }

void dvb_pileup_params_default(DvbPileupParams* p) {
  memset(p, 0, sizeof(*p));
  p->width = 221; p->height = 100; p->reference_band_height = 5;
  p->num_channels = 6;
  for (int i = 0; i < 6; ++i) p->channels[i] = i + 1;
  p->base_color_offset_a_and_g = 40; p->base_color_offset_t_and_c = 30; p->base_color_stride = 70;
  p->allele_supporting_read_alpha = 1.0f; p->allele_unsupporting_read_alpha = 0.6f;
  p->other_allele_supporting_read_alpha = 0.6f; p->reference_matching_read_alpha = 0.2f;
  p->reference_mismatching_read_alpha = 1.0f; p->indel_anchoring_base_char = '*';
  p->reference_base_quality = 60; p->positive_strand_color = 70; p->negative_strand_color = 240;
  p->base_quality_cap = 40; p->mapping_quality_cap = 60;
  p->min_base_quality = 10; p->min_mapping_quality = 10;
  p->random_seed = 2101079370u;
  p->max_reads_per_image = 0;
}

int64_t dvb_image_bytes(const DvbPileupParams* p) {
  return (int64_t)p->height * p->width * (p->num_channels + p->num_alt_channels);
}

int dvb_shuffle_table(int32_t n, uint32_t seed, int32_t shuffle_stdlib, int32_t* out) {
  if (n < 0 || !out || (shuffle_stdlib != DVB_SHUFFLE_LIBCXX && shuffle_stdlib != DVB_SHUFFLE_LIBSTDCXX))
    return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_shuffle_table: bad arguments");
  std::vector<int> idx = ShuffledIndices(n, seed, shuffle_stdlib);
  for (int i = 0; i < n; ++i) out[i] = idx[i];
  return DVB_OK;
}

int dvb_encoder_create(const DvbPileupParams* params, int device, DvbEncoder** out) {
  if (!params || !out) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "null argument");
  *out = nullptr;
  EncDev dev;
  int st = BuildDev(*params, &dev);
  if (st) return st;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return dvb::fail(DVB_ERR_NO_DEVICE, "no CUDA device (this library has no CPU path)");
  if (device < 0 || device >= ndev) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "device %d out of range", device);
  DVB_CUDA(cudaSetDevice(device));
  DvbEncoder* enc = new DvbEncoder();
  enc->params = *params;
  enc->device = device;
  cudaDeviceProp prop;
  DVB_CUDA(cudaGetDeviceProperties(&prop, device));
  enc->num_sms = prop.multiProcessorCount;
  // down-sampling tables for n = max_rows+1 .. cap
  const int cap = std::max(params->max_reads_per_image > 0 ? params->max_reads_per_image : 2048, dev.max_rows + 1);
  std::vector<int> tables;
  tables.reserve((size_t)cap * cap / 2);
  for (int n = dev.max_rows + 1; n <= cap; ++n) {
    std::vector<int> idx = ShuffledIndices(n, params->random_seed, params->shuffle_stdlib);  // fresh generator per call, pileup_image_native.cc:327,343
    tables.insert(tables.end(), idx.begin(), idx.end());
  }
  DVB_CUDA(cudaMalloc(&enc->d_perm, std::max<size_t>(tables.size(), 1) * sizeof(int)));
  DVB_CUDA(cudaMemcpy(enc->d_perm, tables.data(), tables.size() * sizeof(int), cudaMemcpyHostToDevice));
  DVB_CUDA(cudaMalloc(&enc->d_err, sizeof(int)));
  DVB_CUDA(cudaMemset(enc->d_err, 0, sizeof(int)));
  dev.perm = enc->d_perm;
  dev.perm_cap = cap;
  enc->dev = dev;
  enc->smem_bytes = SmemBytes(dev);
  if (enc->smem_bytes > 227 * 1024) {
    delete enc;
    return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "image row too large for shared memory (%d bytes)", enc->smem_bytes);
  }
  enc->fast7 = dev.C == 7 && dev.Cout == 7 && !dev.has_homo && !dev.n_planes;   // the 7-byte fast path knows the four classic pixel terms only
  { const char* e = getenv("DVB_ENC_PREPASS"); enc->prepass = !(e && atoi(e) == 0); }
  int occ = 1;
  if (enc->fast7) {
    DVB_CUDA(cudaFuncSetAttribute(dvb_encode_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, enc->smem_bytes));
    DVB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dvb_encode_kernel<true, false>, kThreads, enc->smem_bytes));
  } else if (dev.has_homo || dev.n_planes) {
    DVB_CUDA(cudaFuncSetAttribute(dvb_encode_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, enc->smem_bytes));
    DVB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dvb_encode_kernel<false, true>, kThreads, enc->smem_bytes));
  } else {
    DVB_CUDA(cudaFuncSetAttribute(dvb_encode_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, enc->smem_bytes));
    DVB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dvb_encode_kernel<false, false>, kThreads, enc->smem_bytes));
  }
  enc->grid_cap = enc->num_sms * std::max(occ, 1);
  DVB_CUDA(cudaStreamCreateWithFlags(&enc->stream, cudaStreamNonBlocking));
  *out = enc;
  return DVB_OK;
}

void dvb_encoder_destroy(DvbEncoder* enc) {
  if (!enc) return;
  cudaSetDevice(enc->device);
  if (enc->d_perm) cudaFree(enc->d_perm);
  if (enc->d_err) cudaFree(enc->d_err);
  enc->d_in.release(); enc->d_out.release(); enc->d_rows.release(); enc->d_recs.release(); enc->d_support.release();
  enc->h_in.release(); enc->h_out.release();
  if (enc->stream) cudaStreamDestroy(enc->stream);
  if (enc->copy_stream) cudaStreamDestroy(enc->copy_stream);
  for (cudaEvent_t e : enc->copy_events) cudaEventDestroy(e);
  delete enc;
}

int dvb_encode_batch_device(DvbEncoder* enc, const DvbBatch* batch, uint8_t* out, int32_t* rows_kept, void* stream) {
  if (!enc || !batch || (!out && batch->n_images > 0)) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "null argument");
  if (batch->ref_stride < enc->dev.W) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "ref_stride < width");
  DVB_CUDA(cudaSetDevice(enc->device));
  return Launch(enc, *batch, out, rows_kept, static_cast<cudaStream_t>(stream));
}

int dvb_encoder_check(DvbEncoder* enc, void* stream) {
  if (!enc) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "null encoder");
  DVB_CUDA(cudaSetDevice(enc->device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int e = 0;
  DVB_CUDA(cudaMemcpyAsync(&e, enc->d_err, sizeof(int), cudaMemcpyDeviceToHost, s));
  DVB_CUDA(cudaStreamSynchronize(s));
  if (e) {
    DVB_CUDA(cudaMemsetAsync(enc->d_err, 0, sizeof(int), s));
    DVB_CUDA(cudaStreamSynchronize(s));
    if (e == DVB_ERR_BAD_CIGAR) return dvb::fail(e, "Unrecognized CIGAR op");
    if (e == DVB_ERR_TOO_MANY_READS)
      return dvb::fail(e, "more than %d reads overlap one candidate (raise max_reads_per_image)", enc->dev.perm_cap);
    return dvb::fail(e, "device error %d", e);
  }
  return DVB_OK;
}

int dvb_encode_batch_host(DvbEncoder* enc, const DvbBatch* hb, uint8_t* out_host, int32_t* rows_kept_host) {
  if (!enc || !hb || (!out_host && hb->n_images > 0)) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "null argument");
  if (hb->n_images == 0) return DVB_OK;
  if (hb->ref_stride < enc->dev.W) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "ref_stride < width");
  DVB_CUDA(cudaSetDevice(enc->device));
  cudaStream_t s = enc->stream;
  DvbBatch db;
  int st = StageHostBatch(enc, hb, &db, s);
  if (st) return st;
  const int64_t NI = hb->n_images;
  const size_t out_bytes = (size_t)NI * enc->dev.image_bytes;
  DVB_CUDA(enc->d_out.reserve(out_bytes));
  DVB_CUDA(enc->d_rows.reserve(NI * 4));
  DVB_CUDA(enc->h_out.reserve(out_bytes + NI * 4));
  st = Launch(enc, db, static_cast<uint8_t*>(enc->d_out.p), static_cast<int32_t*>(enc->d_rows.p), s);
  if (st) return st;
  DVB_CUDA(cudaMemcpyAsync(enc->h_out.p, enc->d_out.p, out_bytes, cudaMemcpyDeviceToHost, s));
  DVB_CUDA(cudaMemcpyAsync(static_cast<char*>(enc->h_out.p) + out_bytes, enc->d_rows.p, NI * 4, cudaMemcpyDeviceToHost, s));
  st = dvb_encoder_check(enc, s);  // synchronises
  if (st) return st;
  memcpy(out_host, enc->h_out.p, out_bytes);
  if (rows_kept_host) memcpy(rows_kept_host, static_cast<char*>(enc->h_out.p) + out_bytes, NI * 4);
  return DVB_OK;
}


// Phases of at most `sub` images; phase k uploads the reads [r0, r1) that no earlier phase uploaded (r1 = 1 + the largest
// read index its pairs reference, never below the previous r1).  Empty when the batch fits one phase.
static int PlanUploadPhases(const DvbBatch* hb, int64_t sub, std::vector<StagePhase>* phases) {
  phases->clear();
  const int64_t NI = hb->n_images;
  if (NI <= sub || hb->n_pairs <= 0) return DVB_OK;
  int64_t r_done = 0;
  for (int64_t i0 = 0; i0 < NI; i0 += sub) {
    const int64_t i1 = std::min(NI, i0 + sub);
    StagePhase ph{i0, i1, hb->pair_begin[i0], hb->pair_begin[i1], r_done, r_done, nullptr};
    if (ph.p0 < 0 || ph.p1 < ph.p0 || ph.p1 > hb->n_pairs) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "pair_begin is not a CSR over n_pairs");
    int64_t rmax = -1;
    for (int64_t p = ph.p0; p < ph.p1; ++p) {
      const int64_t r = hb->pair_read[p];
      if (r < 0 || r >= hb->n_reads) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "pair_read out of range");
      rmax = std::max(rmax, r);
    }
    ph.r1 = std::max(r_done, rmax + 1);
    r_done = ph.r1;
    phases->push_back(ph);
  }
  return DVB_OK;
}

// Test access to the upload plan of dvb_encode_classify_host: out = int64[n][6] = {i0, i1, p0, p1, r0, r1}; returns the
// number of phases (0 = single upload), or a negative status.
int dvb_debug_upload_phases(const DvbBatch* hb, int64_t sub, int64_t* out, int32_t cap) {
  if (!hb || sub < 1) return -DVB_ERR_INVALID_ARGUMENT;
  std::vector<StagePhase> phases;
  const int st = PlanUploadPhases(hb, sub, &phases);
  if (st) return -st;
  for (size_t k = 0; k < phases.size() && (int32_t)k < cap; ++k) {
    const StagePhase& ph = phases[k];
    const int64_t v[6] = {ph.i0, ph.i1, ph.p0, ph.p1, ph.r0, ph.r1};
    memcpy(out + 6 * k, v, sizeof(v));
  }
  return (int)phases.size();
}

int dvb_encode_classify_host(DvbEncoder* enc, DvbCnn* cnn, const DvbBatch* hb, float* probs_host, int32_t* rows_kept_host) {
  if (!enc || !cnn || !hb || (!probs_host && hb->n_images > 0)) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "null argument");
  if (hb->n_images == 0) return DVB_OK;
  if (hb->ref_stride < enc->dev.W) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "ref_stride < width");
  DVB_CUDA(cudaSetDevice(enc->device));
  cudaStream_t s = enc->stream;
  const int64_t NI = hb->n_images;
  // Phases of at most one classifier chunk each: the upload of phase k + 1 (copy stream) overlaps encode + classify of
  // phase k (compute stream).  Reads are uploaded by the first phase that references them; batches whose images reference
  // reads in roughly increasing order (position-sorted candidates over position-sorted reads) pipeline fully, any other
  // order stays correct and simply front-loads the upload.
  const int64_t sub = std::max<int64_t>(1, dvb_cnn_max_batch(cnn));
  std::vector<StagePhase> phases;
  {
    const char* e = getenv("DVB_E2E_PIPELINE");
    if (!(e && atoi(e) == 0)) {
      int st0 = PlanUploadPhases(hb, sub, &phases);
      if (st0) return st0;
    }
    if (!enc->copy_stream) DVB_CUDA(cudaStreamCreateWithFlags(&enc->copy_stream, cudaStreamNonBlocking));
    while (enc->copy_events.size() < phases.size()) {
      cudaEvent_t ev;
      DVB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      enc->copy_events.push_back(ev);
    }
    for (size_t k = 0; k < phases.size(); ++k) phases[k].done = enc->copy_events[k];
  }
  const bool chunked = phases.size() > 1;
  DvbBatch db;
  int st = StageHostBatch(enc, hb, &db, chunked ? enc->copy_stream : s, chunked ? &phases : nullptr);
  if (st) return st;
  const size_t out_bytes = (size_t)NI * enc->dev.image_bytes;
  DVB_CUDA(enc->d_out.reserve(out_bytes));
  DVB_CUDA(enc->d_rows.reserve(NI * 4 + NI * 3 * sizeof(float)));
  DVB_CUDA(enc->h_out.reserve(NI * 4 + NI * 3 * sizeof(float)));
  uint8_t* d_images = static_cast<uint8_t*>(enc->d_out.p);
  int32_t* d_rows = static_cast<int32_t*>(enc->d_rows.p);
  float* d_probs = reinterpret_cast<float*>(d_rows + NI);
  if (!chunked) {
    st = Launch(enc, db, d_images, d_rows, s);
    if (st) return st;
    // the images never leave HBM: the classifier consumes the encoder's output in place, on the same stream
    st = dvb_cnn_forward_device(cnn, d_images, (int32_t)NI, d_probs, s);
    if (st) return st;
  } else {
    for (const StagePhase& ph : phases) {
      DVB_CUDA(cudaStreamWaitEvent(s, ph.done, 0));
      DvbBatch sb = db;           // same device arrays; only the per-image views move (pair / read / allele indices are absolute)
      sb.n_images = (int32_t)(ph.i1 - ph.i0);
      sb.ref_bases = db.ref_bases + (size_t)ph.i0 * db.ref_stride;
      sb.image_start_pos = db.image_start_pos + ph.i0;
      sb.variant_start = db.variant_start + ph.i0;
      sb.pair_begin = db.pair_begin + ph.i0;
      if (db.allele_begin) {
        sb.allele_begin = db.allele_begin + ph.i0;
        sb.image_ref_run = db.image_ref_run + ph.i0;
        if (db.image_group_default) sb.image_group_default = db.image_group_default + ph.i0;
      }
      st = Launch(enc, sb, d_images + (size_t)ph.i0 * enc->dev.image_bytes, d_rows + ph.i0, s);
      if (st) return st;
      st = dvb_cnn_forward_device(cnn, d_images + (size_t)ph.i0 * enc->dev.image_bytes, sb.n_images, d_probs + 3 * ph.i0, s);
      if (st) return st;
    }
  }
  DVB_CUDA(cudaMemcpyAsync(enc->h_out.p, d_rows, NI * 4 + NI * 3 * sizeof(float), cudaMemcpyDeviceToHost, s));
  st = dvb_encoder_check(enc, s);  // synchronises
  if (st) return st;
  if (rows_kept_host) memcpy(rows_kept_host, enc->h_out.p, NI * 4);
  memcpy(probs_host, static_cast<char*>(enc->h_out.p) + NI * 4, NI * 3 * sizeof(float));
  return DVB_OK;
}

int64_t dvb_encoder_launch_count(const DvbEncoder* enc) { return enc ? enc->launches : 0; }

int dvb_encoder_last_pair_support(DvbEncoder* enc, int64_t n_pairs, uint8_t* support, uint8_t* group) {
  if (!enc || n_pairs < 0 || (n_pairs > 0 && !support)) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_encoder_last_pair_support: bad arguments");
  if (enc->last_support_pairs != n_pairs)
    return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_encoder_last_pair_support: the last batch derived support for %lld pairs, not %lld",
                     (long long)enc->last_support_pairs, (long long)n_pairs);
  if (n_pairs == 0) return DVB_OK;
  DVB_CUDA(cudaSetDevice(enc->device));
  DVB_CUDA(cudaDeviceSynchronize());
  const uint8_t* sup = static_cast<const uint8_t*>(enc->d_support.p);
  DVB_CUDA(cudaMemcpy(support, sup + 2 * n_pairs, (size_t)n_pairs, cudaMemcpyDeviceToHost));
  if (group) DVB_CUDA(cudaMemcpy(group, sup + 3 * n_pairs, (size_t)n_pairs, cudaMemcpyDeviceToHost));
  return DVB_OK;
}

}  // extern "C"

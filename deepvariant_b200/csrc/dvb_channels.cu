// dvb_channels.cu — value functions of the plane-backed pileup channels (host code; include/dvb.h "channel planes").
//
// The encoder places one byte per (image, read) pair or per base where the reference's FillReadBase would
// (csrc/dvb_encoder.cu); these functions compute those bytes from the data only the caller holds — the alignment
// records' per-base aux tags and DeepVariantCall's numbers — with the reference's float32 / double expressions:
//   ScaleColorVector(255)                     deepvariant/channels/base_methylation_channel.cc:87-99, base_6ma_channel.cc:87-99
//   HomoPolymerWeighted / HomoPolymerInDelQuality   deepvariant/channels/homopolymer_indel_quality_channel.cc:91-183
//   GetT0QualityValues                        deepvariant/channels/inter_homopolymer_insertion_quality_channel.cc:75-127
//   BaseQualityColor                          deepvariant/channels/channel_utils.cc:42-45
//   AlleleFrequencyColor                      deepvariant/channels/allele_frequency_channel.cc:76-87
//   AlleleSampleProbabilityChannel::ScaleColor   deepvariant/channels/allele_sample_probability_channel.cc:87-101
// One pass per read when the BAM records are decoded (like the rest of the read table), not per pixel.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#include "dvb_common.h"

namespace {

constexpr float kMaxPixelValueAsFloat = 254.0f;   // channels/channel.h:78
constexpr float kMaxQScore = 93.0f;               // channels/channel_utils.h:49

inline uint8_t BaseQualityColor(int base_qual) {  // float * int -> float, / float, narrowed
  return static_cast<uint8_t>(kMaxPixelValueAsFloat * base_qual / kMaxQScore);
}

}  // namespace

extern "C" {

// out[i] = int(254 * (float(min(v, 255)) / 255)): the ML probability bytes of one modification type, laid out per base of the
// aligned sequence (nucleus Read.base_modifications), as the 5mC / 6mA channels colour them.
int dvb_channel_base_modification_plane(const uint8_t* values, int64_t len, uint8_t* out) {
  if (len < 0 || (len > 0 && (!values || !out))) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_channel_base_modification_plane: bad arguments");
  const float max_val = 255.0f;
  for (int64_t i = 0; i < len; ++i) {
    int value = values[i];
    if (static_cast<float>(value) > max_val) value = static_cast<int>(max_val);
    out[i] = static_cast<uint8_t>(static_cast<int>(kMaxPixelValueAsFloat * (static_cast<float>(value) / max_val)));
  }
  return DVB_OK;
}

// Ultima's tp tag: tp[i] is the direction and size of the homopolymer-length error whose probability qual[i] encodes.  For every
// homopolymer of the read the error probabilities of its bases that point in the asked direction (tp < 0 = deletion) are summed
// and the run is coloured with the phred score of the sum.  tp == NULL (tag absent): every run keeps kMaxQScore, i.e. 254.
int dvb_channel_hmer_quality_plane(const uint8_t* seq, const uint8_t* qual, int64_t len, const int8_t* tp, int32_t is_deletion, uint8_t* out) {
  if (len < 0 || (len > 0 && (!seq || !qual || !out))) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_channel_hmer_quality_plane: bad arguments");
  int64_t i = 0;
  while (i < len) {
    int64_t j = i + 1;
    while (j < len && seq[j] == seq[i]) ++j;
    // the reference keeps run lengths in a uint8: runs longer than 255 are handled in pieces of at most 255 bases (its own walk past
    // such a run is undefined)
    const int64_t run = std::min<int64_t>(j - i, 255);
    float directed_error_prob = 0;
    for (int64_t k = 0; k < run && tp; ++k) {
      if (tp[i + k] == 0) continue;
      if ((tp[i + k] < 0) == (is_deletion != 0)) {
        const float error_prob = static_cast<float>(std::pow(10, (qual[i + k] / -10.0)));
        directed_error_prob += error_prob;
      }
    }
    int directed_quality = directed_error_prob == 0 ? static_cast<int>(kMaxQScore) : static_cast<int>(-10 * std::log10(directed_error_prob));
    if (directed_quality > kMaxQScore) directed_quality = static_cast<int>(kMaxQScore);
    const uint8_t color = BaseQualityColor(directed_quality);
    for (int64_t k = 0; k < run; ++k) out[i + k] = color;
    i += run;
  }
  return DVB_OK;
}

// Ultima's t0 tag (phred + 33 text, one character per base): the colour of each base's score; bases beyond the tag and reads
// without it get BaseQualityColor(0) = 0.
int dvb_channel_t0_plane(int64_t len, const char* t0, int64_t t0_len, uint8_t* out) {
  if (len < 0 || t0_len < 0 || (len > 0 && !out) || (t0_len > 0 && !t0)) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_channel_t0_plane: bad arguments");
  for (int64_t i = 0; i < len; ++i) {
    const uint8_t q = i < t0_len ? static_cast<uint8_t>(t0[i] - 33) : static_cast<uint8_t>(0);
    out[i] = BaseQualityColor(q);
  }
  return DVB_OK;
}

int32_t dvb_channel_allele_frequency_color(float allele_frequency, float min_non_zero_allele_frequency) {
  if (allele_frequency <= min_non_zero_allele_frequency) return 0;
  const float log10_af = std::log10(allele_frequency);
  const float log10_min = std::log10(min_non_zero_allele_frequency);
  return static_cast<uint8_t>(((log10_min - log10_af) / log10_min) * static_cast<int>(kMaxPixelValueAsFloat));
}

int32_t dvb_channel_allele_sample_probability_color(int32_t value, float max_val) {
  if (max_val == 0) return 0;
  float value_as_float = static_cast<float>(value);
  value_as_float = std::clamp<float>(value_as_float, 0.0f, max_val);
  const double probability = value_as_float / max_val;
  const double scaled_probability = std::sqrt(probability);
  return static_cast<uint8_t>(static_cast<int>(kMaxPixelValueAsFloat * scaled_probability));
}

}  // extern "C"

// Temporary: CNN entry points until dvb_cnn.cu lands.
#include "dvb_common.h"
extern "C" {
int dvb_cnn_create(const void*, int64_t, int32_t, int32_t, int32_t, int32_t, int32_t, int, DvbCnn** out) { if (out) *out = nullptr; return dvb::fail(DVB_ERR_INTERNAL, "CNN not built yet"); }
void dvb_cnn_destroy(DvbCnn*) {}
int dvb_cnn_forward_device(DvbCnn*, const uint8_t*, int32_t, float*, void*) { return dvb::fail(DVB_ERR_INTERNAL, "CNN not built yet"); }
int dvb_cnn_forward_host(DvbCnn*, const uint8_t*, int32_t, float*) { return dvb::fail(DVB_ERR_INTERNAL, "CNN not built yet"); }
int64_t dvb_cnn_launch_count(const DvbCnn*) { return 0; }
double dvb_cnn_flops_per_image(const DvbCnn*) { return 0; }
}

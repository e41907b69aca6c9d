// placeholder until the encoder lands (keeps every symbol of include/ragmeup_b200.h exported)
#include "rmu_common.h"
extern "C" {
int rmu_encoder_create(const rmu_bert_config*, const float* const*, int, int, rmu_encoder**) { rmu::set_error("encoder not built yet"); return RMU_ERR_UNSUPPORTED; }
void rmu_encoder_destroy(rmu_encoder*) {}
int rmu_encoder_embed(rmu_encoder*, const int32_t*, const int32_t*, const int32_t*, int, int, int, int, int, float*, void*) { return RMU_ERR_UNSUPPORTED; }
int rmu_encoder_classify(rmu_encoder*, const int32_t*, const int32_t*, const int32_t*, int, int, int, float*, void*) { return RMU_ERR_UNSUPPORTED; }
int rmu_encoder_hidden(rmu_encoder*, const int32_t*, const int32_t*, const int32_t*, int, int, int, float*, void*) { return RMU_ERR_UNSUPPORTED; }
int rmu_encoder_embed_host(rmu_encoder*, const int32_t*, const int32_t*, const int32_t*, int, int, int, int, int, float*, void*) { return RMU_ERR_UNSUPPORTED; }
int rmu_encoder_classify_host(rmu_encoder*, const int32_t*, const int32_t*, const int32_t*, int, int, int, float*, void*) { return RMU_ERR_UNSUPPORTED; }
}

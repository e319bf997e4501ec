// tensor-core (mma.sync) specialisation of the attention core for the ViT shape; see attention_vit.cu
#pragma once
#include "kernels.h"
namespace vck {
bool vit_attention_eligible(const AttnDesc& a);
int vit_attention_fwd(const AttnDesc& a, bf16_t* o_hi, bf16_t* o_lo, int64_t ldo, float* lse, stream_t s);
int vit_attention_bwd(const AttnDesc& a, const bf16_t* o_hi, const bf16_t* o_lo, int64_t ldo, const float* lse, const float* dout,
                      int64_t lddo, float* dq, int64_t lddq, float* dk, int64_t lddk, float* dv, int64_t lddv, stream_t s);
int vit_attention_bwd_split(const AttnDesc& a, const bf16_t* o_hi, const bf16_t* o_lo, int64_t ldo, const float* lse,
                            const float* dout, const bf16_t* dout_hi, const bf16_t* dout_lo, int64_t lddo, bf16_t* dq_hi, bf16_t* dq_lo, bf16_t* dk_hi, bf16_t* dk_lo,
                            bf16_t* dv_hi, bf16_t* dv_lo, int64_t lds, stream_t s);
}  // namespace vck

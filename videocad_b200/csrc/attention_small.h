// short-sequence (Tq == Tk <= 32) decoder attention kernels; see attention_small.cu
#pragma once
#include "kernels.h"
namespace vck {
bool attention_small_eligible(const AttnDesc& a);
int attention_small_fwd(const AttnDesc& a, bf16_t* o_hi, bf16_t* o_lo, int64_t ldo, float* lse, stream_t s);
// any of the three output groups may be null: fp32 (dq, dk, dv), split-bf16 (d*_hi / d*_lo, leading dimension ld_split) and the
// accumulated column sums (dbq, dbk, dbv: [nh*d] each, caller-zeroed)
int attention_small_bwd(const AttnDesc& a, const bf16_t* o_hi, const bf16_t* o_lo, int64_t ldo, const float* lse, const float* dout,
                        int64_t lddo, float* dq, int64_t lddq, float* dk, int64_t lddk, float* dv, int64_t lddv, bf16_t* dq_hi,
                        bf16_t* dq_lo, bf16_t* dk_hi, bf16_t* dk_lo, bf16_t* dv_hi, bf16_t* dv_lo, int64_t ld_split, float* dbq,
                        float* dbk, float* dbv, stream_t s);
}  // namespace vck

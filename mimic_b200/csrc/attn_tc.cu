// tcgen05 / TMEM / TMA shifted attention (bf16, d == 128).  Placeholder until the kernels land.
#include "attn_common.cuh"

namespace mimic {
namespace tc {
bool supports(const AttnParams&, bool) { return false; }
int fwd(const AttnParams&, cudaStream_t) { set_error("tcgen05 forward not built"); return MIMIC_ERR_UNSUPPORTED; }
int bwd(const AttnParams&, cudaStream_t) { set_error("tcgen05 backward not built"); return MIMIC_ERR_UNSUPPORTED; }
}  // namespace tc
}  // namespace mimic

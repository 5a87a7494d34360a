// attn.cu — placeholder until the tcgen05 attention kernels land (next commit).
#include "common.cuh"
extern "C" int x2k_attn_fwd(const X2kAttnArgs*, void*) { x2k::set_error("x2k_attn_fwd: not built yet"); return X2K_ERR_UNSUPPORTED; }
extern "C" int x2k_attn_bwd(const X2kAttnArgs*, void*) { x2k::set_error("x2k_attn_bwd: not built yet"); return X2K_ERR_UNSUPPORTED; }

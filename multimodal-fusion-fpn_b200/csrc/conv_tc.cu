// tcgen05 / TMEM implicit-GEMM convolution (placeholder until the kernels land: reports "unsupported" so the
// dispatcher uses the CUDA-core kernels).
#include "common.cuh"

bool ffpn_tc_fwd_supported(const ffpn_conv_desc*) { return false; }
bool ffpn_tc_dgrad_supported(const ffpn_conv_desc*) { return false; }
bool ffpn_tc_wgrad_supported(const ffpn_conv_desc*) { return false; }
size_t ffpn_tc_workspace_bytes(const ffpn_conv_desc*) { return 0; }
int ffpn_conv_fwd_tc(ffpn_ctx* ctx, const ffpn_conv_desc*, bool, const void*, const float*, const float*, int, const float*,
                     const void*, void*, float*, int*, void*, size_t, cudaStream_t) { FFPN_FAIL(ctx, "tcgen05 conv not built"); }
int ffpn_conv_wgrad_tc(ffpn_ctx* ctx, const ffpn_conv_desc*, const void*, const float*, const float*, int, const void*,
                       float*, void*, size_t, cudaStream_t) { FFPN_FAIL(ctx, "tcgen05 wgrad not built"); }

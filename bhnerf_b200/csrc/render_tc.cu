// placeholder: replaced by the tcgen05 kernels
#include "common.cuh"
size_t bh_tc_acts_bytes_per_frame(int n_pad) { return (size_t)n_pad * 1024; }
size_t bh_tc_ws_bytes() { return 1 << 20; }
int bh_tc_prepare_weights(const float*, void*, cudaStream_t) { bh_set_error("TC kernels not built yet"); return 3; }
int bh_tc_fwd(const PackedView&, const FrameConsts&, const void*, const float*, const float*, int, float*, void*, cudaStream_t) { bh_set_error("TC kernels not built yet"); return 3; }
int bh_tc_bwd(const PackedView&, const void*, const float*, const float*, int, const float*, const void*, float*, cudaStream_t) { bh_set_error("TC kernels not built yet"); return 3; }

// NTT kernels of the fused mod-down + rescale jobs (see ntt_launch.cuh); their own translation unit so
// that the instantiations compile in parallel.
#define HEGPU_NTT_INSTANTIATE
#include "ntt_launch.cuh"

template int launch_ntt_inv<FinalInttJob>(hegpu_ctx *, const FinalInttJob &, u32, u64 *, int);
template int launch_ntt_fwd<FinalNttJob>(hegpu_ctx *, const FinalNttJob &, u32, int, u64);

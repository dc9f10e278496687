// NTT kernels of one job resolver (see ntt_launch.cuh); its own translation unit so that the
// instantiations compile in parallel.
#define HEGPU_NTT_INSTANTIATE
#include "ntt_launch.cuh"

template int launch_ntt_fwd<KsLiftJob>(hegpu_ctx *, const KsLiftJob &, u32, int, u64);

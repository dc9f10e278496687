// NTT kernels of one job resolver (see ntt_launch.cuh); its own translation unit so that the
// instantiations compile in parallel.
#define HEGPU_NTT_INSTANTIATE
#include "ntt_launch.cuh"

template int launch_ntt_inv<KsInttJob>(hegpu_ctx *, const KsInttJob &, u32, u64 *, int);

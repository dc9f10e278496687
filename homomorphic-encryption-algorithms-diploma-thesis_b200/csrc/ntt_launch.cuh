// ntt_launch.cuh -- launchers of the NTT kernels, templated on the job resolver.  Every job is
// instantiated in its own translation unit (ntt_inst_*.cu); hegpu.cu only sees the declarations.
#pragma once
#include "internal.cuh"

// ------------------------------------------------------------------------- NTT launchers
template <int LOGL, int SPLIT, int LOGE, class Job>
int launch_fwd_shape(hegpu_ctx *c, const Job &job, u32 jobs, int kind, u64 words_per_job)
{
    Prof pf(c, kind, jobs, (u64)jobs * words_per_job * 8 * c->n);
    auto kern = ntt_fwd_kernel<LOGL, SPLIT, LOGE, Job>;
    TRY(configure_smem(c, (const void *)kern, NttShape<LOGL, LOGE>::SMEM));
    kern<<<jobs << SPLIT, NttShape<LOGL, LOGE>::THREADS, NttShape<LOGL, LOGE>::SMEM, c->stream>>>(job, c->tabs);
    c->launches++;
    CU(cudaGetLastError());
    return HEGPU_OK;
}
template <int LOGL, int LOGE, class Job>
int launch_fwd_park(hegpu_ctx *c, const Job &job, u32 jobs, int kind, u64 words_per_job)
{
    TRY(park_reserve(c, (size_t)jobs << LOGL));
    Prof pf(c, kind, jobs, (u64)jobs * words_per_job * 8 * c->n);
    auto kern = ntt_fwd_park_kernel<LOGL, LOGE, Job>;
    TRY(configure_smem(c, (const void *)kern, NttShape<LOGL, LOGE>::SMEM));
    kern<<<jobs, NttShape<LOGL, LOGE>::THREADS, NttShape<LOGL, LOGE>::SMEM, c->stream>>>(job, c->tabs, c->park);
    c->launches++;
    CU(cudaGetLastError());
    return HEGPU_OK;
}
template <int LOGL, int LOGE, class Job>
int launch_inv_park(hegpu_ctx *c, const Job &job, u32 jobs, int kind)
{
    TRY(park_reserve(c, (size_t)jobs << LOGL));
    Prof pf(c, kind, jobs, (u64)jobs * 16 * c->n);
    auto kern = ntt_inv_park_kernel<LOGL, LOGE, Job>;
    TRY(configure_smem(c, (const void *)kern, NttShape<LOGL, LOGE>::SMEM));
    kern<<<jobs, NttShape<LOGL, LOGE>::THREADS, NttShape<LOGL, LOGE>::SMEM, c->stream>>>(job, c->tabs, c->park);
    c->launches++;
    CU(cudaGetLastError());
    return HEGPU_OK;
}

// two-level park (N = 4 * 2^LOGL): three parked quarters per job
template <int LOGL, int LOGE, class Job>
int launch_fwd_park4(hegpu_ctx *c, const Job &job, u32 jobs, int kind, u64 words_per_job)
{
    TRY(park_reserve(c, (size_t)jobs * (3u << LOGL)));
    Prof pf(c, kind, jobs, (u64)jobs * words_per_job * 8 * c->n);
    auto kern = ntt_fwd_park4_kernel<LOGL, LOGE, Job>;
    TRY(configure_smem(c, (const void *)kern, NttShape<LOGL, LOGE>::SMEM));
    kern<<<jobs, NttShape<LOGL, LOGE>::THREADS, NttShape<LOGL, LOGE>::SMEM, c->stream>>>(job, c->tabs, c->park);
    c->launches++;
    CU(cudaGetLastError());
    return HEGPU_OK;
}
template <int LOGL, int LOGE, class Job>
int launch_inv_park4(hegpu_ctx *c, const Job &job, u32 jobs, int kind)
{
    TRY(park_reserve(c, (size_t)jobs * (3u << LOGL)));
    Prof pf(c, kind, jobs, (u64)jobs * 16 * c->n);
    auto kern = ntt_inv_park4_kernel<LOGL, LOGE, Job>;
    TRY(configure_smem(c, (const void *)kern, NttShape<LOGL, LOGE>::SMEM));
    kern<<<jobs, NttShape<LOGL, LOGE>::THREADS, NttShape<LOGL, LOGE>::SMEM, c->stream>>>(job, c->tabs, c->park);
    c->launches++;
    CU(cudaGetLastError());
    return HEGPU_OK;
}

// words_per_job: algorithmic HBM words per coefficient of one job (2 = read + write)
template <class Job>
int launch_ntt_fwd(hegpu_ctx *c, const Job &job, u32 jobs, int kind, u64 words_per_job)
{
    if (jobs == 0) return HEGPU_OK;
    switch (c->logn) {
    case 12: return launch_fwd_shape<12, 0, 4>(c, job, jobs, kind, words_per_job);
    case 13: return launch_fwd_shape<13, 0, 4>(c, job, jobs, kind, words_per_job);
    case 14:
        if (c->use_park)
            return c->loge == 3 ? launch_fwd_park<13, 3>(c, job, jobs, kind, words_per_job)
                                : launch_fwd_park<13, 4>(c, job, jobs, kind, words_per_job);
        return launch_fwd_shape<14, 0, 4>(c, job, jobs, kind, words_per_job);
    case 15:
        if (c->use_park && c->park32k == 2) return launch_fwd_park4<13, 3>(c, job, jobs, kind, words_per_job);
        if (c->use_park && c->park32k) return launch_fwd_park<14, 4>(c, job, jobs, kind, words_per_job);
        return launch_fwd_shape<14, 1, 4>(c, job, jobs, kind, words_per_job);
    }
    LOGIC("unsupported ring degree");
}

template <int LOGL, int SPLIT, int LOGE, class Job>
int launch_inv_shape(hegpu_ctx *c, const Job &job, u32 jobs, u64 *scratch, int kind)
{
    Prof pf(c, kind, jobs, (u64)jobs * 16 * c->n);
    auto kern = ntt_inv_kernel<LOGL, SPLIT, LOGE, Job>;
    TRY(configure_smem(c, (const void *)kern, NttShape<LOGL, LOGE>::SMEM));
    kern<<<jobs << SPLIT, NttShape<LOGL, LOGE>::THREADS, NttShape<LOGL, LOGE>::SMEM, c->stream>>>(job, c->tabs, scratch);
    c->launches++;
    CU(cudaGetLastError());
    if (SPLIT) {
        const size_t total = (size_t)jobs << LOGL;
        ntt_inv_final_kernel<LOGL, Job><<<ew_grid(c, total), 256, 0, c->stream>>>(job, c->tabs, scratch, jobs);
        c->launches++;
        CU(cudaGetLastError());
    }
    return HEGPU_OK;
}
// scratch: [jobs][N] words, only used for N = 32768
template <class Job>
int launch_ntt_inv(hegpu_ctx *c, const Job &job, u32 jobs, u64 *scratch, int kind)
{
    if (jobs == 0) return HEGPU_OK;
    switch (c->logn) {
    case 12: return launch_inv_shape<12, 0, 4>(c, job, jobs, scratch, kind);
    case 13: return launch_inv_shape<13, 0, 4>(c, job, jobs, scratch, kind);
    case 14:
        if (c->use_park)
            return c->loge == 3 ? launch_inv_park<13, 3>(c, job, jobs, kind) : launch_inv_park<13, 4>(c, job, jobs, kind);
        return launch_inv_shape<14, 0, 4>(c, job, jobs, scratch, kind);
    case 15:
        if (c->use_park && c->park32k == 2) return launch_inv_park4<13, 3>(c, job, jobs, kind);
        if (c->use_park && c->park32k) return launch_inv_park<14, 4>(c, job, jobs, kind);
        return launch_inv_shape<14, 1, 4>(c, job, jobs, scratch, kind);
    }
    LOGIC("unsupported ring degree");
}

#ifndef HEGPU_NTT_INSTANTIATE
extern template int launch_ntt_fwd<PlainJob>(hegpu_ctx *, const PlainJob &, u32, int, u64);
extern template int launch_ntt_inv<PlainJob>(hegpu_ctx *, const PlainJob &, u32, u64 *, int);
extern template int launch_ntt_inv<KsInttJob>(hegpu_ctx *, const KsInttJob &, u32, u64 *, int);
extern template int launch_ntt_fwd<KsLiftJob>(hegpu_ctx *, const KsLiftJob &, u32, int, u64);
extern template int launch_ntt_inv<HalfInttJob>(hegpu_ctx *, const HalfInttJob &, u32, u64 *, int);
extern template int launch_ntt_fwd<KsModDownJob>(hegpu_ctx *, const KsModDownJob &, u32, int, u64);
extern template int launch_ntt_fwd<RescaleJob>(hegpu_ctx *, const RescaleJob &, u32, int, u64);
extern template int launch_ntt_inv<FinalInttJob>(hegpu_ctx *, const FinalInttJob &, u32, u64 *, int);
extern template int launch_ntt_fwd<FinalNttJob>(hegpu_ctx *, const FinalNttJob &, u32, int, u64);
#endif

// Launch helper choosing between the tcgen05 decoder (default) and the fp32 CUDA-core decoder.
#pragma once
#include "field_launch.cuh"
#include "field_tc.cuh"

template <class Src, class Epi, bool SDF_ONLY>
static inline int launch_field_fwd_tc(const FieldDev& d, const Src& src, const Epi& epi, int64_t N, cudaStream_t st,
                                      const unsigned int* n_dev = nullptr) {
    int rc = set_smem(field_fwd_tc_kernel<Src, Epi, SDF_ONLY>, SMEM_TC); if (rc) return rc;
    const int64_t tiles = (N + TC_TP - 1) / TC_TP;
    const int64_t cap = mf_sm_count_cached();
    const int grid = (int)(tiles < cap ? (tiles > 0 ? tiles : 1) : cap);
    field_fwd_tc_kernel<Src, Epi, SDF_ONLY><<<grid, TC_NT, SMEM_TC, st>>>(d, src, epi, N, n_dev, d.tc_img, mf_tc_error_flag());
    MF_LAUNCH_CHECK();
    return MF_OK;
}

template <class Src, class Epi, bool SDF_ONLY>
static inline int launch_field_fwd_auto(const FieldDev& d, const Src& src, const Epi& epi, int64_t N, cudaStream_t st,
                                        const unsigned int* n_dev = nullptr) {
    if (d.impl == 0) return launch_field_fwd_tc<Src, Epi, SDF_ONLY>(d, src, epi, N, st, n_dev);
    return launch_field_fwd<Src, Epi, SDF_ONLY>(d, src, epi, N, st, n_dev);
}

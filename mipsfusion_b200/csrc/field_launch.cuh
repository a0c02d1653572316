// Generic persistent forward kernel of the fused field evaluation, parameterised on the point
// source (where the query points come from) and the epilogue (what is done with the 10 outputs).
#pragma once
#include "field_core.cuh"

// Epilogues see the tile's outputs as OUT[c][m]: c = 0..9 (rgb_raw 3, sdf, entropy, prob 5), m = point in tile,
// row length ld, tp points per tile; (tid, nthreads) enumerate the cooperating threads.
struct EpiRaw {                          // write the (N,10) decoder output
    float* raw;
    __device__ __forceinline__ void store(const float* OUT, int ld, int tp, int64_t tile, int64_t N, int tid, int nthreads) const {
        store_raw_tile(OUT, ld, tp, raw, tile, N, tid, nthreads);
    }
};

template <class Src, class Epi, bool SDF_ONLY>
__global__ void __launch_bounds__(NT, 2) field_fwd_kernel(FieldDev f, Src src, Epi epi, int64_t N,
                                                          const unsigned int* __restrict__ n_dev) {
    extern __shared__ __align__(16) float sm[];
    if (n_dev) N = (int64_t)*n_dev;            // point count produced on the device (compacted lists)
    const int64_t n_tiles = (N + TP - 1) / TP;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        encode_tile(f, src, tile, N, sm);
        __syncthreads();
        mlp_forward_tile<SDF_ONLY>(f.prep, sm, sm + ROW_H1 * LDA);
        __syncthreads();
        epi.store(sm + ROW_OUT * LDA, LDA, TP, tile, N, threadIdx.x, NT);
        __syncthreads();
    }
}

static inline int persistent_grid(int64_t N, int ctas_per_sm) {
    const int64_t tiles = (N + TP - 1) / TP;
    const int64_t cap = (int64_t)mf_sm_count_cached() * ctas_per_sm;
    return (int)(tiles < cap ? (tiles > 0 ? tiles : 1) : cap);
}

template <class K> static inline int set_smem(K kernel, size_t bytes) {
    MF_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return MF_OK;
}

template <class Src, class Epi, bool SDF_ONLY>
static inline int launch_field_fwd(const FieldDev& d, const Src& src, const Epi& epi, int64_t N, cudaStream_t st,
                                   const unsigned int* n_dev = nullptr) {
    // N is the exact point count, or an upper bound when the count lives on the device (n_dev)
    int rc = set_smem(field_fwd_kernel<Src, Epi, SDF_ONLY>, SMEM_FWD); if (rc) return rc;
    field_fwd_kernel<Src, Epi, SDF_ONLY><<<persistent_grid(N, 2), NT, SMEM_FWD, st>>>(d, src, epi, N, n_dev);
    MF_LAUNCH_CHECK();
    return MF_OK;
}

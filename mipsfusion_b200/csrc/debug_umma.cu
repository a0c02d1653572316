// Diagnostics: one 128-point tile through a single tcgen05 layer, built from exactly the primitives the
// decoder kernels use (TMEM A operand written by the owning thread, swizzled K-major bf16 weights in shared
// memory, bf16x3 split, fp32 accumulators read back with tcgen05.ld).  Lets the tensor-core data path be
// validated in isolation against a plain fp32 matmul.
#include "field_tc.cuh"
#include "mf_common.cuh"

// out (128,128) = x (128,K) * w (128,K)^T ; K multiple of 16, <= 128.  passes: 1 = hi*hi only, 3 = bf16x3.
__global__ void __launch_bounds__(128, 1) debug_umma_linear_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                   float* __restrict__ out, int K, int passes, int* err) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* w_hi = base; uint8_t* w_lo = base + 2 * IMG_BLOCK;
    uint64_t* bar = (uint64_t*)(base + 4 * IMG_BLOCK);
    uint32_t* tptr = (uint32_t*)(base + 4 * IMG_BLOCK + 8);
    const int tid = threadIdx.x;
    for (int i = tid; i < 4 * IMG_BLOCK / 4; i += 128) reinterpret_cast<uint32_t*>(base)[i] = 0u;
    __syncthreads();
    for (int idx = tid; idx < 128 * K; idx += 128) {
        const int n = idx / K, k = idx % K;
        const float v = w[n * K + k];
        const __nv_bfloat16 h = __float2bfloat16_rn(v), l = __float2bfloat16_rn(v - __bfloat162float(h));
        const uint32_t off = (uint32_t)(k >> 6) * IMG_BLOCK + sw128_offset(n, k & 63);
        *reinterpret_cast<__nv_bfloat16*>(w_hi + off) = h;
        *reinterpret_cast<__nv_bfloat16*>(w_lo + off) = l;
    }
    umma::fence_proxy_async();
    if ((tid >> 5) == 0) umma::tmem_alloc<512>(tptr);
    if (tid == 0) { umma::mbar_init(bar, 1); umma::fence_barrier_init(); }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem_base = *tptr;
    const uint32_t lane_base = tmem_base + ((uint32_t)((tid >> 5) * 32) << 16);
    // A operand: thread = row; K/2 columns hi at 128.., lo at 192..
    for (int c0 = 0; c0 < K / 2; c0 += 8) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) umma::split2(x[tid * K + 2 * (c0 + t)], x[tid * K + 2 * (c0 + t) + 1], hi[t], lo[t]);
        umma::tmem_st8(lane_base + TM_A_HI + c0, hi);
        umma::tmem_st8(lane_base + TM_A_LO + c0, lo);
    }
    umma::wait_st();
    umma::fence_before_sync();
    __syncthreads();
    if (tid == 0) {
        umma::fence_after_sync();
        constexpr uint32_t idesc = umma::idesc_bf16(128, 128, 0, 0);
        uint32_t acc = 0;
        for (int pass = 0; pass < passes; ++pass)
            for (int ks = 0; ks < K / 16; ++ks) {
                const uint32_t a = tmem_base + (pass == 2 ? TM_A_LO : TM_A_HI) + 8 * ks;
                const uint32_t wb = umma::smem_u32(pass == 1 ? w_lo : w_hi) + (uint32_t)((ks >> 2) * IMG_BLOCK + (ks & 3) * 32);
                umma::mma_ts(tmem_base + TM_D, a, umma::smem_desc_sw128(wb, 16, 1024), idesc, acc);
                acc = 1;
            }
        umma::commit(bar);
    }
    const bool ok = umma::mbar_wait(bar, 0);
    umma::fence_after_sync();
    if (!ok && err) atomicExch(err, 1);
    for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t r[32];
        umma::tmem_ld32(lane_base + TM_D + c0, r);
        umma::wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) out[tid * 128 + c0 + i] = __uint_as_float(r[i]);
    }
    umma::fence_before_sync();
    __syncthreads();
    if ((tid >> 5) == 0) umma::tmem_dealloc<512>(tmem_base);
}

MF_API int mf_debug_umma_linear(const float* x, const float* w, float* out, int K, int passes, void* stream) {
    MF_CHECK_ARG(x && w && out && K >= 16 && K <= 128 && K % 16 == 0 && (passes == 1 || passes == 3));
    const size_t smem = 4 * IMG_BLOCK + 64 + 1024;
    MF_CUDA(cudaFuncSetAttribute(debug_umma_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    debug_umma_linear_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(x, w, out, K, passes, mf_tc_error_flag());
    MF_LAUNCH_CHECK();
    return MF_OK;
}

// ---------------------------------------------------------------------------------------------
// dgrad-style probe: out (128, Kf) = dz (128,128) * w (128, Kf), with dz in TMEM (bf16x3 split) and w read
// as an MN-major B operand from the same [n][k] swizzled image the forward uses.  Kf in {64, 96, 128}.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) debug_umma_dgrad_kernel(const float* __restrict__ dz, const float* __restrict__ w,
                                                                  float* __restrict__ out, int Kf, int* err) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* w_hi = base; uint8_t* w_lo = base + 2 * IMG_BLOCK;
    uint64_t* bar = (uint64_t*)(base + 4 * IMG_BLOCK);
    uint32_t* tptr = (uint32_t*)(base + 4 * IMG_BLOCK + 8);
    const int tid = threadIdx.x;
    for (int i = tid; i < 4 * IMG_BLOCK / 4; i += 128) reinterpret_cast<uint32_t*>(base)[i] = 0u;
    __syncthreads();
    for (int idx = tid; idx < 128 * Kf; idx += 128) {
        const int n = idx / Kf, k = idx % Kf;
        const float v = w[n * Kf + k];
        const __nv_bfloat16 h = __float2bfloat16_rn(v), l = __float2bfloat16_rn(v - __bfloat162float(h));
        const uint32_t off = (uint32_t)(k >> 6) * IMG_BLOCK + sw128_offset(n, k & 63);
        *reinterpret_cast<__nv_bfloat16*>(w_hi + off) = h;
        *reinterpret_cast<__nv_bfloat16*>(w_lo + off) = l;
    }
    umma::fence_proxy_async();
    if ((tid >> 5) == 0) umma::tmem_alloc<512>(tptr);
    if (tid == 0) { umma::mbar_init(bar, 1); umma::fence_barrier_init(); }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem_base = *tptr;
    const uint32_t lane_base = tmem_base + ((uint32_t)((tid >> 5) * 32) << 16);
    for (int c0 = 0; c0 < 64; c0 += 8) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) umma::split2(dz[tid * 128 + 2 * (c0 + t)], dz[tid * 128 + 2 * (c0 + t) + 1], hi[t], lo[t]);
        umma::tmem_st8(lane_base + TM_A_HI + c0, hi);
        umma::tmem_st8(lane_base + TM_A_LO + c0, lo);
    }
    umma::wait_st();
    umma::fence_before_sync();
    __syncthreads();
    if (tid == 0) {
        umma::fence_after_sync();
        const uint32_t idesc = umma::idesc_bf16(128, Kf, 0, 1);            // B is MN-major
        uint32_t acc = 0;
        for (int pass = 0; pass < 3; ++pass)
            for (int ks = 0; ks < 8; ++ks) {                                // K = n = 128
                const uint32_t a = tmem_base + (pass == 2 ? TM_A_LO : TM_A_HI) + 8 * ks;
                const uint32_t wb = umma::smem_u32(pass == 1 ? w_lo : w_hi) + (uint32_t)(ks * 2048);
                umma::mma_ts(tmem_base + TM_D, a, umma::smem_desc_sw128(wb, IMG_BLOCK, 1024), idesc, acc);
                acc = 1;
            }
        umma::commit(bar);
    }
    const bool ok = umma::mbar_wait(bar, 0);
    umma::fence_after_sync();
    if (!ok && err) atomicExch(err, 1);
    for (int c0 = 0; c0 < Kf; c0 += 32) {
        uint32_t r[32];
        umma::tmem_ld32(lane_base + TM_D + c0, r);
        umma::wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) out[tid * Kf + c0 + i] = __uint_as_float(r[i]);
    }
    umma::fence_before_sync();
    __syncthreads();
    if ((tid >> 5) == 0) umma::tmem_dealloc<512>(tmem_base);
}

// ---------------------------------------------------------------------------------------------
// wgrad-style probe: out (128, Kf) [n][k] = sum_p dz[p][n] * x[p][k] over P = 128 points, both operands
// MN-major from [p][feature] bf16 tiles in shared memory, processed as two 64-point half tiles.
// passes = 1: dz_hi*x_hi; 2: + dz_lo*x_hi.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512, 1) debug_umma_wgrad_kernel(const float* __restrict__ dz, const float* __restrict__ x,
                                                                  float* __restrict__ out, int Kf, int passes, int* err) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    constexpr uint32_t BLK = 64 * 128;                   // 64 rows x 128 B
    uint8_t* z_hi = base; uint8_t* z_lo = base + 2 * BLK; uint8_t* x_hi = base + 4 * BLK;
    uint64_t* bar = (uint64_t*)(base + 6 * BLK);
    uint32_t* tptr = (uint32_t*)(base + 6 * BLK + 8);
    const int tid = threadIdx.x, p = tid & 127, q = tid >> 7;
    if ((tid >> 5) == 0) umma::tmem_alloc<512>(tptr);
    if (tid == 0) { umma::mbar_init(bar, 1); umma::fence_barrier_init(); }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem_base = *tptr;
    const uint32_t lane_base = tmem_base + ((uint32_t)(((tid >> 5) & 3) * 32) << 16);
    const uint32_t idesc = umma::idesc_bf16(128, Kf, 1, 1);
    for (int half = 0; half < 2; ++half) {
        if ((p >> 6) == half) {
            const int row = p & 63;
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) umma::split2(dz[p * 128 + 32 * q + 2 * i], dz[p * 128 + 32 * q + 2 * i + 1], hi[i], lo[i]);
            umma::store_row32(z_hi, row, q, hi, BLK);
            umma::store_row32(z_lo, row, q, lo, BLK);
            if (32 * q < Kf) {
                uint32_t xh[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) xh[i] = umma::pack_bf16(x[p * Kf + 32 * q + 2 * i], x[p * Kf + 32 * q + 2 * i + 1]);
                umma::store_row32(x_hi, row, q, xh, BLK);
            }
        }
        umma::fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            umma::fence_after_sync();
            for (int pass = 0; pass < passes; ++pass)
                for (int ks = 0; ks < 4; ++ks)                       // K = 64 points of this half
                    umma::mma_ss(tmem_base + TM_D, umma::desc_mn(pass ? z_lo : z_hi, 16 * ks, BLK), umma::desc_mn(x_hi, 16 * ks, BLK),
                                 idesc, (half | pass | ks) ? 1u : 0u);
            umma::commit(bar);
        }
        const bool ok = umma::mbar_wait(bar, (uint32_t)half);        // buffers are reused by the next half
        umma::fence_after_sync();
        if (!ok && err) atomicExch(err, 1);
        __syncthreads();
    }
    // read-out: thread (lane n = p, q) owns out[n][32q .. 32q+32)
    if (32 * q < Kf) {
        uint32_t r[32];
        umma::tmem_ld32(lane_base + TM_D + 32 * q, r);
        umma::wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) out[p * Kf + 32 * q + i] = __uint_as_float(r[i]);
    }
    umma::fence_before_sync();
    __syncthreads();
    if ((tid >> 5) == 0) umma::tmem_dealloc<512>(tmem_base);
}

MF_API int mf_debug_umma_dgrad(const float* dz, const float* w, float* out, int Kf, void* stream) {
    MF_CHECK_ARG(dz && w && out && (Kf == 64 || Kf == 96 || Kf == 128));
    const size_t smem = 4 * IMG_BLOCK + 64 + 1024;
    MF_CUDA(cudaFuncSetAttribute(debug_umma_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    debug_umma_dgrad_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(dz, w, out, Kf, mf_tc_error_flag());
    MF_LAUNCH_CHECK();
    return MF_OK;
}

MF_API int mf_debug_umma_wgrad(const float* dz, const float* x, float* out, int Kf, int passes, void* stream) {
    MF_CHECK_ARG(dz && x && out && (Kf == 64 || Kf == 96 || Kf == 128) && (passes == 1 || passes == 2));
    const size_t smem = 6 * 64 * 128 + 64 + 1024;
    MF_CUDA(cudaFuncSetAttribute(debug_umma_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    debug_umma_wgrad_kernel<<<1, 512, smem, (cudaStream_t)stream>>>(dz, x, out, Kf, passes, mf_tc_error_flag());
    MF_LAUNCH_CHECK();
    return MF_OK;
}

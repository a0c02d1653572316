// Tensor-core (tcgen05) backward of the fused field evaluation, 128 points per tile:
//   recompute forward -> head backward -> for each 128-wide layer { wgrad, dgrad } -> grid scatter (+ dL/dx).
//
//   forward / dgrad : A operand (activations / gradients, bf16 hi+lo) in tensor memory, written by the thread
//                     that owns the point; B operand = the resident swizzled weight image, read K-major
//                     (forward) or MN-major (dgrad: the same bytes, transposed by the descriptor); bf16x3.
//   wgrad           : dW[n][k] = sum_p dZ[p][n] X[p][k]: both operands MN-major from [point][feature] bf16
//                     tiles in shared memory (two 64-point half tiles), passes dZ_hi X_hi + dZ_lo X_hi +
//                     dZ_hi X_lo; the 128 x K fp32 result is read back from tensor memory and added to this
//                     CTA's private partial of the parameter gradient.
//   narrow heads    : (5 x 128 logits, 3 x 115 colour) on CUDA cores with warp reduce-scatter.
#pragma once
#include "field_tc.cuh"

// ---- shared memory map ---------------------------------------------------------------------------
constexpr int BS_W2 = 0;                               // 64 KB: image bytes [IMG_W2_HI, IMG_W3_HI)
constexpr int BS_W3 = 65536;                           // 64 KB: image bytes [IMG_W3_HI, IMG_F32)
constexpr int BS_R1 = 131072;                          // 32 KB: W1 image (hi, lo)  |  wgrad dZ tiles (hi, lo)
constexpr int BS_R2 = 163840;                          // 32 KB: wgrad X tiles (hi, lo)
constexpr int BS_F32 = 196608;                         // fp32 section of the image
constexpr int BS_PART = BS_F32 + ((F_COUNT * 4 + 127) / 128) * 128;
constexpr int BS_PART_ROWS = 32;                       // 20 logit partial rows + 12 dx partial rows
constexpr int BS_BAR = BS_PART + BS_PART_ROWS * TC_LD * 4;
constexpr int BS_BYTES = BS_BAR + 32;
constexpr size_t SMEM_TC_BWD = BS_BYTES + 1024;
static_assert(SMEM_TC_BWD <= 227 * 1024, "shared memory budget");
constexpr uint32_t HALF_BLK = 64 * 128;                // one 64-feature block of a 64-row tile

// ---- tensor memory map (512 columns) -------------------------------------------------------------
constexpr int TB_D = 0;                                // 128 fp32 accumulator columns (forward, dgrad, wgrad results)
constexpr int TB_OP1_HI = 128, TB_OP1_LO = 192;        // operand 1: e (layer 1), then H1 (kept until wgrad of layer 2)
constexpr int TB_OP2_HI = 256, TB_OP2_LO = 320;        // operand 2: sdf_emb, then dZ3, dH, dZ1
constexpr int TB_G_HI = TB_OP2_HI + 32, TB_G_LO = TB_OP2_LO + 32;   // grid features (dead before dZ3 is written)

struct TbCtx {
    uint8_t *w2, *w3, *r1, *r2;
    const float* fw; float* part; uint64_t* bar; uint32_t* tmem_ptr;
    uint32_t tmem_base, lane_base, phase;
    bool ok;
};

__device__ __forceinline__ void tb_copy(uint8_t* dst, const uint8_t* __restrict__ src, int bytes) {
    for (int i = threadIdx.x; i < bytes / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(dst)[i] = __ldg(reinterpret_cast<const uint4*>(src) + i);
}

__device__ __forceinline__ void tb_setup(TbCtx& c, uint8_t* smem_raw, const uint8_t* __restrict__ img) {
    uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    c.w2 = base + BS_W2; c.w3 = base + BS_W3; c.r1 = base + BS_R1; c.r2 = base + BS_R2;
    c.fw = (const float*)(base + BS_F32); c.part = (float*)(base + BS_PART);
    c.bar = (uint64_t*)(base + BS_BAR); c.tmem_ptr = (uint32_t*)(base + BS_BAR + 8);
    tb_copy(c.w2, img + IMG_W2_HI, 4 * IMG_BLOCK);
    tb_copy(c.w3, img + IMG_W3_HI, 4 * IMG_BLOCK);
    tb_copy(base + BS_F32, img + IMG_F32, F_COUNT * 4);
    umma::fence_proxy_async();
    const int tid = threadIdx.x;
    if ((tid >> 5) == 0) umma::tmem_alloc<512>(c.tmem_ptr);
    if (tid == 0) { umma::mbar_init(c.bar, 1); umma::fence_barrier_init(); }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    c.tmem_base = *c.tmem_ptr;
    c.lane_base = c.tmem_base + ((uint32_t)(((tid >> 5) & 3) * 32) << 16);
    c.phase = 0; c.ok = true;
}

// One tensor-core round: everybody publishes its TMEM / shared-memory writes, thread 0 issues the MMAs given by
// `issue` and commits, everybody waits for completion.
template <class IssueFn>
__device__ __forceinline__ void tb_round(TbCtx& c, IssueFn issue) {
    umma::wait_st();
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    if (threadIdx.x == 0) {
        umma::fence_after_sync();
        issue();
        umma::commit(c.bar);
    }
    c.ok &= umma::mbar_wait(c.bar, c.phase);
    c.phase ^= 1;
    umma::fence_after_sync();
}

// forward layer: D = A W^T, A at TMEM columns a_col(ks, lo), W image K-major at (w_hi, w_lo)
template <class ColFn>
__device__ __forceinline__ void tb_issue_fwd(const TbCtx& c, const uint8_t* w_hi, const uint8_t* w_lo, int KS, ColFn a_col) {
    constexpr uint32_t idesc = umma::idesc_bf16(128, 128, 0, 0);
    const uint32_t wh = umma::smem_u32(w_hi), wl = umma::smem_u32(w_lo);
    uint32_t acc = 0;
    // (rolled: one thread issues, and the kernel's instruction footprint is what the instruction cache has to hold)
#pragma unroll 1
    for (int pass = 0; pass < 3; ++pass)
#pragma unroll 1
        for (int ks = 0; ks < KS; ++ks) {
            const uint32_t wb = (pass == 1 ? wl : wh) + (uint32_t)((ks >> 2) * IMG_BLOCK + (ks & 3) * 32);
            umma::mma_ts(c.tmem_base + TB_D, c.tmem_base + (uint32_t)a_col(ks, pass == 2), umma::smem_desc_sw128(wb, 16, 1024), idesc, acc);
            acc = 1;
        }
}

// dgrad: D[p][k] = sum_n dZ[p][n] W[n][k], dZ (128 features) in operand 2, W image read MN-major, N = n_out columns
__device__ __forceinline__ void tb_issue_dgrad(const TbCtx& c, const uint8_t* w_hi, const uint8_t* w_lo, int n_out) {
    const uint32_t idesc = umma::idesc_bf16(128, n_out, 0, 1);
    const uint32_t wh = umma::smem_u32(w_hi), wl = umma::smem_u32(w_lo);
    uint32_t acc = 0;
#pragma unroll 1
    for (int pass = 0; pass < 3; ++pass)
#pragma unroll 1
        for (int ks = 0; ks < 8; ++ks) {
            const uint32_t a = c.tmem_base + (uint32_t)((pass == 2 ? TB_OP2_LO : TB_OP2_HI) + 8 * ks);
            const uint32_t wb = (pass == 1 ? wl : wh) + (uint32_t)(ks * 2048);
            umma::mma_ts(c.tmem_base + TB_D, a, umma::smem_desc_sw128(wb, IMG_BLOCK, 1024), idesc, acc);
            acc = 1;
        }
}

// wgrad on one 64-point half tile: D[n][k] (+)= sum_p dZ[p][n] X[p][k]
__device__ __forceinline__ void tb_issue_wgrad(const TbCtx& c, int n_out, bool first) {
    const uint32_t idesc = umma::idesc_bf16(128, n_out, 1, 1);
    const uint8_t *z_hi = c.r1, *z_lo = c.r1 + 2 * HALF_BLK, *x_hi = c.r2, *x_lo = c.r2 + 2 * HALF_BLK;
    uint32_t acc = first ? 0u : 1u;
#pragma unroll 1
    for (int pass = 0; pass < 3; ++pass)
#pragma unroll 1
        for (int ks = 0; ks < 4; ++ks) {
            umma::mma_ss(c.tmem_base + TB_D, umma::desc_mn(pass == 1 ? z_lo : z_hi, 16 * ks, HALF_BLK),
                         umma::desc_mn(pass == 2 ? x_lo : x_hi, 16 * ks, HALF_BLK), idesc, acc);
            acc = 1;
        }
}

// Warp reduce-scatter: on return v[0] of lane l holds the sum over the 32 lanes of element l.
__device__ __forceinline__ float rs32(float (&v)[32], int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; ++i) {
            const float send = up ? v[i] : v[i + s];
            const float keep = up ? v[i + s] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    return v[0];
}
// 16 elements over 32 lanes: lanes l and l ^ 16 both end with the sum of element l & 15.
__device__ __forceinline__ float rs16(float (&v)[16], int lane) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], 16);
#pragma unroll
    for (int s = 8; s >= 1; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; ++i) {
            const float send = up ? v[i] : v[i + s];
            const float keep = up ? v[i + s] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    return v[0];
}

// the 16 layer-1 input slots owned by thread q of a point (12 frequency features, xyz for q = 0, zero padding)
__device__ __forceinline__ void tb_e_slots(const float (&x)[3], int q, float (&e)[16]) {
#pragma unroll
    for (int jj = 0; jj < 12; ++jj) {
        const int j = q * 12 + jj, d = j >> 4, k = (j & 15) >> 1, s = j & 1;
        e[jj] = sin_reduced(freq_arg(x[d], k, s));
    }
    e[12] = q == 0 ? x[0] : 0.f; e[13] = q == 0 ? x[1] : 0.f; e[14] = q == 0 ? x[2] : 0.f; e[15] = 0.f;
}

__device__ __forceinline__ void tb_load32(const TbCtx& c, int col, float (&v)[32]) {
    uint32_t r[32];
    umma::tmem_ld32(c.lane_base + (uint32_t)col, r);
    umma::wait_ld();
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 32 features [32q, 32q+32) of this thread's point -> operand region (hi at hi_col, lo at lo_col), bf16 pairs
__device__ __forceinline__ void tb_store_op32(const TbCtx& c, int hi_col, int lo_col, int q, const float (&v)[32]) {
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) umma::split2(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
    umma::tmem_st16(c.lane_base + (uint32_t)(hi_col + 16 * q), hi);
    umma::tmem_st16(c.lane_base + (uint32_t)(lo_col + 16 * q), lo);
}

// wgrad of one 128-wide layer.  z: this thread's 32 gradient values (features [32q,32q+32) of point p).
// load_x(hi, lo) fills the packed bf16 pairs of this thread's slice of the activation operand (or returns false
// when the thread has no slice).  X has n_x features (64, 96 or 128).  The result D[n][k] is added to gW, whose
// element (n, k) lives at gW[kmap(k) * 128 + n] (transposed, so that a warp's reductions coalesce; kmap(k) < 0: skip;
// reduce_partials_kernel restores the (out, in) layout).
template <class LoadX, class KMap>
__device__ __forceinline__ void tb_wgrad_layer(TbCtx& c, int p, int q, const float (&z)[32], LoadX load_x, int n_x,
                                               float* __restrict__ gW, int ldw, KMap kmap) {
    uint8_t *z_hi = c.r1, *z_lo = c.r1 + 2 * HALF_BLK, *x_hi = c.r2, *x_lo = c.r2 + 2 * HALF_BLK;
    for (int half = 0; half < 2; ++half) {
        if ((p >> 6) == half) {                          // warp-uniform; packed operands are transient registers
            const int row = p & 63;
            {
                uint32_t zh[16], zl[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) umma::split2(z[2 * i], z[2 * i + 1], zh[i], zl[i]);
                umma::store_row32(z_hi, row, q, zh, HALF_BLK);
                umma::store_row32(z_lo, row, q, zl, HALF_BLK);
            }
            {
                uint32_t xh[16], xl[16];
                if (load_x(xh, xl)) { umma::store_row32(x_hi, row, q, xh, HALF_BLK); umma::store_row32(x_lo, row, q, xl, HALF_BLK); }
            }
        }
        tb_round(c, [&]() { tb_issue_wgrad(c, n_x, half == 0); });
    }
    // read-out: thread (n = p, q) owns D[n][32q .. 32q+32); added to the CTA-private partial with reductions
    // (fire-and-forget, no load latency)
    if (32 * q < n_x) {
#pragma unroll
        for (int c0 = 0; c0 < 32; c0 += 8) {
            uint32_t r[8];
            umma::tmem_ld8(c.lane_base + (uint32_t)(TB_D + 32 * q + c0), r);
            umma::wait_ld();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int k = kmap(32 * q + c0 + i);
                if (k >= 0) atomicAdd(&gW[k * D_H + p], __uint_as_float(r[i]));   // [k][n]: lanes (n) are contiguous
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// The kernel.  part: [gridDim.x][MF_MLP_PARAMS] per-CTA partial parameter gradients (zeroed here).
// ---------------------------------------------------------------------------------------------
// PARAMS = false (with WANT_DX): input gradients only -- the gradient pose refinement of tracking (mipsfusion.py:501-556) needs
// d loss / d rays and nothing else: no wgrad products, no narrow-head reductions, no table scatter, W1 stays resident.
template <class Src, bool WANT_DX, bool PARAMS = true>
__global__ void __launch_bounds__(TC_NT, 1) field_bwd_tc_kernel(FieldDev f, Src src, const float* __restrict__ d_raw,
                                                                float* __restrict__ grad_grid, float* __restrict__ part,
                                                                float* __restrict__ d_pts, int64_t N_all, ActiveMap am,
                                                                int* __restrict__ err, long long* __restrict__ prof) {
    extern __shared__ uint8_t smem_raw[];
    const int tid = threadIdx.x, p = tid & (TC_TP - 1), q = tid >> 7, lane = tid & 31;
#define TB_MARK(k) do { if (prof && tid == 0 && blockIdx.x == 0 && tile == (int64_t)gridDim.x) prof[k] = clock64(); } while (0)
    float* gpart = PARAMS ? part + (size_t)blockIdx.x * MF_MLP_PARAMS : nullptr;
    if (PARAMS)
        for (int i = tid; i < MF_MLP_PARAMS; i += TC_NT) gpart[i] = 0.f;
    TbCtx c;
    tb_setup(c, smem_raw, f.tc_img);
    if (!PARAMS) { tb_copy(c.r1, f.tc_img + IMG_W1_HI, 2 * IMG_BLOCK); umma::fence_proxy_async(); }   // (published by the first round's barrier)
    const float2* grid2 = reinterpret_cast<const float2*>(f.grid);

    const int64_t N = am.n(N_all);                     // active points only (ascending point indices in am.idx)
    const int64_t n_tiles = (N + TC_TP - 1) / TC_TP;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t slot = tile * TC_TP + p;
        const bool valid = slot < N;
        const int64_t i = valid ? am(slot) : 0;
        TB_MARK(0);
        // ---- W1 image into region 1 (it doubles as the wgrad dZ tile later in the tile) ----
        if (PARAMS) tb_copy(c.r1, f.tc_img + IMG_W1_HI, 2 * IMG_BLOCK);
        // ---- upstream gradient of the colour outputs of this point (the sdf-head part is loaded at the heads) ----
        float g[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) g[k] = valid ? d_raw[i * MF_RAW_DIM + k] : 0.f;
        // ---- encode (or reload the operand words the forward kernel cached) ----
        float x[3] = {0.f, 0.f, 0.f};
        if (valid) src.point(i, f, x);
        // the forward cached the operand words of point i at [i / 128][q][word][i % 128]
        const uint32_t* fin = f.feat ? f.feat + (size_t)(i >> 7) * FEAT_TILE_WORDS + (size_t)q * FEAT_WORDS * TC_TP + (i & (TC_TP - 1)) : nullptr;
        {
            float e[16];
            uint32_t hi[8], lo[8];
            if (fin) {
#pragma unroll
                for (int t = 0; t < 8; ++t) { hi[t] = __ldg(fin + t * TC_TP); lo[t] = __ldg(fin + (8 + t) * TC_TP); }
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    e[2 * t] = __uint_as_float(hi[t] << 16) + __uint_as_float(lo[t] << 16);
                    e[2 * t + 1] = __uint_as_float(hi[t] & 0xffff0000u) + __uint_as_float(lo[t] & 0xffff0000u);
                }
            } else {
                tb_e_slots(x, q, e);
#pragma unroll
                for (int t = 0; t < 8; ++t) umma::split2(e[2 * t], e[2 * t + 1], hi[t], lo[t]);
            }
            umma::tmem_st8(c.lane_base + TB_OP1_HI + 8 * q, hi);
            umma::tmem_st8(c.lane_base + TB_OP1_LO + 8 * q, lo);
            // colour head, e part: dWr[c][64 + e_index(slot)] += dRGB[c] e[slot]
#pragma unroll 1
            for (int ch = 0; ch < (PARAMS ? 3 : 0); ++ch) {
                float t[16];
#pragma unroll
                for (int s = 0; s < 16; ++s) t[s] = g[ch] * e[s];
                const float r = rs16(t, lane);
                const int ei = tc_e_slot_to_index(16 * q + (lane & 15));
                if (lane < 16 && ei >= 0) atomicAdd(&gpart[OFF_WR + ch * D_RGB_IN + 64 + ei], r);
            }
        }
        {
            uint32_t hi[4], lo[4];
            if (fin) {
#pragma unroll
                for (int t = 0; t < 4; ++t) { hi[t] = __ldg(fin + (16 + t) * TC_TP); lo[t] = __ldg(fin + (20 + t) * TC_TP); }
            } else {
                float gf[8];
#pragma unroll
                for (int ll = 0; ll < 4; ++ll) {
                    float2 vv = make_float2(0.f, 0.f);
                    if (valid) vv = grid_level_fwd(x, grid2, level_info(f, q * 4 + ll), nullptr);
                    gf[2 * ll] = vv.x; gf[2 * ll + 1] = vv.y;
                }
#pragma unroll
                for (int t = 0; t < 4; ++t) umma::split2(gf[2 * t], gf[2 * t + 1], hi[t], lo[t]);
            }
            umma::tmem_st4(c.lane_base + TB_G_HI + 4 * q, hi);
            umma::tmem_st4(c.lane_base + TB_G_LO + 4 * q, lo);
        }
        TB_MARK(1);
        float v[32];
        uint32_t mask1 = 0, mask3 = 0;
        // ---- forward layer 1 ----
        tb_round(c, [&]() { tb_issue_fwd(c, c.r1, c.r1 + IMG_BLOCK, 4, [](int ks, bool lo) { return (lo ? TB_OP1_LO : TB_OP1_HI) + 8 * ks; }); });
        tb_load32(c, TB_D + 32 * q, v);
        {
            const float4* b4 = reinterpret_cast<const float4*>(c.fw + F_B1 + 32 * q);
#pragma unroll
            for (int k4 = 0; k4 < 8; ++k4) {
                const float4 b = b4[k4];
                v[4 * k4] = fmaxf(v[4 * k4] + b.x, 0.f); v[4 * k4 + 1] = fmaxf(v[4 * k4 + 1] + b.y, 0.f);
                v[4 * k4 + 2] = fmaxf(v[4 * k4 + 2] + b.z, 0.f); v[4 * k4 + 3] = fmaxf(v[4 * k4 + 3] + b.w, 0.f);
            }
#pragma unroll
            for (int k = 0; k < 32; ++k) mask1 |= (v[k] > 0.f ? 1u : 0u) << k;
        }
        tb_store_op32(c, TB_OP1_HI, TB_OP1_LO, q, v);                        // H1 stays in operand 1 until wgrad of layer 2
        TB_MARK(2);
        // ---- forward layer 2 ----
        tb_round(c, [&]() { tb_issue_fwd(c, c.w2, c.w2 + 2 * IMG_BLOCK, 8, [](int ks, bool lo) { return (lo ? TB_OP1_LO : TB_OP1_HI) + 8 * ks; }); });
        tb_load32(c, TB_D + 32 * q, v);
        {
            const float4* b4 = reinterpret_cast<const float4*>(c.fw + F_B2 + 32 * q);
#pragma unroll
            for (int k4 = 0; k4 < 8; ++k4) {
                const float4 b = b4[k4];
                v[4 * k4] += b.x; v[4 * k4 + 1] += b.y; v[4 * k4 + 2] += b.z; v[4 * k4 + 3] += b.w;
            }
        }
        if (q < 2) {
            tb_store_op32(c, TB_OP2_HI, TB_OP2_LO, q, v);                    // sdf_emb
        } else {                                                             // colour head, rgb_emb part
#pragma unroll 1
            for (int ch = 0; ch < (PARAMS ? 3 : 0); ++ch) {
                float t[32];
#pragma unroll
                for (int k = 0; k < 32; ++k) t[k] = g[ch] * v[k];
                atomicAdd(&gpart[OFF_WR + ch * D_RGB_IN + 32 * (q - 2) + lane], rs32(t, lane));
            }
        }
        TB_MARK(3);
        // ---- forward layer 3 ----
        tb_round(c, [&]() {
            tb_issue_fwd(c, c.w3, c.w3 + 2 * IMG_BLOCK, 6, [](int ks, bool lo) {
                return ks < 4 ? (lo ? TB_OP2_LO : TB_OP2_HI) + 8 * ks : (lo ? TB_G_LO : TB_G_HI) + 8 * (ks - 4);
            });
        });
        tb_load32(c, TB_D + 32 * q, v);
        {
            float s[N_CLASS] = {0.f, 0.f, 0.f, 0.f, 0.f};
            const float4* b4 = reinterpret_cast<const float4*>(c.fw + F_BS1 + 32 * q);
#pragma unroll
            for (int k4 = 0; k4 < 8; ++k4) {
                const float4 b = b4[k4];
                v[4 * k4] = fmaxf(v[4 * k4] + b.x, 0.f); v[4 * k4 + 1] = fmaxf(v[4 * k4 + 1] + b.y, 0.f);
                v[4 * k4 + 2] = fmaxf(v[4 * k4 + 2] + b.z, 0.f); v[4 * k4 + 3] = fmaxf(v[4 * k4 + 3] + b.w, 0.f);
            }
#pragma unroll
            for (int k = 0; k < 32; ++k) mask3 |= (v[k] > 0.f ? 1u : 0u) << k;
#pragma unroll
            for (int ch = 0; ch < N_CLASS; ++ch) {
                const float4* w4 = reinterpret_cast<const float4*>(c.fw + F_WS2 + ch * 128 + 32 * q);
#pragma unroll
                for (int k4 = 0; k4 < 8; ++k4) {
                    const float4 w = w4[k4];
                    s[ch] = fmaf(w.x, v[4 * k4], s[ch]); s[ch] = fmaf(w.y, v[4 * k4 + 1], s[ch]);
                    s[ch] = fmaf(w.z, v[4 * k4 + 2], s[ch]); s[ch] = fmaf(w.w, v[4 * k4 + 3], s[ch]);
                }
            }
#pragma unroll
            for (int ch = 0; ch < N_CLASS; ++ch) c.part[(q * 5 + ch) * TC_LD + p] = s[ch];
        }
        TB_MARK(4);
        __syncthreads();
        // ---- heads: softmax forward + backward (all four threads of a point, redundantly) ----
        float dz4[N_CLASS];
        {
            float gs[7];                                   // d loss / d (sdf, entropy, prob[5])
#pragma unroll
            for (int k = 0; k < 7; ++k) gs[k] = valid ? d_raw[i * MF_RAW_DIM + 3 + k] : 0.f;
            float zl[N_CLASS], pr[N_CLASS], dp[N_CLASS];
            float mx = -INFINITY, se = 0.f, dot = 0.f;
#pragma unroll
            for (int ch = 0; ch < N_CLASS; ++ch) {
                zl[ch] = c.fw[F_BS2 + ch] + ((c.part[ch * TC_LD + p] + c.part[(5 + ch) * TC_LD + p]) + (c.part[(10 + ch) * TC_LD + p] + c.part[(15 + ch) * TC_LD + p]));
                mx = fmaxf(mx, zl[ch]);
            }
#pragma unroll
            for (int ch = 0; ch < N_CLASS; ++ch) { pr[ch] = expf(zl[ch] - mx); se += pr[ch]; }
#pragma unroll
            for (int ch = 0; ch < N_CLASS; ++ch) {
                pr[ch] = pr[ch] / se;
                const float qq = pr[ch] + 1e-5f;
                dp[ch] = gs[2 + ch] + gs[0] * (0.5f * (float)ch) - gs[1] * (log2f(qq) + pr[ch] / (qq * 0.6931471805599453f));
                dot = fmaf(pr[ch], dp[ch], dot);
            }
#pragma unroll
            for (int ch = 0; ch < N_CLASS; ++ch) dz4[ch] = pr[ch] * (dp[ch] - dot);
        }
        // sdf_linear.2 weight gradient: dW4[c][32q + l] += sum_p dz4[c] h3[l]; biases from the q == 0 warps
#pragma unroll 1
        for (int ch = 0; ch < (PARAMS ? N_CLASS : 0); ++ch) {
            float t[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) t[k] = dz4[ch] * v[k];
            atomicAdd(&gpart[OFF_WS2 + ch * D_H + 32 * q + lane], rs32(t, lane));
            if (q == 0) { const float b = warp_sum(dz4[ch]); if (lane == 0) atomicAdd(&gpart[OFF_BS2 + ch], b); }
        }
        if (PARAMS && q == 0) {
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) { const float b = warp_sum(g[ch]); if (lane == 0) atomicAdd(&gpart[OFF_BR + ch], b); }
        }
        // dZ3 = (Ws2^T dz4) * relu'(h3)
        {
#pragma unroll
            for (int k = 0; k < 32; ++k) v[k] = 0.f;
#pragma unroll
            for (int ch = 0; ch < N_CLASS; ++ch) {
                const float4* w4 = reinterpret_cast<const float4*>(c.fw + F_WS2 + ch * 128 + 32 * q);
                const float dzc = dz4[ch];
#pragma unroll
                for (int k4 = 0; k4 < 8; ++k4) {
                    const float4 w = w4[k4];
                    v[4 * k4] = fmaf(w.x, dzc, v[4 * k4]); v[4 * k4 + 1] = fmaf(w.y, dzc, v[4 * k4 + 1]);
                    v[4 * k4 + 2] = fmaf(w.z, dzc, v[4 * k4 + 2]); v[4 * k4 + 3] = fmaf(w.w, dzc, v[4 * k4 + 3]);
                }
            }
#pragma unroll
            for (int k = 0; k < 32; ++k) v[k] = ((mask3 >> k) & 1u) ? v[k] : 0.f;
        }
        if (PARAMS) { float t[32];
#pragma unroll
          for (int k = 0; k < 32; ++k) t[k] = v[k];
          atomicAdd(&gpart[OFF_BS1 + 32 * q + lane], rs32(t, lane)); }
        TB_MARK(5);
        // ---- layer 3: wgrad (X = [sdf_emb (operand 2), grid features]), then dgrad ----
        if (PARAMS) tb_wgrad_layer(c, p, q, v,
            [&](uint32_t (&xh)[16], uint32_t (&xl)[16]) {
                if (q < 2) { umma::tmem_ld16(c.lane_base + TB_OP2_HI + 16 * q, xh); umma::tmem_ld16(c.lane_base + TB_OP2_LO + 16 * q, xl); }
                else if (q == 2) { umma::tmem_ld16(c.lane_base + TB_G_HI, xh); umma::tmem_ld16(c.lane_base + TB_G_LO, xl); }
                umma::wait_ld();
                return q < 3;
            },
            D_SDF_IN, gpart + OFF_WS1, D_SDF_IN, [](int k) { return k < D_SDF_IN ? k : -1; });
        TB_MARK(6);
        tb_store_op32(c, TB_OP2_HI, TB_OP2_LO, q, v);                        // dZ3 (overwrites sdf_emb + grid features)
        tb_round(c, [&]() { tb_issue_dgrad(c, c.w3, c.w3 + 2 * IMG_BLOCK, D_SDF_IN); });
        // d grid features: D columns [64 + 8q, 64 + 8q + 8) -> scatter into the table (+ dL/dx through the grid)
        float dx[3] = {0.f, 0.f, 0.f};
        {
            uint32_t r[8];
            umma::tmem_ld8(c.lane_base + TB_D + 64 + 8 * q, r);
            umma::wait_ld();
            if (valid) {
#pragma unroll 1
                for (int ll = 0; ll < 4; ++ll) {
                    const float2 dy = ll == 0 ? make_float2(__uint_as_float(r[0]), __uint_as_float(r[1]))
                                    : ll == 1 ? make_float2(__uint_as_float(r[2]), __uint_as_float(r[3]))
                                    : ll == 2 ? make_float2(__uint_as_float(r[4]), __uint_as_float(r[5]))
                                              : make_float2(__uint_as_float(r[6]), __uint_as_float(r[7]));
                    grid_level_bwd<WANT_DX, PARAMS>(x, dy, grid2, grad_grid, level_info(f, q * 4 + ll), dx);
                }
            }
        }
        TB_MARK(7);
        // dH = [d sdf_emb (dgrad of layer 3), d rgb_emb (colour head)]
        if (q < 2) {
            tb_load32(c, TB_D + 32 * q, v);
        } else {
            const float* wr = c.fw + F_WR_EMB + 32 * (q - 2);
#pragma unroll
            for (int k = 0; k < 32; ++k) v[k] = fmaf(wr[128 + k], g[2], fmaf(wr[64 + k], g[1], wr[k] * g[0]));
        }
        if (PARAMS) { float t[32];
#pragma unroll
          for (int k = 0; k < 32; ++k) t[k] = v[k];
          atomicAdd(&gpart[OFF_B2 + 32 * q + lane], rs32(t, lane)); }
        TB_MARK(8);
        // ---- layer 2: wgrad (X = H1 from operand 1), then dgrad ----
        if (PARAMS) tb_wgrad_layer(c, p, q, v,
            [&](uint32_t (&xh)[16], uint32_t (&xl)[16]) {
                umma::tmem_ld16(c.lane_base + TB_OP1_HI + 16 * q, xh); umma::tmem_ld16(c.lane_base + TB_OP1_LO + 16 * q, xl);
                umma::wait_ld();
                return true;
            },
            D_H, gpart + OFF_W2, D_H, [](int k) { return k; });
        TB_MARK(9);
        tb_store_op32(c, TB_OP2_HI, TB_OP2_LO, q, v);                        // dH
        tb_round(c, [&]() { tb_issue_dgrad(c, c.w2, c.w2 + 2 * IMG_BLOCK, D_H); });
        tb_load32(c, TB_D + 32 * q, v);
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = ((mask1 >> k) & 1u) ? v[k] : 0.f;   // dZ1
        if (PARAMS) { float t[32];
#pragma unroll
          for (int k = 0; k < 32; ++k) t[k] = v[k];
          atomicAdd(&gpart[OFF_B1 + 32 * q + lane], rs32(t, lane)); }
        TB_MARK(10);
        // ---- layer 1: wgrad (X = e, 64 slots: this thread owns slots [16q, 16q+16), recomputed here) ----
        if (PARAMS) {
            uint8_t *z_hi = c.r1, *z_lo = c.r1 + 2 * HALF_BLK, *x_hi = c.r2, *x_lo = c.r2 + 2 * HALF_BLK;
            for (int half = 0; half < 2; ++half) {
                if ((p >> 6) == half) {
                    const int row = p & 63;
                    {
                        uint32_t zh[16], zl[16];
#pragma unroll
                        for (int k = 0; k < 16; ++k) umma::split2(v[2 * k], v[2 * k + 1], zh[k], zl[k]);
                        umma::store_row32(z_hi, row, q, zh, HALF_BLK);
                        umma::store_row32(z_lo, row, q, zl, HALF_BLK);
                    }
                    {
                        uint32_t xh[8], xl[8];
                        if (fin) {
#pragma unroll
                            for (int t = 0; t < 8; ++t) { xh[t] = __ldg(fin + t * TC_TP); xl[t] = __ldg(fin + (8 + t) * TC_TP); }
                        } else {
                            float e[16];
                            tb_e_slots(x, q, e);
#pragma unroll
                            for (int t = 0; t < 8; ++t) umma::split2(e[2 * t], e[2 * t + 1], xh[t], xl[t]);
                        }
                        umma::store_row16(x_hi, row, q, xh);
                        umma::store_row16(x_lo, row, q, xl);
                    }
                }
                tb_round(c, [&]() { tb_issue_wgrad(c, 64, half == 0); });
            }
            if (q < 2) {
#pragma unroll
                for (int c0 = 0; c0 < 32; c0 += 8) {
                    uint32_t r[8];
                    umma::tmem_ld8(c.lane_base + (uint32_t)(TB_D + 32 * q + c0), r);
                    umma::wait_ld();
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int ei = tc_e_slot_to_index(32 * q + c0 + k);
                        if (ei >= 0) atomicAdd(&gpart[OFF_W1 + ei * D_H + p], __uint_as_float(r[k]));
                    }
                }
            }
        }
        TB_MARK(11);
        if (WANT_DX) {
            // ---- dgrad of layer 1: dE (64 slots) = dZ1 W1, plus the colour head's direct use of e ----
            tb_store_op32(c, TB_OP2_HI, TB_OP2_LO, q, v);                    // dZ1
            umma::wait_st();
            __syncthreads();                                                 // wgrad MMAs done (waited) -> region 1 reusable
            if (PARAMS) tb_copy(c.r1, f.tc_img + IMG_W1_HI, 2 * IMG_BLOCK);
            tb_round(c, [&]() { tb_issue_dgrad(c, c.r1, c.r1 + IMG_BLOCK, 64); });
            uint32_t r[16];
            umma::tmem_ld16(c.lane_base + TB_D + 16 * q, r);
            umma::wait_ld();
            const float* wre = c.fw + F_WR_E + 16 * q;
#pragma unroll
            for (int s = 0; s < 16; ++s) {
                const float de = __uint_as_float(r[s]) + fmaf(wre[128 + s], g[2], fmaf(wre[64 + s], g[1], wre[s] * g[0]));
                if (s < 12) {
                    const int j = q * 12 + s, d = j >> 4, k = (j & 15) >> 1, ph = j & 1;
                    const float cs = cos_reduced(freq_arg(x[d], k, ph)) * ldexpf(3.14159274101257324f, k) * de;
                    if (d == 0) dx[0] += cs; else if (d == 1) dx[1] += cs; else dx[2] += cs;
                } else if (q == 0 && s < 15) {
                    dx[s - 12] += de;
                }
            }
            float* DXP = c.part + 20 * TC_LD;
            DXP[(q * 3 + 0) * TC_LD + p] = dx[0]; DXP[(q * 3 + 1) * TC_LD + p] = dx[1]; DXP[(q * 3 + 2) * TC_LD + p] = dx[2];
            __syncthreads();
            if (q == 0 && valid) {
                float t[3], dp[3];
#pragma unroll
                for (int d = 0; d < 3; ++d)
                    t[d] = (DXP[d * TC_LD + p] + DXP[(3 + d) * TC_LD + p]) + (DXP[(6 + d) * TC_LD + p] + DXP[(9 + d) * TC_LD + p]);
                src.dx_to_dp(f, t, dp);
                d_pts[i * 3 + 0] = dp[0]; d_pts[i * 3 + 1] = dp[1]; d_pts[i * 3 + 2] = dp[2];
            }
        }
        TB_MARK(12);
        umma::fence_before_sync();
        __syncthreads();           // accumulator / operand reads of this tile are done before the next tile reuses them
    }

    if (!c.ok && err) atomicExch(err, 1);
    umma::fence_before_sync();
    __syncthreads();
    if ((tid >> 5) == 0) umma::tmem_dealloc<512>(c.tmem_base);
}

// Tensor-core (tcgen05) implementation of the fused field forward:
//   normalise -> hash-grid + frequency encode -> MLP_reg decoder, 128 points per tile.
//
// The three 128-wide decoder layers run as tcgen05.mma with
//   * the activations (A operand) in tensor memory, written by the thread that owns the point
//     (thread <-> TMEM lane <-> point),
//   * the weights (B operand) resident in shared memory for the whole persistent CTA,
//   * fp32 accumulators in tensor memory, read back with tcgen05.ld for bias / ReLU / heads.
// fp32 fidelity comes from a bf16 "x3" split: x = hi + lo on both operands and
// D = A_hi W_hi + A_hi W_lo + A_lo W_hi (the lo*lo term is below 2^-16 relative), three MMAs per
// 16-wide K step.  The narrow heads (115->3 colour, 128->5 logits) stay on CUDA cores.
#pragma once
#include "grid_encode.cuh"
#include "umma.cuh"

constexpr int TC_TP = 128;               // points per tile = UMMA M
constexpr int TC_NT = 512;               // threads per CTA: 4 threads per point
constexpr int TC_LD = TC_TP + 4;         // row length of the small fp32 staging arrays
// Encoded-feature cache (optional): the packed bf16 hi/lo operand words a thread writes to tensor memory at
// encode time, kept in HBM so the backward need not redo the gathers.  Per tile [q 4][word 24][p 128] uint32:
// words 0..7 e_hi, 8..15 e_lo, 16..19 grid_hi, 20..23 grid_lo (384 B per point).
constexpr int FEAT_WORDS = 24;
constexpr int FEAT_TILE_WORDS = 4 * FEAT_WORDS * TC_TP;

// ---- weight image (built by mf_mlp_prepare, copied once per CTA into shared memory) -----------
// bf16 K-major blocks of 128 rows (output n) x 64 k, 128-byte swizzle, 16 KB each.
constexpr int IMG_BLOCK = 16384;
constexpr int IMG_W1_HI = 0;                          // pts_linear.0 : K' = 64 (permuted e, see tc_e_slot)
constexpr int IMG_W1_LO = IMG_W1_HI + IMG_BLOCK;
constexpr int IMG_W2_HI = IMG_W1_LO + IMG_BLOCK;      // pts_linear.2 : K = 128 (2 blocks)
constexpr int IMG_W2_LO = IMG_W2_HI + 2 * IMG_BLOCK;
constexpr int IMG_W3_HI = IMG_W2_LO + 2 * IMG_BLOCK;  // sdf_linear.0 : K = 96 (2 blocks, second half empty)
constexpr int IMG_W3_LO = IMG_W3_HI + 2 * IMG_BLOCK;
constexpr int IMG_F32 = IMG_W3_LO + 2 * IMG_BLOCK;    // = 163840: fp32 section
// fp32 section (float offsets)
constexpr int F_B1 = 0, F_B2 = 128, F_BS1 = 256;
constexpr int F_WR_EMB = 384;                         // rgb_linear.0 weight on rgb_emb: [3][64]
constexpr int F_WR_E = F_WR_EMB + 192;                // rgb_linear.0 weight on e, in slot order: [3][64]
constexpr int F_BR = F_WR_E + 192;                    // [4]
constexpr int F_WS2 = F_BR + 4;                       // sdf_linear.2 weight [5][128]
constexpr int F_BS2 = F_WS2 + 640;                    // [8]
constexpr int F_COUNT = F_BS2 + 8;
constexpr int IMG_BYTES = IMG_F32 + F_COUNT * 4;      // 169,552
static_assert(IMG_BYTES % 16 == 0, "image is copied with 16-byte loads");

// Layer-1 input slot order: thread q of a point owns slots [16q, 16q+16): 12 frequency features
// j = 12q + jj (original e index 3 + j), and for q = 0 the raw xyz in slots 12..14; the rest is zero.
__host__ __device__ inline int tc_e_slot_to_index(int slot) {       // -> index into e = [xyz, freq] or -1
    const int q = slot >> 4, jj = slot & 15;
    if (jj < 12) return 3 + 12 * q + jj;
    if (q == 0 && jj < 15) return jj - 12;
    return -1;
}

__host__ __device__ inline uint32_t sw128_offset(int row, int k) {   // byte offset inside a 16 KB block, k in [0,64)
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((k >> 3) ^ (row & 7)) & 7) << 4) + (k & 7) * 2);
}

// ---- tensor-memory column map (512 columns allocated) ------------------------------------------
constexpr int TM_D = 0;                  // 128 fp32 accumulator columns
constexpr int TM_A_HI = 128;             // activations, bf16 pairs: layer 1: 32 cols, layer 2: 64 cols, layer 3 (sdf_emb): 32 cols
constexpr int TM_A_LO = 192;
constexpr int TM_G_HI = 256;             // grid features (K = 32): 16 cols
constexpr int TM_G_LO = 272;
constexpr int TM_COLS = 512;

// ---- shared memory map ---------------------------------------------------------------------------
constexpr int TCS_IMG = 0;
constexpr int TCS_PART = ((IMG_BYTES + 127) / 128) * 128;            // fp32 partial sums: 40 rows x TC_LD
constexpr int TCS_OUT = TCS_PART + 40 * TC_LD * 4;                   // outputs: 10 rows x TC_LD
constexpr int TCS_BAR = TCS_OUT + 10 * TC_LD * 4;                    // mbarrier (8 B) + tmem pointer (4 B) + error flag
constexpr int TCS_BYTES = TCS_BAR + 32;
constexpr size_t SMEM_TC = TCS_BYTES + 1024;                         // + slack to align the image to 1024 B

struct TcCtx {
    uint8_t* img; float* part; float* out; uint64_t* bar; uint32_t* tmem_ptr;
    uint32_t tmem_base, lane_base, phase;
    const float* fw;                      // fp32 section of the image
};

__device__ __forceinline__ void tc_setup(TcCtx& c, uint8_t* smem_raw, const uint8_t* __restrict__ img_g) {
    uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    c.img = base + TCS_IMG; c.part = (float*)(base + TCS_PART); c.out = (float*)(base + TCS_OUT);
    c.bar = (uint64_t*)(base + TCS_BAR); c.tmem_ptr = (uint32_t*)(base + TCS_BAR + 8);
    c.fw = (const float*)(c.img + IMG_F32);
    const int tid = threadIdx.x;
    for (int i = tid; i < IMG_BYTES / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(c.img)[i] = __ldg(reinterpret_cast<const uint4*>(img_g) + i);
    umma::fence_proxy_async();            // generic-proxy stores -> visible to the tensor core (async proxy)
    if ((tid >> 5) == 0) umma::tmem_alloc<TM_COLS>(c.tmem_ptr);
    if (tid == 0) { umma::mbar_init(c.bar, 1); umma::fence_barrier_init(); }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    c.tmem_base = *c.tmem_ptr;
    c.lane_base = c.tmem_base + ((uint32_t)(((tid >> 5) & 3) * 32) << 16);
    c.phase = 0;
}

__device__ __forceinline__ void tc_teardown(TcCtx& c) {
    umma::fence_before_sync();
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) umma::tmem_dealloc<TM_COLS>(c.tmem_base);
}

// One decoder layer on the tensor core: D = A W^T with the bf16x3 split.  KS 16-wide K steps; the A
// operand of step ks lives at TMEM column a_hi_col(ks) / a_lo_col(ks) (8 columns per step).
template <class ColFn>
__device__ __forceinline__ void tc_layer_mma(const TcCtx& c, int img_hi, int img_lo, int KS, ColFn a_col) {
    constexpr uint32_t idesc = umma::idesc_bf16(128, 128, 0, 0);
    const uint32_t d = c.tmem_base + TM_D;
    const uint32_t w_hi = umma::smem_u32(c.img + img_hi), w_lo = umma::smem_u32(c.img + img_lo);
    uint32_t acc = 0;
#pragma unroll 1
    for (int pass = 0; pass < 3; ++pass) {          // hi*hi, hi*lo, lo*hi
        const bool a_is_lo = pass == 2, w_is_lo = pass == 1;
#pragma unroll 1
        for (int ks = 0; ks < KS; ++ks) {
            const uint32_t a = c.tmem_base + (uint32_t)a_col(ks, a_is_lo);
            const uint32_t wb = (w_is_lo ? w_lo : w_hi) + (uint32_t)((ks >> 2) * IMG_BLOCK + (ks & 3) * 32);
            umma::mma_ts(d, a, umma::smem_desc_sw128(wb, 16, 1024), idesc, acc);
            acc = 1;
        }
    }
}

// All threads: make this thread's TMEM stores visible, rendezvous, thread 0 issues the layer's MMAs and
// commits; everybody waits for completion.  Returns false if the completion wait timed out.
template <class ColFn>
__device__ __forceinline__ bool tc_run_layer(TcCtx& c, int img_hi, int img_lo, int KS, ColFn a_col) {
    umma::wait_st();
    umma::fence_before_sync();
    __syncthreads();
    if (threadIdx.x == 0) {
        umma::fence_after_sync();
        tc_layer_mma(c, img_hi, img_lo, KS, a_col);
        umma::commit(c.bar);
    }
    const bool ok = umma::mbar_wait(c.bar, c.phase);
    c.phase ^= 1;
    umma::fence_after_sync();
    return ok;
}

// Thread (p, q): accumulator columns [32q, 32q+32) of point p as fp32.
__device__ __forceinline__ void tc_load_acc(const TcCtx& c, int q, float (&v)[32]) {
    uint32_t r[32];
    umma::tmem_ld32(c.lane_base + TM_D + 32 * q, r);
    umma::wait_ld();
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// Thread (p, q): store 32 activations (features [32q, 32q+32)) as bf16 hi/lo pairs: 16 + 16 columns.
__device__ __forceinline__ void tc_store_act32(const TcCtx& c, int q, const float (&v)[32]) {
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) umma::split2(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
    umma::tmem_st16(c.lane_base + TM_A_HI + 16 * q, hi);
    umma::tmem_st16(c.lane_base + TM_A_LO + 16 * q, lo);
}

// ---------------------------------------------------------------------------------------------
// Encode phase: thread (p = tid & 127, q = tid >> 7) computes 4 grid levels and 12 frequency features of
// point p, writes them straight into tensor memory as layer-1 / layer-3 operands, and accumulates the
// part of the colour head that reads e directly (model/decoder.py:61).
// ---------------------------------------------------------------------------------------------
template <class Src, bool SDF_ONLY>
__device__ __forceinline__ void tc_encode_tile(const TcCtx& c, const FieldDev& f, const Src& src, int64_t tile, int64_t N) {
    const int tid = threadIdx.x, p = tid & (TC_TP - 1), q = tid >> 7;
    const int64_t i = tile * TC_TP + p;
    const bool valid = i < N;
    float x[3] = {0.f, 0.f, 0.f};
    if (valid) src.point(i, f, x);
    float e[16];
#pragma unroll
    for (int jj = 0; jj < 12; ++jj) {
        const int j = q * 12 + jj, d = j >> 4, k = (j & 15) >> 1, s = j & 1;
        e[jj] = sin_reduced(freq_arg(x[d], k, s));
    }
    e[12] = q == 0 ? x[0] : 0.f; e[13] = q == 0 ? x[1] : 0.f; e[14] = q == 0 ? x[2] : 0.f; e[15] = 0.f;
    {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) umma::split2(e[2 * t], e[2 * t + 1], hi[t], lo[t]);
        umma::tmem_st8(c.lane_base + TM_A_HI + 8 * q, hi);
        umma::tmem_st8(c.lane_base + TM_A_LO + 8 * q, lo);
        if (f.feat) {
            uint32_t* fo = f.feat + (size_t)tile * FEAT_TILE_WORDS + (size_t)q * FEAT_WORDS * TC_TP + p;
#pragma unroll
            for (int t = 0; t < 8; ++t) { fo[t * TC_TP] = hi[t]; fo[(8 + t) * TC_TP] = lo[t]; }
        }
    }
    if (!SDF_ONLY) {
        const float* wre = c.fw + F_WR_E + 16 * q;
        float r[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int t = 0; t < 16; ++t)
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) r[ch] = fmaf(wre[ch * 64 + t], e[t], r[ch]);
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) c.part[(q * 3 + ch) * TC_LD + p] = r[ch];        // rows 0..11
    }
    const float2* grid2 = reinterpret_cast<const float2*>(f.grid);
    float g[8];
#pragma unroll
    for (int ll = 0; ll < 4; ++ll) {
        float2 v = make_float2(0.f, 0.f);
        if (valid) v = grid_level_fwd(x, grid2, level_info(f, q * 4 + ll), nullptr);
        g[2 * ll] = v.x; g[2 * ll + 1] = v.y;
    }
    {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) umma::split2(g[2 * t], g[2 * t + 1], hi[t], lo[t]);
        umma::tmem_st4(c.lane_base + TM_G_HI + 4 * q, hi);
        umma::tmem_st4(c.lane_base + TM_G_LO + 4 * q, lo);
        if (f.feat) {
            uint32_t* fo = f.feat + (size_t)tile * FEAT_TILE_WORDS + (size_t)q * FEAT_WORDS * TC_TP + p;
#pragma unroll
            for (int t = 0; t < 4; ++t) { fo[(16 + t) * TC_TP] = hi[t]; fo[(20 + t) * TC_TP] = lo[t]; }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Decoder forward of one tile.  On exit c.out holds OUT[c][p] (10 x 128).  Returns false on MMA timeout.
// ---------------------------------------------------------------------------------------------
template <bool SDF_ONLY>
__device__ __forceinline__ bool tc_mlp_forward_tile(TcCtx& c) {
    const int tid = threadIdx.x, p = tid & (TC_TP - 1), q = tid >> 7;
    bool ok = true;
    float v[32];
    // ---- pts_linear.0 + ReLU: K' = 64 (4 steps) ----
    ok &= tc_run_layer(c, IMG_W1_HI, IMG_W1_LO, 4, [](int ks, bool lo) { return (lo ? TM_A_LO : TM_A_HI) + 8 * ks; });
    tc_load_acc(c, q, v);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i] + c.fw[F_B1 + 32 * q + i], 0.f);
    tc_store_act32(c, q, v);
    // ---- pts_linear.2: K = 128 (8 steps), no activation ----
    ok &= tc_run_layer(c, IMG_W2_HI, IMG_W2_LO, 8, [](int ks, bool lo) { return (lo ? TM_A_LO : TM_A_HI) + 8 * ks; });
    tc_load_acc(c, q, v);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] += c.fw[F_B2 + 32 * q + i];
    if (q < 2) {
        tc_store_act32(c, q, v);                         // sdf_emb = h[:64] -> layer-3 operand, K steps 0..3
    } else if (!SDF_ONLY) {                              // rgb_emb = h[64:] -> colour head partial sums
        const float* wr = c.fw + F_WR_EMB + 32 * (q - 2);
        float r[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 32; ++i)
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) r[ch] = fmaf(wr[ch * 64 + i], v[i], r[ch]);
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) c.part[(12 + (q - 2) * 3 + ch) * TC_LD + p] = r[ch];   // rows 12..17
    }
    // ---- sdf_linear.0 + ReLU on [sdf_emb (64), grid (32)]: K = 96 (6 steps) ----
    ok &= tc_run_layer(c, IMG_W3_HI, IMG_W3_LO, 6, [](int ks, bool lo) {
        return ks < 4 ? (lo ? TM_A_LO : TM_A_HI) + 8 * ks : (lo ? TM_G_LO : TM_G_HI) + 8 * (ks - 4);
    });
    tc_load_acc(c, q, v);
    {
        float s[N_CLASS] = {0.f, 0.f, 0.f, 0.f, 0.f};
        const float* ws2 = c.fw + F_WS2 + 32 * q;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const float h = fmaxf(v[i] + c.fw[F_BS1 + 32 * q + i], 0.f);
#pragma unroll
            for (int ch = 0; ch < N_CLASS; ++ch) s[ch] = fmaf(ws2[ch * 128 + i], h, s[ch]);
        }
#pragma unroll
        for (int ch = 0; ch < N_CLASS; ++ch) c.part[(18 + q * 5 + ch) * TC_LD + p] = s[ch];    // rows 18..37
    }
    umma::fence_before_sync();               // accumulator reads done before the next tile's MMAs overwrite D
    __syncthreads();
    if (q == 0) {
        const float* P = c.part;
        float zl[N_CLASS], rgb[3];
#pragma unroll
        for (int ch = 0; ch < N_CLASS; ++ch)
            zl[ch] = c.fw[F_BS2 + ch] + ((P[(18 + ch) * TC_LD + p] + P[(23 + ch) * TC_LD + p]) + (P[(28 + ch) * TC_LD + p] + P[(33 + ch) * TC_LD + p]));
#pragma unroll
        for (int ch = 0; ch < 3; ++ch)
            rgb[ch] = SDF_ONLY ? 0.f
                               : c.fw[F_BR + ch] + (((P[ch * TC_LD + p] + P[(3 + ch) * TC_LD + p]) + (P[(6 + ch) * TC_LD + p] + P[(9 + ch) * TC_LD + p])) +
                                                    (P[(12 + ch) * TC_LD + p] + P[(15 + ch) * TC_LD + p]));
        float mx = zl[0];
#pragma unroll
        for (int ch = 1; ch < N_CLASS; ++ch) mx = fmaxf(mx, zl[ch]);
        float pr[N_CLASS], se = 0.f;
#pragma unroll
        for (int ch = 0; ch < N_CLASS; ++ch) { pr[ch] = expf(zl[ch] - mx); se += pr[ch]; }
        float ent = 0.f, ex = 0.f;
#pragma unroll
        for (int ch = 0; ch < N_CLASS; ++ch) {
            pr[ch] = pr[ch] / se;
            ent += pr[ch] * log2f(pr[ch] + 1e-5f);
            ex += pr[ch] * (float)ch;
        }
        float* O = c.out;
        O[0 * TC_LD + p] = rgb[0]; O[1 * TC_LD + p] = rgb[1]; O[2 * TC_LD + p] = rgb[2];
        O[3 * TC_LD + p] = (ex / 4.0f - 0.5f) * 2.0f;
        O[4 * TC_LD + p] = -1.0f * ent;
#pragma unroll
        for (int ch = 0; ch < N_CLASS; ++ch) O[(5 + ch) * TC_LD + p] = pr[ch];
    }
    return ok;
}

template <class Src, class Epi, bool SDF_ONLY>
__global__ void __launch_bounds__(TC_NT, 1) field_fwd_tc_kernel(FieldDev f, Src src, Epi epi, int64_t N,
                                                                const unsigned int* __restrict__ n_dev,
                                                                const uint8_t* __restrict__ img, int* __restrict__ err,
                                                                long long* __restrict__ prof) {
    extern __shared__ uint8_t smem_raw[];
#define TC_MARK(k) do { if (prof && threadIdx.x == 0 && blockIdx.x == 0 && tile == (int64_t)gridDim.x) prof[32 + (k)] = clock64(); } while (0)
    if (n_dev) N = (int64_t)*n_dev;
    TcCtx c;
    tc_setup(c, smem_raw, img);
    bool ok = true;
    const int64_t n_tiles = (N + TC_TP - 1) / TC_TP;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        TC_MARK(0);
        tc_encode_tile<Src, SDF_ONLY>(c, f, src, tile, N);
        TC_MARK(1);
        ok &= tc_mlp_forward_tile<SDF_ONLY>(c);
        TC_MARK(2);
        __syncthreads();
        epi.store(c.out, TC_LD, TC_TP, tile, N, threadIdx.x, TC_NT);
        __syncthreads();
        TC_MARK(3);
    }
    if (!ok && err) atomicExch(err, 1);
    tc_teardown(c);
}

// Tensor-core forward, dual-pipeline version: one persistent CTA per SM holds the weight image once and runs
// TWO independent 256-thread pipelines (warps 0-7 and 8-15), each working on its own 128-point tile with its
// own half of tensor memory (256 columns), its own named barrier and its own mbarrier.  The pipelines drift
// apart by construction, so one pipeline's gather phase (LSU bound) overlaps the other's MMA / epilogue phases
// (tensor / ALU bound) -- the effect of two CTAs per SM without a second copy of the 160 KB weight image.
//
// Per pipeline: thread (p, h), p = point / TMEM lane, h in {0,1} owns half of the columns.
//   TMEM (relative to the pipeline's base column): D [0,128) | A_hi [128,192) | A_lo [192,256).
//   The grid features stay in registers (16 packed words) until the layer-2 MMAs are done, then go to the
//   upper part of the operand region for layer 3.
#pragma once
#include "field_tc.cuh"

constexpr int T2_GT = 256;                       // threads per pipeline
constexpr int T2_D = 0, T2_A_HI = 128, T2_A_LO = 192, T2_COLS = 256;
constexpr int T2_PART_ROWS = 20;                 // 6 colour-e partials, 3 colour-emb partials, 10 logit partials (+1)
constexpr int T2S_PART = ((IMG_BYTES + 127) / 128) * 128;
constexpr int T2S_GROUP_BYTES = (T2_PART_ROWS + MF_RAW_DIM) * TC_LD * 4;
constexpr int T2S_BAR = T2S_PART + 2 * T2S_GROUP_BYTES;
constexpr size_t SMEM_TC2 = T2S_BAR + 64 + 1024;
static_assert(SMEM_TC2 <= 227 * 1024, "shared memory budget");

struct T2Ctx {
    uint8_t* img; const float* fw; float* part; float* out; uint64_t* bar;
    uint32_t tmem, lane_base, phase;
    int g, tid, p, h;
    bool ok;
};

__device__ __forceinline__ void t2_group_sync(int g) { asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "r"(T2_GT) : "memory"); }

template <int N_OUT = 128, class ColFn>
__device__ __forceinline__ void t2_run_layer(T2Ctx& c, int img_hi, int img_lo, int KS, ColFn a_col) {
    umma::wait_st();
    umma::fence_before_sync();
    t2_group_sync(c.g);
    if (c.tid == 0) {
        umma::fence_after_sync();
        constexpr uint32_t idesc = umma::idesc_bf16(128, N_OUT, 0, 0);     // N_OUT < 128: only the first N_OUT output features
        const uint32_t w_hi = umma::smem_u32(c.img + img_hi), w_lo = umma::smem_u32(c.img + img_lo);
        uint32_t acc = 0;
#pragma unroll 1
        for (int pass = 0; pass < 3; ++pass) {
#pragma unroll 1
            for (int ks = 0; ks < KS; ++ks) {
                const uint32_t wb = (pass == 1 ? w_lo : w_hi) + (uint32_t)((ks >> 2) * IMG_BLOCK + (ks & 3) * 32);
                umma::mma_ts(c.tmem + T2_D, c.tmem + (uint32_t)a_col(ks, pass == 2), umma::smem_desc_sw128(wb, 16, 1024), idesc, acc);
                acc = 1;
            }
        }
        umma::commit(c.bar);
    }
    c.ok &= umma::mbar_wait(c.bar, c.phase);
    c.phase ^= 1;
    umma::fence_after_sync();
}

__device__ __forceinline__ void t2_load32(const T2Ctx& c, int col, float (&v)[32]) {
    uint32_t r[32];
    umma::tmem_ld32(c.lane_base + (uint32_t)col, r);
    umma::wait_ld();
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// 32 features starting at feature f0 (multiple of 32) -> operand region
__device__ __forceinline__ void t2_store32(const T2Ctx& c, int f0, const float (&v)[32]) {
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) umma::split2(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
    umma::tmem_st16(c.lane_base + (uint32_t)(T2_A_HI + f0 / 2), hi);
    umma::tmem_st16(c.lane_base + (uint32_t)(T2_A_LO + f0 / 2), lo);
}

template <class Src, class Epi, bool SDF_ONLY>
__global__ void __launch_bounds__(2 * T2_GT, 1) field_fwd_tc2_kernel(FieldDev f, Src src, Epi epi, int64_t N,
                                                                     const unsigned int* __restrict__ n_dev,
                                                                     const uint8_t* __restrict__ img, int* __restrict__ err) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint32_t tmem_ptr_s;
    if (n_dev) N = (int64_t)*n_dev;
    uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    T2Ctx c;
    c.g = threadIdx.x >> 8; c.tid = threadIdx.x & (T2_GT - 1); c.p = c.tid & 127; c.h = c.tid >> 7;
    c.img = base; c.fw = (const float*)(base + IMG_F32);
    c.part = (float*)(base + T2S_PART + c.g * T2S_GROUP_BYTES);
    c.out = c.part + T2_PART_ROWS * TC_LD;
    c.bar = (uint64_t*)(base + T2S_BAR) + c.g;
    for (int i = threadIdx.x; i < IMG_BYTES / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(c.img)[i] = __ldg(reinterpret_cast<const uint4*>(img) + i);
    umma::fence_proxy_async();
    if ((threadIdx.x >> 5) == 0) umma::tmem_alloc<512>(&tmem_ptr_s);
    if (c.tid == 0) { umma::mbar_init(c.bar, 1); umma::fence_barrier_init(); }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    c.tmem = tmem_ptr_s + (uint32_t)(c.g * T2_COLS);
    c.lane_base = c.tmem + ((uint32_t)(((c.tid >> 5) & 3) * 32) << 16);
    c.phase = 0; c.ok = true;
    const int p = c.p, h = c.h;
    const float2* grid2 = reinterpret_cast<const float2*>(f.grid);

    const int64_t n_tiles = (N + TC_TP - 1) / TC_TP;
    // Stagger the two pipelines by one encode phase so that they stay out of phase (identical pipelines started
    // together would gather at the same time and run their MMAs at the same time): pipeline 1 waits once, on
    // named barrier 3, until pipeline 0 has finished its first encode.
    bool staggered = false;
    if (c.g == 1) asm volatile("bar.sync 3, 512;" ::: "memory");
    for (int64_t tile = 2 * (int64_t)blockIdx.x + c.g; tile < n_tiles; tile += 2 * (int64_t)gridDim.x) {
        const int64_t i = tile * TC_TP + p;
        const bool valid = i < N;
        float x[3] = {0.f, 0.f, 0.f};
        if (valid) src.point(i, f, x);
        // ---- encode: this thread owns slot groups 2h, 2h+1 (32 slots) and grid levels [8h, 8h+8) ----
#pragma unroll
        for (int sg = 0; sg < 2; ++sg) {
            const int qq = 2 * h + sg;
            float e[16];
#pragma unroll
            for (int jj = 0; jj < 12; ++jj) {
                const int j = qq * 12 + jj, d = j >> 4, k = (j & 15) >> 1, s = j & 1;
                e[jj] = sin_reduced(freq_arg(x[d], k, s));
            }
            e[12] = qq == 0 ? x[0] : 0.f; e[13] = qq == 0 ? x[1] : 0.f; e[14] = qq == 0 ? x[2] : 0.f; e[15] = 0.f;
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) umma::split2(e[2 * t], e[2 * t + 1], hi[t], lo[t]);
            umma::tmem_st8(c.lane_base + T2_A_HI + 8 * qq, hi);
            umma::tmem_st8(c.lane_base + T2_A_LO + 8 * qq, lo);
            if (f.feat) {
                uint32_t* fo = f.feat + (size_t)tile * FEAT_TILE_WORDS + (size_t)qq * FEAT_WORDS * TC_TP + p;
#pragma unroll
                for (int t = 0; t < 8; ++t) { fo[t * TC_TP] = hi[t]; fo[(8 + t) * TC_TP] = lo[t]; }
            }
            if (!SDF_ONLY) {
                const float* wre = c.fw + F_WR_E + 16 * qq;
                float r[3] = {0.f, 0.f, 0.f};
#pragma unroll
                for (int t = 0; t < 16; ++t)
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) r[ch] = fmaf(wre[ch * 64 + t], e[t], r[ch]);
                if (sg == 0) {
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) c.part[(h * 3 + ch) * TC_LD + p] = r[ch];
                } else {
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) c.part[(h * 3 + ch) * TC_LD + p] += r[ch];
                }
            }
        }
        uint32_t ghi[8], glo[8];                         // 16 grid features, packed; live until layer 3
#pragma unroll
        for (int sg = 0; sg < 2; ++sg) {
            float gf[8];
#pragma unroll
            for (int ll = 0; ll < 4; ++ll) {
                float2 v2 = make_float2(0.f, 0.f);
                if (valid) v2 = grid_level_fwd(x, grid2, level_info(f, 8 * h + 4 * sg + ll), nullptr);
                gf[2 * ll] = v2.x; gf[2 * ll + 1] = v2.y;
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) umma::split2(gf[2 * t], gf[2 * t + 1], ghi[4 * sg + t], glo[4 * sg + t]);
            if (f.feat) {
                uint32_t* fo = f.feat + (size_t)tile * FEAT_TILE_WORDS + (size_t)(2 * h + sg) * FEAT_WORDS * TC_TP + p;
#pragma unroll
                for (int t = 0; t < 4; ++t) { fo[(16 + t) * TC_TP] = ghi[4 * sg + t]; fo[(20 + t) * TC_TP] = glo[4 * sg + t]; }
            }
        }
        if (c.g == 0 && !staggered) { asm volatile("bar.arrive 3, 512;" ::: "memory"); staggered = true; }
        float v[32];
        // ---- pts_linear.0 + ReLU ----
        t2_run_layer(c, IMG_W1_HI, IMG_W1_LO, 4, [](int ks, bool lo) { return (lo ? T2_A_LO : T2_A_HI) + 8 * ks; });
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            const int f0 = 64 * h + 32 * half;
            t2_load32(c, T2_D + f0, v);
            {
                const float4* b4 = reinterpret_cast<const float4*>(c.fw + F_B1 + f0);
#pragma unroll
                for (int k4 = 0; k4 < 8; ++k4) {
                    const float4 b = b4[k4];
                    v[4 * k4] = fmaxf(v[4 * k4] + b.x, 0.f); v[4 * k4 + 1] = fmaxf(v[4 * k4 + 1] + b.y, 0.f);
                    v[4 * k4 + 2] = fmaxf(v[4 * k4 + 2] + b.z, 0.f); v[4 * k4 + 3] = fmaxf(v[4 * k4 + 3] + b.w, 0.f);
                }
            }
            t2_store32(c, f0, v);
        }
        // ---- pts_linear.2 (SDF only: just the 64 sdf_emb outputs, 32 per thread) ----
        if (SDF_ONLY) {
            t2_run_layer<64>(c, IMG_W2_HI, IMG_W2_LO, 8, [](int ks, bool lo) { return (lo ? T2_A_LO : T2_A_HI) + 8 * ks; });
            const int f0 = 32 * h;
            t2_load32(c, T2_D + f0, v);
            const float4* b4 = reinterpret_cast<const float4*>(c.fw + F_B2 + f0);
#pragma unroll
            for (int k4 = 0; k4 < 8; ++k4) {
                const float4 b = b4[k4];
                v[4 * k4] += b.x; v[4 * k4 + 1] += b.y; v[4 * k4 + 2] += b.z; v[4 * k4 + 3] += b.w;
            }
            t2_store32(c, f0, v);
            umma::tmem_st8(c.lane_base + T2_A_HI + 32 + 8 * h, ghi);
            umma::tmem_st8(c.lane_base + T2_A_LO + 32 + 8 * h, glo);
        } else {
        t2_run_layer(c, IMG_W2_HI, IMG_W2_LO, 8, [](int ks, bool lo) { return (lo ? T2_A_LO : T2_A_HI) + 8 * ks; });
        {
            float r[3] = {0.f, 0.f, 0.f};
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                const int f0 = 64 * h + 32 * half;
                t2_load32(c, T2_D + f0, v);
                {
                    const float4* b4 = reinterpret_cast<const float4*>(c.fw + F_B2 + f0);
#pragma unroll
                    for (int k4 = 0; k4 < 8; ++k4) {
                        const float4 b = b4[k4];
                        v[4 * k4] += b.x; v[4 * k4 + 1] += b.y; v[4 * k4 + 2] += b.z; v[4 * k4 + 3] += b.w;
                    }
                }
                if (h == 0) {
                    t2_store32(c, f0, v);                    // sdf_emb -> features [0,64) of the layer-3 operand
                } else if (!SDF_ONLY) {
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) {
                        const float4* w4 = reinterpret_cast<const float4*>(c.fw + F_WR_EMB + ch * 64 + 32 * half);
#pragma unroll
                        for (int k4 = 0; k4 < 8; ++k4) {
                            const float4 w = w4[k4];
                            r[ch] = fmaf(w.x, v[4 * k4], r[ch]); r[ch] = fmaf(w.y, v[4 * k4 + 1], r[ch]);
                            r[ch] = fmaf(w.z, v[4 * k4 + 2], r[ch]); r[ch] = fmaf(w.w, v[4 * k4 + 3], r[ch]);
                        }
                    }
                }
            }
            if (h == 1 && !SDF_ONLY) {
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) c.part[(6 + ch) * TC_LD + p] = r[ch];
            }
            // grid features -> features [64, 96) of the layer-3 operand (operand columns 32 + 8h ..)
            umma::tmem_st8(c.lane_base + T2_A_HI + 32 + 8 * h, ghi);
            umma::tmem_st8(c.lane_base + T2_A_LO + 32 + 8 * h, glo);
        }
        }
        // ---- sdf_linear.0 + ReLU, logits ----
        t2_run_layer(c, IMG_W3_HI, IMG_W3_LO, 6, [](int ks, bool lo) { return (lo ? T2_A_LO : T2_A_HI) + 8 * ks; });
        {
            float s[N_CLASS] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                const int f0 = 64 * h + 32 * half;
                t2_load32(c, T2_D + f0, v);
                {
                    const float4* b4 = reinterpret_cast<const float4*>(c.fw + F_BS1 + f0);
#pragma unroll
                    for (int k4 = 0; k4 < 8; ++k4) {
                        const float4 b = b4[k4];
                        v[4 * k4] = fmaxf(v[4 * k4] + b.x, 0.f); v[4 * k4 + 1] = fmaxf(v[4 * k4 + 1] + b.y, 0.f);
                        v[4 * k4 + 2] = fmaxf(v[4 * k4 + 2] + b.z, 0.f); v[4 * k4 + 3] = fmaxf(v[4 * k4 + 3] + b.w, 0.f);
                    }
                }
#pragma unroll
                for (int ch = 0; ch < N_CLASS; ++ch) {
                    const float4* w4 = reinterpret_cast<const float4*>(c.fw + F_WS2 + ch * 128 + f0);
#pragma unroll
                    for (int k4 = 0; k4 < 8; ++k4) {
                        const float4 w = w4[k4];
                        s[ch] = fmaf(w.x, v[4 * k4], s[ch]); s[ch] = fmaf(w.y, v[4 * k4 + 1], s[ch]);
                        s[ch] = fmaf(w.z, v[4 * k4 + 2], s[ch]); s[ch] = fmaf(w.w, v[4 * k4 + 3], s[ch]);
                    }
                }
            }
#pragma unroll
            for (int ch = 0; ch < N_CLASS; ++ch) c.part[(9 + h * 5 + ch) * TC_LD + p] = s[ch];
        }
        umma::fence_before_sync();
        t2_group_sync(c.g);
        if (h == 0) {
            const float* P = c.part;
            float zl[N_CLASS], rgb[3];
#pragma unroll
            for (int ch = 0; ch < N_CLASS; ++ch) zl[ch] = c.fw[F_BS2 + ch] + (P[(9 + ch) * TC_LD + p] + P[(14 + ch) * TC_LD + p]);
#pragma unroll
            for (int ch = 0; ch < 3; ++ch)
                rgb[ch] = SDF_ONLY ? 0.f : c.fw[F_BR + ch] + ((P[ch * TC_LD + p] + P[(3 + ch) * TC_LD + p]) + P[(6 + ch) * TC_LD + p]);
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
        t2_group_sync(c.g);
        epi.store(c.out, TC_LD, TC_TP, tile, N, c.tid, T2_GT);
        t2_group_sync(c.g);
    }
    if (c.g == 0 && !staggered) asm volatile("bar.arrive 3, 512;" ::: "memory");     // pipeline 0 had no tile at all
    if (!c.ok && err) atomicExch(err, 1);
    umma::fence_before_sync();
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) umma::tmem_dealloc<512>(tmem_ptr_s);
}

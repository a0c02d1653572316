// Tensor-core (tcgen05) backward of the fused field evaluation, four roles.  NOT the default: correct (tests/test_gpu_tensorcore.py)
// but measured slower than the three-role kernel of field_tc_bwd2.cuh (0.293 vs 0.267 ms at C1, scripts/prof_bwd3.py); kept behind
// mf_set_bwd_impl(3) as the record of the experiment -- see the note at the end of this comment.  One persistent CTA per SM,
// 20 warps, 128 points per tile:
//
//   warps 12-19 CHAIN    the serial dependency chain of a tile and nothing else: operand words from the feature cache, forward
//                        layers 1-3, heads (softmax forward / backward), dZ3, dgrad3, dH, dgrad2, dZ1.  Two threads per point.
//                        Whatever a wgrad product needs is left where it already is -- dZ3 / dZ1 in region R1 and x3 / dH in
//                        region R2 of tensor memory, H1 / H3 / rgb_emb / U parked in the CTA's scratch -- and announced with
//                        one mbarrier arrival (z3, zh, z1).
//   warps 4-11  COPY     the five wgrad products of the tile (see field_tc_bwd2.cuh for the products): per 32-point quarter the
//                        two warps of the quarter read the operands (tcgen05.ld from R1 / R2, L2 loads from the scratch and the
//                        feature cache), write them as bf16 hi / lo [point][feature] rows into one of two staging buffers and
//                        issue the quarter's MMAs; after the last product of a layer all eight warps read the 128 x K result
//                        out of tensor memory into the CTA's partial gradient.  As soon as a tensor-memory operand is in
//                        registers the warp tells the chain (rd0, rd2, rd3; rd4 for the scratch words), which waits for that
//                        only where it overwrites the region: the chain of tile k+1 runs against the products of tile k.
//   warps 0-3   SCATTER  as before: d(grid features) from the dgrad3 accumulator -> 128 reductions per point.
// (warp order: the scheduler prefers the highest eligible warp id -> the latency-critical chain first.)
//
// Shared memory (231,968 of 232,448 bytes): W1 | W2 | W3 images resident (160 KB: the staging buffers are busy with tile k
// while the chain needs W1 for tile k+1, so the per-tile TMA copy of the three-role kernel is gone) | two 32 KB staging buffers
// | biases and the colour head (2.3 KB).  To make room the logit partial sums of the two threads of a point are exchanged
// through eight spare accumulator columns of tensor memory, and sdf_linear.2.weight (2.5 KB) is read from the weight image in
// global memory (uniform 16-byte loads, L1 resident).
//
// Tensor memory as in field_tc_bwd2.cuh: D | R1 | R2 | DW.
//
// Why it is slower (clock64 stamps, B200): moving the products off the chain does shorten the chain's own tile to ~45 k cycles,
// but the products do not get faster on dedicated warps -- ten staging rounds per tile through two 32 KB buffers (quarters 0 / 1,
// then 2 / 3, each round: loads, 32 16-byte stores per thread, proxy fence, pair barrier, issue, MMA completion before the buffer
// turns around) plus three read-outs take 75-84 k cycles -- and the hand-off regions R1 / R2 are single-buffered (tensor memory
// is full): the chain of tile k+1 cannot load its operands before COPY has taken dZ1 of tile k, which COPY reaches only after
// the products of layers 3 and 2.  The two roles therefore overlap for ~20 k cycles per tile instead of a whole tile, and the
// tile period is ~107 k cycles (three-role kernel: 70 k).  Decoupling needs a second copy of dZ3 / dH / dZ1 (tensor memory or
// shared memory that does not exist) or products that do not pass through shared memory.  Also measured here: 16-byte
// reductions for the read-out (row-major partials) and polling with nanosleep back-off changed nothing.
#pragma once
#include <type_traits>
#include "field_tc_bwd2.cuh"

namespace b3 {

constexpr int B3_NT = 640, CHAIN_NT = 256, COPY_NT = 256, SC_NT = 128;
constexpr uint32_t QBLK = b2::QBLK;
// ---- shared memory map ----
constexpr int S_ST = IMG_F32;                          // images at [0, IMG_F32) with their IMG_* offsets; then the staging buffers
constexpr int ST_A = b2::ST_A, ST_B = b2::ST_B, ST_BYTES = b2::ST_BYTES;
constexpr int S_F32 = S_ST + 2 * ST_BYTES;             // fp32: b1 | b2 | bs1 | Wr(rgb_emb) (= image floats [0, 576)) | bs2 (8)
constexpr int G_BS2 = F_WR_EMB + 192, G_COUNT = G_BS2 + 8;
static_assert(F_B1 == 0 && F_B2 == 128 && F_BS1 == 256 && F_WR_EMB == 384, "the first 576 floats of the image's fp32 section are copied as one block");
constexpr int S_BAR = S_F32 + G_COUNT * 4;
constexpr int S_BYTES = S_BAR + 256;
constexpr size_t SMEM = S_BYTES;
static_assert(SMEM <= 227 * 1024, "shared memory budget");
constexpr int SCR_U = b2::SCR_U;
// 640 threads start with 96 registers each (61,440: the CTA's pool; setmaxnreg only moves registers inside it)
constexpr int REGS_CHAIN = 112, REGS_COPY = 80;        // 256 x 112 + 256 x 80 + 128 x 96 = 61,440
using b2::T_D; using b2::T_R1_HI; using b2::T_R1_LO; using b2::T_R2_HI; using b2::T_R2_LO; using b2::T_DW; using b2::T_G_HI; using b2::T_G_LO;
using b2::SCR_H1; using b2::SCR_H3; using b2::SCR_RGB; using b2::SCR_CTA_WORDS; using b2::N_X3;
// ---- mbarriers ----
enum { B_MMA = 0, B_DG3, B_DCONS, B_DWRDY, B_Z3, B_ZH, B_Z1, B_RD0, B_RD2, B_RD3, B_RD4, B_FREE0, B_ISS0 = B_FREE0 + 4, B_COUNT = B_ISS0 + 4 };
static_assert(8 * B_COUNT + 8 <= 256, "barrier block");

__device__ __forceinline__ void chain_sync() { asm volatile("bar.sync 1, %0;" ::"n"(CHAIN_NT) : "memory"); }
__device__ __forceinline__ void copy_sync() { asm volatile("bar.sync 2, %0;" ::"n"(COPY_NT) : "memory"); }
// 32 features = half CC of block h of a staging tile (b2::put32 with the block offset added on the fly)
template <int TILE_OFF, int CC>
__device__ __forceinline__ void put32h(const uint32_t (&PA)[8], uint32_t hq, const uint32_t (&pk)[16]) {
#pragma unroll
    for (int c_ = 0; c_ < 4; ++c_) b2::sts128<TILE_OFF>(PA[4 * CC + c_] + hq, pk[4 * c_], pk[4 * c_ + 1], pk[4 * c_ + 2], pk[4 * c_ + 3]);
}
// one arrival per warp once every lane's preceding tensor-memory / shared-memory / global accesses are done
__device__ __forceinline__ void warp_arrive(uint64_t* bar, int lane) {
    umma::fence_before_sync();
    __syncwarp();
    if (lane == 0) umma::mbar_arrive(bar);
}

}  // namespace b3

// part: [gridDim.x][MF_MLP_PARAMS] per-CTA partial parameter gradients (zeroed here); scratch: [gridDim.x][SCR_CTA_WORDS] uint32;
// f.feat (the forward's feature cache) is required.
template <class Src>
__global__ void __launch_bounds__(b3::B3_NT, 1) field_bwd_tc3_kernel(FieldDev f, Src src, const float* __restrict__ d_raw,
                                                                  float* __restrict__ grad_grid, float* __restrict__ part,
                                                                  uint32_t* __restrict__ scratch, int64_t N_all, ActiveMap am,
                                                                  int* __restrict__ err, long long* __restrict__ prof) {
    using namespace b3;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* gpart = part + (size_t)blockIdx.x * MF_MLP_PARAMS;
    for (int i = tid; i < MF_MLP_PARAMS; i += B3_NT) gpart[i] = 0.f;

    b2::Ctx c;
    c.base = smem_raw;
    c.ok = (umma::smem_u32(smem_raw) & 1023u) == 0u;   // the swizzled tiles need a 1 KB aligned base (no slack left to align by hand)
    c.fw = (const float*)(c.base + S_F32); c.part = nullptr; c.bars = (uint64_t*)(c.base + S_BAR);
    uint32_t* tmem_ptr = (uint32_t*)(c.base + S_BAR + 8 * B_COUNT);
    for (int i = tid; i < IMG_F32 / 16; i += B3_NT)
        reinterpret_cast<uint4*>(c.base)[i] = __ldg(reinterpret_cast<const uint4*>(f.tc_img) + i);
    for (int i = tid; i < (G_BS2 * 4) / 16 + 2; i += B3_NT) {
        const int src16 = i < (G_BS2 * 4) / 16 ? i : (F_BS2 * 4) / 16 + (i - (G_BS2 * 4) / 16);
        reinterpret_cast<uint4*>(c.base + S_F32)[i] = __ldg(reinterpret_cast<const uint4*>(f.tc_img + IMG_F32) + src16);
    }
    umma::fence_proxy_async();
    if (warp == 0) umma::tmem_alloc<512>(tmem_ptr);
    if (tid == 0) {
        umma::mbar_init(c.bars + B_MMA, 1); umma::mbar_init(c.bars + B_DG3, 1); umma::mbar_init(c.bars + B_DCONS, SC_NT);
        umma::mbar_init(c.bars + B_DWRDY, 4);
        umma::mbar_init(c.bars + B_Z3, 1); umma::mbar_init(c.bars + B_ZH, 1); umma::mbar_init(c.bars + B_Z1, CHAIN_NT / 32);
        umma::mbar_init(c.bars + B_RD0, COPY_NT / 32); umma::mbar_init(c.bars + B_RD2, COPY_NT / 32);
        umma::mbar_init(c.bars + B_RD3, COPY_NT / 32); umma::mbar_init(c.bars + B_RD4, COPY_NT / 32);
        for (int q = 0; q < 4; ++q) { umma::mbar_init(c.bars + B_FREE0 + q, 1); umma::mbar_init(c.bars + B_ISS0 + q, 1); }
        umma::fence_barrier_init();
    }
    umma::fence_before_sync();
    __syncthreads();                                   // also orders the gpart zero-fill before COPY's reductions
    umma::fence_after_sync();
    c.tmem = *tmem_ptr;
    c.lane_base = c.tmem + ((uint32_t)((warp & 3) * 32) << 16);

    const int64_t N = am.n(N_all);                     // active points only (ascending point indices in am.idx)
    const int64_t n_tiles = (N + TC_TP - 1) / TC_TP;
    const int p = tid & (TC_TP - 1);
    uint32_t* scr_p = scratch + (size_t)blockIdx.x * SCR_CTA_WORDS + p;
#define B3_MARK(slot) do { if (prof && blockIdx.x == 0 && k == 1 && p == 0 && h == 0) prof[slot] = clock64(); } while (0)

    if (tid >= SC_NT + COPY_NT) {
        // =============================== CHAIN ===============================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_CHAIN));
        const int ctid = tid - (SC_NT + COPY_NT), h = ctid >> 7;
        uint32_t ph_mma = 0;
        uint32_t k = 0;
        auto round = [&](auto issue) {                 // publish TMEM stores, rendezvous, thread 0 issues + commits, all wait
            umma::wait_st();
            umma::fence_before_sync();
            chain_sync();
            if (ctid == 0) { umma::fence_after_sync(); issue(); umma::commit(c.bars + B_MMA); }
            c.ok &= umma::mbar_wait_spin(c.bars + B_MMA, ph_mma);
            ph_mma ^= 1;
            umma::fence_after_sync();
        };
        const float4* ws2g = reinterpret_cast<const float4*>(f.tc_img + IMG_F32) + F_WS2 / 4;      // sdf_linear.2.weight [5][128]
        static_assert(F_WS2 % 4 == 0, "16-byte loads of sdf_linear.2.weight");

        int64_t i_next = 0;
        { const int64_t s0 = (int64_t)blockIdx.x * TC_TP + p; if (s0 < N) i_next = am(s0); }
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++k) {
            const int64_t slot = tile * TC_TP + p;
            const bool valid = slot < N;
            const int64_t i = i_next;                                       // (loaded one tile ahead)
            { const int64_t s1 = slot + (int64_t)gridDim.x * TC_TP; i_next = s1 < N ? am(s1) : 0; }
            const uint32_t pk = (k - 1u) & 1u;                              // parity of the previous tile's hand-backs
            B3_MARK(0);
            float g[3], gs[7];                                 // d loss / d (rgb 3 | sdf, entropy, prob[5])
            {
                const float2* gr = reinterpret_cast<const float2*>(d_raw + i * MF_RAW_DIM);       // 40-byte rows: 8-byte aligned
                float2 t2[5];
#pragma unroll
                for (int j = 0; j < 5; ++j) t2[j] = valid ? __ldg(gr + j) : make_float2(0.f, 0.f);
                g[0] = t2[0].x; g[1] = t2[0].y; g[2] = t2[1].x;
                gs[0] = t2[1].y; gs[1] = t2[2].x; gs[2] = t2[2].y; gs[3] = t2[3].x; gs[4] = t2[3].y; gs[5] = t2[4].x; gs[6] = t2[4].y;
            }
            // operand words the forward kernel cached: [i / 128][slot group 4][word 24][i % 128]
            const uint32_t* fin = f.feat + (size_t)(i >> 7) * FEAT_TILE_WORDS + (i & (TC_TP - 1));
            // ---- layer-1 operand e (slot groups 2h, 2h+1) -> R1, grid features (levels 8h .. 8h+7) -> R2 ----
#pragma unroll
            for (int sg = 0; sg < 2; ++sg) {
                const int qq = 2 * h + sg;
                uint32_t hi[8], lo[8], gh[4], gl[4];
                const uint32_t* fq = fin + (size_t)qq * FEAT_WORDS * TC_TP;
#pragma unroll
                for (int t = 0; t < 8; ++t) { hi[t] = __ldg(fq + t * TC_TP); lo[t] = __ldg(fq + (8 + t) * TC_TP); }
#pragma unroll
                for (int t = 0; t < 4; ++t) { gh[t] = __ldg(fq + (16 + t) * TC_TP); gl[t] = __ldg(fq + (20 + t) * TC_TP); }
                if (sg == 0 && k > 0) {                        // COPY has taken dZ1 (R1) and dH (R2) of the previous tile
                    c.ok &= umma::mbar_wait_spin(c.bars + B_RD3, pk);
                    c.ok &= umma::mbar_wait_spin(c.bars + B_RD2, pk);
                    umma::fence_after_sync();
                }
                umma::tmem_st8(c.lane_base + T_R1_HI + 8 * qq, hi);
                umma::tmem_st8(c.lane_base + T_R1_LO + 8 * qq, lo);
                umma::tmem_st4(c.lane_base + T_G_HI + 4 * qq, gh);
                umma::tmem_st4(c.lane_base + T_G_LO + 4 * qq, gl);
            }
            B3_MARK(1);
            float v[32];
            uint32_t mask1a = 0, mask1b = 0, mask3a = 0, mask3b = 0;  // (scalars: the cc loops are rolled, an indexed array would live in local memory)
            // ---- forward layer 1 ----
            round([&]() { b2::issue_fwd<4>(c, IMG_W1_HI, IMG_W1_LO, [](int ks, bool lo) { return (lo ? T_R1_LO : T_R1_HI) + 8 * ks; }); });
#pragma unroll 1
            for (int cc = 0; cc < 2; ++cc) {
                const int f0 = 64 * h + 32 * cc;
                b2::ld32f(c.lane_base + T_D + f0, v);
                const float4* b4 = reinterpret_cast<const float4*>(c.fw + F_B1 + f0);
                uint32_t m = 0;
#pragma unroll
                for (int k4 = 0; k4 < 8; ++k4) {
                    const float4 b = b4[k4];
                    v[4 * k4] = fmaxf(v[4 * k4] + b.x, 0.f); v[4 * k4 + 1] = fmaxf(v[4 * k4 + 1] + b.y, 0.f);
                    v[4 * k4 + 2] = fmaxf(v[4 * k4 + 2] + b.z, 0.f); v[4 * k4 + 3] = fmaxf(v[4 * k4 + 3] + b.w, 0.f);
                }
#pragma unroll
                for (int kk = 0; kk < 32; ++kk) m |= (v[kk] > 0.f ? 1u : 0u) << kk;
                if (cc == 0) mask1a = m; else mask1b = m;
                uint32_t hi[16], lo[16];
                b2::split32(v, hi, lo);
                b2::st_op(c, T_R1_HI, T_R1_LO, f0, hi, lo);                  // H1: A operand of layer 2 ...
                b2::scr_store(scr_p, SCR_H1, 64, f0 / 2, hi, lo);            // ... and, parked, the X operand of its wgrad
            }
            B3_MARK(2);
            // ---- forward layer 2 ----
            round([&]() { b2::issue_fwd<8>(c, IMG_W2_HI, IMG_W2_LO, [](int ks, bool lo) { return (lo ? T_R1_LO : T_R1_HI) + 8 * ks; }); });
            if (k > 0) c.ok &= umma::mbar_wait_spin(c.bars + B_RD4, pk);    // COPY has taken rgb_emb and U of the previous tile
#pragma unroll 1
            for (int cc = 0; cc < 2; ++cc) {
                const int f0 = 64 * h + 32 * cc;
                b2::ld32f(c.lane_base + T_D + f0, v);
                const float4* b4 = reinterpret_cast<const float4*>(c.fw + F_B2 + f0);
#pragma unroll
                for (int k4 = 0; k4 < 8; ++k4) {
                    const float4 b = b4[k4];
                    v[4 * k4] += b.x; v[4 * k4 + 1] += b.y; v[4 * k4 + 2] += b.z; v[4 * k4 + 3] += b.w;
                }
                uint32_t hi[16], lo[16];
                b2::split32(v, hi, lo);
                if (h == 0) b2::st_op(c, T_R2_HI, T_R2_LO, f0, hi, lo);      // sdf_emb -> layer-3 operand, features [0, 64)
                else b2::scr_store(scr_p, SCR_RGB, 32, 16 * cc, hi, lo);     // rgb_emb: A operand of the colour-head wgrad
            }
            B3_MARK(3);
            // ---- forward layer 3 ----
            round([&]() {
                b2::issue_fwd<6>(c, IMG_W3_HI, IMG_W3_LO, [](int ks, bool lo) {
                    return ks < 4 ? (lo ? T_R2_LO : T_R2_HI) + 8 * ks : (lo ? T_G_LO : T_G_HI) + 8 * (ks - 4);
                });
            });
            float s[N_CLASS] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
            for (int cc = 0; cc < 2; ++cc) {
                const int f0 = 64 * h + 32 * cc;
                b2::ld32f(c.lane_base + T_D + f0, v);
                const float4* b4 = reinterpret_cast<const float4*>(c.fw + F_BS1 + f0);
                uint32_t m = 0;
#pragma unroll
                for (int k4 = 0; k4 < 8; ++k4) {
                    const float4 b = b4[k4];
                    v[4 * k4] = fmaxf(v[4 * k4] + b.x, 0.f); v[4 * k4 + 1] = fmaxf(v[4 * k4 + 1] + b.y, 0.f);
                    v[4 * k4 + 2] = fmaxf(v[4 * k4 + 2] + b.z, 0.f); v[4 * k4 + 3] = fmaxf(v[4 * k4 + 3] + b.w, 0.f);
                }
#pragma unroll
                for (int kk = 0; kk < 32; ++kk) m |= (v[kk] > 0.f ? 1u : 0u) << kk;
                if (cc == 0) mask3a = m; else mask3b = m;
#pragma unroll
                for (int ch = 0; ch < N_CLASS; ++ch) {
                    const float4* w4 = ws2g + (ch * 128 + f0) / 4;
#pragma unroll
                    for (int k4 = 0; k4 < 8; ++k4) {
                        const float4 w = __ldg(w4 + k4);
                        s[ch] = fmaf(w.x, v[4 * k4], s[ch]); s[ch] = fmaf(w.y, v[4 * k4 + 1], s[ch]);
                        s[ch] = fmaf(w.z, v[4 * k4 + 2], s[ch]); s[ch] = fmaf(w.w, v[4 * k4 + 3], s[ch]);
                    }
                }
                uint32_t hi[16], lo[16];
                b2::split32(v, hi, lo);
                b2::scr_store(scr_p, SCR_H3, 64, f0 / 2, hi, lo);            // H3: A operand of the sdf_linear.2 wgrad
            }
            // the two threads of a point exchange their logit partial sums through accumulator columns [64 h, 64 h + 8) of the
            // point's lane (columns this thread itself has just consumed; the next MMA into D is issued two rendezvous later)
            float so[N_CLASS];
            {
                uint32_t sx[8] = {__float_as_uint(s[0]), __float_as_uint(s[1]), __float_as_uint(s[2]), __float_as_uint(s[3]), __float_as_uint(s[4]), 0u, 0u, 0u};
                umma::tmem_st8(c.lane_base + T_D + 64 * h, sx);
                umma::wait_st();
                umma::fence_before_sync();
                chain_sync();
                umma::fence_after_sync();
                uint32_t ox[8];
                umma::tmem_ld8(c.lane_base + T_D + 64 * (1 - h), ox);
                umma::wait_ld();
#pragma unroll
                for (int ch = 0; ch < N_CLASS; ++ch) so[ch] = __uint_as_float(ox[ch]);
            }
            B3_MARK(4);
            // ---- heads: softmax forward + backward (both threads of a point, redundantly) ----
            float dz4[N_CLASS];
            {
                float zl[N_CLASS], pr[N_CLASS], dp[N_CLASS];
                float mx = -INFINITY, se = 0.f, dot = 0.f;
#pragma unroll
                for (int ch = 0; ch < N_CLASS; ++ch) {
                    zl[ch] = c.fw[G_BS2 + ch] + (s[ch] + so[ch]);
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
            // U = [dlogits 5 | dRGB 3 | 0 x 8] as bf16 hi / lo words (B operand of the two narrow-head wgrads) -> scratch:
            // the h = 0 thread parks the hi words, h = 1 the lo words
            {
                uint32_t uh[4], ul[4];
                umma::split2(dz4[0], dz4[1], uh[0], ul[0]); umma::split2(dz4[2], dz4[3], uh[1], ul[1]);
                umma::split2(dz4[4], g[0], uh[2], ul[2]); umma::split2(g[1], g[2], uh[3], ul[3]);
#pragma unroll
                for (int t = 0; t < 4; ++t) __stcg(scr_p + (size_t)(SCR_U + 4 * h + t) * TC_TP, h ? ul[t] : uh[t]);
            }
            // ---- dZ3 = (Ws2^T dz4) * relu'(h3) -> R1 (H1 is parked; its columns are free since the layer-2 MMAs completed) ----
#pragma unroll 1
            for (int cc = 0; cc < 2; ++cc) {
                const int f0 = 64 * h + 32 * cc;
#pragma unroll
                for (int kk = 0; kk < 32; ++kk) v[kk] = 0.f;
#pragma unroll
                for (int ch = 0; ch < N_CLASS; ++ch) {
                    const float4* w4 = ws2g + (ch * 128 + f0) / 4;
                    const float dzc = dz4[ch];
#pragma unroll
                    for (int k4 = 0; k4 < 8; ++k4) {
                        const float4 w = __ldg(w4 + k4);
                        v[4 * k4] = fmaf(w.x, dzc, v[4 * k4]); v[4 * k4 + 1] = fmaf(w.y, dzc, v[4 * k4 + 1]);
                        v[4 * k4 + 2] = fmaf(w.z, dzc, v[4 * k4 + 2]); v[4 * k4 + 3] = fmaf(w.w, dzc, v[4 * k4 + 3]);
                    }
                }
                const uint32_t m = cc == 0 ? mask3a : mask3b;
#pragma unroll
                for (int kk = 0; kk < 32; ++kk) v[kk] = ((m >> kk) & 1u) ? v[kk] : 0.f;
                uint32_t zh[16], zl[16];
                b2::split32(v, zh, zl);
                b2::st_op(c, T_R1_HI, T_R1_LO, f0, zh, zl);
            }
            B3_MARK(5);
            // ---- hand layer 3 to COPY (dZ3 in R1, x3 in R2, H3 / U parked) and run its dgrad (whose completion also releases
            //      SCATTER on the grid-feature columns of D) ----
            umma::wait_st();
            umma::fence_before_sync();
            __threadfence_block();
            chain_sync();
            if (ctid == 0) {
                umma::fence_after_sync();
                umma::mbar_arrive(c.bars + B_Z3);
                b2::issue_dgrad<D_SDF_IN>(c, T_R1_HI, T_R1_LO, IMG_W3_HI, IMG_W3_LO);
                umma::commit(c.bars + B_DG3);
            }
            c.ok &= umma::mbar_wait_spin(c.bars + B_DG3, k & 1u);
            c.ok &= umma::mbar_wait_spin(c.bars + B_RD0, k & 1u);           // COPY holds dZ3 and x3 in registers: R2 may be overwritten
            umma::fence_after_sync();
            B3_MARK(6);
            // ---- dH = [d sdf_emb (dgrad of layer 3), d rgb_emb (colour head)] -> R2 ----
#pragma unroll 1
            for (int cc = 0; cc < 2; ++cc) {
                const int f0 = 64 * h + 32 * cc;
                if (h == 0) {
                    b2::ld32f(c.lane_base + T_D + f0, v);
                } else {
                    const float* wr = c.fw + F_WR_EMB + 32 * cc;
#pragma unroll
                    for (int kk = 0; kk < 32; ++kk) v[kk] = fmaf(wr[128 + kk], g[2], fmaf(wr[64 + kk], g[1], wr[kk] * g[0]));
                }
                uint32_t zh[16], zl[16];
                b2::split32(v, zh, zl);
                b2::st_op(c, T_R2_HI, T_R2_LO, f0, zh, zl);
            }
            // ---- hand layer 2 to COPY, dgrad of layer 2 (D is overwritten: SCATTER must have taken its columns) ----
            umma::wait_st();
            umma::fence_before_sync();
            chain_sync();
            if (ctid == 0) {
                umma::fence_after_sync();
                umma::mbar_arrive(c.bars + B_ZH);
                c.ok &= umma::mbar_wait_spin(c.bars + B_DCONS, k & 1u);
                umma::fence_after_sync();
                b2::issue_dgrad<D_H>(c, T_R2_HI, T_R2_LO, IMG_W2_HI, IMG_W2_LO);
                umma::commit(c.bars + B_MMA);
            }
            B3_MARK(7);
            c.ok &= umma::mbar_wait_spin(c.bars + B_MMA, ph_mma);
            ph_mma ^= 1;
            umma::fence_after_sync();
            B3_MARK(8);
            // ---- dZ1 = dgrad2 * relu'(h1) -> R1 (COPY took dZ3 before rd0), hand layer 1 to COPY ----
#pragma unroll 1
            for (int cc = 0; cc < 2; ++cc) {
                const int f0 = 64 * h + 32 * cc;
                b2::ld32f(c.lane_base + T_D + f0, v);
                const uint32_t m = cc == 0 ? mask1a : mask1b;
#pragma unroll
                for (int kk = 0; kk < 32; ++kk) v[kk] = ((m >> kk) & 1u) ? v[kk] : 0.f;
                uint32_t zh[16], zl[16];
                b2::split32(v, zh, zl);
                b2::st_op(c, T_R1_HI, T_R1_LO, f0, zh, zl);
            }
            umma::wait_st();
            warp_arrive(c.bars + B_Z1, lane);
            B3_MARK(9);
        }
    } else if (tid >= SC_NT) {
        // =============================== COPY ===============================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_COPY));
        const int cp = tid - SC_NT, h = cp >> 7;
        const int row = p & (b2::QROWS - 1), quarter = p >> 5;
        // destination of this lane's row of the colour-head product [e | rgb_emb]^T U inside rgb_linear.0.weight (or -1)
        const int kin = p >= 64 ? p - 64 : (tc_e_slot_to_index(p) >= 0 ? 64 + tc_e_slot_to_index(p) : -1);
        uint32_t PA[8];                                // chunk addresses of this thread's row in block 0 of a tile; block h: + hq
#pragma unroll
        for (int j = 0; j < 8; ++j)
            PA[j] = umma::smem_u32(c.base) + (uint32_t)(S_ST + (quarter & 1) * ST_BYTES) + (uint32_t)((row >> 3) * 1024 + (row & 7) * 128) +
                    (uint32_t)(((j ^ row) & 7) << 4);
        const uint32_t hq = (uint32_t)h * QBLK;
        const uint64_t dA0 = umma::desc_mn(c.base + S_ST + (quarter & 1) * ST_BYTES + ST_A, 0, QBLK);
        const uint64_t dB0 = umma::desc_mn(c.base + S_ST + (quarter & 1) * ST_BYTES + ST_B, 0, QBLK);
        constexpr uint32_t LO_STEP = (2 * QBLK) >> 4, KS_STEP = (16 * 128) >> 4;
        // a staging buffer is free once the MMAs of the quarter that used it before have completed (field_tc_bwd2.cuh)
        auto prod_begin = [&](uint32_t prod) {
            c.ok &= umma::mbar_wait_spin(c.bars + B_FREE0 + ((quarter + 2) & 3), quarter >= 2 ? (prod & 1u) : ((prod + 1u) & 1u));
        };
        auto prod_end = [&](int PK, uint32_t prod) {
            umma::fence_proxy_async();
            asm volatile("bar.sync %0, 64;" ::"r"(3 + quarter) : "memory");
            if (h == 0 && lane == 0) {
                umma::fence_after_sync();
                if (quarter != 0) c.ok &= umma::mbar_wait_spin(c.bars + B_ISS0 + quarter - 1, prod & 1u);
                const int n_out = PK == 0 ? N_X3 : (PK == 2 ? D_H : (PK == 3 ? 64 : 16));
                const uint32_t idesc = umma::idesc_bf16(128, n_out, 1, 1);
                const uint32_t dcol = c.tmem + (uint32_t)(T_DW + (PK == 1 ? N_X3 : (PK == 4 ? 64 : 0)));
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    const int pass = j >> 1, ks = j & 1;
                    umma::mma_ss(dcol, dA0 + (uint64_t)((pass == 1 ? LO_STEP : 0) + ks * KS_STEP),
                                 dB0 + (uint64_t)((pass == 2 ? LO_STEP : 0) + ks * KS_STEP), idesc, (j == 0 && quarter == 0) ? 0u : 1u);
                }
                umma::mbar_arrive(c.bars + B_ISS0 + quarter);
                umma::commit(c.bars + B_FREE0 + quarter);
                if (PK == 1 || PK == 2 || PK == 4) umma::commit(c.bars + B_DWRDY);       // (covers this thread's MMAs of P - 1 too)
            }
        };
        // A tile <- this thread's 64 features (two 32-feature groups) of a tensor-memory operand region; the buffer is claimed
        // once the first group is in registers
        auto a_from_tmem = [&](int r_hi, int r_lo, uint32_t prod) {
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                uint32_t wh[16], wl[16];
                umma::tmem_ld16(c.lane_base + (uint32_t)(r_hi + 16 * (2 * h + cc)), wh);
                umma::tmem_ld16(c.lane_base + (uint32_t)(r_lo + 16 * (2 * h + cc)), wl);
                umma::wait_ld();
                if (cc == 0) prod_begin(prod);
                if (cc == 0) { put32h<ST_A, 0>(PA, hq, wh); put32h<ST_A + 2 * (int)QBLK, 0>(PA, hq, wl); }
                else { put32h<ST_A, 1>(PA, hq, wh); put32h<ST_A + 2 * (int)QBLK, 1>(PA, hq, wl); }
            }
        };
        // tile at TILE_OFF <- this thread's 64 parked features (words 32 h .. 32 h + 31 of a 128-word hi | lo scratch region)
        auto from_scratch = [&](auto tile_off, int region, uint32_t prod, bool claim) {
            constexpr int TO = decltype(tile_off)::value;
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                uint32_t wh[16], wl[16];
                b2::scr_load(scr_p, region, 64, 32 * h + 16 * cc, wh, wl);
                if (cc == 0 && claim) prod_begin(prod);
                if (cc == 0) { put32h<TO, 0>(PA, hq, wh); put32h<TO + 2 * (int)QBLK, 0>(PA, hq, wl); }
                else { put32h<TO, 1>(PA, hq, wh); put32h<TO + 2 * (int)QBLK, 1>(PA, hq, wl); }
            }
        };
        // B tile <- U (features 0 .. 15 of block 0): the h = 0 thread writes the hi part, h = 1 the lo part
        auto put_u = [&]() {
            uint32_t u4[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) u4[t] = __ldcg(scr_p + (size_t)(SCR_U + 4 * h + t) * TC_TP);
            if (h == 0) { b2::sts128<ST_B>(PA[0], u4[0], u4[1], u4[2], u4[3]); b2::sts128<ST_B>(PA[1], 0u, 0u, 0u, 0u); }
            else { b2::sts128<ST_B + 2 * (int)QBLK>(PA[0], u4[0], u4[1], u4[2], u4[3]); b2::sts128<ST_B + 2 * (int)QBLK>(PA[1], 0u, 0u, 0u, 0u); }
        };
        // read-out of the finished layer: lane n owns DW[n][*], the two warps of a lane quadrant take alternate 16-column chunks;
        // every element has one owner thread (reductions without return into the CTA's partial)
        uint32_t lay = 0;
        auto read_out = [&](int L) {
            c.ok &= umma::mbar_wait_spin(c.bars + B_DWRDY, lay & 1u);
            umma::fence_after_sync();
            const int n_cols = L == 2 ? 80 : 128;
#pragma unroll 1
            for (int c0 = 16 * h; c0 < n_cols; c0 += 32) {
                uint32_t r[16];
                umma::tmem_ld16(c.lane_base + (uint32_t)(T_DW + c0), r);
                umma::wait_ld();
                if (L == 1 || (L == 0 && c0 < D_SDF_IN)) {             // dW2 / dWs1: column kc -> [kc][n]
                    float* dst = gpart + (L == 1 ? OFF_W2 : OFF_WS1) + c0 * D_H + p;
#pragma unroll
                    for (int j = 0; j < 16; ++j) atomicAdd(dst + j * D_H, __uint_as_float(r[j]));
                } else if (L == 0) {
                    if (c0 == D_SDF_IN) {                              // ones column: dbs1
                        atomicAdd(&gpart[OFF_BS1 + p], __uint_as_float(r[0]));
                    } else {                                           // H3^T U: dWs2[c][n]
#pragma unroll
                        for (int j = 0; j < N_CLASS; ++j) atomicAdd(&gpart[OFF_WS2 + j * D_H + p], __uint_as_float(r[j]));
                    }
                } else if (c0 < 64) {                                  // dW1: slot kc -> e index (slot 15 = ones column: db1)
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int ei = tc_e_slot_to_index(c0 + j);
                        if (ei >= 0) atomicAdd(&gpart[OFF_W1 + ei * D_H + p], __uint_as_float(r[j]));
                        else if (c0 + j == 15) atomicAdd(&gpart[OFF_B1 + p], __uint_as_float(r[j]));
                    }
                } else if (p == 15) {                                  // the ones row of [e | rgb_emb]: sums of U over the points
#pragma unroll
                    for (int j = 0; j < N_CLASS; ++j) atomicAdd(&gpart[OFF_BS2 + j], __uint_as_float(r[j]));
#pragma unroll
                    for (int j = 0; j < 3; ++j) atomicAdd(&gpart[OFF_BR + j], __uint_as_float(r[N_CLASS + j]));
                } else if (kin >= 0) {                                 // [e | rgb_emb]^T dRGB: rgb_linear.0.weight
#pragma unroll
                    for (int j = 0; j < 3; ++j) atomicAdd(&gpart[OFF_WR + j * D_RGB_IN + kin], __uint_as_float(r[N_CLASS + j]));
                }
            }
            umma::fence_before_sync();
            copy_sync();                                // DW fully read: the next layer's first MMA may overwrite it
            ++lay;
        };
        using OffA = std::integral_constant<int, ST_A>;
        using OffB = std::integral_constant<int, ST_B>;

        uint32_t k = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++k) {
            const int64_t slot = tile * TC_TP + p;
            const int64_t i = slot < N ? am(slot) : 0;
            const uint32_t* fin = f.feat + (size_t)(i >> 7) * FEAT_TILE_WORDS + (i & (TC_TP - 1));
            const uint32_t prod0 = 5u * k;
            // ---------------- layer 3: P0 = dZ3^T [sdf_emb | grid | 1] -> DW[0, 112), P1 = H3^T U -> DW[112, 128) ----------------
            c.ok &= umma::mbar_wait_spin(c.bars + B_Z3, k & 1u);
            umma::fence_after_sync();
            B3_MARK(16);
            a_from_tmem(T_R1_HI, T_R1_LO, prod0);
            if (h == 0) {                                               // sdf_emb: features [0, 64) of x3
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    uint32_t wh[16], wl[16];
                    umma::tmem_ld16(c.lane_base + T_R2_HI + 16 * cc, wh); umma::tmem_ld16(c.lane_base + T_R2_LO + 16 * cc, wl);
                    umma::wait_ld();
                    if (cc == 0) { put32h<ST_B, 0>(PA, hq, wh); put32h<ST_B + 2 * (int)QBLK, 0>(PA, hq, wl); }
                    else { put32h<ST_B, 1>(PA, hq, wh); put32h<ST_B + 2 * (int)QBLK, 1>(PA, hq, wl); }
                }
            } else {                                                    // grid features [64, 96), feature 96 = 1.0 (bias column), zeros
                uint32_t wh[16], wl[16];
                umma::tmem_ld16(c.lane_base + T_G_HI, wh); umma::tmem_ld16(c.lane_base + T_G_LO, wl);
                umma::wait_ld();
                put32h<ST_B, 0>(PA, hq, wh); put32h<ST_B + 2 * (int)QBLK, 0>(PA, hq, wl);
#pragma unroll
                for (int t = 0; t < 16; ++t) { wh[t] = 0u; wl[t] = 0u; }
                wh[0] = 0x00003F80u;
                put32h<ST_B, 1>(PA, hq, wh); put32h<ST_B + 2 * (int)QBLK, 1>(PA, hq, wl);
            }
            warp_arrive(c.bars + B_RD0, lane);                           // R1 (dZ3) and R2 (x3) are in shared memory
            prod_end(0, prod0);
            B3_MARK(17);
            from_scratch(OffA{}, SCR_H3, prod0 + 1u, true);
            put_u();
            prod_end(1, prod0 + 1u);
            B3_MARK(18);
            read_out(0);
            B3_MARK(19);
            // ---------------- layer 2: P2 = dH^T H1 -> DW[0, 128) ----------------
            c.ok &= umma::mbar_wait_spin(c.bars + B_ZH, k & 1u);
            umma::fence_after_sync();
            B3_MARK(20);
            a_from_tmem(T_R2_HI, T_R2_LO, prod0 + 2u);
            warp_arrive(c.bars + B_RD2, lane);
            from_scratch(OffB{}, SCR_H1, 0u, false);
            prod_end(2, prod0 + 2u);
            B3_MARK(21);
            read_out(1);
            B3_MARK(22);
            // ---------------- layer 1: P3 = dZ1^T e (ones in slot 15: bias column) -> DW[0, 64),
            //                  P4 = [e | rgb_emb (parked)]^T U -> DW[64, 80) (the ones row gives both head biases).  Thread h owns
            //                  slot groups 2h, 2h+1 of e (its half of block 0) and half h of rgb_emb (block 1) ----------------
            c.ok &= umma::mbar_wait_spin(c.bars + B_Z1, k & 1u);
            umma::fence_after_sync();
            B3_MARK(23);
            a_from_tmem(T_R1_HI, T_R1_LO, prod0 + 3u);
            warp_arrive(c.bars + B_RD3, lane);
            // e words of slot group 2h + sg (slot 15 = 1.0 for the bias column / ones row) -> 16 features of block 0 of a tile
            auto put_e = [&](auto tile_off, int sg) {
                constexpr int TO = decltype(tile_off)::value;
                uint32_t eh[8], el[8];
                const uint32_t* fq = fin + (size_t)(2 * h + sg) * FEAT_WORDS * TC_TP;
#pragma unroll
                for (int t = 0; t < 8; ++t) { eh[t] = __ldg(fq + t * TC_TP); el[t] = __ldg(fq + (8 + t) * TC_TP); }
                if (h == 0 && sg == 0) { eh[7] = (eh[7] & 0x0000ffffu) | 0x3F800000u; el[7] &= 0x0000ffffu; }
#pragma unroll
                for (int c_ = 0; c_ < 2; ++c_) {
                    const uint32_t a = h ? PA[4 + 2 * sg + c_] : PA[2 * sg + c_];
                    b2::sts128<TO>(a, eh[4 * c_], eh[4 * c_ + 1], eh[4 * c_ + 2], eh[4 * c_ + 3]);
                    b2::sts128<TO + 2 * (int)QBLK>(a, el[4 * c_], el[4 * c_ + 1], el[4 * c_ + 2], el[4 * c_ + 3]);
                }
            };
            put_e(OffB{}, 0); put_e(OffB{}, 1);
            prod_end(3, prod0 + 3u);
            B3_MARK(24);
            {
                uint32_t rh[16], rl[16];
                b2::scr_load(scr_p, SCR_RGB, 32, 16 * h, rh, rl);
                prod_begin(prod0 + 4u);
#pragma unroll
                for (int c_ = 0; c_ < 4; ++c_) {
                    const uint32_t a = h ? PA[4 + c_] : PA[c_];
                    b2::sts128<ST_A + (int)QBLK>(a, rh[4 * c_], rh[4 * c_ + 1], rh[4 * c_ + 2], rh[4 * c_ + 3]);
                    b2::sts128<ST_A + 3 * (int)QBLK>(a, rl[4 * c_], rl[4 * c_ + 1], rl[4 * c_ + 2], rl[4 * c_ + 3]);
                }
            }
            put_e(OffA{}, 0); put_e(OffA{}, 1);
            put_u();
            warp_arrive(c.bars + B_RD4, lane);                           // rgb_emb and U of this tile are in shared memory
            prod_end(4, prod0 + 4u);
            B3_MARK(25);
            read_out(2);
            B3_MARK(26);
        }
    } else {
        // =============================== SCATTER ===============================
        uint32_t k = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++k) {
            const int64_t slot = tile * TC_TP + p;
            const bool valid = slot < N;
            const int64_t i = valid ? am(slot) : 0;
            float x[3] = {0.f, 0.f, 0.f};
            if (valid) src.point(i, f, x);
            c.ok &= umma::mbar_wait_spin(c.bars + B_DG3, k & 1u);
            umma::fence_after_sync();
            uint32_t r[32];
            umma::tmem_ld32(c.lane_base + T_D + 64, r);
            umma::wait_ld();
            umma::fence_before_sync();
            umma::mbar_arrive(c.bars + B_DCONS);
            if (prof && blockIdx.x == 0 && k == 1 && p == 0) prof[28] = clock64();
            if (valid) {
                float dx[3];
#pragma unroll 1
                for (int gq = 0; gq < 4; ++gq) {         // rolled over groups of 4 levels: the roles share the instruction cache
                    float dy8[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        dy8[j] = __uint_as_float(gq == 0 ? r[j] : (gq == 1 ? r[8 + j] : (gq == 2 ? r[16 + j] : r[24 + j])));
#pragma unroll
                    for (int ll = 0; ll < 4; ++ll)
                        grid_level_bwd<false>(x, make_float2(dy8[2 * ll], dy8[2 * ll + 1]), nullptr, grad_grid, level_info(f, gq * 4 + ll), dx);
                }
            }
            if (prof && blockIdx.x == 0 && k == 1 && p == 0) prof[29] = clock64();
        }
    }
#undef B3_MARK
    if (!c.ok && err) atomicExch(err, 1);
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc<512>(c.tmem);
}

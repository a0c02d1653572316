// N2: marching cubes on a dense SDF volume, on the device, bit-identical to the reference's NumpyMarchingCubes
// (/root/reference/external/NumpyMarchingCubes/marching_cubes/src/marching_cubes.cpp, called from utils/utils.py:78,159).
//
// The reference is a sequential scan: per integer cell centre it tri-linearly samples 8 corners at +-0.5 (each the mean of 8
// voxels, :93-113), classifies (:160-167), emits triangles in table order (:219-227), then clusters the triangle-soup vertices
// greedily on a 1e-5 lattice in soup order (:316-390) and drops degenerate and duplicate faces (:246-288).  The sequential
// parts are restated as order-free fixed points so that every step is a flat kernel:
//   mc_nodes_kernel    corner ("dual node") values once per node instead of 8 x per cell; NaN marks an invalid node
//   mc_classify_kernel triangle count per cell (u8) + totals per block of 1,024 cells      } offsets = exclusive scan, which
//   mc_emit_kernel     triangle soup written at scan offsets -> the reference's order      } reproduces the i,j,k scan order
//   mc_keys / mc_insert lattice key per soup vertex; open-addressing table keyed by the 96-bit lattice cell.  A slot stores only
//                      the index of the vertex that claimed it (32-bit CAS); its key is re-read from keys[owner], so no
//                      multi-word atomics are needed.  first[slot] = atomicMin over the vertices of the cell.
//   mc_resolve_kernel  the greedy clustering as a fixed point: a lattice cell becomes a REPRESENTATIVE iff no cell of its
//                      27-neighbourhood that was first touched earlier is one ("lexicographically first independent set");
//                      cells whose earlier neighbours are all decided get decided, rounds run until none is left (the cell with
//                      the smallest first-touch index among the undecided is always decidable, so each round makes progress).
//   mc_lookup_kernel   soup vertex -> representative: first hit in the reference's (di,dj,dk) probe order among representatives
//                      inserted before it.
//   mc_face_* kernels  degenerate faces dropped, duplicates resolved by atomicMin of the face index on the sorted triple.
// All float arithmetic uses the round-to-nearest intrinsics so nothing is contracted into FMAs (the reference is x86-64 SSE2).
#include "mf_common.cuh"

#define MC_TABLE_DECL __constant__
#define MC_NTRI_DECL __device__ const
#include "mc_tables.h"

namespace {

constexpr uint32_t EMPTY = 0xFFFFFFFFu;
constexpr int CB = 256;                 // threads per classify block (4 cells each: 1,024 cells per block)
constexpr int SCAN_T = 512, SCAN_I = 8, SCAN_B = SCAN_T * SCAN_I;   // exclusive scan: 4096 items per block

__device__ __forceinline__ bool voxel_valid(float d, float trunc) { return d != -INFINITY && fabsf(d) < trunc; }

// ---- volume-shaped ("padded") layouts --------------------------------------------------------------------------------------
// Nodes and per-cell triangle counts keep the VOLUME's strides (plane stride PL = ny * nz, row stride nz): node (a,b,c) lives at
// a * PL + b * nz + c like voxel (a,b,c), the entries with b = ny-1 or c = nz-1 are never read.  A thread then owns 4 consecutive
// flat indices of a plane: two 16-byte loads + two scalars per plane feed 4 nodes (or the 8 corners of 4 cells), stores are 16
// bytes, and no kernel divides by a row length except once per thread.  VEC = the 16-byte path (nz % 4 == 0 and aligned bases).
template <bool VEC>
__device__ __forceinline__ void load5(const float* __restrict__ p, float* r) {
    if (VEC) { float4 v = __ldg(reinterpret_cast<const float4*>(p)); r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w; }
    else { r[0] = __ldg(p); r[1] = __ldg(p + 1); r[2] = __ldg(p + 2); r[3] = __ldg(p + 3); }
    r[4] = __ldg(p + 4);
}
// the 8 corner values of element i of a thread's 4, in the reference's order 000,100,010,001,110,011,101,111 (first index = plane)
#define MC_CORNERS(i) {r00[i], r10[i], r01[i], r00[i + 1], r11[i], r01[i + 1], r10[i + 1], r11[i + 1]}

__device__ __forceinline__ float node_value(const float* v, float trunc) {
    bool ok = true;
    float dist = 0.0f;
#pragma unroll
    for (int q = 0; q < 8; q++) {
        ok = ok && voxel_valid(v[q], trunc);
        dist = __fadd_rn(dist, __fmul_rn(0.125f, v[q]));          // (0.5 * 0.5 * 0.5) * d: exact product, the reference's sum order
    }
    return ok ? dist : __int_as_float(0x7fc00000);
}

// ---- dual nodes: node (a,b,c) = trilerp at (a+.5, b+.5, c+.5) (:93-113).  grid (ceil(PL / 1024), nx - 1) ----
template <bool VEC>
__global__ void __launch_bounds__(256) mc_nodes_kernel(const float* __restrict__ vol, float* __restrict__ node, int PL, int ny, int nz,
                                                       float trunc) {
    int j = (blockIdx.x * 256 + threadIdx.x) * 4, a = blockIdx.y;
    if (j >= PL) return;
    const float* p0 = vol + (int64_t)a * PL + j;
    float* out = node + (int64_t)a * PL + j;
    if (j + 4 + nz < PL) {                                        // every read stays inside the two planes
        float r00[5], r01[5], r10[5], r11[5];
        load5<VEC>(p0, r00); load5<VEC>(p0 + nz, r01); load5<VEC>(p0 + PL, r10); load5<VEC>(p0 + PL + nz, r11);
        float o[4];
#pragma unroll
        for (int i = 0; i < 4; i++) { const float v[8] = MC_CORNERS(i); o[i] = node_value(v, trunc); }
        if (VEC) *reinterpret_cast<float4*>(out) = make_float4(o[0], o[1], o[2], o[3]);
        else { out[0] = o[0]; out[1] = o[1]; out[2] = o[2]; out[3] = o[3]; }
        return;
    }
    for (int i = 0; i < 4 && j + i < PL; i++) {                   // tail of the plane: rows ny-2 (last elements) and ny-1
        int b = (j + i) / nz, c = (j + i) - b * nz;
        float val = __int_as_float(0x7fc00000);
        if (b < ny - 1 && c < nz - 1) {
            const float* p = p0 + i;
            const float v[8] = {__ldg(p), __ldg(p + PL), __ldg(p + nz), __ldg(p + 1), __ldg(p + PL + nz), __ldg(p + nz + 1), __ldg(p + PL + 1),
                                __ldg(p + PL + nz + 1)};
            val = node_value(v, trunc);
        }
        out[i] = val;
    }
}

// corner numbering 0:000 1:100 2:010 3:001 4:110 5:011 6:101 7:111
__device__ __forceinline__ int cube_index(const float* d, float iso, bool* ok) {
    const int bit[8] = {8, 4, 1, 128, 2, 16, 64, 32};             // :160-167
    int cube = 0;
    bool all = true;
#pragma unroll
    for (int q = 0; q < 8; q++) {
        all = all && (d[q] == d[q]);
        if (d[q] < iso) cube += bit[q];
    }
    *ok = all;
    return cube;
}
// the threshold tests of marching_cubes.cpp:170-188 (thresh = 10)
__device__ __forceinline__ bool cell_within_thresh(const float* d) {
    const float thresh = 10.0f;
    bool ok = true;
#pragma unroll
    for (int x = 0; x < 8; x++) {
        ok = ok && !(fabsf(d[x]) > thresh);
#pragma unroll
        for (int y = x + 1; y < 8; y++) {
            if (__fmul_rn(d[x], d[y]) < 0.0f) ok = ok && !(__fadd_rn(fabsf(d[x]), fabsf(d[y])) > thresh);
            else ok = ok && !(fabsf(__fsub_rn(d[x], d[y])) > thresh);
        }
    }
    return ok;
}
__device__ __forceinline__ int cell_triangles(const float* d, float iso, int* cube_out, const uint8_t* __restrict__ ntri_of_case) {
    bool ok;
    int cube = cube_index(d, iso, &ok);
    *cube_out = cube;
    if (!ok || cube == 0 || cube == 255) return 0;
    // every |d| <= 5  =>  none of the thresh = 10 tests can fire (|dx| + |dy| <= 10 and |dx - dy| <= 10 also after rounding)
    float m = fmaxf(fmaxf(fmaxf(fabsf(d[0]), fabsf(d[1])), fmaxf(fabsf(d[2]), fabsf(d[3]))),
                    fmaxf(fmaxf(fabsf(d[4]), fabsf(d[5])), fmaxf(fabsf(d[6]), fabsf(d[7]))));
    if (m > 5.0f && !cell_within_thresh(d)) return 0;
    return ntri_of_case[cube];
}

// ---- triangle count per cell.  Cell (a,b,c) (centre (a+1,b+1,c+1)) uses nodes (a..a+1, b..b+1, c..c+1).  grid (ceil(PL / 1024), nx - 2);
// block number blockIdx.y * gridDim.x + blockIdx.x and the flat index inside the block order the triangles like the reference's
// i, j, k loops (:424-431). ----
template <bool VEC>
__global__ void __launch_bounds__(CB) mc_classify_kernel(const float* __restrict__ node, uint8_t* __restrict__ ntri_cell,
                                                         uint32_t* __restrict__ block_tris, int PL, int ny, int nz, float iso) {
    __shared__ uint8_t ntri_of_case[256];
    ntri_of_case[threadIdx.x] = MC_NTRI[threadIdx.x];
    __syncthreads();
    int j = (blockIdx.x * CB + threadIdx.x) * 4, a = blockIdx.y;
    int n[4] = {0, 0, 0, 0};
    if (j + 4 + nz < PL) {                                        // beyond: rows >= ny-2, no cells
        const float* p0 = node + (int64_t)a * PL + j;
        float r00[5], r01[5], r10[5], r11[5];
        load5<VEC>(p0, r00); load5<VEC>(p0 + nz, r01); load5<VEC>(p0 + PL, r10); load5<VEC>(p0 + PL + nz, r11);
        int b = j / nz, c = j - b * nz;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (b < ny - 2 && c < nz - 2) { const float d[8] = MC_CORNERS(i); int cube; n[i] = cell_triangles(d, iso, &cube, ntri_of_case); }
            if (++c == nz) { c = 0; b++; }
        }
    }
    if (j < PL) {
        uint8_t* o = ntri_cell + (int64_t)a * PL + j;
        if (VEC) *reinterpret_cast<uchar4*>(o) = make_uchar4((uint8_t)n[0], (uint8_t)n[1], (uint8_t)n[2], (uint8_t)n[3]);
        else for (int i = 0; i < 4 && j + i < PL; i++) o[i] = (uint8_t)n[i];
    }
    int s = n[0] + n[1] + n[2] + n[3];
    int64_t blk = (int64_t)blockIdx.y * gridDim.x + blockIdx.x;
    if (!__syncthreads_or(s)) { if (threadIdx.x == 0) block_tris[blk] = 0u; return; }
    __shared__ int ws[CB / 32];
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < CB / 32; w++) t += ws[w];
        block_tris[blk] = (uint32_t)t;
    }
}
// blocks that hold triangles, in any order (their soup offsets come from the scan)
__global__ void mc_active_blocks_kernel(const uint32_t* __restrict__ block_tris, uint32_t* list, uint32_t* count, int64_t n_blocks) {
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b < n_blocks && block_tris[b]) list[atomicAdd(count, 1u)] = (uint32_t)b;
}

struct F3 { float x, y, z; };
__device__ __forceinline__ F3 vertex_interp(float iso, F3 p1, F3 p2, float d1, float d2) {      // :115-136
    if (fabsf(__fsub_rn(iso, d1)) < 0.00001f) return p1;
    if (fabsf(__fsub_rn(iso, d2)) < 0.00001f) return p2;
    if (fabsf(__fsub_rn(d1, d2)) < 0.00001f) return p1;
    float mu = __fdiv_rn(__fsub_rn(iso, d1), __fsub_rn(d2, d1));
    F3 r;
    r.x = __fadd_rn(p1.x, __fmul_rn(mu, __fsub_rn(p2.x, p1.x)));
    r.y = __fadd_rn(p1.y, __fmul_rn(mu, __fsub_rn(p2.y, p1.y)));
    r.z = __fadd_rn(p1.z, __fmul_rn(mu, __fsub_rn(p2.z, p1.z)));
    return r;
}

// ---- triangle soup.  Persistent CTAs; every WARP takes blocks (1,024 cells) that hold triangles from the list: a lane reads its
// 32 cell counts as two 16-byte words, the warp scans the lane totals with shuffles, every triangle gets a (cell slot, number
// within the cell) tag in the warp's shared-memory strip, and then one LANE PER TRIANGLE interpolates its three vertices (a lane
// per cell would idle: few of a block's cells touch the surface).  No block-wide barrier. ----
constexpr int EMIT_WARPS = 4;
__global__ void __launch_bounds__(EMIT_WARPS * 32) mc_emit_kernel(const float* __restrict__ node, const uint8_t* __restrict__ ntri_cell,
                                                                  const uint32_t* __restrict__ block_off, const uint32_t* __restrict__ active,
                                                                  const uint32_t* __restrict__ n_active, float* __restrict__ soup, int PL, int nz,
                                                                  int blocks_per_plane, float iso, int vec) {
    __shared__ uint16_t tag_all[EMIT_WARPS][CB * 4 * 5];         // (cell slot << 3) | triangle number; at most 5 triangles per cell
    // corner q has +0.5 along x iff bit q of 0xD2, y: 0xB4, z: 0xE8; edge e joins corners (E1 >> 4e) & 15 and (E2 >> 4e) & 15:
    // {2,4},{4,1},{1,0},{0,2},{5,7},{7,6},{6,3},{3,5},{2,5},{4,7},{1,6},{0,3} (:205-216)
    const uint64_t E1 = 0x014236750142ull, E2 = 0x367553672014ull;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint16_t* tag = tag_all[w];
    const uint32_t total = *n_active, stride = gridDim.x * EMIT_WARPS;
    for (uint32_t it = blockIdx.x * EMIT_WARPS + w; it < total; it += stride) {
        uint32_t blk = active[it];
        int a = (int)(blk / (uint32_t)blocks_per_plane), chunk = (int)(blk - (uint32_t)a * blocks_per_plane);
        int j0 = chunk * CB * 4, j = j0 + lane * 32;             // this lane's 32 cells
        const uint8_t* src = ntri_cell + (int64_t)a * PL + j;
        uint32_t wd[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (vec && j + 32 <= PL) {
            uint4 u0 = __ldg(reinterpret_cast<const uint4*>(src)), u1 = __ldg(reinterpret_cast<const uint4*>(src) + 1);
            wd[0] = u0.x; wd[1] = u0.y; wd[2] = u0.z; wd[3] = u0.w; wd[4] = u1.x; wd[5] = u1.y; wd[6] = u1.z; wd[7] = u1.w;
        } else {
            for (int i = 0; i < 32 && j + i < PL; i++) wd[i >> 2] |= (uint32_t)src[i] << (8 * (i & 3));
        }
        int tot = 0;
#pragma unroll
        for (int q = 0; q < 8; q++) tot += (int)((wd[q] * 0x01010101u) >> 24);     // byte sum (each byte <= 5)
        int incl = tot;
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        int block_total = __shfl_sync(0xffffffffu, incl, 31);
        __syncwarp();                                             // the previous block's tags are no longer read
        if (tot) {
            int t0 = incl - tot;
            for (int q = 0; q < 8; q++) {
                if (!wd[q]) continue;
                for (int i = 0; i < 4; i++) {
                    int n = (int)((wd[q] >> (8 * i)) & 255);
                    for (int t = 0; t < n; t++) tag[t0++] = (uint16_t)(((lane * 32 + q * 4 + i) << 3) | t);
                }
            }
        }
        __syncwarp();
        float* out0 = soup + (int64_t)block_off[blk] * 9;
        for (int t = lane; t < block_total; t += 32) {
            int slot = tag[t] >> 3, tn = tag[t] & 7;
            int jj = j0 + slot, b = jj / nz, c = jj - b * nz;
            const float* p = node + (int64_t)a * PL + jj;
            float d[8] = {__ldg(p), __ldg(p + PL), __ldg(p + nz), __ldg(p + 1), __ldg(p + PL + nz), __ldg(p + nz + 1), __ldg(p + PL + 1),
                          __ldg(p + PL + nz + 1)};
            bool ok;
            int cube = cube_index(d, iso, &ok);
            float fx = (float)(a + 1), fy = (float)(b + 1), fz = (float)(c + 1);        // cell centre; corners at +-0.5
            uint64_t row = MC_TRI_PACKED[cube] >> (12 * tn);
            float* out = out0 + (int64_t)t * 9;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                int e = (int)((row >> (4 * k)) & 15);
                int q1 = (int)((E1 >> (4 * e)) & 15), q2 = (int)((E2 >> (4 * e)) & 15);
                F3 p1 = {__fadd_rn(fx, (0xD2 >> q1) & 1 ? 0.5f : -0.5f), __fadd_rn(fy, (0xB4 >> q1) & 1 ? 0.5f : -0.5f), __fadd_rn(fz, (0xE8 >> q1) & 1 ? 0.5f : -0.5f)};
                F3 p2 = {__fadd_rn(fx, (0xD2 >> q2) & 1 ? 0.5f : -0.5f), __fadd_rn(fy, (0xB4 >> q2) & 1 ? 0.5f : -0.5f), __fadd_rn(fz, (0xE8 >> q2) & 1 ? 0.5f : -0.5f)};
                float d1 = d[0], d2 = d[0];
#pragma unroll
                for (int q = 1; q < 8; q++) { d1 = q1 == q ? d[q] : d1; d2 = q2 == q ? d[q] : d2; }
                F3 v = vertex_interp(iso, p1, p2, d1, d2);
                out[3 * k] = v.x; out[3 * k + 1] = v.y; out[3 * k + 2] = v.z;
            }
        }
    }
}

// ---- exclusive scan of uint32 (in place allowed), block sums to `sums` ----
__global__ void __launch_bounds__(SCAN_T) scan_block_kernel(const uint32_t* in, uint32_t* out, uint32_t* sums, int64_t n) {
    __shared__ uint32_t ws[SCAN_T / 32];
    int64_t base = (int64_t)blockIdx.x * SCAN_B + (int64_t)threadIdx.x * SCAN_I;
    uint32_t v[SCAN_I], tot = 0;
#pragma unroll
    for (int i = 0; i < SCAN_I; i++) { v[i] = base + i < n ? in[base + i] : 0u; tot += v[i]; }
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t incl = tot;
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) ws[w] = incl;
    __syncthreads();
    uint32_t pre = 0, all = 0;
    for (int i = 0; i < SCAN_T / 32; i++) { if (i < w) pre += ws[i]; all += ws[i]; }
    uint32_t run = pre + incl - tot;
#pragma unroll
    for (int i = 0; i < SCAN_I; i++) { if (base + i < n) out[base + i] = run; run += v[i]; }
    if (threadIdx.x == 0) sums[blockIdx.x] = all;
}
__global__ void scan_add_kernel(uint32_t* __restrict__ out, const uint32_t* __restrict__ block_pre, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] += block_pre[i / SCAN_B];
}
// scratch: >= scan_scratch_items(n) uint32.  total (device, int64) receives the sum.  Recursion depth <= 3 for n < 2^32.
int64_t scan_scratch_items(int64_t n) {
    int64_t t = 0;
    while (true) { int64_t nb = (n + SCAN_B - 1) / SCAN_B; if (nb < 1) nb = 1; t += nb + 1; if (nb == 1) break; n = nb; }
    return t;
}
__global__ void scan_total_kernel(const uint32_t* sum0, int64_t* total) { *total = (int64_t)sum0[0]; }
cudaError_t exclusive_scan(const uint32_t* in, uint32_t* out, int64_t n, uint32_t* scratch, int64_t* total, cudaStream_t st) {
    int64_t nb = (n + SCAN_B - 1) / SCAN_B;
    if (nb < 1) nb = 1;
    scan_block_kernel<<<(unsigned)nb, SCAN_T, 0, st>>>(in, out, scratch, n);
    if (nb == 1) {
        if (total) scan_total_kernel<<<1, 1, 0, st>>>(scratch, total);
        return cudaGetLastError();
    }
    cudaError_t e = exclusive_scan(scratch, scratch, nb, scratch + nb + 1, total, st);
    if (e != cudaSuccess) return e;
    scan_add_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(out, scratch, n);
    return cudaGetLastError();
}

// ---- lattice table ----
struct K3 { int x, y, z; };
__device__ __forceinline__ bool keq(K3 a, K3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
__device__ __forceinline__ uint32_t khash(K3 k) {
    uint64_t h = (uint64_t)(uint32_t)k.x * 0x9E3779B97F4A7C15ull;
    h ^= ((uint64_t)(uint32_t)k.y + 0x7F4A7C15ull) * 0xC2B2AE3D27D4EB4Full;
    h ^= h >> 29;
    h += (uint64_t)(uint32_t)k.z * 0x165667B19E3779F9ull;
    h ^= h >> 32; h *= 0xD6E8FEB86659FD93ull; h ^= h >> 32;
    return (uint32_t)h;
}
struct VertKey {                                   // lattice cell of a soup vertex (precomputed)
    const int* keys;
    __device__ __forceinline__ K3 operator()(uint32_t v) const { return K3{keys[3 * (int64_t)v], keys[3 * (int64_t)v + 1], keys[3 * (int64_t)v + 2]}; }
};
struct FaceKey {                                   // sorted vertex triple of a soup triangle
    const uint32_t* lookup;
    __device__ __forceinline__ K3 operator()(uint32_t t) const {
        uint32_t a = lookup[3 * (int64_t)t], b = lookup[3 * (int64_t)t + 1], c = lookup[3 * (int64_t)t + 2], s;
        if (a > b) { s = a; a = b; b = s; }
        if (b > c) { s = b; b = c; c = s; }
        if (a > b) { s = a; a = b; b = s; }
        return K3{(int)a, (int)b, (int)c};
    }
};
template <class KeyOf>
__device__ __forceinline__ uint32_t table_insert(uint32_t* owner, uint32_t* first, uint32_t mask, K3 key, uint32_t idx, KeyOf keyof) {
    uint32_t s = khash(key) & mask;
    while (true) {
        uint32_t o = *(volatile uint32_t*)(owner + s);
        if (o == EMPTY) { o = atomicCAS(owner + s, EMPTY, idx); if (o == EMPTY) o = idx; }
        if (o == idx || keq(keyof(o), key)) { atomicMin(first + s, idx); return s; }
        s = (s + 1) & mask;
    }
}
template <class KeyOf>
__device__ __forceinline__ uint32_t table_find(const uint32_t* __restrict__ owner, uint32_t mask, K3 key, KeyOf keyof) {
    uint32_t s = khash(key) & mask;
    while (true) {
        uint32_t o = owner[s];
        if (o == EMPTY) return EMPTY;
        if (keq(keyof(o), key)) return s;
        s = (s + 1) & mask;
    }
}

__device__ __forceinline__ int lattice_coord(float v) {                 // (int)(v / 1e-5f + 0.5f * sgn(v)), :351
    int sg = (0.0f < v) - (v < 0.0f);
    return __float2int_rz(__fadd_rn(__fdiv_rn(v, 0.00001f), __fmul_rn(0.5f, (float)sg)));
}
__global__ void mc_keys_kernel(const float* __restrict__ soup, int* __restrict__ keys, int64_t n3) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n3) keys[i] = lattice_coord(soup[i]);
}
__global__ void mc_insert_kernel(const int* __restrict__ keys, uint32_t* owner, uint32_t* first, uint32_t* __restrict__ slot_of, uint32_t mask,
                                 int64_t nv) {
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    VertKey ko{keys};
    slot_of[v] = table_insert(owner, first, mask, ko((uint32_t)v), (uint32_t)v, ko);
}
// dense list of the occupied slots (one per lattice cell, appended by the vertex that touched the cell first; order is irrelevant)
__global__ void mc_entries_kernel(const uint32_t* __restrict__ slot_of, const uint32_t* __restrict__ first, uint32_t* list, uint32_t* count,
                                  int64_t nv) {
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    uint32_t s = slot_of[v];
    if (first[s] == (uint32_t)v) list[atomicAdd(count, 1u)] = s;
}

enum : uint8_t { ST_UNDECIDED = 0, ST_REP = 1, ST_NONREP = 2 };
// One round of the fixed point over the slots of list_in; still-undecided slots are appended to list_out.
__global__ void mc_resolve_kernel(const int* __restrict__ keys, const uint32_t* __restrict__ owner, const uint32_t* __restrict__ first,
                                  uint8_t* status, uint32_t mask, const uint32_t* __restrict__ list_in, const uint32_t* n_in,
                                  uint32_t* list_out, uint32_t* n_out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)*n_in) return;
    uint32_t s = list_in[i];
    uint32_t o = owner[s];
    if (o == EMPTY) return;
    VertKey ko{keys};
    K3 k = ko(o);
    uint32_t f = first[s];
    // all 26 home slots are read first (independent loads in flight); an empty home slot means the neighbour cell does not exist
    // (linear probing, no deletions), which is the case for almost every probe
    uint32_t home[27];
#pragma unroll
    for (int q = 0; q < 27; q++)
        home[q] = q == 13 ? EMPTY : owner[khash(K3{k.x + q / 9 - 1, k.y + (q / 3) % 3 - 1, k.z + q % 3 - 1}) & mask];
    bool pending = false, dominated = false;
#pragma unroll
    for (int q = 0; q < 27; q++) {
        if (home[q] == EMPTY || dominated) continue;
        uint32_t n = table_find(owner, mask, K3{k.x + q / 9 - 1, k.y + (q / 3) % 3 - 1, k.z + q % 3 - 1}, ko);
        if (n == EMPTY || first[n] > f) continue;
        uint8_t st = *(volatile uint8_t*)(status + n);
        if (st == ST_REP) dominated = true;
        else if (st == ST_UNDECIDED) pending = true;
    }
    if (dominated) status[s] = ST_NONREP;
    else if (!pending) status[s] = ST_REP;
    else list_out[atomicAdd(n_out, 1u)] = s;
}
// flag[v] = 1 iff v opened a representative cell (the reference's new_verts.push_back, :360)
__global__ void mc_repflag_kernel(const uint32_t* __restrict__ entries, const uint32_t* n_entries, const uint32_t* __restrict__ first,
                                  const uint8_t* __restrict__ status, uint32_t* flag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)*n_entries) return;
    uint32_t s = entries[i];
    if (status[s] == ST_REP) flag[first[s]] = 1u;
}
__global__ void mc_lookup_kernel(const float* __restrict__ soup, const int* __restrict__ keys, const uint32_t* __restrict__ owner,
                                 const uint32_t* __restrict__ first, const uint8_t* __restrict__ status, const uint32_t* __restrict__ rep_id,
                                 uint32_t* lookup, float* __restrict__ verts, uint32_t mask, int64_t nv) {
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    VertKey ko{keys};
    K3 k = ko((uint32_t)v);
    uint32_t hit = EMPTY;
    {   // own cell first: if it is a representative it is the only one in the 27-neighbourhood (two adjacent cells cannot both
        // be representatives), so the probe order does not matter -- the common case costs one probe instead of up to 27
        uint32_t n = lookup[v];                                  // on entry lookup[v] holds the slot of v's own cell (mc_insert_kernel)
        if (status[n] == ST_REP) hit = first[n];
    }
    for (int di = -1; di <= 1 && hit == EMPTY; di++) for (int dj = -1; dj <= 1 && hit == EMPTY; dj++) for (int dk = -1; dk <= 1; dk++) {
        uint32_t n = table_find(owner, mask, K3{k.x + di, k.y + dj, k.z + dk}, ko);
        if (n == EMPTY || status[n] != ST_REP) continue;
        uint32_t f = first[n];
        if (f <= (uint32_t)v) { hit = f; break; }       // inserted before v (or v itself opened it: only possible for its own cell)
    }
    if (hit == EMPTY) { lookup[v] = 0u; return; }          // unreachable: a non-representative cell has an earlier representative neighbour
    uint32_t id = rep_id[hit];
    lookup[v] = id;
    if (hit == (uint32_t)v) { verts[3 * (int64_t)id] = soup[3 * v]; verts[3 * (int64_t)id + 1] = soup[3 * v + 1]; verts[3 * (int64_t)id + 2] = soup[3 * v + 2]; }
}

__device__ __forceinline__ bool face_degenerate(const uint32_t* lookup, int64_t t) {
    uint32_t a = lookup[3 * t], b = lookup[3 * t + 1], c = lookup[3 * t + 2];
    return a == b || a == c || b == c;
}
__global__ void mc_face_insert_kernel(const uint32_t* __restrict__ lookup, uint32_t* owner, uint32_t* first, uint32_t mask, int64_t nt) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nt || face_degenerate(lookup, t)) return;
    FaceKey ko{lookup};
    table_insert(owner, first, mask, ko((uint32_t)t), (uint32_t)t, ko);
}
__global__ void mc_face_keep_kernel(const uint32_t* __restrict__ lookup, const uint32_t* __restrict__ owner, const uint32_t* __restrict__ first,
                                    uint32_t* keep, uint32_t mask, int64_t nt) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nt) return;
    uint32_t k = 0;
    if (!face_degenerate(lookup, t)) {
        FaceKey ko{lookup};
        uint32_t s = table_find(owner, mask, ko((uint32_t)t), ko);
        k = (s != EMPTY && first[s] == (uint32_t)t) ? 1u : 0u;
    }
    keep[t] = k;
}
__global__ void mc_face_write_kernel(const uint32_t* __restrict__ lookup, const uint32_t* __restrict__ keep_off, const int64_t* n_faces,
                                     uint32_t* __restrict__ faces, int64_t nt) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nt) return;
    int64_t o = keep_off[t], nxt = t + 1 < nt ? (int64_t)keep_off[t + 1] : *n_faces;
    if (nxt == o) return;
    faces[3 * o] = lookup[3 * t]; faces[3 * o + 1] = lookup[3 * t + 1]; faces[3 * o + 2] = lookup[3 * t + 2];
}

inline int64_t align256(int64_t b) { return (b + 255) / 256 * 256; }
// > 1.25 x the number of keys that can ever be inserted (keeps linear probing finite even if every key were distinct; a marching-cubes
// soup shares each vertex between ~6 triangles and each face is unique, so the usual load is 0.15-0.6)
inline uint32_t table_capacity(int64_t n) { uint64_t c = 1024; while (c < (uint64_t)n + (uint64_t)n / 4 + 1) c <<= 1; return (uint32_t)c; }

struct CountLayout {
    int64_t PL, mx, cx, bpp, n_blocks;          // plane stride, node planes, cell planes, blocks per plane
    int64_t off_node, off_ntri, off_btris, off_boff, off_active, off_nactive, off_scan, total;
    bool any;
    CountLayout(int64_t nx, int64_t ny, int64_t nz) {
        any = nx > 2 && ny > 2 && nz > 2;
        PL = ny * nz; mx = nx - 1; cx = nx - 2;
        bpp = (PL + CB * 4 - 1) / (CB * 4);
        n_blocks = any ? cx * bpp : 0;
        int64_t o = 0;
        off_node = o; o += any ? align256(mx * PL * 4 + 64) : 0;
        off_ntri = o; o += any ? align256(cx * PL + 64) : 0;
        off_btris = o; o += align256((n_blocks + 1) * 4);
        off_boff = o; o += align256((n_blocks + 1) * 4);
        off_active = o; o += align256((n_blocks + 1) * 4);
        off_nactive = o; o += 256;
        off_scan = o; o += align256(scan_scratch_items(n_blocks) * 4);
        total = o + 256;
    }
};
struct MeshLayout {
    int64_t nt, nv; uint32_t cap, capf;
    int64_t off_soup, off_keys, off_owner, off_first, off_status, off_flag, off_lookup, off_la, off_lb, off_lc, off_cnt, off_fowner,
        off_ffirst, off_keep, off_scan, total;
    explicit MeshLayout(int64_t n_tris) {
        nt = n_tris; nv = 3 * n_tris; cap = table_capacity(nv); capf = table_capacity(nt);
        int64_t o = 0;
        off_soup = o; o += align256(nv * 12);
        off_keys = o; o += align256(nv * 12);
        off_owner = o; o += align256((int64_t)cap * 4);
        off_first = o; o += align256((int64_t)cap * 4);
        off_status = o; o += align256((int64_t)cap);
        off_flag = o; o += align256((nv + 1) * 4);
        off_lookup = o; o += align256((nv + 1) * 4);
        off_la = o; o += align256((nv + 1) * 4);
        off_lb = o; o += align256((nv + 1) * 4);
        off_lc = o; o += align256((nv + 1) * 4);
        off_cnt = o; o += 256;
        off_fowner = o; o += align256((int64_t)capf * 4);
        off_ffirst = o; o += align256((int64_t)capf * 4);
        off_keep = o; o += align256((nt + 1) * 4);
        off_scan = o; o += align256(scan_scratch_items(nv > 0 ? nv : 1) * 4);
        total = o + 256;
    }
};

}  // namespace

MF_API int64_t mf_mcubes_count_workspace_size(int64_t nx, int64_t ny, int64_t nz) {
    if (nx < 0 || ny < 0 || nz < 0) return 0;
    return CountLayout(nx, ny, nz).total;
}

MF_API int mf_mcubes_count(const float* volume, int64_t nx, int64_t ny, int64_t nz, float isovalue, float truncation, void* workspace,
                           int64_t* n_tris, void* stream) {
    MF_CHECK_ARG(n_tris != nullptr && workspace != nullptr);
    MF_CHECK_ARG(nx >= 0 && ny >= 0 && nz >= 0 && nx < 2048 && ny < 2048 && nz < 2048);
    MF_CHECK_ARG(truncation == truncation && truncation < INFINITY && isovalue == isovalue);
    cudaStream_t st = (cudaStream_t)stream;
    CountLayout L(nx, ny, nz);
    if (!L.any) { MF_CUDA(cudaMemsetAsync(n_tris, 0, sizeof(int64_t), st)); return MF_OK; }
    MF_CHECK_ARG(volume != nullptr);
    char* ws = (char*)workspace;
    float* node = (float*)(ws + L.off_node);
    uint8_t* ntri = (uint8_t*)(ws + L.off_ntri);
    uint32_t* btris = (uint32_t*)(ws + L.off_btris);
    uint32_t* boff = (uint32_t*)(ws + L.off_boff);
    const bool vec = nz % 4 == 0 && ((uintptr_t)volume % 16) == 0 && ((uintptr_t)ws % 16) == 0;
    dim3 gn((unsigned)L.bpp, (unsigned)L.mx), gc((unsigned)L.bpp, (unsigned)L.cx);
    mf_ktimer_begin(0, st);
    if (vec) mc_nodes_kernel<true><<<gn, 256, 0, st>>>(volume, node, (int)L.PL, (int)ny, (int)nz, truncation);
    else mc_nodes_kernel<false><<<gn, 256, 0, st>>>(volume, node, (int)L.PL, (int)ny, (int)nz, truncation);
    mf_ktimer_end(0, st);
    MF_LAUNCH_CHECK();
    mf_ktimer_begin(1, st);
    if (vec) mc_classify_kernel<true><<<gc, CB, 0, st>>>(node, ntri, btris, (int)L.PL, (int)ny, (int)nz, isovalue);
    else mc_classify_kernel<false><<<gc, CB, 0, st>>>(node, ntri, btris, (int)L.PL, (int)ny, (int)nz, isovalue);
    mf_ktimer_end(1, st);
    MF_LAUNCH_CHECK();
    MF_CUDA(cudaMemsetAsync(ws + L.off_nactive, 0, 256, st));
    mc_active_blocks_kernel<<<(unsigned)((L.n_blocks + 255) / 256), 256, 0, st>>>(btris, (uint32_t*)(ws + L.off_active), (uint32_t*)(ws + L.off_nactive),
                                                                               L.n_blocks);
    MF_LAUNCH_CHECK();
    MF_CUDA(exclusive_scan(btris, boff, L.n_blocks, (uint32_t*)(ws + L.off_scan), n_tris, st));
    return MF_OK;
}

MF_API int64_t mf_mcubes_mesh_workspace_size(int64_t n_tris) {
    if (n_tris < 0 || n_tris > 1400000000ll) return 0;
    return MeshLayout(n_tris).total;
}

MF_API int mf_mcubes_mesh(const void* count_workspace, int64_t nx, int64_t ny, int64_t nz, float isovalue, int64_t n_tris,
                          void* mesh_workspace, float* verts, uint32_t* faces, int64_t* counts, void* stream) {
    MF_CHECK_ARG(counts != nullptr);
    MF_CHECK_ARG(n_tris >= 0 && n_tris <= 1400000000ll);
    cudaStream_t st = (cudaStream_t)stream;
    if (n_tris == 0) { MF_CUDA(cudaMemsetAsync(counts, 0, 3 * sizeof(int64_t), st)); return MF_OK; }
    MF_CHECK_ARG(count_workspace != nullptr && mesh_workspace != nullptr && verts != nullptr && faces != nullptr);
    CountLayout C(nx, ny, nz);
    MF_CHECK_ARG(C.any);
    MeshLayout M(n_tris);
    const char* cw = (const char*)count_workspace;
    char* ws = (char*)mesh_workspace;
    const float* node = (const float*)(cw + C.off_node);
    const uint8_t* ntri = (const uint8_t*)(cw + C.off_ntri);
    const uint32_t* boff = (const uint32_t*)(cw + C.off_boff);
    const uint32_t* active = (const uint32_t*)(cw + C.off_active);
    const uint32_t* n_active = (const uint32_t*)(cw + C.off_nactive);
    float* soup = (float*)(ws + M.off_soup);
    int* keys = (int*)(ws + M.off_keys);
    uint32_t *owner = (uint32_t*)(ws + M.off_owner), *first = (uint32_t*)(ws + M.off_first);
    uint8_t* status = (uint8_t*)(ws + M.off_status);
    uint32_t *flag = (uint32_t*)(ws + M.off_flag), *lookup = (uint32_t*)(ws + M.off_lookup);
    uint32_t* entries = (uint32_t*)(ws + M.off_la);
    uint32_t* lists[2] = {(uint32_t*)(ws + M.off_lb), (uint32_t*)(ws + M.off_lc)};
    uint32_t* cnt = (uint32_t*)(ws + M.off_cnt);                 // [0]: lattice cells, [1], [2]: undecided lists (ping-pong)
    uint32_t *fowner = (uint32_t*)(ws + M.off_fowner), *ffirst = (uint32_t*)(ws + M.off_ffirst), *keep = (uint32_t*)(ws + M.off_keep);
    uint32_t* scan = (uint32_t*)(ws + M.off_scan);
    const uint32_t mask = M.cap - 1, fmask = M.capf - 1;
    const int T = 256;
    auto nb = [](int64_t n) { return (unsigned)((n + 255) / 256); };

    const bool vec = C.PL % 16 == 0 && ((uintptr_t)cw % 16) == 0;                     // 16-byte reads of the per-cell counts
    int64_t emit_ctas = (C.n_blocks + EMIT_WARPS - 1) / EMIT_WARPS;
    if (emit_ctas > (int64_t)mf_sm_count_cached() * 4) emit_ctas = (int64_t)mf_sm_count_cached() * 4;
    mc_emit_kernel<<<(unsigned)emit_ctas, EMIT_WARPS * 32, 0, st>>>(node, ntri, boff, active, n_active, soup, (int)C.PL, (int)nz, (int)C.bpp,
                                                                   isovalue, vec ? 1 : 0);
    MF_LAUNCH_CHECK();
    mc_keys_kernel<<<nb(M.nv * 3), T, 0, st>>>(soup, keys, M.nv * 3);
    MF_CUDA(cudaMemsetAsync(owner, 0xFF, (size_t)M.cap * 4, st));
    MF_CUDA(cudaMemsetAsync(first, 0xFF, (size_t)M.cap * 4, st));
    MF_CUDA(cudaMemsetAsync(status, 0, (size_t)M.cap, st));
    MF_CUDA(cudaMemsetAsync(flag, 0, (size_t)(M.nv + 1) * 4, st));
    MF_CUDA(cudaMemsetAsync(cnt, 0, 256, st));
    mc_insert_kernel<<<nb(M.nv), T, 0, st>>>(keys, owner, first, lookup, mask, M.nv);
    MF_LAUNCH_CHECK();
    mc_entries_kernel<<<nb(M.nv), T, 0, st>>>(lookup, first, entries, cnt, M.nv);
    MF_LAUNCH_CHECK();
    // fixed point of the greedy clustering; the host reads the number of still-undecided cells after every round.  Round 1 runs
    // over all lattice cells (their number is only known on the device: the grid covers the upper bound nv, surplus warps exit).
    int64_t rounds = 1;
    uint32_t pending = 0;
    mc_resolve_kernel<<<nb(M.nv), T, 0, st>>>(keys, owner, first, status, mask, entries, cnt, lists[0], cnt + 1);
    MF_LAUNCH_CHECK();
    MF_CUDA(cudaMemcpyAsync(&pending, cnt + 1, 4, cudaMemcpyDeviceToHost, st));
    MF_CUDA(cudaStreamSynchronize(st));
    int cur = 0;
    while (pending > 0) {
        if (rounds > 100000) { mf_set_error("mf_mcubes_mesh: clustering did not converge"); return MF_ERR_INVALID; }
        MF_CUDA(cudaMemsetAsync(cnt + 1 + (cur ^ 1), 0, 4, st));
        mc_resolve_kernel<<<nb(pending), T, 0, st>>>(keys, owner, first, status, mask, lists[cur], cnt + 1 + cur, lists[cur ^ 1], cnt + 1 + (cur ^ 1));
        MF_LAUNCH_CHECK();
        cur ^= 1;
        MF_CUDA(cudaMemcpyAsync(&pending, cnt + 1 + cur, 4, cudaMemcpyDeviceToHost, st));
        MF_CUDA(cudaStreamSynchronize(st));
        rounds++;
    }
    mc_repflag_kernel<<<nb(M.nv), T, 0, st>>>(entries, cnt, first, status, flag);
    MF_LAUNCH_CHECK();
    MF_CUDA(exclusive_scan(flag, flag, M.nv, scan, counts, st));                       // counts[0] = vertices
    mc_lookup_kernel<<<nb(M.nv), T, 0, st>>>(soup, keys, owner, first, status, flag, lookup, verts, mask, M.nv);
    MF_LAUNCH_CHECK();
    MF_CUDA(cudaMemsetAsync(fowner, 0xFF, (size_t)M.capf * 4, st));
    MF_CUDA(cudaMemsetAsync(ffirst, 0xFF, (size_t)M.capf * 4, st));
    mc_face_insert_kernel<<<nb(M.nt), T, 0, st>>>(lookup, fowner, ffirst, fmask, M.nt);
    MF_LAUNCH_CHECK();
    mc_face_keep_kernel<<<nb(M.nt), T, 0, st>>>(lookup, fowner, ffirst, keep, fmask, M.nt);
    MF_LAUNCH_CHECK();
    MF_CUDA(exclusive_scan(keep, keep, M.nt, scan, counts + 1, st));                   // counts[1] = faces
    mc_face_write_kernel<<<nb(M.nt), T, 0, st>>>(lookup, keep, counts + 1, faces, M.nt);
    MF_LAUNCH_CHECK();
    MF_CUDA(cudaMemcpyAsync(counts + 2, &rounds, sizeof(int64_t), cudaMemcpyHostToDevice, st));
    MF_CUDA(cudaStreamSynchronize(st));                                                 // (`rounds` is a stack variable)
    return MF_OK;
}

// ---- mesh visibility filter of the Mesher (model/Mesher.py:221-281; helper_functions/geometry_helper.py:216-222) ----
// A vertex is seen if, for some keyframe, it projects more than `edge` pixels inside the image, lies in front of the camera
// (camera z < 0) and nearer than the keyframe's largest stored depth.  fp32 in the reference's operation order (products, then
// left-to-right sums, no contraction): cam = (p0 R00 + p1 R01) + p2 R02 + t; uv_h = K (-cam_x, cam_y, cam_z); uv = uv_h / (z_h + 1e-5).
namespace {
__global__ void mc_seen_mask_kernel(const float* __restrict__ pts, int64_t n, const float* __restrict__ w2c, const float* __restrict__ max_depth,
                                    int k, const float* __restrict__ Kmat, float u_hi, float v_hi, float lo, uint8_t* __restrict__ seen) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float p0 = pts[3 * i], p1 = pts[3 * i + 1], p2 = pts[3 * i + 2];
    float K9[9];
#pragma unroll
    for (int q = 0; q < 9; q++) K9[q] = __ldg(Kmat + q);
    bool s = false;
    for (int j = 0; j < k && !s; j++) {
        const float* m = w2c + 12 * j;                            // rows of [R | t]
        float c[3];
#pragma unroll
        for (int r = 0; r < 3; r++)
            c[r] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(p0, __ldg(m + 4 * r)), __fmul_rn(p1, __ldg(m + 4 * r + 1))), __fmul_rn(p2, __ldg(m + 4 * r + 2))),
                             __ldg(m + 4 * r + 3));
        const float xn = -c[0];
        float h[3];
#pragma unroll
        for (int r = 0; r < 3; r++)
            h[r] = __fadd_rn(__fadd_rn(__fmul_rn(K9[3 * r], xn), __fmul_rn(K9[3 * r + 1], c[1])), __fmul_rn(K9[3 * r + 2], c[2]));
        const float zc = __fadd_rn(h[2], 1e-5f);
        const float u = __fdiv_rn(h[0], zc), v = __fdiv_rn(h[1], zc);
        const bool m1 = (u < u_hi) && (u > lo) && (v < v_hi) && (v > lo) && (c[2] < 0.0f);
        const float cz = fabsf(c[2]);
        s = m1 && (cz > 0.0f) && (cz < __ldg(max_depth + j));
    }
    seen[i] = s ? 1 : 0;
}
__global__ void mc_face_seen_kernel(const uint8_t* __restrict__ seen, const int64_t* __restrict__ faces, int64_t nf, uint8_t* __restrict__ keep) {
    int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nf) return;
    keep[f] = (seen[faces[3 * f]] | seen[faces[3 * f + 1]] | seen[faces[3 * f + 2]]) ? 1 : 0;     // dropped only if ALL three are unseen
}
}  // namespace

MF_API int mf_mesh_seen_mask(const float* points, int64_t n, const float* w2c, const float* max_depth, int k, const float* K, int img_w,
                             int img_h, int edge, uint8_t* seen, void* stream) {
    MF_CHECK_ARG(n >= 0 && k >= 0 && seen != nullptr);
    if (n == 0) return MF_OK;
    MF_CHECK_ARG(points != nullptr && K != nullptr && (k == 0 || (w2c != nullptr && max_depth != nullptr)));
    mc_seen_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(points, n, w2c, max_depth, k, K, (float)(img_w - edge),
                                                                                      (float)(img_h - edge), (float)edge, seen);
    MF_LAUNCH_CHECK();
    return MF_OK;
}

MF_API int mf_mesh_face_mask(const uint8_t* seen, const int64_t* faces, int64_t n_faces, uint8_t* keep, void* stream) {
    MF_CHECK_ARG(n_faces >= 0);
    if (n_faces == 0) return MF_OK;
    MF_CHECK_ARG(seen != nullptr && faces != nullptr && keep != nullptr);
    mc_face_seen_kernel<<<(unsigned)((n_faces + 255) / 256), 256, 0, (cudaStream_t)stream>>>(seen, faces, n_faces, keep);
    MF_LAUNCH_CHECK();
    return MF_OK;
}

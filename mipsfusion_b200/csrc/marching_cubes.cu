// N2: marching cubes on a dense SDF volume, on the device, bit-identical to the reference's NumpyMarchingCubes
// (/root/reference/external/NumpyMarchingCubes/marching_cubes/src/marching_cubes.cpp, called from utils/utils.py:78,159).
//
// The reference is a sequential scan: per integer cell centre it tri-linearly samples 8 corners at +-0.5 (each the mean of 8
// voxels, :93-113), classifies (:160-167), emits triangles in table order (:219-227), then clusters the triangle-soup vertices
// greedily on a 1e-5 lattice in soup order (:316-390) and drops degenerate and duplicate faces (:246-288).  The sequential
// parts are restated as order-free fixed points so that every step is a flat kernel:
//   mc_nodes_kernel    corner ("dual node") values once per node instead of 8 x per cell; NaN marks an invalid node
//   mc_classify_kernel triangle count per cell (u8) + per-256-cell block totals           } offsets = exclusive scan, which
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
#include "mc_tables.h"

namespace {

constexpr uint32_t EMPTY = 0xFFFFFFFFu;
constexpr int CB = 256;                 // cells per block in classify / emit
constexpr int SCAN_T = 512, SCAN_I = 8, SCAN_B = SCAN_T * SCAN_I;   // exclusive scan: 4096 items per block

__device__ __forceinline__ bool voxel_valid(float d, float trunc) { return d != -INFINITY && fabsf(d) < trunc; }

// ---- dual nodes: node (a,b,c) = trilerp at (a+.5, b+.5, c+.5): all weights are 0.5, summed in the reference's corner order ----
__global__ void __launch_bounds__(256) mc_nodes_kernel(const float* __restrict__ vol, float* __restrict__ node, int ny, int nz,
                                                       int my, int mz, int64_t n_nodes, float trunc) {
    int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n_nodes) return;
    int c = (int)(id % mz);
    int64_t r = id / mz;
    int b = (int)(r % my), a = (int)(r / my);
    const float* p = vol + ((int64_t)a * ny + b) * nz + c;
    int64_t sx = (int64_t)ny * nz, sy = nz;
    float v[8] = {__ldg(p), __ldg(p + sx), __ldg(p + sy), __ldg(p + 1), __ldg(p + sx + sy), __ldg(p + sy + 1), __ldg(p + sx + 1),
                  __ldg(p + sx + sy + 1)};
    bool ok = true;
    float dist = 0.0f;
#pragma unroll
    for (int q = 0; q < 8; q++) {
        ok = ok && voxel_valid(v[q], trunc);
        dist = __fadd_rn(dist, __fmul_rn(0.125f, v[q]));          // (0.5 * 0.5 * 0.5) * d, exact product
    }
    node[id] = ok ? dist : __int_as_float(0x7fc00000);
}

struct Cell { float d[8]; int cube; bool ok; };
// corner numbering 0:000 1:100 2:010 3:001 4:110 5:011 6:101 7:111
__device__ __forceinline__ Cell load_cell(const float* __restrict__ node, int a, int b, int c, int my, int mz, float iso) {
    Cell ce;
    const float* p = node + ((int64_t)a * my + b) * mz + c;
    int64_t sx = (int64_t)my * mz, sy = mz;
    ce.d[0] = __ldg(p); ce.d[1] = __ldg(p + sx); ce.d[2] = __ldg(p + sy); ce.d[3] = __ldg(p + 1);
    ce.d[4] = __ldg(p + sx + sy); ce.d[5] = __ldg(p + sy + 1); ce.d[6] = __ldg(p + sx + 1); ce.d[7] = __ldg(p + sx + sy + 1);
    ce.ok = true;
    const int bit[8] = {8, 4, 1, 128, 2, 16, 64, 32};
    ce.cube = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) {
        ce.ok = ce.ok && (ce.d[q] == ce.d[q]);
        if (ce.d[q] < iso) ce.cube += bit[q];
    }
    return ce;
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
__device__ __forceinline__ int case_edges_and_count(int cube, int* ntri) {
    uint64_t row = MC_TRI_PACKED[cube];
    int m = 0, n = 0;
    for (int i = 0; i < 16; i++) {
        int e = (int)((row >> (4 * i)) & 15);
        if (e == 15) break;
        m |= 1 << e; n++;
    }
    *ntri = n / 3;
    return m;
}
__device__ __forceinline__ int cell_triangles(const Cell& ce) {
    if (!ce.ok || ce.cube == 0 || ce.cube == 255) return 0;
    int ntri;
    int em = case_edges_and_count(ce.cube, &ntri);
    if (em == 255) return 0;                                     // :194
    if (!cell_within_thresh(ce.d)) return 0;
    return ntri;
}
__device__ __forceinline__ void cell_coords(int64_t id, int cy, int cz, int& a, int& b, int& c) {
    c = (int)(id % cz);
    int64_t r = id / cz;
    b = (int)(r % cy); a = (int)(r / cy);
}

__global__ void __launch_bounds__(CB) mc_classify_kernel(const float* __restrict__ node, uint8_t* __restrict__ ntri_cell,
                                                         uint32_t* __restrict__ block_tris, int cy, int cz, int my, int mz,
                                                         int64_t n_cells, float iso) {
    int64_t id = (int64_t)blockIdx.x * CB + threadIdx.x;
    int n = 0;
    if (id < n_cells) {
        int a, b, c;
        cell_coords(id, cy, cz, a, b, c);
        Cell ce = load_cell(node, a, b, c, my, mz, iso);
        n = cell_triangles(ce);
        ntri_cell[id] = (uint8_t)n;
    }
    __shared__ int ws[CB / 32];
    int s = n;
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < CB / 32; w++) t += ws[w];
        block_tris[blockIdx.x] = (uint32_t)t;
    }
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

__global__ void __launch_bounds__(CB) mc_emit_kernel(const float* __restrict__ node, const uint8_t* __restrict__ ntri_cell,
                                                     const uint32_t* __restrict__ block_off, float* __restrict__ soup, int cy, int cz,
                                                     int my, int mz, int64_t n_cells, float iso) {
    int64_t id = (int64_t)blockIdx.x * CB + threadIdx.x;
    int n = id < n_cells ? (int)ntri_cell[id] : 0;
    // block-exclusive scan of n
    __shared__ int ws[CB / 32];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, incl = n;
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) ws[w] = incl;
    __syncthreads();
    int base = 0;
    for (int i = 0; i < w; i++) base += ws[i];
    if (n == 0) return;
    int64_t tri0 = (int64_t)block_off[blockIdx.x] + base + incl - n;
    int a, b, c;
    cell_coords(id, cy, cz, a, b, c);
    Cell ce = load_cell(node, a, b, c, my, mz, iso);
    // cell centre is (a+1, b+1, c+1); corner q at centre +- 0.5
    // corner q has +0.5 along x iff bit q of 0xD2, y: 0xB4, z: 0xE8; edge e joins corners (E1 >> 4e) & 15 and (E2 >> 4e) & 15:
    // {2,4},{4,1},{1,0},{0,2},{5,7},{7,6},{6,3},{3,5},{2,5},{4,7},{1,6},{0,3} (:205-216)
    const uint64_t E1 = 0x014236750142ull, E2 = 0x367553672014ull;
    float fx = (float)(a + 1), fy = (float)(b + 1), fz = (float)(c + 1);
    uint64_t row = MC_TRI_PACKED[ce.cube];
    float* out = soup + tri0 * 9;
    for (int t = 0; t < n * 3; t++) {
        int e = (int)((row >> (4 * t)) & 15);
        int q1 = (int)((E1 >> (4 * e)) & 15), q2 = (int)((E2 >> (4 * e)) & 15);
        F3 p1 = {__fadd_rn(fx, (0xD2 >> q1) & 1 ? 0.5f : -0.5f), __fadd_rn(fy, (0xB4 >> q1) & 1 ? 0.5f : -0.5f), __fadd_rn(fz, (0xE8 >> q1) & 1 ? 0.5f : -0.5f)};
        F3 p2 = {__fadd_rn(fx, (0xD2 >> q2) & 1 ? 0.5f : -0.5f), __fadd_rn(fy, (0xB4 >> q2) & 1 ? 0.5f : -0.5f), __fadd_rn(fz, (0xE8 >> q2) & 1 ? 0.5f : -0.5f)};
        F3 v = vertex_interp(iso, p1, p2, ce.d[q1], ce.d[q2]);
        out[3 * t] = v.x; out[3 * t + 1] = v.y; out[3 * t + 2] = v.z;
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
__device__ __forceinline__ void table_insert(uint32_t* owner, uint32_t* first, uint32_t mask, K3 key, uint32_t idx, KeyOf keyof) {
    uint32_t s = khash(key) & mask;
    while (true) {
        uint32_t o = *(volatile uint32_t*)(owner + s);
        if (o == EMPTY) { o = atomicCAS(owner + s, EMPTY, idx); if (o == EMPTY) o = idx; }
        if (o == idx || keq(keyof(o), key)) { atomicMin(first + s, idx); return; }
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
__global__ void mc_insert_kernel(const int* __restrict__ keys, uint32_t* owner, uint32_t* first, uint32_t mask, int64_t nv) {
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    VertKey ko{keys};
    table_insert(owner, first, mask, ko((uint32_t)v), (uint32_t)v, ko);
}

enum : uint8_t { ST_UNDECIDED = 0, ST_REP = 1, ST_NONREP = 2 };
// One round of the fixed point.  list_in == nullptr: every occupied slot; undecided slots are appended to list_out.
__global__ void mc_resolve_kernel(const int* __restrict__ keys, const uint32_t* __restrict__ owner, const uint32_t* __restrict__ first,
                                  uint8_t* status, uint32_t mask, const uint32_t* __restrict__ list_in, const uint32_t* n_in,
                                  uint32_t* list_out, uint32_t* n_out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t s;
    if (list_in) { if (i >= (int64_t)*n_in) return; s = list_in[i]; }
    else { if (i > (int64_t)mask) return; s = (uint32_t)i; }
    uint32_t o = owner[s];
    if (o == EMPTY) return;
    VertKey ko{keys};
    K3 k = ko(o);
    uint32_t f = first[s];
    bool pending = false, dominated = false;
    for (int di = -1; di <= 1 && !dominated; di++) for (int dj = -1; dj <= 1 && !dominated; dj++) for (int dk = -1; dk <= 1; dk++) {
        if (!(di | dj | dk)) continue;
        uint32_t n = table_find(owner, mask, K3{k.x + di, k.y + dj, k.z + dk}, ko);
        if (n == EMPTY || first[n] > f) continue;
        uint8_t st = *(volatile uint8_t*)(status + n);
        if (st == ST_REP) { dominated = true; break; }
        if (st == ST_UNDECIDED) pending = true;
    }
    if (dominated) status[s] = ST_NONREP;
    else if (!pending) status[s] = ST_REP;
    else list_out[atomicAdd(n_out, 1u)] = s;
}
// flag[v] = 1 iff v opened a representative cell (the reference's new_verts.push_back, :360)
__global__ void mc_repflag_kernel(const uint32_t* __restrict__ owner, const uint32_t* __restrict__ first, const uint8_t* __restrict__ status,
                                  uint32_t* flag, uint32_t mask) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s > (int64_t)mask || owner[s] == EMPTY) return;
    if (status[s] == ST_REP) flag[first[s]] = 1u;
}
__global__ void mc_lookup_kernel(const float* __restrict__ soup, const int* __restrict__ keys, const uint32_t* __restrict__ owner,
                                 const uint32_t* __restrict__ first, const uint8_t* __restrict__ status, const uint32_t* __restrict__ rep_id,
                                 uint32_t* __restrict__ lookup, float* __restrict__ verts, uint32_t mask, int64_t nv) {
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    VertKey ko{keys};
    K3 k = ko((uint32_t)v);
    uint32_t hit = EMPTY;
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
inline uint32_t table_capacity(int64_t n) { uint64_t c = 1024; while (c < (uint64_t)n * 2) c <<= 1; return (uint32_t)c; }

struct CountLayout {
    int64_t mx, my, mz, cx, cy, cz, n_nodes, n_cells, n_blocks;
    int64_t off_node, off_ntri, off_btris, off_boff, off_scan, total;
    CountLayout(int64_t nx, int64_t ny, int64_t nz) {
        mx = nx - 1; my = ny - 1; mz = nz - 1; cx = nx - 2; cy = ny - 2; cz = nz - 2;
        bool any = cx > 0 && cy > 0 && cz > 0;
        n_nodes = any ? mx * my * mz : 0; n_cells = any ? cx * cy * cz : 0;
        n_blocks = (n_cells + CB - 1) / CB;
        int64_t o = 0;
        off_node = o; o += align256(n_nodes * 4);
        off_ntri = o; o += align256(n_cells);
        off_btris = o; o += align256((n_blocks + 1) * 4);
        off_boff = o; o += align256((n_blocks + 1) * 4);
        off_scan = o; o += align256(scan_scratch_items(n_blocks) * 4);
        total = o + 256;
    }
};
struct MeshLayout {
    int64_t nt, nv; uint32_t cap, capf;
    int64_t off_soup, off_keys, off_owner, off_first, off_status, off_flag, off_lookup, off_la, off_lb, off_cnt, off_fowner, off_ffirst,
        off_keep, off_scan, total;
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
    if (L.n_cells == 0) { MF_CUDA(cudaMemsetAsync(n_tris, 0, sizeof(int64_t), st)); return MF_OK; }
    MF_CHECK_ARG(volume != nullptr);
    char* ws = (char*)workspace;
    float* node = (float*)(ws + L.off_node);
    uint8_t* ntri = (uint8_t*)(ws + L.off_ntri);
    uint32_t* btris = (uint32_t*)(ws + L.off_btris);
    uint32_t* boff = (uint32_t*)(ws + L.off_boff);
    mc_nodes_kernel<<<(unsigned)((L.n_nodes + 255) / 256), 256, 0, st>>>(volume, node, (int)ny, (int)nz, (int)L.my, (int)L.mz, L.n_nodes, truncation);
    MF_LAUNCH_CHECK();
    mf_ktimer_begin(0, st);
    mc_classify_kernel<<<(unsigned)L.n_blocks, CB, 0, st>>>(node, ntri, btris, (int)L.cy, (int)L.cz, (int)L.my, (int)L.mz, L.n_cells, isovalue);
    mf_ktimer_end(0, st);
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
    MeshLayout M(n_tris);
    const char* cw = (const char*)count_workspace;
    char* ws = (char*)mesh_workspace;
    const float* node = (const float*)(cw + C.off_node);
    const uint8_t* ntri = (const uint8_t*)(cw + C.off_ntri);
    const uint32_t* boff = (const uint32_t*)(cw + C.off_boff);
    float* soup = (float*)(ws + M.off_soup);
    int* keys = (int*)(ws + M.off_keys);
    uint32_t *owner = (uint32_t*)(ws + M.off_owner), *first = (uint32_t*)(ws + M.off_first);
    uint8_t* status = (uint8_t*)(ws + M.off_status);
    uint32_t *flag = (uint32_t*)(ws + M.off_flag), *lookup = (uint32_t*)(ws + M.off_lookup);
    uint32_t* lists[2] = {(uint32_t*)(ws + M.off_la), (uint32_t*)(ws + M.off_lb)};
    uint32_t* cnt = (uint32_t*)(ws + M.off_cnt);                 // [0], [1]: list lengths (ping-pong)
    uint32_t *fowner = (uint32_t*)(ws + M.off_fowner), *ffirst = (uint32_t*)(ws + M.off_ffirst), *keep = (uint32_t*)(ws + M.off_keep);
    uint32_t* scan = (uint32_t*)(ws + M.off_scan);
    const uint32_t mask = M.cap - 1, fmask = M.capf - 1;
    const int T = 256;
    auto nb = [](int64_t n) { return (unsigned)((n + 255) / 256); };

    mc_emit_kernel<<<(unsigned)C.n_blocks, CB, 0, st>>>(node, ntri, boff, soup, (int)C.cy, (int)C.cz, (int)C.my, (int)C.mz, C.n_cells, isovalue);
    MF_LAUNCH_CHECK();
    mc_keys_kernel<<<nb(M.nv * 3), T, 0, st>>>(soup, keys, M.nv * 3);
    MF_CUDA(cudaMemsetAsync(owner, 0xFF, (size_t)M.cap * 4, st));
    MF_CUDA(cudaMemsetAsync(first, 0xFF, (size_t)M.cap * 4, st));
    MF_CUDA(cudaMemsetAsync(status, 0, (size_t)M.cap, st));
    MF_CUDA(cudaMemsetAsync(flag, 0, (size_t)(M.nv + 1) * 4, st));
    MF_CUDA(cudaMemsetAsync(cnt, 0, 256, st));
    mc_insert_kernel<<<nb(M.nv), T, 0, st>>>(keys, owner, first, mask, M.nv);
    MF_LAUNCH_CHECK();
    // fixed point of the greedy clustering; the host reads the number of still-undecided cells after every round
    int64_t rounds = 0;
    uint32_t pending = 0;
    mc_resolve_kernel<<<nb((int64_t)M.cap), T, 0, st>>>(keys, owner, first, status, mask, nullptr, nullptr, lists[0], cnt);
    MF_LAUNCH_CHECK();
    MF_CUDA(cudaMemcpyAsync(&pending, cnt, 4, cudaMemcpyDeviceToHost, st));
    MF_CUDA(cudaStreamSynchronize(st));
    rounds = 1;
    int cur = 0;
    while (pending > 0) {
        if (rounds > 100000) { mf_set_error("mf_mcubes_mesh: clustering did not converge"); return MF_ERR_INVALID; }
        MF_CUDA(cudaMemsetAsync(cnt + (cur ^ 1), 0, 4, st));
        mc_resolve_kernel<<<nb(pending), T, 0, st>>>(keys, owner, first, status, mask, lists[cur], cnt + cur, lists[cur ^ 1], cnt + (cur ^ 1));
        MF_LAUNCH_CHECK();
        cur ^= 1;
        MF_CUDA(cudaMemcpyAsync(&pending, cnt + cur, 4, cudaMemcpyDeviceToHost, st));
        MF_CUDA(cudaStreamSynchronize(st));
        rounds++;
    }
    mc_repflag_kernel<<<nb((int64_t)M.cap), T, 0, st>>>(owner, first, status, flag, mask);
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

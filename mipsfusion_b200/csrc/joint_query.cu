// Joint multi-submap SDF / colour query with entropy- and distance-weighted blending (a14).
// Reference: model/Mesher.py:464-528 (geometry), :606-663 (colour); vis/math_helper.py:47-96;
// helper_functions/geometry_helper.py:93-99 (world -> submap frame).
#include "field_launch.cuh"
#include "field_tc_launch.cuh"

constexpr int64_t JQ_CHUNK = 16 << 20;        // points per internal chunk (bounds the compacted index list)

struct PointSetDev {
    const double* pts; const double* ax; const double* ay; const double* az;
    int nx, ny, nz;
    __device__ __forceinline__ void get(int64_t gi, double p[3]) const {
        if (pts) { p[0] = pts[gi * 3]; p[1] = pts[gi * 3 + 1]; p[2] = pts[gi * 3 + 2]; return; }
        const int iz = (int)(gi % nz); const int64_t t = gi / nz;          // np.meshgrid(x, y, z).ravel(): (iy, ix, iz)
        const int ix = (int)(t % nx); const int iy = (int)(t / nx);
        p[0] = ax[ix]; p[1] = ay[iy]; p[2] = az[iz];
    }
};

struct SubmapDev {
    float w2l[12]; double bmin[3], bmax[3]; float centroid[3];
    __device__ __forceinline__ bool contains(const double p[3]) const {    // open3d AABB test: inclusive, fp64
        return p[0] >= bmin[0] && p[0] <= bmax[0] && p[1] >= bmin[1] && p[1] <= bmax[1] && p[2] >= bmin[2] && p[2] <= bmax[2];
    }
    __device__ __forceinline__ float dist(const double p[3]) const {       // np.linalg.norm(pts_f32 - centroid)
        const float dx = (float)p[0] - centroid[0], dy = (float)p[1] - centroid[1], dz = (float)p[2] - centroid[2];
        return sqrtf((dx * dx + dy * dy) + dz * dz);
    }
};

static PointSetDev ps_to_dev(const mf_point_set* ps) { return PointSetDev{ps->pts, ps->ax, ps->ay, ps->az, ps->nx, ps->ny, ps->nz}; }
static SubmapDev sm_to_dev(const mf_submap* s) {
    SubmapDev d;
    for (int k = 0; k < 12; ++k) d.w2l[k] = s->w2l[k];
    for (int k = 0; k < 3; ++k) { d.bmin[k] = s->aabb_min[k]; d.bmax[k] = s->aabb_max[k]; d.centroid[k] = s->centroid[k]; }
    return d;
}

__global__ void jq_maxdist_kernel(PointSetDev ps, SubmapDev sub, int64_t g_begin, int64_t g_count, unsigned int* __restrict__ out) {
    float mx = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < g_count; i += (int64_t)gridDim.x * blockDim.x) {
        double p[3]; ps.get(g_begin + i, p);
        if (sub.contains(p)) mx = fmaxf(mx, sub.dist(p));
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0 && mx > 0.f) atomicMax(out, __float_as_uint(mx));     // non-negative floats order as uints
}

// compaction of the points of [c_begin, c_begin+c_count) that fall inside the submap's AABB
__global__ void jq_compact_kernel(PointSetDev ps, SubmapDev sub, int64_t g_begin, int64_t c_begin, int64_t c_count, int M, int m,
                                  uint8_t* __restrict__ contain, unsigned int* __restrict__ counter, int* __restrict__ list) {
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < c_count; base += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = base + threadIdx.x;
        bool in = false;
        if (i < c_count) {
            double p[3]; ps.get(g_begin + c_begin + i, p);
            in = sub.contains(p);
            if (contain) contain[(c_begin + i) * M + m] = in ? 1 : 0;
        }
        const unsigned int ballot = __ballot_sync(0xffffffffu, in);
        const int lane = threadIdx.x & 31;
        unsigned int wbase = 0;
        if (lane == 0 && ballot) wbase = atomicAdd(counter, (unsigned int)__popc(ballot));
        wbase = __shfl_sync(0xffffffffu, wbase, 0);
        if (in) list[wbase + __popc(ballot & ((1u << lane) - 1u))] = (int)i;
    }
}

struct SrcJoint {                       // compacted list entry -> world point -> submap frame -> normalised
    PointSetDev ps; SubmapDev sub; const int* list; int64_t origin;       // origin = g_begin + c_begin
    __device__ __forceinline__ void point(int64_t i, const FieldDev& f, float x[3]) const {
        double p[3]; ps.get(origin + list[i], p);
        const float w[3] = {(float)p[0], (float)p[1], (float)p[2]};       // .astype(np.float32)  (Mesher.py:476)
        float l[3];
#pragma unroll
        for (int j = 0; j < 3; ++j)                                        // sum(pts[:,None,:] * rot, -1) + trans
            l[j] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w[0], sub.w2l[j * 4]), __fmul_rn(w[1], sub.w2l[j * 4 + 1])),
                                       __fmul_rn(w[2], sub.w2l[j * 4 + 2])), sub.w2l[j * 4 + 3]);
        normalize_point(f, l, x);
    }
};

struct EpiJoint {
    PointSetDev ps; SubmapDev sub; const int* list; int64_t origin; int64_t c_begin;
    const uint8_t* vis; int M, m; const float* max_dist; int color; float* acc; uint8_t* mask_any;
    __device__ __forceinline__ void store(const float* OUT, int ld, int tp, int64_t tile, int64_t N, int tid, int nthreads) const {
        const int t = tid;
        const int64_t i = tile * tp + t;
        if (t >= tp || i >= N) return;
        const int64_t li = c_begin + list[i];                             // index relative to g_begin
        if (vis && !vis[li * M + m]) return;                              // Mesher.py:509-512
        double p[3]; ps.get(origin + list[i], p);
        const float ent = fminf(fmaxf(OUT[4 * ld + t], 0.f), 10000.f);   // np.clip(entropy, 0, 1e4)
        const float d = sub.dist(p);
        const float sigma = max_dist[m] / 3.0f;                            // convert_dist_to_weight, math_helper.py:66-72
        const float k1 = 1.0f / (sigma * 2.50662827463100050f);           // pdf_gauss, vis/math_helper.py:47-51
        const float m1 = d / sigma;
        const float w = expf(-10.0f * ent) * (k1 * expf(-0.5f * m1 * m1));
        mask_any[li] = 1;
        if (color) {
            float* a = acc + li * 4;
#pragma unroll
            for (int k = 0; k < 3; ++k) a[k] = fmaf(w, sigmoidf_(OUT[k * ld + t]), a[k]);
            a[3] += w;
        } else {
            float* a = acc + li * 2;
            a[0] = fmaf(w, OUT[3 * ld + t], a[0]);
            a[1] += w;
        }
    }
};

__global__ void jq_finalize_kernel(const float* __restrict__ acc, const uint8_t* __restrict__ mask_any, int color, int64_t n,
                                   float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (color) {
        const float w = acc[i * 4 + 3];
        const bool ok = mask_any[i] && w > 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) out[i * 3 + k] = ok ? acc[i * 4 + k] / w : 0.f;
    } else {
        const float w = acc[i * 2 + 1];
        out[i] = mask_any[i] ? (w > 0.f ? acc[i * 2] / w : 0.f) : -1.0f;
    }
}

static int check_ps(const mf_point_set* ps, int64_t g_begin, int64_t g_count) {
    if (!ps) { mf_set_error("null point set"); return MF_ERR_INVALID; }
    int64_t total;
    if (ps->pts) total = INT64_MAX;
    else {
        if (!ps->ax || !ps->ay || !ps->az || ps->nx <= 0 || ps->ny <= 0 || ps->nz <= 0) { mf_set_error("bad grid spec"); return MF_ERR_INVALID; }
        total = (int64_t)ps->nx * ps->ny * ps->nz;
    }
    if (g_begin < 0 || g_count < 0 || g_begin + g_count > total) { mf_set_error("point range out of bounds"); return MF_ERR_INVALID; }
    return MF_OK;
}

MF_API int64_t mf_joint_query_scratch_size(int64_t g_count) {
    const int64_t chunk = g_count < JQ_CHUNK ? g_count : JQ_CHUNK;
    return 256 + (chunk + 64) * (int64_t)sizeof(int);
}

MF_API int mf_joint_query_maxdist(const mf_point_set* ps, const mf_submap* submaps, int M, int64_t g_begin, int64_t g_count,
                                  float* max_dist, void* stream) {
    int rc = check_ps(ps, g_begin, g_count); if (rc) return rc;
    MF_CHECK_ARG(submaps && M > 0 && max_dist);
    if (g_count == 0) return MF_OK;
    const PointSetDev p = ps_to_dev(ps);
    const int64_t want = (g_count + 255) / 256;
    const unsigned blocks = (unsigned)(want < 8 * mf_sm_count_cached() ? want : 8 * mf_sm_count_cached());
    for (int m = 0; m < M; ++m) {
        jq_maxdist_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(p, sm_to_dev(&submaps[m]), g_begin, g_count, (unsigned int*)max_dist + m);
        MF_LAUNCH_CHECK();
    }
    return MF_OK;
}

MF_API int mf_joint_query_accumulate(const mf_point_set* ps, const mf_submap* submaps, int M, int m_begin, int m_count,
                                     const float* max_dist, const uint8_t* vis, int color, int64_t g_begin, int64_t g_count,
                                     float* acc, uint8_t* mask_any, uint8_t* contain, void* scratch, void* stream) {
    int rc = check_ps(ps, g_begin, g_count); if (rc) return rc;
    MF_CHECK_ARG(submaps && M > 0 && m_begin >= 0 && m_count >= 0 && m_begin + m_count <= M);
    MF_CHECK_ARG(max_dist && acc && mask_any && scratch);
    if (g_count == 0 || m_count == 0) return MF_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const PointSetDev p = ps_to_dev(ps);
    unsigned int* counter = (unsigned int*)scratch;
    int* list = (int*)((char*)scratch + 256);
    for (int m = m_begin; m < m_begin + m_count; ++m) {
        FieldDev d; rc = mf_field_to_dev(&submaps[m].field, &d); if (rc) return rc;
        const SubmapDev sub = sm_to_dev(&submaps[m]);
        for (int64_t c_begin = 0; c_begin < g_count; c_begin += JQ_CHUNK) {
            const int64_t c_count = (g_count - c_begin) < JQ_CHUNK ? (g_count - c_begin) : JQ_CHUNK;
            MF_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned int), st));
            const int64_t want = (c_count + 255) / 256;
            const unsigned blocks = (unsigned)(want < 8 * mf_sm_count_cached() ? want : 8 * mf_sm_count_cached());
            jq_compact_kernel<<<blocks, 256, 0, st>>>(p, sub, g_begin, c_begin, c_count, M, m, contain, counter, list);
            MF_LAUNCH_CHECK();
            SrcJoint src{p, sub, list, g_begin + c_begin};
            EpiJoint epi{p, sub, list, g_begin + c_begin, c_begin, vis, M, m, max_dist, color, acc, mask_any};
            // launches with few tiles per CTA cannot fill the producer -> consumer pipeline: dual-pipeline kernel for those; at
            // mesh resolution (512^3: ~30 tiles per CTA and chunk) the producer / consumer kernel is 12 % faster (0.165 vs 0.188 s)
            const bool dual = c_count < (1 << 20);
            if (color) rc = launch_field_fwd_auto<SrcJoint, EpiJoint, false>(d, src, epi, c_count, st, counter, dual);
            else rc = launch_field_fwd_auto<SrcJoint, EpiJoint, true>(d, src, epi, c_count, st, counter, dual);
            if (rc) return rc;
        }
    }
    return MF_OK;
}

MF_API int mf_joint_query_finalize(const float* acc, const uint8_t* mask_any, int color, int64_t g_count, float* out, void* stream) {
    MF_CHECK_ARG(g_count >= 0);
    if (g_count == 0) return MF_OK;
    MF_CHECK_ARG(acc && mask_any && out);
    jq_finalize_kernel<<<(unsigned)((g_count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(acc, mask_any, color, g_count, out);
    MF_LAUNCH_CHECK();
    return MF_OK;
}

// ---------------------------------------------------------------------------------------------
// N3: containment of a frame's surface points in the submaps' axis-aligned boxes
// (Manager.find_highest_containing_ratio / compute_containing_ratio, Manager.py:159-244; pts_in_bbox, geometry_helper.py:193-203).
// One thread per sampled pixel: world point = t + (R d_cam) * depth with the reference's fp32 operation order (three products
// summed left to right, then multiply, then add: torch.sum(d[..., None, :] * R, -1), rays_o + rays_d * depth), strict
// comparisons against every box, integer counts (deterministic).
// counts: [0, k) points inside box j; [k, 2k) points inside box j with depth > 0; [2k] points with depth > 0.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) containment_kernel(const float* __restrict__ dirs_cam, const float* __restrict__ depth,
                                                          const float* __restrict__ pose, const float* __restrict__ pts_in,
                                                          const float* __restrict__ xyz_min, const float* __restrict__ xyz_max,
                                                          int k, int64_t n, int64_t n_depth, uint8_t* __restrict__ mask,
                                                          unsigned long long* __restrict__ counts) {
    // n_depth > 0: every direction is paired with every depth (n = n_dirs * n_depth points; see mf_containment)
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool on = i < n;
    float p[3] = {0.f, 0.f, 0.f};
    bool valid = false;
    if (on) {
        if (pts_in) {
            p[0] = pts_in[i * 3]; p[1] = pts_in[i * 3 + 1]; p[2] = pts_in[i * 3 + 2];
            valid = true;
        } else {
            const int64_t id = n_depth > 0 ? i / n_depth : i, ip = n_depth > 0 ? i % n_depth : i;
            const float dx = dirs_cam[id * 3], dy = dirs_cam[id * 3 + 1], dz = dirs_cam[id * 3 + 2], dep = depth[ip];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const float dw = __fadd_rn(__fadd_rn(__fmul_rn(dx, pose[j * 4]), __fmul_rn(dy, pose[j * 4 + 1])), __fmul_rn(dz, pose[j * 4 + 2]));
                p[j] = __fadd_rn(pose[j * 4 + 3], __fmul_rn(dw, dep));
            }
            valid = dep > 0.f;
        }
    }
    const int lane = threadIdx.x & 31;
    for (int j = 0; j < k; ++j) {
        bool in = on;
#pragma unroll
        for (int d = 0; d < 3; ++d) in = in && (p[d] > xyz_min[j * 3 + d]) && (p[d] < xyz_max[j * 3 + d]);
        if (mask && on) mask[i * k + j] = in ? 1 : 0;
        if (counts) {
            const unsigned b_all = __ballot_sync(0xffffffffu, in), b_val = __ballot_sync(0xffffffffu, in && valid);
            if (lane == 0) {
                if (b_all) atomicAdd(&counts[j], (unsigned long long)__popc(b_all));
                if (b_val) atomicAdd(&counts[k + j], (unsigned long long)__popc(b_val));
            }
        }
    }
    if (counts) {
        const unsigned b = __ballot_sync(0xffffffffu, on && valid);
        if (lane == 0 && b) atomicAdd(&counts[2 * k], (unsigned long long)__popc(b));
    }
}

MF_API int mf_containment(const float* dirs_cam, const float* depth, const float* pose_c2w, const float* pts, const float* xyz_min,
                          const float* xyz_max, int k, int64_t n, int cross, uint8_t* mask, int64_t* counts, void* stream) {
    MF_CHECK_ARG(n >= 0 && k >= 1);
    cudaStream_t st = (cudaStream_t)stream;
    if (counts) MF_CUDA(cudaMemsetAsync(counts, 0, (size_t)(2 * k + 1) * sizeof(int64_t), st));
    if (n == 0) return MF_OK;                          // (an empty point set has no buffers to check)
    MF_CHECK_ARG(xyz_min && xyz_max && (mask || counts) && !(cross && pts));
    MF_CHECK_ARG(pts || (dirs_cam && depth && pose_c2w));
    const int64_t n_depth = cross ? n : 0;
    if (cross) n = n * n;
    containment_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dirs_cam, depth, pose_c2w, pts, xyz_min, xyz_max, k, n, n_depth, mask,
                                                                    reinterpret_cast<unsigned long long*>(counts));
    MF_LAUNCH_CHECK();
    return MF_OK;
}

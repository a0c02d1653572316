// TEST INFRASTRUCTURE -- CPU restatement of the reference's marching cubes (N2), used only by tests/, smoke() and bench.py's CPU
// baseline; the product path never links or calls it.
//
// Restates /root/reference/external/NumpyMarchingCubes/marching_cubes/src/marching_cubes.cpp (called from utils/utils.py:78,159
// as mcubes.marching_cubes(raw, isolevel, truncation=3.0)):
//   * get_voxel (:70-91) / trilerp (:93-113): the corners of the cell centred on the integer position (i,j,k) sit at +-0.5, so each
//     corner value is the mean of the 8 voxels around it, summed in the order 000,100,010,001,110,011,101,111 with weights
//     ((1-wx)(1-wy)(1-wz)) d evaluated left to right in float; a corner is valid iff all 8 voxels are inside the volume, not -inf
//     and |d| < truncation.  Here the corner ("dual node") values are formed once per node instead of once per cell.
//   * extract_isosurface_at_position (:138-243): all 8 corners valid, case index from (010,110,100,000,011,111,101,001) < iso,
//     the pairwise / absolute threshold tests (thresh = 10, :441), cases with edge mask 0 or 255 dropped (:194), vertexInterp
//     (:115-136) on the 12 edges in the reference's corner order, triangles in table order, cells scanned i, j, k (k fastest, :418-433).
//   * merge_close_vertices(approx = true, thresh = 1e-5) (:316-390): greedy first-come clustering on the integer lattice
//     (int)(v / thresh + 0.5 sgn v) with the 27-neighbourhood probed in (di,dj,dk) order; remove_degenerate_faces (:264-288);
//     remove_duplicate_faces (:246-262, first occurrence of each sorted index triple kept).
// Pinned against the reference binary itself (oracle/_ref/_mcubes_ref.so, built from the reference's own sources by oracle/Makefile)
// in tests/test_marching_cubes.py and against tests/golden/mcubes.npz generated from that binary.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <algorithm>
#include <limits>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../mipsfusion_b200/csrc/mc_tables.h"

namespace {

struct P3 { float x, y, z; };
struct K3 {
    int x, y, z;
    bool operator==(const K3& o) const { return x == o.x && y == o.y && z == o.z; }
};
struct K3Hash {
    size_t operator()(const K3& k) const {
        uint64_t h = (uint64_t)(uint32_t)k.x * 0x9E3779B97F4A7C15ull;
        h ^= (uint64_t)(uint32_t)k.y * 0xC2B2AE3D27D4EB4Full + (h >> 29);
        h ^= (uint64_t)(uint32_t)k.z * 0x165667B19E3779F9ull + (h << 7);
        return (size_t)(h ^ (h >> 31));
    }
};

inline int sgn(float v) { return (0.0f < v) - (v < 0.0f); }

P3 interp(float iso, P3 p1, P3 p2, float d1, float d2) {           // marching_cubes.cpp:115-136
    if (fabsf(iso - d1) < 0.00001f) return p1;
    if (fabsf(iso - d2) < 0.00001f) return p2;
    if (fabsf(d1 - d2) < 0.00001f) return p1;
    float mu = (iso - d1) / (d2 - d1);
    P3 r;
    r.x = p1.x + mu * (p2.x - p1.x);
    r.y = p1.y + mu * (p2.y - p1.y);
    r.z = p1.z + mu * (p2.z - p1.z);
    return r;
}

}  // namespace

extern "C" {

// vol (nx,ny,nz) C-contiguous float32.  Outputs are malloc'd: verts (nv,3) float32, faces (nf,3) uint32; soup_tris (optional)
// receives the number of triangles before merging.  Returns 0.
int mcubes_oracle(const float* vol, long nx, long ny, long nz, float iso, float truncation, float** verts_out, long* nv_out,
                  uint32_t** faces_out, long* nf_out, long* soup_tris) {
    const float thresh = 10.0f;
    const float NEG_INF = -std::numeric_limits<float>::infinity();
    // ---- dual nodes: node (a,b,c) sits at (a+.5, b+.5, c+.5), a in [0,nx-1) ----
    long mx = nx - 1, my = ny - 1, mz = nz - 1;
    std::vector<float> node;
    std::vector<uint8_t> nvalid;
    if (mx > 0 && my > 0 && mz > 0) {
        node.resize((size_t)mx * my * mz);
        nvalid.resize((size_t)mx * my * mz);
        static const int ORD[8][3] = {{0,0,0},{1,0,0},{0,1,0},{0,0,1},{1,1,0},{0,1,1},{1,0,1},{1,1,1}};
        for (long a = 0; a < mx; a++) for (long b = 0; b < my; b++) for (long c = 0; c < mz; c++) {
            // weight = pos - (int)pos with pos = a + 0.5 >= 0.5
            float px = (float)a + 0.5f, py = (float)b + 0.5f, pz = (float)c + 0.5f;
            float wx = px - (float)(int)px, wy = py - (float)(int)py, wz = pz - (float)(int)pz;
            float dist = 0.0f; bool ok = true;
            for (int q = 0; q < 8 && ok; q++) {
                float d = vol[((a + ORD[q][0]) * ny + (b + ORD[q][1])) * nz + (c + ORD[q][2])];
                if (!(d != NEG_INF && fabsf(d) < truncation)) { ok = false; break; }
                float fx = ORD[q][0] ? wx : (1.0f - wx), fy = ORD[q][1] ? wy : (1.0f - wy), fz = ORD[q][2] ? wz : (1.0f - wz);
                dist += fx * fy * fz * d;
            }
            size_t id = ((size_t)a * my + b) * mz + c;
            node[id] = dist; nvalid[id] = ok;
        }
    }
    // ---- case tables derived from the packed triangle list ----
    int edge_mask[256];
    for (int c = 0; c < 256; c++) {
        int m = 0;
        for (int i = 0; i < 16; i++) { int e = (int)((MC_TRI_PACKED[c] >> (4 * i)) & 15); if (e == 15) break; m |= 1 << e; }
        edge_mask[c] = m;
    }
    // corner numbering: 0:000 1:100 2:010 3:001 4:110 5:011 6:101 7:111 (offsets in units of a node step)
    static const int COFF[8][3] = {{0,0,0},{1,0,0},{0,1,0},{0,0,1},{1,1,0},{0,1,1},{1,0,1},{1,1,1}};
    static const int CASE_BIT[8] = {8, 4, 1, 128, 2, 16, 64, 32};           // :160-167
    static const int EDGE[12][2] = {{2,4},{4,1},{1,0},{0,2},{5,7},{7,6},{6,3},{3,5},{2,5},{4,7},{1,6},{0,3}};   // :205-216
    std::vector<P3> soup;
    for (long i = 1; i + 1 < nx; i++) for (long j = 1; j + 1 < ny; j++) for (long k = 1; k + 1 < nz; k++) {
        float d[8]; P3 p[8]; bool ok = true;
        for (int q = 0; q < 8; q++) {
            size_t id = ((size_t)(i - 1 + COFF[q][0]) * my + (j - 1 + COFF[q][1])) * mz + (k - 1 + COFF[q][2]);
            if (!nvalid[id]) { ok = false; break; }
            d[q] = node[id];
            p[q].x = (float)i + (COFF[q][0] ? 0.5f : -0.5f);
            p[q].y = (float)j + (COFF[q][1] ? 0.5f : -0.5f);
            p[q].z = (float)k + (COFF[q][2] ? 0.5f : -0.5f);
        }
        if (!ok) continue;
        int cube = 0;
        for (int q = 0; q < 8; q++) if (d[q] < iso) cube += CASE_BIT[q];
        for (int a = 0; a < 8 && ok; a++) for (int b = 0; b < 8; b++) {
            if (d[a] * d[b] < 0.0f) { if (fabsf(d[a]) + fabsf(d[b]) > thresh) { ok = false; break; } }
            else if (fabsf(d[a] - d[b]) > thresh) { ok = false; break; }
        }
        for (int q = 0; q < 8 && ok; q++) if (fabsf(d[q]) > thresh) ok = false;
        if (!ok) continue;
        int em = edge_mask[cube];
        if (em == 0 || em == 255) continue;
        P3 vl[12];
        for (int e = 0; e < 12; e++) if (em & (1 << e)) vl[e] = interp(iso, p[EDGE[e][0]], p[EDGE[e][1]], d[EDGE[e][0]], d[EDGE[e][1]]);
        for (int t = 0; t < 16; t += 3) {
            int e0 = (int)((MC_TRI_PACKED[cube] >> (4 * t)) & 15);
            if (e0 == 15) break;
            soup.push_back(vl[e0]);
            soup.push_back(vl[(MC_TRI_PACKED[cube] >> (4 * (t + 1))) & 15]);
            soup.push_back(vl[(MC_TRI_PACKED[cube] >> (4 * (t + 2))) & 15]);
        }
    }
    if (soup_tris) *soup_tris = (long)(soup.size() / 3);
    // ---- greedy lattice merge (approx = true, thresh 1e-5) ----
    const float mt = 0.00001f;
    size_t numV = soup.size();
    std::vector<uint32_t> lookup(numV);
    std::vector<P3> verts; verts.reserve(numV / 4 + 16);
    std::unordered_map<K3, uint32_t, K3Hash> grid; grid.reserve(numV / 2 + 16);
    for (size_t v = 0; v < numV; v++) {
        const P3& q = soup[v];
        K3 c = {(int)(q.x / mt + 0.5f * sgn(q.x)), (int)(q.y / mt + 0.5f * sgn(q.y)), (int)(q.z / mt + 0.5f * sgn(q.z))};
        uint32_t nn = 0xFFFFFFFFu;
        for (int di = -1; di <= 1 && nn == 0xFFFFFFFFu; di++) for (int dj = -1; dj <= 1 && nn == 0xFFFFFFFFu; dj++)
            for (int dk = -1; dk <= 1; dk++) {
                auto it = grid.find(K3{c.x + di, c.y + dj, c.z + dk});
                if (it != grid.end()) { nn = it->second; break; }
            }
        if (nn == 0xFFFFFFFFu) { nn = (uint32_t)verts.size(); grid.emplace(c, nn); verts.push_back(q); }
        lookup[v] = nn;
    }
    // ---- faces: degenerate ones dropped, then duplicates (first occurrence of the sorted triple kept) ----
    struct TriHash { size_t operator()(const K3& k) const { return K3Hash()(k); } };
    std::unordered_set<K3, K3Hash> seen; seen.reserve(numV / 3 + 16);
    std::vector<uint32_t> faces; faces.reserve(numV);
    for (size_t t = 0; t < numV / 3; t++) {
        uint32_t a = lookup[3 * t], b = lookup[3 * t + 1], c = lookup[3 * t + 2];
        if (a == b || a == c || b == c) continue;
        uint32_t s[3] = {a, b, c};
        std::sort(s, s + 3);
        if (!seen.insert(K3{(int)s[0], (int)s[1], (int)s[2]}).second) continue;
        faces.push_back(a); faces.push_back(b); faces.push_back(c);
    }
    *nv_out = (long)verts.size();
    *nf_out = (long)(faces.size() / 3);
    *verts_out = (float*)malloc(std::max<size_t>(1, verts.size()) * sizeof(P3));
    *faces_out = (uint32_t*)malloc(std::max<size_t>(1, faces.size()) * sizeof(uint32_t));
    if (!verts.empty()) memcpy(*verts_out, verts.data(), verts.size() * sizeof(P3));
    if (!faces.empty()) memcpy(*faces_out, faces.data(), faces.size() * sizeof(uint32_t));
    return 0;
}

void mcubes_oracle_free(void* p) { free(p); }

}  // extern "C"

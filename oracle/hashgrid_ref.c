/* Oracle (plain C): tiny-cuda-nn 1.7 HashGrid index + interpolation arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY -- never linked into the product library.
 * Independent restatement (real uint32 wrap-around, real fmaf) of the algorithm
 * that oracle/hashgrid.py emulates in int64; the two are cross-checked in
 * tests/test_oracle_hashgrid.py.  **parity unpinned** (tinycudann==1.7 is an
 * un-vendored third-party dependency of the reference, environment.yaml:74;
 * call sites model/encodings.py:14-25, model/scene_rep.py:122).
 *
 * Build: make -C oracle   (gcc -O2 -shared -fPIC -> oracle/libhashgrid_ref.so)
 */
#include <math.h>
#include <stdint.h>

#define PRIME1 2654435761u
#define PRIME2 805459861u

/* tcnn grid_scale / grid_resolution + the offset table of GridEncodingTemplated. */
int hg_level_table(int log2_T, int n_levels, int base_res, double per_level_scale_d,
                   float* scale, uint32_t* res, uint32_t* size, uint32_t* offset) {
    float per_level_scale = (float)per_level_scale_d;   /* JSON -> float */
    float log2_pls = log2f(per_level_scale);
    uint32_t off = 0;
    for (int l = 0; l < n_levels; ++l) {
        /* exp2f := correctly rounded fp32 (via fp64), see oracle/hashgrid.py */
        float arg = (float)l * log2_pls;
        float s = (float)exp2((double)arg) * (float)base_res - 1.0f;
        uint32_t r = (uint32_t)ceilf(s) + 1u;
        uint32_t max_params = UINT32_MAX / 2;
        uint32_t n = (powf((float)r, 3.0f) > (float)max_params) ? max_params : r * r * r;
        n = (n + 7u) / 8u * 8u;
        uint32_t cap = 1u << log2_T;
        if (n > cap) n = cap;
        scale[l] = s; res[l] = r; size[l] = n; offset[l] = off;
        off += n;
    }
    offset[n_levels] = off;
    return 0;
}

static inline uint32_t grid_index(const uint32_t p[3], uint32_t res, uint32_t size) {
    uint32_t stride = 1, index = 0;
    for (int d = 0; d < 3 && stride <= size; ++d) {
        index += p[d] * stride;
        stride *= res;
    }
    if (size < stride) index = (p[0] * 1u) ^ (p[1] * PRIME1) ^ (p[2] * PRIME2);
    return index % size;
}

/* idx: (N, L, 8) uint32; w: (N, L, 8) float (either may be NULL);
 * out: (N, L*F) float if params != NULL. */
void hg_eval(const float* x, int64_t N, int n_levels, int F,
             const float* scale, const uint32_t* res, const uint32_t* size, const uint32_t* offset,
             const float* params, uint32_t* idx, float* w, float* out) {
    for (int64_t i = 0; i < N; ++i) {
        for (int l = 0; l < n_levels; ++l) {
            float f[3]; uint32_t g[3];
            for (int d = 0; d < 3; ++d) {
                float pos = fmaf(scale[l], x[i * 3 + d], 0.5f);
                float t = floorf(pos);
                g[d] = (uint32_t)(int)t;
                f[d] = pos - t;
            }
            float acc[8] = {0};
            for (int c = 0; c < 8; ++c) {
                float wc = 1.0f; uint32_t p[3];
                for (int d = 0; d < 3; ++d) {
                    if ((c >> d) & 1) { wc *= f[d]; p[d] = g[d] + 1u; }
                    else              { wc *= 1.0f - f[d]; p[d] = g[d]; }
                }
                uint32_t k = grid_index(p, res[l], size[l]);
                if (idx) idx[(i * n_levels + l) * 8 + c] = k;
                if (w)   w[(i * n_levels + l) * 8 + c] = wc;
                if (params && out)
                    for (int ft = 0; ft < F; ++ft)
                        acc[ft] = fmaf(wc, params[((uint64_t)offset[l] + k) * F + ft], acc[ft]);
            }
            if (params && out)
                for (int ft = 0; ft < F; ++ft) out[i * (n_levels * F) + l * F + ft] = acc[ft];
        }
    }
}

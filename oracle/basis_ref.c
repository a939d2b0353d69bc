/* CPU restatement of the basis layer in plain C (TEST INFRASTRUCTURE ONLY).
 *
 * Same FP32 operation order as the CUDA kernels (st_dadk_b200/csrc/basis.cuh) so that
 * knot-support index sets compare bit-exactly: d2 = fl(fl(dx*dx)+fl(dy*dy)) < fl(th*th).
 * Follows stnf/models/st_interp.py:433-491 (spatial), :583-596 (temporal).
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (see oracle/Makefile); never linked into the product.
 * Parity: pinned against tests/golden/basis_values.npz and kat.json by tests/test_oracle.py.
 */
#include <math.h>
#include <stdint.h>

static float phi_eval(int fn, float d2, float th2, float inv_th) {
    if (fn == 1) { /* gaussian: exp(-r^2/2) */
        float r = sqrtf(d2) * inv_th;
        return expf(-0.5f * r * r);
    }
    if (!(d2 < th2)) return 0.0f;
    float r = sqrtf(d2) * inv_th;
    if (r >= 1.0f) return 0.0f;
    float u = 1.0f - r;
    if (fn == 2) return u; /* triangular */
    float u2 = u * u;
    float u6 = u2 * u2 * u2;
    return u6 * ((35.0f * r + 18.0f) * r + 3.0f) * (1.0f / 3.0f);
}

/* phi (n x k) row-major; thetap[k] = bandwidth*calibration already applied. */
void ref_spatial_basis_f32(const float* coords, int64_t n, const float* centers, const float* thetap,
                           int k, int fn, float* phi) {
    for (int64_t i = 0; i < n; ++i) {
        float x = coords[2 * i], y = coords[2 * i + 1];
        for (int j = 0; j < k; ++j) {
            float dx = x - centers[2 * j], dy = y - centers[2 * j + 1];
            float a = dx * dx, b = dy * dy;
            float d2 = a + b;
            float th = thetap[j];
            phi[i * k + j] = phi_eval(fn, d2, th * th, 1.0f / th);
        }
    }
}

/* support mask (n x k) bytes: 1 iff d2 < th2 (all ones for the non-compact gaussian). */
void ref_support_mask(const float* coords, int64_t n, const float* centers, const float* thetap,
                      int k, int fn, uint8_t* mask) {
    for (int64_t i = 0; i < n; ++i) {
        float x = coords[2 * i], y = coords[2 * i + 1];
        for (int j = 0; j < k; ++j) {
            float dx = x - centers[2 * j], dy = y - centers[2 * j + 1];
            float a = dx * dx, b = dy * dy;
            float d2 = a + b;
            float th = thetap[j];
            mask[i * k + j] = (fn == 1) ? 1 : (d2 < th * th);
        }
    }
}

void ref_temporal_basis_f32(const float* t, int64_t n, const float* centers, const float* bw, int k, float* psi) {
    for (int64_t i = 0; i < n; ++i)
        for (int j = 0; j < k; ++j) {
            float s = (t[i] - centers[j]) * (1.0f / bw[j]);
            psi[i * k + j] = expf(-0.5f * s * s);
        }
}

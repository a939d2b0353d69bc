// Large-knot-count regime (K_s ~ 1e5, BASELINE config 4): the spatial part of block 1 walks only the knots inside
// each point's compact support instead of generating K_s dense operand columns.
//
//   sparse_spatial_fwd_kernel  : zs[n, :] = sum_{j in supp(s_n)} phi_j(s_n) * W1t[p + j, :]      (N x n_out, FP32)
//   sparse_spatial_wgrad_kernel: dW1t[p + j, :] += phi_j(s_n) * dz1[n, :] for j in supp(s_n)       (vector atomics)
//
// W1t is the knot-major storage of the first Linear layer ((n_in, n_out)-contiguous: one knot = one contiguous row,
// 1 KB at n_out = 256), so a gather is a coalesced row read.  The dense temporal / covariate columns still run on
// the tensor cores (layer_fwd/bwd with a k_s = 0 basis); zs enters their epilogue as an addend before LayerNorm.
//
// One warp per point.  Uniform lattices only (knot j of a level sits at (ix, iy) = (j / side, j % side)): the
// candidate window per level is closed form, [ceil((x - th) g) - 1, floor((x + th) g) + 1] x the same in y with
// g = side - 1 (one lattice step of margin; the exact FP32 support predicate of phi_eval decides membership, so index
// sets are the same as the dense path's).  Lanes evaluate candidates in parallel and compact the active ones into a
// per-warp list with ballots; then all lanes stream the listed rows, each lane owning 4-column groups.
#pragma once
#include "common.cuh"

namespace stdadk {

constexpr int SP_MAX_LEVELS = 8;
constexpr int SP_MAX_ACTIVE = 192;   // >= sum over levels of knots in a support disk (21 per level for 2.5 spacings)
constexpr int SP_WARPS = 8;

struct Lattice {
    int n_levels;
    int side[SP_MAX_LEVELS];
    int offset[SP_MAX_LEVELS];
    float thetap[SP_MAX_LEVELS];
};

// Fills the warp's active list; returns its length (warp-uniform).
__device__ __forceinline__ int sparse_collect(const Lattice& Lt, const float4* __restrict__ knots, int fn, float x, float y,
                                              int lane, int* s_idx, float* s_phi) {
    int cnt = 0;
    for (int l = 0; l < Lt.n_levels; ++l) {
        const int side = Lt.side[l];
        const float g = (float)(side - 1), th = Lt.thetap[l];
        int ix0 = max(0, (int)ceilf((x - th) * g) - 1), ix1 = min(side - 1, (int)floorf((x + th) * g) + 1);
        int iy0 = max(0, (int)ceilf((y - th) * g) - 1), iy1 = min(side - 1, (int)floorf((y + th) * g) + 1);
        if (side == 1) { ix0 = ix1 = iy0 = iy1 = 0; }
        const int ny = iy1 - iy0 + 1, ncand = (ix1 - ix0 + 1) * ny;
        for (int base = 0; base < ncand; base += 32) {
            const int c = base + lane;
            float val = 0.0f;
            int j = 0;
            if (c < ncand) {
                j = Lt.offset[l] + (ix0 + c / ny) * side + (iy0 + c % ny);
                float4 kn = __ldg(&knots[j]);
                val = phi_eval(fn, x - kn.x, y - kn.y, kn.z, kn.w);
            }
            const bool active = val > 0.0f;
            const unsigned m = __ballot_sync(0xffffffffu, active);
            const int pos = cnt + __popc(m & ((1u << lane) - 1u));
            if (active && pos < SP_MAX_ACTIVE) {
                s_idx[pos] = j;
                s_phi[pos] = val;
            }
            cnt += __popc(m);
        }
    }
    __syncwarp();
    return min(cnt, SP_MAX_ACTIVE);
}

struct SparseK {
    PointsP pts;
    Lattice lat;
    const float4* knots;
    const float* w1t;        // (n_in, n_out) contiguous; spatial rows start at row p_cov
    float* zs;               // fwd: out (n_rows x n_out)
    const float* dz_img;     // wgrad: image (n_rows x n_out)
    float* dw1t;             // wgrad: += into the same geometry as w1t
    int n_out, p_cov, fn, _pad;
};

__global__ void __launch_bounds__(SP_WARPS * 32) sparse_spatial_fwd_kernel(const __grid_constant__ SparseK P) {
    __shared__ int s_idx[SP_WARPS][SP_MAX_ACTIVE];
    __shared__ float s_phi[SP_WARPS][SP_MAX_ACTIVE];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long n_warps = (long long)gridDim.x * SP_WARPS;
    const int ngroups = P.n_out >> 2;
    for (long long r = (long long)blockIdx.x * SP_WARPS + warp; r < P.pts.n_rows; r += n_warps) {
        float x, y, t;
        load_point(P.pts, P.pts.row_begin + r, x, y, t);
        const int cnt = sparse_collect(P.lat, P.knots, P.fn, x, y, lane, s_idx[warp], s_phi[warp]);
        float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
        for (int e = 0; e < cnt; ++e) {
            const float f = s_phi[warp][e];
            const float4* row = reinterpret_cast<const float4*>(P.w1t + (size_t)(P.p_cov + s_idx[warp][e]) * P.n_out);
            if (lane < ngroups) {
                float4 w = __ldg(row + lane);
                acc0.x = fmaf(f, w.x, acc0.x); acc0.y = fmaf(f, w.y, acc0.y);
                acc0.z = fmaf(f, w.z, acc0.z); acc0.w = fmaf(f, w.w, acc0.w);
            }
            if (lane + 32 < ngroups) {
                float4 w = __ldg(row + lane + 32);
                acc1.x = fmaf(f, w.x, acc1.x); acc1.y = fmaf(f, w.y, acc1.y);
                acc1.z = fmaf(f, w.z, acc1.z); acc1.w = fmaf(f, w.w, acc1.w);
            }
        }
        float4* out = reinterpret_cast<float4*>(P.zs + (size_t)r * P.n_out);
        if (lane < ngroups) out[lane] = acc0;
        if (lane + 32 < ngroups) out[lane + 32] = acc1;
        __syncwarp();
    }
}

// element group (row r, columns 4*g4 .. 4*g4+3) of an operand image
__device__ __forceinline__ float4 image_load4(const float* img, int slabs, long long r, int g4) {
    const long long tile = r / TILE_M;
    const uint32_t row = (uint32_t)(r - tile * TILE_M);
    const int slab = g4 >> 3, chunk = g4 & 7;
    const uint8_t* base = reinterpret_cast<const uint8_t*>(img + ((size_t)tile * slabs + slab) * SLAB_FLOATS);
    return *reinterpret_cast<const float4*>(base + swz_off(row, (uint32_t)chunk));
}

__global__ void __launch_bounds__(SP_WARPS * 32) sparse_spatial_wgrad_kernel(const __grid_constant__ SparseK P) {
    __shared__ int s_idx[SP_WARPS][SP_MAX_ACTIVE];
    __shared__ float s_phi[SP_WARPS][SP_MAX_ACTIVE];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long n_warps = (long long)gridDim.x * SP_WARPS;
    const int ngroups = P.n_out >> 2;
    const int slabs = pad32(P.n_out) / SLAB_K;
    for (long long r = (long long)blockIdx.x * SP_WARPS + warp; r < P.pts.n_rows; r += n_warps) {
        float x, y, t;
        load_point(P.pts, P.pts.row_begin + r, x, y, t);
        const int cnt = sparse_collect(P.lat, P.knots, P.fn, x, y, lane, s_idx[warp], s_phi[warp]);
        float4 d0 = make_float4(0.f, 0.f, 0.f, 0.f), d1 = d0;
        if (lane < ngroups) d0 = image_load4(P.dz_img, slabs, r, lane);
        if (lane + 32 < ngroups) d1 = image_load4(P.dz_img, slabs, r, lane + 32);
        for (int e = 0; e < cnt; ++e) {
            const float f = s_phi[warp][e];
            float4* row = reinterpret_cast<float4*>(P.dw1t + (size_t)(P.p_cov + s_idx[warp][e]) * P.n_out);
            if (lane < ngroups) atomicAdd(row + lane, make_float4(f * d0.x, f * d0.y, f * d0.z, f * d0.w));
            if (lane + 32 < ngroups) atomicAdd(row + lane + 32, make_float4(f * d1.x, f * d1.y, f * d1.z, f * d1.w));
        }
        __syncwarp();
    }
}

}  // namespace stdadk

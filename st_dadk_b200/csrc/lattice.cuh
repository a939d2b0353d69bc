// Block 1 (basis + Linear + LayerNorm/ReLU/dropout [+ head]) for the FIXED UNIFORM LATTICE of upstream's default
// configuration (SpatialBasisEmbedding._init_uniform, st_interp.py:152-185): the basis operand is built by WALKING each
// point's compact support instead of testing all K_s knots.
//
// layer_fwd_kernel<BASIS> evaluates every knot of a row (227 distance tests + a warp vote per 4 knots): with rows of a
// warp that are not neighbours (training batches are random samples) the vote never skips, and the kernel spends ~390 of
// its ~640 warp instructions per row there (ncu, profiles/).  Here, per row and level:
//   * the lattice lines within theta' of the point are a closed-form window of at most 6 x 6 knots (support radius =
//     2.5 lattice spacings); the 6 + 6 squared coordinate differences are formed once, a candidate costs one add and
//     one compare, and only knots inside the support (~16 of 36) get the Wendland / triangular polynomial;
//   * the values are appended to a per-row list in shared memory as 32-bit words: the TF32-rounded value in the upper
//     19 bits, the operand column in the (zero) lower 13 -- no information is lost and a row's support is ~200 bytes;
//   * the 128-row x 32-column operand slabs are zero-filled and the list entries scattered into them; temporal columns
//     (dense by nature) are evaluated as before.
// The support predicate is the library's exact one (d2 = fl(fl(dx^2) + fl(dy^2)) < theta'^2 on the same knot
// coordinates), so index sets and values equal the dense generator's bit for bit.
//
// One CTA per SM, persistent over 128-row tiles: 16 worker warps (thread = row x 1/4 of the work), a producer warp
// (W1 slabs by TMA) and an MMA-issuer warp.  Level l of a row is walked by worker column group l; slab s is assembled by
// column group s % 4; the epilogue is fwd_epilogue<4> (layer.cuh).
#pragma once
#include "layer.cuh"

namespace stdadk {

constexpr int LT_CG = 4;
constexpr int LT_NW = 128 * LT_CG;
constexpr int LT_NT = LT_NW + 64;
constexpr int LT_MAX_LEVELS = 4;
constexpr int LT_LIST = 24;              // entries per (row, level): a support disk of radius 2.5 spacings holds <= 21 knots
constexpr int LT_WIN = 6;                // lattice lines per axis inside a support
constexpr int LT_AXIS = 64;              // knots per axis and level, at most

struct LatP {
    int n_levels, ks_aligned;            // ks_aligned: first operand column that is not a purely spatial 4-column chunk
    int side[LT_MAX_LEVELS];
    int offset[LT_MAX_LEVELS];           // operand column of the level's first knot (p_cov = 0)
    float th2[LT_MAX_LEVELS], inv_th[LT_MAX_LEVELS], thg[LT_MAX_LEVELS];   // theta'^2, 1/theta', theta' * (side - 1)
};

struct LatSmem {
    uint32_t a_off, b_off, bar_off, tmem_off, vec_off, headw_off, knots_off, tknots_off, list_off, cnt_off, axis_off, red_off,
        total;
};
__host__ __device__ inline LatSmem plan_lattice(int n_pad, int q, int k_s, int k_t, int ns) {
    LatSmem s;
    uint32_t o = 0;
    s.a_off = o; o += ns * SLAB_BYTES;
    s.b_off = o; o += ns * (uint32_t)n_pad * 128u;
    s.bar_off = o; o += 128;
    s.tmem_off = o; o += 16;
    s.vec_off = o; o += 3u * n_pad * 4u;
    s.headw_off = o; o += (uint32_t)(q > 0 ? (q * n_pad + STDADK_MAX_Q) * 4 : 0);
    o = (o + 15u) & ~15u;
    s.knots_off = o; o += (uint32_t)k_s * 16u;
    s.tknots_off = o; o += (uint32_t)k_t * 8u;
    o = (o + 15u) & ~15u;
    s.list_off = o; o += TILE_M * LT_MAX_LEVELS * LT_LIST * 4u;          // 48 KB
    s.cnt_off = o; o += TILE_M * LT_MAX_LEVELS * 4u;
    s.axis_off = o; o += 2u * LT_MAX_LEVELS * LT_AXIS * 4u;
    s.red_off = o; o += LT_CG * TILE_M * RED_STRIDE * 4u;                // LayerNorm / head partials (24 KB)
    s.total = o + 1024;
    return s;
}

template <int FN>
__global__ void __launch_bounds__(LT_NT, 1) layer_fwd_lattice_kernel(const __grid_constant__ FwdK P,
                                                                     const __grid_constant__ LatP Lt, int ns) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align_smem(smem_raw);
    const LatSmem sp = plan_lattice(P.n_pad, P.has_head ? P.head.q : 0, P.basis.k_s, P.basis.k_t, ns);
    float* sA = reinterpret_cast<float*>(smem + sp.a_off);
    float* sB = reinterpret_cast<float*>(smem + sp.b_off);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + sp.bar_off);
    uint64_t* empty = full + 4;
    uint64_t* accf = full + 8;
    uint64_t* kbar = accf + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + sp.tmem_off);
    float* sbias = reinterpret_cast<float*>(smem + sp.vec_off);
    float* sgam = sbias + P.n_pad;
    float* sbet = sgam + P.n_pad;
    float* shw = reinterpret_cast<float*>(smem + sp.headw_off);
    float* shb = shw + (P.has_head ? P.head.q * P.n_pad : 0);
    float4* sk = reinterpret_cast<float4*>(smem + sp.knots_off);
    float2* st = reinterpret_cast<float2*>(smem + sp.tknots_off);
    uint32_t* slist = reinterpret_cast<uint32_t*>(smem + sp.list_off);
    int* scnt = reinterpret_cast<int*>(smem + sp.cnt_off);
    float* sax = reinterpret_cast<float*>(smem + sp.axis_off);           // [level][LT_AXIS] x, then [level][LT_AXIS] y
    float* say = sax + LT_MAX_LEVELS * LT_AXIS;
    float* red = reinterpret_cast<float*>(smem + sp.red_off);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_pad = P.n_pad, n_out = P.L.n_out, k_slabs = P.k_slabs;
    const bool has_ln = P.L.gamma != nullptr;
    const size_t b_stage_floats = (size_t)n_pad * SLAB_K;

    if (tid == LT_NW) {
        for (int s = 0; s < ns; ++s) {
            mbar_init(&full[s], 1 + LT_NW);
            mbar_init(&empty[s], 1);
        }
        mbar_init(accf, 1);
        mbar_init(kbar, 1);
        mbar_fence_init();
        stage_knots_async(P.basis, sk, st, kbar);
    }
    if (warp == 4 * LT_CG) {
        __syncwarp();
        tmem_alloc(tmem_slot, (uint32_t)P.tmem_cols);
    }
    for (int i = tid; i < n_pad; i += LT_NT) {
        const bool ok = i < n_out;
        sbias[i] = ok ? P.L.bias[i] : 0.0f;
        sgam[i] = (ok && has_ln) ? P.L.gamma[i] : 1.0f;
        sbet[i] = (ok && has_ln) ? P.L.beta[i] : 0.0f;
    }
    if (P.has_head) {
        for (int i = tid; i < P.head.q * n_pad; i += LT_NT) {
            const int k = i / n_pad, c = i - k * n_pad;
            shw[i] = c < n_out ? P.head.w[(size_t)k * n_out + c] : 0.0f;
        }
        if (tid < P.head.q) shb[tid] = P.head.b[tid];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4 * LT_CG) {
        // ---------------- producer: W1 slabs of every tile
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x)
                for (int s = 0; s < k_slabs; ++s, ++it) {
                    const uint32_t stage = it % (uint32_t)ns, use = it / (uint32_t)ns;
                    if (use > 0) mbar_wait(&empty[stage], (use - 1) & 1u);
                    issue_slab_copies(P.L.w_img, k_slabs, s, n_pad, nullptr, nullptr, sB + stage * b_stage_floats, &full[stage]);
                }
        }
        __syncwarp();
    } else if (warp == 4 * LT_CG + 1) {
        // ---------------- MMA issuer
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32((uint32_t)n_pad, 0, 0);
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
                for (int s = 0; s < k_slabs; ++s, ++it) {
                    const uint32_t stage = it % (uint32_t)ns, use = it / (uint32_t)ns;
                    mbar_wait(&full[stage], use & 1u);
                    tc_fence_after();
                    issue_slab_mma(tmem_base, sA + (size_t)stage * SLAB_FLOATS, sB + stage * b_stage_floats, idesc, s == 0);
                    umma_commit(&empty[stage]);
                }
                umma_commit(accf);
            }
        }
        __syncwarp();
    } else {
        // ---------------- workers
        const int q4 = warp & 3, cg = warp >> 2;
        const int row = q4 * 32 + lane;
        mbar_wait(kbar, 0);
        // lattice coordinates per level and axis, taken from the staged knot table (so they ARE the knots' coordinates)
        for (int i = tid; i < Lt.n_levels * LT_AXIS; i += LT_NW) {
            const int l = i / LT_AXIS, a = i - l * LT_AXIS;
            if (a < Lt.side[l]) {
                sax[l * LT_AXIS + a] = sk[Lt.offset[l] + a * Lt.side[l]].x;
                say[l * LT_AXIS + a] = sk[Lt.offset[l] + a].y;
            }
        }
        worker_barrier(LT_NW);
        uint32_t it = 0, tcount = 0;
        uint32_t* my_list = slist + ((size_t)row * LT_MAX_LEVELS + cg) * LT_LIST;
        for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++tcount) {
            const long long lrow = (long long)tile * TILE_M + row;
            const bool rvalid = lrow < P.pts.n_rows;
            const long long grow = P.pts.row_begin + lrow;
            float x = 0.f, y = 0.f, t = 0.f;
            if (rvalid) load_point(P.pts, grow, x, y, t);
            // ---- phase A: column group l walks level l of this row into the row's list
            if (cg < Lt.n_levels) {
                const int l = cg, side = Lt.side[l], off = Lt.offset[l];
                const float th2 = Lt.th2[l], ith = Lt.inv_th[l], thg = Lt.thg[l] + 1e-4f, g = (float)(side - 1);
                const float fx = x * g, fy = y * g;
                const int ix0 = max(0, (int)ceilf(fx - thg)), ix1 = min(side - 1, (int)floorf(fx + thg));
                const int iy0 = max(0, (int)ceilf(fy - thg)), iy1 = min(side - 1, (int)floorf(fy + thg));
                float dx2[LT_WIN], dy2[LT_WIN];
#pragma unroll
                for (int i = 0; i < LT_WIN; ++i) {
                    const float cx = sax[l * LT_AXIS + min(ix0 + i, side - 1)], cy = say[l * LT_AXIS + min(iy0 + i, side - 1)];
                    const float dx = x - cx, dy = y - cy;
                    dx2[i] = (ix0 + i <= ix1) ? __fmul_rn(dx, dx) : 1e30f;
                    dy2[i] = (iy0 + i <= iy1) ? __fmul_rn(dy, dy) : 1e30f;
                }
                int cnt = 0;
                if (rvalid) {
#pragma unroll
                    for (int i = 0; i < LT_WIN; ++i) {
                        const int cbase = off + (ix0 + i) * side + iy0;
#pragma unroll
                        for (int j = 0; j < LT_WIN; ++j) {
                            const float d2 = __fadd_rn(dx2[i], dy2[j]);
                            if (d2 < th2 && cnt < LT_LIST) {
                                const float val = to_tf32(phi_from_d2<FN>(d2, th2, ith));
                                my_list[cnt++] = (__float_as_uint(val) & 0xFFFFE000u) | (uint32_t)(cbase + j);
                            }
                        }
                    }
                }
                scnt[row * LT_MAX_LEVELS + l] = cnt;
            }
            worker_barrier(LT_NW);
            // ---- phase B: column group s % 4 assembles operand slab s of its row
            for (int s = 0; s < k_slabs; ++s, ++it) {
                const uint32_t stage = it % (uint32_t)ns, use = it / (uint32_t)ns;
                // every thread waits for the stage to be free before it arrives for its next use (an arrival must never
                // land in a phase that has not started), the owner group then assembles the slab
                if (use > 0) mbar_wait(&empty[stage], (use - 1) & 1u);
                if ((s & (LT_CG - 1)) == cg) {
                    const uint32_t slab_saddr = smem_u32(sA + (size_t)stage * SLAB_FLOATS);
                    const int c_lo = s * SLAB_K, c_hi = min(c_lo + SLAB_K, Lt.ks_aligned);
#pragma unroll 1
                    for (int c = 0; c < 8; ++c) {
                        const int f = c_lo + 4 * c;
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (f + 4 > Lt.ks_aligned && rvalid) v = feature_chunk(P.basis, sk, st, f, x, y, t, nullptr);
                        st_shared_v4(slab_saddr + swz_off((uint32_t)row, (uint32_t)c), v.x, v.y, v.z, v.w);
                    }
                    for (int l = 0; l < Lt.n_levels; ++l) {
                        if (Lt.offset[l] >= c_hi || Lt.offset[l] + Lt.side[l] * Lt.side[l] <= c_lo) continue;
                        const uint32_t* lst = slist + ((size_t)row * LT_MAX_LEVELS + l) * LT_LIST;
                        const int n = scnt[row * LT_MAX_LEVELS + l];
                        for (int e = 0; e < n; ++e) {
                            const uint32_t ent = lst[e];
                            const int col = (int)(ent & 0x1FFFu);
                            if (col >= c_lo && col < c_hi) {
                                const uint32_t cc = (uint32_t)(col - c_lo);
                                asm volatile("st.shared.u32 [%0], %1;" ::"r"(slab_saddr + swz_off((uint32_t)row, cc >> 2) + (cc & 3u) * 4u),
                                             "r"(ent & 0xFFFFE000u)
                                             : "memory");
                            }
                        }
                    }
                    fence_proxy_async_smem();
                }
                mbar_arrive(&full[stage]);
            }
            // ---- epilogue
            mbar_wait(accf, tcount & 1u);
            tc_fence_after();
            EpiCtx E{sbias, sgam, sbet, shw, shb, red, tmem_base + ((uint32_t)(q4 * 32) << 16), tile, row, cg, lane, lrow, grow,
                     rvalid, true};
            fwd_epilogue<LT_CG>(P, E);
            tc_fence_before();
            worker_barrier(LT_NW);      // every thread has read the accumulator and the lists before the next tile reuses them
        }
    }
    __syncthreads();
    if (warp == 4 * LT_CG) tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
}

}  // namespace stdadk

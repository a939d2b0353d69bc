// Space-time FIELD prediction (plot_spatial_mse / plot_temporal_series of upstream, scripts/train_st_interp.py:1233-1248,
// :1380-1394, and the dense grid of BASELINE config 3): every one of S sites is predicted at T time steps.
//
// The first Linear layer is linear in its input [phi(s) | psi(t)], so its pre-activation separates:
//     x1(s, t) = W1[:, spatial] phi(s)  +  ( W1[:, temporal] psi(t) + b1 )  =  zs(s) + zt(t)
// zs depends on the site only, zt on the time step only.  A CTA therefore takes a tile of 128 SITES, evaluates the basis
// and runs the block-1 GEMM ONCE (zs stays in tensor memory, 256 columns), and then loops over the time steps of its
// chunk: per (site, time) point only the add, LayerNorm/ReLU and blocks 2.. remain -- 0.20 instead of 0.35 MFLOP and no
// basis evaluation per point.  zt (T x n_1 floats) comes from a tiny pre-kernel and is streamed row by row into shared
// memory with bulk async copies.  The function computed is STInterpMLP.forward (st_interp.py:827-882) on the T x S points;
// only the order of two FP32 additions differs from the generic kernel (zs + zt instead of one accumulation).
//
// Roles as in predict.cuh: 16 worker warps (thread = row x 1/4 of the columns), one producer warp (weight slabs + zt
// rows), one MMA-issuer warp.
// The shared-memory pipe is what bounds this kind of kernel (weight slabs written by TMA and read back by the tensor
// core every time step), so the hidden activations do NOT go through it: the normalised output H of a block is written
// to TENSOR MEMORY (tcgen05.st, 32-column slabs, one mbarrier each) and the next block's MMAs take their A operand from
// there.  TMEM: columns [0, 256) = accumulator (zs of the tile first, then blocks 2..), [256, 512) = H.  zs itself is
// parked in shared memory in the slots the block-1 operand occupied -- every thread reads back only what it wrote.
// Per time step: E1 (zs + zt -> LN -> ReLU -> H) -> MMA block 2 (starts per slab) -> E2 -> MMA block 3 -> E3 + head.
#pragma once
#include "predict.cuh"

namespace stdadk {

constexpr int FD_ZT_RING = 3;           // zt rows staged ahead in shared memory (<= 1 KB each)
constexpr int FD_ZT_AHEAD = 2;          // rows requested before they are needed (< ring size)

struct FieldK {
    BasisP basis;                       // k_t = 0, p_cov = 0: the block-1 operand holds the spatial columns only
    const float* sites;                 // (S, 2) explicit sites, or NULL for the lattice (nx, ny): site = i * ny + j
    int nx, ny;
    long long n_sites;                  // S
    long long site_begin, site_end;     // sites of this launch
    int k_begin, k_end;                 // time steps of this launch
    int k_chunk, n_kchunks;             // a work unit = (site tile, k_chunk consecutive time steps)
    int n_site_tiles, n_units;
    PredLayerP L[PF_MAX_LAYERS];        // L[0].w_img = image of W1[:, spatial] (n_1 x k_s); L[0].bias is NOT used (it is in zt)
    PredSmem sm;
    uint32_t zt_off, hscr_off;          // zt ring, head scratch
    const float* zt;                    // (T, n_pad of block 1): W1[:, temporal] psi(t_k) + b1, zero in the padding columns
    const float* head_w;
    const float* head_b;
    float* yhat;                        // row k * out_k_stride + site - row_base
    long long row_base, out_k_stride;
    int n_layers, q;
    unsigned long long* dbg;            // optional cycle counters (-DSTDADK_PF_DEBUG builds), NULL in production
};

__host__ inline uint32_t plan_field(FieldK& K) {
    uint32_t o = 0;
    K.sm.h_off = o; o += PF_HSLABS * SLAB_BYTES;                       // 128 KB
    K.sm.w_off = o; o += PF_WST * MAX_N * 128u;                        // 64 KB
    K.sm.bar_off = o; o += 320;
    K.sm.tmem_off = o; o += 16;
    for (int l = 0; l < K.n_layers; ++l) {
        K.L[l].prm_off = o;
        o += 3u * (uint32_t)K.L[l].n_pad * 4u;
    }
    K.sm.headw_off = o; o += (uint32_t)(K.q * K.L[K.n_layers - 1].n_pad + STDADK_MAX_Q) * 4u;
    o = (o + 15u) & ~15u;
    K.sm.knots_off = o; o += (uint32_t)K.basis.k_s * 16u;
    K.sm.tknots_off = o;
    o = (o + 15u) & ~15u;
    // LayerNorm partials, single buffer (8 KB): between two uses every thread passes either an accumulator wait that
    // needs all threads' H-slab arrivals, or the head barrier -- see the kernel
    K.sm.red_off = o; o += PF_CG * TILE_M * 16u;
    K.zt_off = o; o += FD_ZT_RING * (uint32_t)K.L[0].n_pad * 4u;
    K.hscr_off = o; o += (PF_CG - 1) * TILE_M * (uint32_t)K.q * 4u;    // head partials of column groups 1..3
    o = (o + 15u) & ~15u;
    K.sm.total = o + 1024;
    return K.sm.total;
}

// zt[k, c] = b1[c] + sum_j W1[c, col0 + j] * psi_j(t_k), t_k = k / (T - 1); FP32 throughout (T x n_1 outputs of k_t terms)
__global__ void field_zt_kernel(const float* __restrict__ w1, long long w_row_stride, long long w_col_stride, int col0,
                                const float* __restrict__ b1, const float2* __restrict__ tknots, int k_t, int n_out, int n_pad,
                                int T, float* __restrict__ zt) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= T * n_pad) return;
    const int k = idx / n_pad, c = idx - k * n_pad;
    float acc = 0.0f;
    if (c < n_out) {
        const float t = T > 1 ? __fdiv_rn((float)k, (float)(T - 1)) : 0.0f;     // as load_point()
        acc = b1[c];
        for (int j = 0; j < k_t; ++j) {
            const float2 tk = tknots[j];
            acc = fmaf(w1[(long long)c * w_row_stride + (long long)(col0 + j) * w_col_stride], psi_eval(t, tk.x, tk.y), acc);
        }
    }
    zt[idx] = acc;
}

// one thread's 32 consecutive columns of its row <-> its 128-byte row of an operand slab (conflict-free swizzle), FP32 as is
__device__ __forceinline__ void pf_store_raw(const float (&v)[32], uint32_t slab_saddr, uint32_t rowoff, uint32_t rx) {
#pragma unroll
    for (int c = 0; c < 8; ++c)
        st_shared_v4(slab_saddr + rowoff + (((uint32_t)c ^ rx) << 4), v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
}
__device__ __forceinline__ void pf_load_raw(float (&v)[32], uint32_t slab_saddr, uint32_t rowoff, uint32_t rx) {
#pragma unroll
    for (int c = 0; c < 8; ++c)
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(v[4 * c]), "=f"(v[4 * c + 1]), "=f"(v[4 * c + 2]), "=f"(v[4 * c + 3])
                     : "r"(slab_saddr + rowoff + (((uint32_t)c ^ rx) << 4)));
}
__device__ __forceinline__ void pf_round_tf32(float (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = to_tf32(v[i]);
}

__global__ void __launch_bounds__(PF_NT, 1) predict_field_kernel(const __grid_constant__ FieldK P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align_smem(smem_raw);
    float* sH = reinterpret_cast<float*>(smem + P.sm.h_off);
    float* sW = reinterpret_cast<float*>(smem + P.sm.w_off);
    uint64_t* wfull = reinterpret_cast<uint64_t*>(smem + P.sm.bar_off);
    uint64_t* wempty = wfull + PF_WST;
    uint64_t* hfull = wempty + PF_WST;              // [8] H slab s written (block-1 operand, or a block's normalised output)
    uint64_t* accf = hfull + PF_HSLABS;             // accumulator ready (zs, then block 2, block 3, ... of every time step)
    uint64_t* kbar = accf + 1;                      // knot table has landed
    uint64_t* ztfull = kbar + 1;                    // [FD_ZT_RING]
    uint64_t* ztempty = ztfull + FD_ZT_RING;        // [FD_ZT_RING]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + P.sm.tmem_off);
    float* shw = reinterpret_cast<float*>(smem + P.sm.headw_off);
    float4* sk = reinterpret_cast<float4*>(smem + P.sm.knots_off);
    float2* st = reinterpret_cast<float2*>(smem + P.sm.tknots_off);
    float4* red = reinterpret_cast<float4*>(smem + P.sm.red_off);
    float* szt = reinterpret_cast<float*>(smem + P.zt_off);
    float* hscr = reinterpret_cast<float*>(smem + P.hscr_off);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nl = P.n_layers;
    const int last_pad = P.L[nl - 1].n_pad, last_out = P.L[nl - 1].n_out;
    const int pad0 = P.L[0].n_pad;
    float* shb = shw + P.q * last_pad;

    if (tid == PF_NW) {
        for (int s = 0; s < PF_WST; ++s) {
            mbar_init(&wfull[s], 1);
            mbar_init(&wempty[s], 1);
        }
        for (int s = 0; s < PF_HSLABS; ++s) mbar_init(&hfull[s], TILE_M / 32);
        mbar_init(accf, 1);
        mbar_init(kbar, 1);
        for (int s = 0; s < FD_ZT_RING; ++s) {
            mbar_init(&ztfull[s], 1);
            mbar_init(&ztempty[s], 1);
        }
        mbar_fence_init();
        stage_knots_async(P.basis, sk, st, kbar);
    }
    if (warp == 4 * PF_CG) {
        __syncwarp();
        tmem_alloc(tmem_slot, 512u);
    }
    for (int l = 0; l < nl; ++l) {
        const PredLayerP& Ly = P.L[l];
        float* prm = reinterpret_cast<float*>(smem + Ly.prm_off);
        for (int i = tid; i < Ly.n_pad; i += PF_NT) {
            const bool ok = i < Ly.n_out;
            prm[i] = (ok && l > 0) ? Ly.bias[i] : 0.0f;                 // block 1's bias is part of zt
            prm[Ly.n_pad + i] = (ok && Ly.has_ln) ? Ly.gamma[i] : 1.0f;
            prm[2 * Ly.n_pad + i] = (ok && Ly.has_ln) ? Ly.beta[i] : 0.0f;
        }
    }
    for (int i = tid; i < P.q * last_pad; i += PF_NT) {
        const int k = i / last_pad, c = i - k * last_pad;
        shw[i] = c < last_out ? P.head_w[(size_t)k * last_out + c] : 0.0f;
    }
    if (tid < P.q) shb[tid] = P.head_b[tid];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t zt_row_bytes = (uint32_t)pad0 * 4u;

    if (warp == 4 * PF_CG) {
        // ---------------- producer: weight slabs in consumption order + one zt row per time step
        if (lane == 0) {
            uint32_t stage = 0, phase = 0, zi = 0;
            auto push_w = [&](const PredLayerP& Ly) {
                for (int s = 0; s < Ly.k_slabs; ++s) {
                    mbar_wait(&wempty[stage], phase ^ 1u);
                    issue_slab_copies(Ly.w_img, Ly.k_slabs, s, Ly.n_pad, nullptr, nullptr,
                                      sW + (size_t)stage * MAX_N * SLAB_K, &wfull[stage]);
                    if (++stage == PF_WST) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            };
            auto push_zt = [&](int k) {          // rows are requested FD_ZT_AHEAD steps early: E1 never waits for HBM
                const uint32_t zs = zi % FD_ZT_RING, zph = (zi / FD_ZT_RING) & 1u;
                ++zi;
                mbar_wait(&ztempty[zs], zph ^ 1u);
                mbar_arrive_expect_tx(&ztfull[zs], zt_row_bytes);
                bulk_g2s(szt + (size_t)zs * pad0, P.zt + (size_t)k * pad0, zt_row_bytes, &ztfull[zs]);
            };
            for (int u = blockIdx.x; u < P.n_units; u += gridDim.x) {
                const int kc = u % P.n_kchunks;
                const int k0 = P.k_begin + kc * P.k_chunk, k1 = min(P.k_end, k0 + P.k_chunk);
                for (int k = k0; k < min(k1, k0 + FD_ZT_AHEAD); ++k) push_zt(k);
                push_w(P.L[0]);
                for (int k = k0; k < k1; ++k) {
                    if (k + FD_ZT_AHEAD < k1) push_zt(k + FD_ZT_AHEAD);
                    for (int l = 1; l < nl; ++l) push_w(P.L[l]);
                }
            }
        }
        __syncwarp();
    } else if (warp == 4 * PF_CG + 1) {
        // ---------------- MMA issuer
        if (lane == 0) {
            uint32_t wstage = 0, wphase = 0, hph = 0;
            PF_DBG(unsigned long long w_a = 0, w_w = 0; const long long t_begin = clock64();)
            auto run_block = [&](const PredLayerP& Ly, bool a_in_tmem) {
                const uint32_t idesc = umma_idesc_tf32((uint32_t)Ly.n_pad, 0, 0);
                for (int s = 0; s < Ly.k_slabs; ++s) {
                    PF_TIMED_WAIT(w_a, mbar_wait(&hfull[s], (hph >> s) & 1u));
                    hph ^= 1u << s;
                    PF_TIMED_WAIT(w_w, mbar_wait(&wfull[wstage], wphase));
                    tc_fence_after();
                    const float* sB = sW + (size_t)wstage * MAX_N * SLAB_K;
                    if (a_in_tmem) {          // A = H slab s in tensor memory: 32 columns, four K = 8 steps
                        const uint64_t bd = umma_desc_sw128(smem_u32(sB), 16, 1024);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_tf32_ts(tmem_base, tmem_base + MAX_N + (uint32_t)(s * SLAB_K + k * 8), bd + (uint64_t)(k * 2), idesc,
                                         (s == 0 && k == 0) ? 0u : 1u);
                    } else {
                        issue_slab_mma(tmem_base, sH + (size_t)s * SLAB_FLOATS, sB, idesc, s == 0);
                    }
                    umma_commit(&wempty[wstage]);
                    if (++wstage == PF_WST) {
                        wstage = 0;
                        wphase ^= 1u;
                    }
                }
                umma_commit(accf);
            };
            for (int u = blockIdx.x; u < P.n_units; u += gridDim.x) {
                const int kc = u % P.n_kchunks;
                const int k0 = P.k_begin + kc * P.k_chunk, k1 = min(P.k_end, k0 + P.k_chunk);
                run_block(P.L[0], false);                                        // zs of this site tile (A = phi slabs in SMEM)
                for (int k = k0; k < k1; ++k)
                    for (int l = 1; l < nl; ++l) run_block(P.L[l], true);
            }
            PF_DBG(if (P.dbg) {
                atomicAdd(&P.dbg[0], w_a);
                atomicAdd(&P.dbg[1], w_w);
                atomicAdd(&P.dbg[2], (unsigned long long)(clock64() - t_begin));
            })
        }
        __syncwarp();
    } else {
        // ---------------- workers: thread = (row = site of the tile, column group)
        const int q4 = warp & 3, cg = warp >> 2;
        const int row = q4 * 32 + lane;
        const uint32_t rowoff = (uint32_t)row * 128u, rx = (uint32_t)row & 7u;
        const uint32_t trow = tmem_base + ((uint32_t)(q4 * 32) << 16);
        const uint32_t sH_addr = smem_u32(sH);
        uint32_t accw = 0, zi = 0;              // accumulator-ready phases waited for; zt rows consumed
#ifdef STDADK_PF_DEBUG
        unsigned long long w_acc = 0, w_zt = 0, w_bar = 0, ph_gen = 0, ph_ld = 0, ph_norm = 0, ph_head = 0;
        const long long t_begin = clock64();
        long long t_last = t_begin;
#endif
        mbar_wait(kbar, 0);
        for (int u = blockIdx.x; u < P.n_units; u += gridDim.x) {
            const int tile = u / P.n_kchunks, kc = u - tile * P.n_kchunks;
            const int k0 = P.k_begin + kc * P.k_chunk, k1 = min(P.k_end, k0 + P.k_chunk);
            const long long site = P.site_begin + (long long)tile * TILE_M + row;
            const bool rvalid = site < P.site_end;
            // ---- block-1 operand of the tile: phi of this row's site, slabs cg, cg + 4 (H is free: the last block of the
            // previous unit has completed, its accumulator was waited for below)
            {
                float x = 0.f, y = 0.f;
                if (rvalid) {
                    if (P.sites) {
                        const float2 c = *reinterpret_cast<const float2*>(P.sites + 2 * site);
                        x = c.x;
                        y = c.y;
                    } else {
                        const long long i = site / P.ny, j = site - i * P.ny;
                        x = P.nx > 1 ? __fdiv_rn((float)i, (float)(P.nx - 1)) : 0.0f;
                        y = P.ny > 1 ? __fdiv_rn((float)j, (float)(P.ny - 1)) : 0.0f;
                    }
                }
                for (int s = cg; s < P.L[0].k_slabs; s += PF_CG) {
                    pf_gen_slab(P.basis, sk, st, s, x, y, 0.0f, nullptr, sH_addr + (uint32_t)s * SLAB_BYTES + rowoff, rx);
                    fence_proxy_async_smem();
                    mbar_arrive_warp(&hfull[s]);
                }
            }
            PF_PHASE(ph_gen);
            PF_TIMED_WAIT(w_acc, mbar_wait(accf, accw & 1u));            // zs complete
            PF_DBG(if (P.dbg) t_last = clock64();)
            ++accw;
            tc_fence_after();
            // park this thread's 64 columns of zs in the (now dead) operand slabs cg, cg + 4 of its own row; the accumulator
            // columns are then free for blocks 2.. (the barrier below: every thread has read them before an MMA writes)
            {
                float z[32];
                if (32 * cg < pad0) {
                    tmem_ld32(trow + (uint32_t)(32 * cg), z);
                    pf_store_raw(z, sH_addr + (uint32_t)cg * SLAB_BYTES, rowoff, rx);
                }
                if (32 * cg + 128 < pad0) {
                    tmem_ld32(trow + (uint32_t)(32 * cg + 128), z);
                    pf_store_raw(z, sH_addr + (uint32_t)(cg + 4) * SLAB_BYTES, rowoff, rx);
                }
            }
            tc_fence_before();
            worker_barrier(PF_NW);
            for (int k = k0; k < k1; ++k, ++zi) {
                const uint32_t zs_slot = zi % FD_ZT_RING, zph = (zi / FD_ZT_RING) & 1u;
                for (int l = 0; l < nl; ++l) {
                    const PredLayerP& Ly = P.L[l];
                    const float* sb = reinterpret_cast<const float*>(smem + Ly.prm_off);
                    const float* sg = sb + Ly.n_pad;
                    const float* sbt = sg + Ly.n_pad;
                    const int c0a = 32 * cg, c0b = 32 * cg + 128;
                    const bool ha = c0a < Ly.n_pad, hb = c0b < Ly.n_pad;
                    const int nva = min(32, Ly.n_out - c0a), nvb = min(32, Ly.n_out - c0b);
                    float4* redl = red;
                    const float* addv = szt + (size_t)zs_slot * pad0;      // block 1: "bias" = zt[k]
                    if (l == 0) {
                        PF_TIMED_WAIT(w_zt, mbar_wait(&ztfull[zs_slot], zph));
                        PF_DBG(if (P.dbg) t_last = clock64();)
                    } else {
                        PF_TIMED_WAIT(w_acc, mbar_wait(accf, accw & 1u));
                        PF_DBG(if (P.dbg) t_last = clock64();)
                        ++accw;
                        tc_fence_after();
                        addv = sb;
                    }
                    float va[32], vb[32];
                    if (l == 0) {                 // zs from this thread's own shared-memory slots
                        if (ha) pf_load_raw(va, sH_addr + (uint32_t)cg * SLAB_BYTES, rowoff, rx);
                        if (hb) pf_load_raw(vb, sH_addr + (uint32_t)(cg + 4) * SLAB_BYTES, rowoff, rx);
                    } else {
                        if (ha) tmem_ld32_issue(trow + (uint32_t)c0a, va);
                        if (hb) tmem_ld32_issue(trow + (uint32_t)c0b, vb);
                    }
                    if (!ha) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) va[i] = 0.0f;
                    }
                    if (!hb) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) vb[i] = 0.0f;
                    }
                    if (l != 0) {
                        tmem_ld_wait(va);
                        tmem_ld_wait(vb);
                    }
                    bool have = false;
                    float K = 0.0f, S1 = 0.0f, S2 = 0.0f;
                    if (ha) pf_bias_stats(va, addv + c0a, nva, have, K, S1, S2);
                    if (hb) pf_bias_stats(vb, addv + c0b, nvb, have, K, S1, S2);
                    float rstd = 1.0f, nmr = 0.0f;
                    if (Ly.has_ln) {
                        const float cntv = (float)((ha ? max(nva, 0) : 0) + (hb ? max(nvb, 0) : 0));
                        redl[cg * TILE_M + row] = make_float4(K, S1, S2, cntv);
                    }
                    tc_fence_before();
                    PF_PHASE(ph_ld);
                    PF_TIMED_WAIT(w_bar, worker_barrier(PF_NW));   // LN partials visible; everyone has read this accumulator / zt row
                    PF_DBG(if (P.dbg) t_last = clock64();)
                    if (l == 0 && tid == 0) mbar_arrive(&ztempty[zs_slot]);
                    if (Ly.has_ln && ha) {
                        const float inv_n = 1.0f / (float)Ly.n_out;
                        float4 part[PF_CG];
                        float tot = 0.0f;
#pragma unroll
                        for (int g = 0; g < PF_CG; ++g) {
                            part[g] = redl[g * TILE_M + row];
                            tot += fmaf(part[g].w, part[g].x, part[g].y);
                        }
                        const float mean = tot * inv_n;
                        float m2 = 0.0f;
#pragma unroll
                        for (int g = 0; g < PF_CG; ++g) {
                            const float dk = mean - part[g].x;
                            m2 += part[g].z - 2.0f * dk * part[g].y + part[g].w * dk * dk;
                        }
                        rstd = 1.0f / sqrtf(fmaxf(m2 * inv_n, 0.0f) + Ly.eps);
                        nmr = -mean * rstd;
                    }
                    if (l + 1 < nl) {
                        if (ha) {
                            pf_normalize(va, sg + c0a, sbt + c0a, Ly.has_ln != 0, rstd, nmr, nva);
                            pf_round_tf32(va);
                            tmem_st32(trow + MAX_N + (uint32_t)c0a, va);
                            tmem_st_wait();
                            tc_fence_before();
                            mbar_arrive_warp(&hfull[cg]);
                        }
                        if (hb) {
                            pf_normalize(vb, sg + c0b, sbt + c0b, Ly.has_ln != 0, rstd, nmr, nvb);
                            pf_round_tf32(vb);
                            tmem_st32(trow + MAX_N + (uint32_t)c0b, vb);
                            tmem_st_wait();
                            tc_fence_before();
                            mbar_arrive_warp(&hfull[cg + 4]);
                        }
                        PF_PHASE(ph_norm);
                    } else {
                        float yh[STDADK_MAX_Q];
#pragma unroll
                        for (int kk = 0; kk < STDADK_MAX_Q; ++kk) yh[kk] = 0.0f;
                        if (ha) {
                            pf_normalize(va, sg + c0a, sbt + c0a, Ly.has_ln != 0, rstd, nmr, nva);
                            pf_head_partial(va, shw, Ly.n_pad, c0a, P.q, yh);
                        }
                        if (hb) {
                            pf_normalize(vb, sg + c0b, sbt + c0b, Ly.has_ln != 0, rstd, nmr, nvb);
                            pf_head_partial(vb, shw, Ly.n_pad, c0b, P.q, yh);
                        }
                        if (cg > 0) {
                            float* mine = hscr + ((size_t)(cg - 1) * TILE_M + row) * P.q;
#pragma unroll
                            for (int kk = 0; kk < STDADK_MAX_Q; ++kk)
                                if (kk < P.q) mine[kk] = yh[kk];
                        }
                        worker_barrier(PF_NW);
                        if (cg == 0 && rvalid) {
                            float* dst = P.yhat + ((long long)k * P.out_k_stride + site - P.row_base) * P.q;
#pragma unroll
                            for (int kk = 0; kk < STDADK_MAX_Q; ++kk)
                                if (kk < P.q) {
                                    float a = yh[kk] + shb[kk];
#pragma unroll
                                    for (int g = 0; g < PF_CG - 1; ++g) a += hscr[((size_t)g * TILE_M + row) * P.q + kk];
                                    dst[kk] = a;
                                }
                        }
                        // the head scratch is read above by the cg == 0 warps only after the barrier, and rewritten only
                        // after the NEXT step's two LayerNorm barriers: no further synchronisation is needed
                        PF_PHASE(ph_head);
                    }
                }
            }
        }
        PF_DBG(if (P.dbg && tid == 0) {
            atomicAdd(&P.dbg[3], w_acc);
            atomicAdd(&P.dbg[4], w_zt);
            atomicAdd(&P.dbg[5], (unsigned long long)(clock64() - t_begin));
            atomicAdd(&P.dbg[6], w_bar);
            atomicAdd(&P.dbg[8], ph_gen);
            atomicAdd(&P.dbg[9], ph_ld);
            atomicAdd(&P.dbg[10], ph_norm);
            atomicAdd(&P.dbg[11], ph_head);
        })
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 4 * PF_CG) tmem_dealloc(tmem_base, 512u);
}

}  // namespace stdadk

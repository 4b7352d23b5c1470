// vfk_split.cuh -- the lane-split shape of the fused control-cycle kernel: L lanes cooperate on ONE instance.
//
// The one-thread-per-instance kernel (vfk_kernels.cuh) keeps the whole 6 x N Jacobian of its instance in registers.  For a
// 17-joint chain that is 102 registers before anything else (248 in all: 2 CTAs per SM, 8 warps, latency-bound at 0.58 of
// the HBM roofline), and in FP64 even 7 joints cost 234.  Here the JOINTS of an instance are split over L = 2 (or 4) lanes:
//
//   * lane h of an instance owns joints [h * NL, (h + 1) * NL), NL = ceil(N / L): their angles, their Jacobian columns,
//     their share of J J^T, J x and J^T y, their limits, their outputs -- a warp takes 32 / L instances of a tile;
//   * forward kinematics runs as L partial chains in parallel, each from the identity (lane 0: from the base frame); the
//     partial frames are combined by a shuffle scan (frame product is associative), which hands every lane the frame its
//     partial chain starts from and all lanes the flange frame; joint axes and origins are then moved to the base frame;
//   * the 21 + 6 partial sums of J J^T and J x are added across the L lanes by xor-shuffles; the 6 x 6 Cholesky and the
//     triangular solves are repeated by every lane (same instructions whether one lane or 32 execute them);
//   * the obstacle pairs of a chunk are dealt round-robin to the L lanes and the three partial repulsor sums shuffled
//     together (north_star's "reduced with warp shuffles"); the all-or-nothing limit check and the leading-joint clamp
//     combine their per-lane verdicts the same way.
//
// Staging: a warp needs the rows of HALF (a quarter) of a tile -- 32 / L lanes of every 32-lane row.  Those are strided
// pieces, so the 1-D bulk copies of the full-tile kernel would become one copy per row; instead the arrays are described to
// the TMA unit as 3-D tensors [tile][row][lane] (cuTensorMapEncodeTiled on the host, one descriptor per array per launch)
// and ONE cp.async.bulk.tensor (SASS UTMALDG) per stage gathers the box {32 / L lanes, rows of the chunk, 1 tile} into a
// dense shared-memory tile.  Same mbarrier ring as the full-tile kernel, one tile-part ahead of the arithmetic.
//
// Scope: the production ("lean") shape -- identity tool frame and IK weights, projector nullspace on the built-in
// limit-avoidance gradient sharing the IK factor, q / qdot the only outputs -- for chains in the generic pattern.  Anything
// else runs the one-thread-per-instance kernel.  Robot constants are indexed by a lane-dependent joint number here, which
// the constant bank serialises; they are copied once per CTA into a shared-memory table instead.
#pragma once
#include <cuda.h>

#include "vfk_kernels.cuh"

namespace vfk {

struct alignas(64) SplitMaps {
    CUtensorMap q, goal, obst;
};

constexpr int kSplitTabJoints = kMaxJ + 3;       // L * NL can exceed N by up to L - 1 padding joints

template <typename T>
struct SplitTab {
    typename WideOf<T>::type tip[kSplitTabJoints][12];
    T q_lo[kSplitTabJoints], q_hi[kSplitTabJoints], ns_scale[kSplitTabJoints], ns_mid[kSplitTabJoints];
    int32_t prismatic[kSplitTabJoints];
};

__host__ __device__ constexpr uint32_t round128(uint32_t x) { return (x + 127u) & ~127u; }

template <typename T, int N, int L>
struct SplitShape {
    static constexpr int NL = (N + L - 1) / L;                                  // joints per lane
    static constexpr int SUB = 32 / L;                                          // instances per warp
    static constexpr uint32_t kRowBytes = SUB * 16;                             // one obstacle plane row of the warp's instances
    static constexpr int kRowsPerChunk = kChunk / 2 * ObstPairs<T>::kPlanes;    // FP32: 8, FP64: 16
    static constexpr uint32_t kStage = kRowsPerChunk * kRowBytes;
    static constexpr uint32_t kQ = round128(N * SUB * sizeof(T));
    static constexpr uint32_t kG = round128(13 * SUB * sizeof(T));
    static constexpr uint32_t kQg = kQ + kG;
    static constexpr int kBars = kMaxStages + 2;
    static constexpr uint32_t kTab = round128(sizeof(SplitTab<T>));
    __host__ __device__ static constexpr uint32_t warp_bytes(int n_stages) { return n_stages * kStage + 2 * kQg; }
};

// One box of a 3-D tensor, global -> shared, completing on an mbarrier (SASS UTMALDG).
__device__ __forceinline__ void tma_box3(void* dst_smem, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_u32(dst_smem)),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}

template <typename W>
__device__ __forceinline__ void frame_mul(const W (&Ra)[9], const W (&pa)[3], const W (&Rb)[9], const W (&pb)[3], W (&R)[9], W (&p)[3]) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int cc = 0; cc < 3; ++cc)
            R[3 * r + cc] = fma(Ra[3 * r + 0], Rb[cc], fma(Ra[3 * r + 1], Rb[3 + cc], Ra[3 * r + 2] * Rb[6 + cc]));
        p[r] = fma(Ra[3 * r + 0], pb[0], fma(Ra[3 * r + 1], pb[1], fma(Ra[3 * r + 2], pb[2], pa[r])));
    }
}

template <typename T, int N, int L, int MINB>
__global__ void __launch_bounds__(kBlock, MINB)
vfk_split_kernel(const __grid_constant__ KConst<T> c, const __grid_constant__ KArgs<T> a, const __grid_constant__ SplitMaps maps) {
    using SH = SplitShape<T, N, L>;
    using W = typename WideOf<T>::type;
    using WN = typename WideNE<T>::type;
    constexpr int NL = SH::NL, SUB = SH::SUB;
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);     // tells the compiler it is warp-uniform: no divergence guards around the shuffles
    const int h = lane / SUB;                                    // which part of the chain this lane owns
    const int sl = lane % SUB;                                   // instance within the warp
    const int j0 = h * NL;                                       // first joint of this lane
    const int nc = a.n_comp;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem) + warp * SH::kBars;
    SplitTab<T>* tab = reinterpret_cast<SplitTab<T>*>(smem + kSmemHeader);
    unsigned char* region = smem + kSmemHeader + SH::kTab + (size_t)warp * SH::warp_bytes(a.n_stages);

    // per-joint robot constants -> shared memory (joints >= nc: identity tip, zero everything: padding)
    for (int j = threadIdx.x; j < L * NL; j += kBlock) {
        const bool real = j < nc;
#pragma unroll
        for (int k = 0; k < 12; ++k) tab->tip[j][k] = real ? c.tip[j][k] : ((k == 0 || k == 4 || k == 8) ? W(1) : W(0));
        tab->q_lo[j] = real ? c.q_lo[j] : T(0);
        tab->q_hi[j] = real ? c.q_hi[j] : T(0);
        tab->ns_scale[j] = real ? c.ns_q0_scale[j] : T(0);
        tab->ns_mid[j] = real ? c.ns_mid[j] : T(0);
        tab->prismatic[j] = real ? ((c.prismatic_mask >> j) & 1) : 0;
    }
    __syncthreads();

    const int64_t n_tiles = (a.n + 31) >> 5;
    const int64_t n_units = n_tiles * L;                         // a unit = the 32 / L instances one warp takes at a time
    const int64_t stride = (int64_t)gridDim.x * (kBlock / 32);
    int64_t unit = (int64_t)blockIdx.x * (kBlock / 32) + warp;
    if (unit >= n_units) return;

    const bool resident = a.n_chunks <= a.n_stages;
    const int S = resident ? a.n_chunks : a.n_stages;
    const int U = resident ? a.n_chunks : a.n_chunks * a.k_cycles;
    const uint32_t qg_bytes = (uint32_t)(nc + 13) * SUB * sizeof(T);

    auto issue_qg = [&](int64_t u, int buf) {
        if (lane == 0) {
            unsigned char* dst = region + (size_t)a.n_stages * SH::kStage + (size_t)buf * SH::kQg;
            const int tile = (int)(u / L), col = (int)(u % L) * SUB;
            mbar_arrive_expect_tx(&bars[kMaxStages + buf], qg_bytes);
            tma_box3(dst, &maps.q, col, 0, tile, &bars[kMaxStages + buf]);
            tma_box3(dst + SH::kQ, &maps.goal, col, 0, tile, &bars[kMaxStages + buf]);
        }
    };
    auto issue_obst = [&](int64_t u, int chunk, int stage) {
        if (lane == 0) {
            const int tile = (int)(u / L), col = (int)(u % L) * (int)(SH::kRowBytes / sizeof(T));
            mbar_arrive_expect_tx(&bars[stage], SH::kStage);     // the box is always whole: rows past the tile's last are zero-filled
            tma_box3(region + (size_t)stage * SH::kStage, &maps.obst, col, chunk * SH::kRowsPerChunk, tile, &bars[stage]);
        }
    };

    if (lane == 0) {
        for (int s = 0; s < kMaxStages + 2; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    __syncwarp();
    issue_qg(unit, 0);
    int64_t p_unit = unit;
    int p_u = 0, p_chunk = 0;
    for (int u = 0; u < S; ++u) {
        issue_obst(p_unit, p_chunk, u);
        if (++p_chunk == a.n_chunks) p_chunk = 0;
        if (++p_u == U) { p_u = 0; p_unit += stride; }
    }
    int c_stage = 0;
    uint32_t c_phase = 0;

    for (int it = 0; unit < n_units; unit += stride, ++it) {
        const int64_t tile = unit / L;
        const int slot = (int)(unit % L) * SUB + sl;             // instance within the tile
        const bool active = (tile << 5) + slot < a.n;
        const int64_t tN = tile * (nc * 32) + slot;
        __syncwarp();
        if (unit + stride < n_units) issue_qg(unit + stride, (it + 1) & 1);
        mbar_wait(&bars[kMaxStages + (it & 1)], (uint32_t)(it >> 1) & 1u);
        const unsigned char* qgb = region + (size_t)a.n_stages * SH::kStage + (size_t)(it & 1) * SH::kQg;
        const T* qs = reinterpret_cast<const T*>(qgb) + sl;
        const T* gs = reinterpret_cast<const T*>(qgb + SH::kQ) + sl;

        T q[NL];
#pragma unroll
        for (int k = 0; k < NL; ++k) q[k] = (j0 + k < nc) ? qs[(j0 + k) * SUB] : T(0);
        T g[13];
#pragma unroll
        for (int k = 0; k < 13; ++k) g[k] = gs[k * SUB];

        for (int cyc = 0; cyc < a.k_cycles; ++cyc) {
            const bool last = (cyc == a.k_cycles - 1);

            // 1. forward kinematics of this lane's part of the chain, from the identity (lane 0: from the base frame)
            T Jl[NL][3], Ja[NL][3];                              // first: joint origin / axis in the part's own frame
            T Rt[9];
            Pos<T> pt;
            {
                W R[9], p[3];
#pragma unroll
                for (int k = 0; k < 9; ++k) R[k] = (h == 0) ? c.base[k] : ((k % 4 == 0) ? W(1) : W(0));
#pragma unroll
                for (int k = 0; k < 3; ++k) p[k] = (h == 0) ? c.base[9 + k] : W(0);
                static_for<0, NL>([&](auto kc) {
                    constexpr int k = decltype(kc)::value;
                    const int j = j0 + k;
                    Ja[k][0] = (T)R[2]; Ja[k][1] = (T)R[5]; Ja[k][2] = (T)R[8];
                    Jl[k][0] = (T)p[0]; Jl[k][1] = (T)p[1]; Jl[k][2] = (T)p[2];
                    const W* tp = tab->tip[j];
                    const bool prism = tab->prismatic[j] != 0;
                    const W qj = (W)q[k];
                    const W qrot = prism ? W(0) : qj, qtr = prism ? qj : W(0);
                    W s, co;
                    sincos_wide<(sizeof(T) == sizeof(W)) ? 7 : 5>(qrot, &s, &co);
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        const W x = R[3 * r + 0], y = R[3 * r + 1];
                        R[3 * r + 0] = fma(co, x, s * y);
                        R[3 * r + 1] = fma(co, y, -s * x);
                        p[r] = fma(R[3 * r + 2], qtr, p[r]);
                    }
                    const W t0 = tp[9], t1 = tp[10], t2 = tp[11];
#pragma unroll
                    for (int r = 0; r < 3; ++r) p[r] = fma(R[3 * r + 0], t0, fma(R[3 * r + 1], t1, fma(R[3 * r + 2], t2, p[r])));
                    if (c.all_xtwist) {                          // every tip rotation is RotX(alpha) (alpha = 0 included)
                        const W ca = tp[4], sa = tp[7];
#pragma unroll
                        for (int r = 0; r < 3; ++r) {
                            const W y = R[3 * r + 1], z = R[3 * r + 2];
                            R[3 * r + 1] = fma(ca, y, sa * z);
                            R[3 * r + 2] = fma(ca, z, -sa * y);
                        }
                    } else {
                        W Rn[9];
#pragma unroll
                        for (int r = 0; r < 3; ++r)
#pragma unroll
                            for (int cc = 0; cc < 3; ++cc)
                                Rn[3 * r + cc] = fma(R[3 * r + 0], tp[cc], fma(R[3 * r + 1], tp[3 + cc], R[3 * r + 2] * tp[6 + cc]));
#pragma unroll
                        for (int k2 = 0; k2 < 9; ++k2) R[k2] = Rn[k2];
                    }
                });

                // inclusive scan of the partial frames over the L lanes of the instance: P_h = X_0 X_1 ... X_h
                W Rx[9], px[3];
#pragma unroll
                for (int d = 1; d < L; d <<= 1) {
                    W Ry[9], py[3];
#pragma unroll
                    for (int k = 0; k < 9; ++k) Ry[k] = __shfl_up_sync(0xffffffffu, R[k], d * SUB);
#pragma unroll
                    for (int k = 0; k < 3; ++k) py[k] = __shfl_up_sync(0xffffffffu, p[k], d * SUB);
                    frame_mul<W>(Ry, py, R, p, Rx, px);
                    const bool take = h >= d;
#pragma unroll
                    for (int k = 0; k < 9; ++k) R[k] = take ? Rx[k] : R[k];
#pragma unroll
                    for (int k = 0; k < 3; ++k) p[k] = take ? px[k] : p[k];
                }
                // the frame this lane's part starts from (lane 0: identity -- its part already began at the base) ...
                T Rs[9], ps[3];
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    const W v = __shfl_up_sync(0xffffffffu, R[k], SUB);
                    Rs[k] = (h == 0) ? ((k % 4 == 0) ? T(1) : T(0)) : (T)v;
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const W v = __shfl_up_sync(0xffffffffu, p[k], SUB);
                    ps[k] = (h == 0) ? T(0) : (T)v;
                }
                // ... and the flange frame, from the last lane of the instance
                W pe[3];
#pragma unroll
                for (int k = 0; k < 9; ++k) Rt[k] = (T)__shfl_sync(0xffffffffu, R[k], (L - 1) * SUB + sl);
#pragma unroll
                for (int k = 0; k < 3; ++k) pe[k] = __shfl_sync(0xffffffffu, p[k], (L - 1) * SUB + sl);
                pt.set(pe);
                // joint axes and origins to the base frame, then the Jacobian columns [z x (p_e - p_j); z] / [z; 0]
                static_for<0, NL>([&](auto kc) {
                    constexpr int k = decltype(kc)::value;
                    const int j = j0 + k;
                    const T zx = Ja[k][0], zy = Ja[k][1], zz = Ja[k][2], ox = Jl[k][0], oy = Jl[k][1], oz = Jl[k][2];
                    T z[3], o[3];
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        z[r] = fma(Rs[3 * r + 0], zx, fma(Rs[3 * r + 1], zy, Rs[3 * r + 2] * zz));
                        o[r] = fma(Rs[3 * r + 0], ox, fma(Rs[3 * r + 1], oy, fma(Rs[3 * r + 2], oz, ps[r])));
                    }
                    const T ex = (pt.hi[0] - o[0]) + pt.lo[0], ey = (pt.hi[1] - o[1]) + pt.lo[1], ez = (pt.hi[2] - o[2]) + pt.lo[2];   // p_e - o
                    const bool prism = tab->prismatic[j] != 0;
                    const bool real = j < nc;
                    const T lx = z[1] * ez - z[2] * ey, ly = z[2] * ex - z[0] * ez, lz = z[0] * ey - z[1] * ex;
                    Jl[k][0] = real ? (prism ? z[0] : lx) : T(0);
                    Jl[k][1] = real ? (prism ? z[1] : ly) : T(0);
                    Jl[k][2] = real ? (prism ? z[2] : lz) : T(0);
                    Ja[k][0] = (real && !prism) ? z[0] : T(0);
                    Ja[k][1] = (real && !prism) ? z[1] : T(0);
                    Ja[k][2] = (real && !prism) ? z[2] : T(0);
                });
            }

            // 2-4. field: attractor (every lane), this lane's share of the repulsor sum, saturation
            T tw[6];
            {
                T V[3], S0, w[3], acc[3] = {T(0), T(0), T(0)};
                attract<T>(c, g, Rt, pt, V, S0, w);
                [[maybe_unused]] NegPos2 np2;
                [[maybe_unused]] float2 acc2[3] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
                if constexpr (sizeof(T) == 4) np2.set(pt);
                for (int ch = 0; ch < a.n_chunks; ++ch) {
                    const int stage = resident ? ch : c_stage;
                    mbar_wait(&bars[stage], resident ? (uint32_t)(it & 1) : c_phase);
                    const unsigned char* sb = region + (size_t)stage * SH::kStage;
                    const int n_here = ch < a.n_full ? kChunk : a.n_rem;
                    const int n_pairs = (n_here + 1) >> 1;
                    if constexpr (sizeof(T) == 4) {
                        // stage = [pair][plane][SUB lanes] of float4; the pairs of the chunk are dealt round-robin to the lanes
                        const float4* pl = reinterpret_cast<const float4*>(sb) + sl;
                        auto deal = [&](auto order_c) {
#pragma unroll
                            for (int i = 0; i < (kChunk / 2 + L - 1) / L; ++i) {
                                const int pr = i * L + h;
                                if (pr < n_pairs)
                                    repel2<decltype(order_c)::value>(pl[(2 * pr) * SUB], pl[(2 * pr + 1) * SUB], c.obst_safe_inv, c.obst_order, np2, acc2);
                            }
                        };
                        if (kF32PowChain && c.order_int == 20) deal(std::integral_constant<int, 20>{});
                        else deal(std::integral_constant<int, 0>{});
                    } else {
                        // stage = [pair][4 planes][SUB lanes] of double2 {slot 0, slot 1}
                        const double2* pl = reinterpret_cast<const double2*>(sb) + sl;
                        auto deal = [&](auto order_c) {
                            constexpr int ORD = decltype(order_c)::value;
#pragma unroll
                            for (int i = 0; i < (kChunk / 2 + L - 1) / L; ++i) {
                                const int pr = i * L + h;
                                if (pr < n_pairs) {
                                    const double2 X = pl[(4 * pr) * SUB], Y = pl[(4 * pr + 1) * SUB], Z = pl[(4 * pr + 2) * SUB], Rr = pl[(4 * pr + 3) * SUB];
                                    Vec4<T> o0, o1;
                                    o0.x = X.x; o0.y = Y.x; o0.z = Z.x; o0.w = Rr.x;
                                    o1.x = X.y; o1.y = Y.y; o1.z = Z.y; o1.w = Rr.y;
                                    repel<T, ORD>(o0, c.obst_safe_inv, c.obst_order, pt, acc);
                                    repel<T, ORD>(o1, c.obst_safe_inv, c.obst_order, pt, acc);       // zero radius in a padding slot: adds 0
                                }
                            }
                        };
                        if (c.order_int == 20) deal(std::integral_constant<int, 20>{});
                        else if (c.order_int == 5) deal(std::integral_constant<int, 5>{});
                        else if (c.order_int == 2) deal(std::integral_constant<int, 2>{});
                        else deal(std::integral_constant<int, 0>{});
                    }
                    if (!resident || last) {
                        __syncwarp();
                        if (p_unit < n_units) issue_obst(p_unit, p_chunk, stage);
                        if (++p_chunk == a.n_chunks) p_chunk = 0;
                        if (++p_u == U) { p_u = 0; p_unit += stride; }
                    }
                    if (!resident && ++c_stage == S) { c_stage = 0; c_phase ^= 1u; }
                }
                if constexpr (sizeof(T) == 4) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) acc[k] = acc2[k].x + acc2[k].y;
                }
#pragma unroll
                for (int off = SUB; off < 32; off <<= 1) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], off);
                }
                V[0] = fma(c.obst_force, acc[0], V[0]); V[1] = fma(c.obst_force, acc[1], V[1]); V[2] = fma(c.obst_force, acc[2], V[2]);
                T v[3];
                saturate<T>(c, V, S0, v);
                tw[0] = v[0]; tw[1] = v[1]; tw[2] = v[2]; tw[3] = w[0]; tw[4] = w[1]; tw[5] = w[2];
            }

            // 5. damped least squares: this lane's columns into J J^T and J x, summed over the instance's lanes
            T x[NL];
#pragma unroll
            for (int k = 0; k < NL; ++k) x[k] = tab->ns_scale[j0 + k] * (q[k] - tab->ns_mid[j0 + k]);
            WN A[21], invd[6], Jx[6];
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                Jx[r] = WN(0);
#pragma unroll
                for (int s = 0; s <= r; ++s) A[tri(r, s)] = WN(0);
            }
            static_for<0, NL>([&](auto kc) {
                constexpr int k = decltype(kc)::value;
                const WN col[6] = {(WN)Jl[k][0], (WN)Jl[k][1], (WN)Jl[k][2], (WN)Ja[k][0], (WN)Ja[k][1], (WN)Ja[k][2]};
                axpy6(Jx, col, (WN)x[k]);
                syr6(A, col);
            });
#pragma unroll
            for (int off = SUB; off < 32; off <<= 1) {
#pragma unroll
                for (int i = 0; i < 21; ++i) A[i] += __shfl_xor_sync(0xffffffffu, A[i], off);
#pragma unroll
                for (int r = 0; r < 6; ++r) Jx[r] += __shfl_xor_sync(0xffffffffu, Jx[r], off);
            }
#pragma unroll
            for (int r = 0; r < 6; ++r) A[tri(r, r)] += c.ik_lambda2;
            chol6<WN>(A, invd);
            T yv[6], yn[6];
            {
                WN y[6];
#pragma unroll
                for (int r = 0; r < 6; ++r) y[r] = (WN)tw[r];
                chol6_fwd<WN>(A, invd, y);
                chol6_bwd<WN>(A, invd, y);
                chol6_fwd<WN>(A, invd, Jx);
                chol6_bwd<WN>(A, invd, Jx);
#pragma unroll
                for (int r = 0; r < 6; ++r) { yv[r] = (T)y[r]; yn[r] = (T)Jx[r]; }
            }

            // 6. nullspace projector on the limit-avoidance gradient, all-or-nothing lookahead check over ALL joints; 8-9. mixer, clamp
            T mix[NL];
            bool bad = false;
#pragma unroll
            for (int k = 0; k < NL; ++k) {
                const T cj[6] = {Jl[k][0], Jl[k][1], Jl[k][2], Ja[k][0], Ja[k][1], Ja[k][2]};
                const T raw = dot6(cj, yn, x[k], true);
                const T d = fma(c.ns_lookahead, raw, q[k]);
                bad = bad || (d < tab->q_lo[j0 + k]) || (d > tab->q_hi[j0 + k]);
                x[k] = raw;                                              // x is not needed any more
                mix[k] = dot6(cj, yv, T(0), false) * c.mixer_w[0];
            }
            {
                int b = bad ? 1 : 0;
#pragma unroll
                for (int off = SUB; off < 32; off <<= 1) b |= __shfl_xor_sync(0xffffffffu, b, off);
                bad = b != 0;
            }
            T lead = T(0);
#pragma unroll
            for (int k = 0; k < NL; ++k) {
                const T ns = bad ? T(0) : x[k] * c.ns_gain;
                mix[k] = fma(ns, c.mixer_w[1], mix[k]);
                lead = Prec<T>::fmax_(lead, Prec<T>::fabs_(mix[k]));
            }
#pragma unroll
            for (int off = SUB; off < 32; off <<= 1) lead = Prec<T>::fmax_(lead, __shfl_xor_sync(0xffffffffu, lead, off));
            T ratio = T(1);
            if (lead > c.max_vel) ratio = Prec<T>::div(c.max_vel, lead);

            if (last && active) {
#pragma unroll
                for (int k = 0; k < NL; ++k)
                    if (j0 + k < nc) a.qdot[tN + (j0 + k) * 32] = mix[k] * ratio;
            }
            // 10. plant
            if (c.integrate) {
#pragma unroll
                for (int k = 0; k < NL; ++k) q[k] = fma(c.dt, mix[k] * ratio, q[k]);
            }
        }
        if (active && c.integrate) {
#pragma unroll
            for (int k = 0; k < NL; ++k)
                if (j0 + k < nc) a.q[tN + (j0 + k) * 32] = q[k];
        }
    }
}

}  // namespace vfk

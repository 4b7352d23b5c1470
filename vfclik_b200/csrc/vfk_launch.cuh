// vfk_launch.cuh -- launch plan and feature dispatch of the fused cycle kernel (included by vfk_cycle_*.cu).
#pragma once
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "vfk_ctx.cuh"
#include "vfk_split.cuh"

using namespace vfk;

// Launch plan: every warp is a persistent worker with its own obstacle ring (n_stages stages of kChunk
// obstacles x 32 instances) and two q/goal buffers.  Two stages measured best on B200 for both K = 1
// (106 us vs 110 us with three at the headline shape) and K = 100 (streaming the ring from L2 every cycle at
// full occupancy beats keeping four chunks resident at 2 CTAs/SM): gpurun_out t03, DESIGN.md section 4.1.
template <typename T, int N, bool EXT, bool TAB>
static void plan_stages(int n_obst, int k_cycles, int* n_chunks, int* n_stages, size_t* smem_bytes) {
    using WS = WarpStage<T, N, EXT>;
    (void)k_cycles;
    *n_chunks = (n_obst + kChunk - 1) / kChunk;
    // long FP32 chains run 2 CTAs per SM on registers, so a third stage is free: config 5 0.684 -> 0.708 of the HBM roofline
    // (4: 0.697; 6 and 8 cost a resident CTA: 0.44); everywhere else a third stage costs occupancy (config 4: 1.009 -> 0.988)
    int stages = (sizeof(T) == 4 && N >= 10) ? 3 : 2;
    if (const char* e = getenv("VFK_STAGES")) {
        const int v = atoi(e);
        if (v >= 1 && v <= kMaxStages) stages = v;
    }
    if (*n_chunks > 0 && *n_chunks < stages) stages = *n_chunks;
    if (*n_chunks == 0) stages = 0;
    *n_stages = stages;
    *smem_bytes = kSmemHeader + kTabBytes<T, TAB> + (size_t)(kBlock / 32) * WS::warp_bytes(stages);
}

// Lanes per instance of the split shape for (precision, joints, pattern); 0 = the one-thread-per-instance kernel only.
template <typename T, int N, class PAT>
constexpr int kSplitLanes = (N >= 10 || (sizeof(T) == 8 && N == 7)) ? 2 : 0;

// Launch plan of one kernel instantiation: resident CTAs per SM for a given dynamic shared-memory size, per device.  One
// instance per instantiation (function-local static at the call site); a handful of (device, smem) entries behind a mutex, so
// concurrent host threads and sessions with different obstacle counts neither race nor evict one another.
struct PlanCache {
    std::mutex mu;
    struct Entry { int device; size_t smem; int per_sm; } e[16];
    int used = 0;
    template <typename K>
    cudaError_t get(K kern, int device, size_t smem, int* per_sm) {
        std::lock_guard<std::mutex> lock(mu);
        for (int i = 0; i < used; ++i)
            if (e[i].device == device && e[i].smem == smem) { *per_sm = e[i].per_sm; return cudaSuccess; }
        cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 << 10);
        if (err == cudaSuccess) err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, kern, kBlock, smem);
        if (err != cudaSuccess) return err;
        const int slot = used < 16 ? used++ : 15;
        e[slot] = Entry{device, smem, *per_sm};
        return cudaSuccess;
    }
};

constexpr int64_t kCoopMaxInstances = 4096;      // <= 128 tiles: 1024 cooperative warps instead of 128 solo ones

// The LEAN kernel's preconditions (see vfk_kernels.cuh).
template <typename T>
static bool is_lean(const KConst<T>& c, const vfk_buffers* b) {
    return c.tool_identity && c.unit_weights && c.share_factor && c.ik_mode == 0 && c.ns_mode == VFK_NS_PROJECTOR && !c.need_jp && !b->ns_in &&
           !c.shoulder_clamp &&
           !b->jp_ref && !b->jp_lo && !b->q_cmded && !b->ext_cmd[0] && !b->ext_cmd[1] && !b->ext_cmd[2] && !b->qdot_vf && !b->qdot_ns &&
           !b->qdot_jp && !b->cmd && !b->pose && !b->twist && !b->flags && !(b->aux && b->n_aux > 0) && b->qdot && !getenv("VFK_NO_LEAN");
}

// HIOCC: the single-cycle lean FP32 launch asks for one more resident CTA per SM (4 x 128 threads at 128 registers, no
// spills).  Measured on B200, 1 M instances: 109.5 us vs 111.2 us per launch at K = 1 (more warps hide more HBM latency),
// but 1.6 % slower at K = 100, where the kernel is issue-bound and the tighter register budget costs instructions.
template <typename T, int N, class PAT, bool EXT, bool LEAN, int G = 1, bool HIOCC = false>
static int launch_cycle(vfk_ctx* h, const KConst<T>& c, const vfk_buffers* b, int64_t n, int n_obst, int k_cycles,
                        cudaStream_t st, const vfk_io* io) {
    KArgs<T> a;
    memset(&a, 0, sizeof a);
    a.q = static_cast<T*>(b->q);
    a.goal = static_cast<const T*>(b->goal);
    a.obst = static_cast<const Vec4<T>*>(b->obst);
    a.obst_ext = static_cast<const Vec2<T>*>(b->obst_ext);
    a.aux = b->n_aux > 0 ? static_cast<const T*>(b->aux) : nullptr;
    a.n_aux = a.aux ? b->n_aux : 0;
    a.jp_ref = static_cast<T*>(b->jp_ref);
    a.jp_lo = static_cast<const T*>(b->jp_lo);
    a.jp_hi = static_cast<const T*>(b->jp_hi);
    a.ns_in = static_cast<const T*>(b->ns_in);
    a.ns_lastvec = static_cast<T*>(b->ns_lastvec);
    a.q_cmded = static_cast<const T*>(b->q_cmded);
    for (int e = 0; e < 3; ++e) a.ext_cmd[e] = static_cast<const T*>(b->ext_cmd[e]);
    a.qdot_vf = static_cast<T*>(b->qdot_vf);
    a.qdot_ns = static_cast<T*>(b->qdot_ns);
    a.qdot_jp = static_cast<T*>(b->qdot_jp);
    a.qdot = static_cast<T*>(b->qdot);
    a.cmd = static_cast<T*>(b->cmd);
    a.pose = static_cast<T*>(b->pose);
    a.twist = static_cast<T*>(b->twist);
    a.flags = b->flags;
    if (io && io->q_src) { a.q_src = static_cast<const T*>(io->q_src); a.q_src_ld = io->q_src_ld; }
    if (io && io->qdot) { a.qdot = static_cast<T*>(io->qdot); a.qdot_ld = io->qdot_ld; a.qdot_dev = static_cast<T*>(b->qdot); }
    a.n = n;
    a.n_comp = h->chain.n_joints;
    a.n_obst = n_obst;
    a.n_obst_p = (n_obst + 1) & ~1;
    a.k_cycles = k_cycles;
    size_t smem = 0;
    // sin / cos of the FP32 mode's wide chain from the shared-memory table (vfk_kernels.cuh: TAB) in EVERY instantiation of the
    // throughput shape, so that K fused cycles stay bit-identical to K single-cycle launches.  Before the L2 prefetch of the
    // obstacle block the HBM-bound single-cycle lean launch of the short chains lost 3.4 % with the table; with the prefetch
    // it is unchanged (104.9 vs 104.7 us) and the K-fused launch gains 4.7 % (profiles/r02t_sincos_table_ab.txt).
    // -DVFK_TAB_MIN_JOINTS=10 restricts the table to the long chains again.
#ifndef VFK_TAB_MIN_JOINTS
#define VFK_TAB_MIN_JOINTS 1
#endif
    constexpr bool TAB = kCanTab<T> && N >= VFK_TAB_MIN_JOINTS && G == 1;
    plan_stages<T, N, EXT, TAB>(n_obst, k_cycles, &a.n_chunks, &a.n_stages, &smem);
    a.n_full = n_obst / kChunk;
    a.n_rem = n_obst % kChunk;
#ifndef VFK_MINB_F32
#define VFK_MINB_F32 3
#endif
#ifndef VFK_MINB_LEAN
#define VFK_MINB_LEAN 3
#endif
#ifndef VFK_MINB_F64
#define VFK_MINB_F64 2
#endif
#ifndef VFK_MINB_LEAN_K1
#define VFK_MINB_LEAN_K1 4
#endif
#ifndef VFK_MINB_LEAN_BIG
#define VFK_MINB_LEAN_BIG 2
#endif
    constexpr int MINB0 = (HIOCC ? VFK_MINB_LEAN_K1
                                 : ((sizeof(T) == 4) ? (N <= 10 ? (LEAN ? VFK_MINB_LEAN : VFK_MINB_F32) : (LEAN ? VFK_MINB_LEAN_BIG : 2))
                                                     : (N <= 7 ? VFK_MINB_F64 : 1))) * (128 / kBlock);
#ifdef VFK_MINB_LEAN_BIG_ABS          // experiment: resident CTAs of kBlock threads for the long-chain lean FP32 kernel, as given
    constexpr int MINB = (sizeof(T) == 4 && N > 10 && LEAN) ? VFK_MINB_LEAN_BIG_ABS : MINB0;
#else
    constexpr int MINB = MINB0;
#endif
    auto kern = vfk_cycle_kernel<T, N, PAT, EXT, LEAN, MINB, G, TAB>;
    // per (instantiation, device, smem size): opt in to > 48 KB of dynamic shared memory and ask the occupancy once
    static PlanCache plans;
    int per_sm = 0;
    VFK_CUDA(h, plans.get(kern, h->device, smem, &per_sm));
    if (per_sm < 1) return fail(h, VFK_ERR_CUDA, "kernel does not fit an SM with %zu bytes of shared memory", smem);
    const int64_t units = (n + 31) / 32 * G;                 // G = 8: a tile is spread over 8 warps (4 instances each)
    const int64_t want = (units + kBlock / 32 - 1) / (kBlock / 32);
    const int64_t cap = (int64_t)h->sm_count * per_sm;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    kern<<<grid, kBlock, smem, st>>>(c, a);
    VFK_CUDA(h, cudaGetLastError());
    return 1;
}

// ------------------------------------------------------------------------------ lane-split shape (vfk_split.cuh)
typedef CUresult (*vfk_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static vfk_encode_tiled_fn encode_tiled_entry() {
    static vfk_encode_tiled_fn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<vfk_encode_tiled_fn>(p);
    }();
    return fn;
}

// [tiles][rows][inner] array of T, box {box_inner, box_rows, 1}: the 32 / L lanes a warp takes of `box_rows` rows of one tile
template <typename T>
static int encode_map3(vfk_ctx* h, CUtensorMap* m, const void* base, uint64_t inner, uint64_t rows, uint64_t tiles, uint32_t box_inner,
                       uint32_t box_rows) {
    vfk_encode_tiled_fn enc = encode_tiled_entry();
    if (!enc) return fail(h, VFK_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t gdim[3] = {inner, rows, tiles};
    const cuuint64_t gstr[2] = {inner * sizeof(T), inner * rows * sizeof(T)};
    const cuuint32_t box[3] = {box_inner, box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(m, sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void*>(base),
                           gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(h, VFK_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return VFK_OK;
}

// May this call run in the lane-split shape?  Lean call on blocked device buffers, DH-form chain of exactly the instantiated
// length.  Default: ON for the FP64 17-joint instantiation -- there the one-thread-per-instance kernel cannot hold its
// Jacobian (102 doubles for 17 joints) and runs 1 CTA per SM with 568 B of spills, while two lanes per instance fit 255
// registers without spills at 2 CTAs per SM: 173 us against 217 us per launch (256 k x 17 joints x 64 obstacles, 0.60 against
// 0.48 of the HBM roofline).  OFF elsewhere, where it measured slower (FP32 config 5: 184 vs 154 us; 10 joints: 98 vs 70 us in
// FP64, 95 vs 70 us in FP32; FP64 config 2: 29.6 vs
// 23.4 us).  Four lanes per instance (the kernel is written for L = 2 and 4) measured 231 us on the
// FP64 17-joint shape: every lane repeats the attractor, the Cholesky and the solves.  VFK_SPLIT=1 / 0 forces the shape on
// (where instantiated) / off.
template <typename T>
static bool split_ok(vfk_ctx* h, int n_kernel, const KConst<T>& c, const vfk_buffers* b, const vfk_io* io) {
    const char* e = getenv("VFK_SPLIT");
    const bool on = e ? atoi(e) != 0 : (sizeof(T) == 8 && n_kernel == 17);
    return on && h->dh_chain && h->chain.n_joints == n_kernel && is_lean<T>(c, b) && !(io && (io->q_src || io->qdot));
}

template <typename T, int N, int L>
static int launch_split(vfk_ctx* h, const KConst<T>& c, const vfk_buffers* b, int64_t n, int n_obst, int k_cycles, cudaStream_t st) {
    using SH = SplitShape<T, N, L>;
    KArgs<T> a;
    memset(&a, 0, sizeof a);
    a.q = static_cast<T*>(b->q);
    a.qdot = static_cast<T*>(b->qdot);
    a.n = n;
    a.n_comp = h->chain.n_joints;
    a.n_obst = n_obst;
    a.n_obst_p = (n_obst + 1) & ~1;
    a.k_cycles = k_cycles;
    a.n_chunks = (n_obst + kChunk - 1) / kChunk;
    a.n_full = n_obst / kChunk;
    a.n_rem = n_obst % kChunk;
    int stages = 2;
    if (const char* e = getenv("VFK_STAGES")) {
        const int v = atoi(e);
        if (v >= 1 && v <= kMaxStages) stages = v;
    }
    if (a.n_chunks < stages) stages = a.n_chunks;
    a.n_stages = stages;
    const size_t smem = kSmemHeader + SH::kTab + (size_t)(kBlock / 32) * SH::warp_bytes(stages);
    const uint64_t tiles = (uint64_t)((n + 31) / 32);
    SplitMaps maps;
    memset(&maps, 0, sizeof maps);
    int rc;
    if ((rc = encode_map3<T>(h, &maps.q, b->q, 32, (uint64_t)a.n_comp, tiles, SH::SUB, (uint32_t)a.n_comp)) != VFK_OK) return rc;
    if ((rc = encode_map3<T>(h, &maps.goal, b->goal, 32, 13, tiles, SH::SUB, 13)) != VFK_OK) return rc;
    if (n_obst > 0) {
        const uint64_t rows = (uint64_t)a.n_obst_p / 2 * ObstPairs<T>::kPlanes;                   // plane rows of 32 lanes x 16 bytes per tile
        if ((rc = encode_map3<T>(h, &maps.obst, b->obst, 32 * 16 / sizeof(T), rows, tiles, SH::kRowBytes / sizeof(T), SH::kRowsPerChunk)) != VFK_OK)
            return rc;
    }
#ifndef VFK_MINB_SPLIT_F32
#define VFK_MINB_SPLIT_F32 4
#endif
#ifndef VFK_MINB_SPLIT_F64
#define VFK_MINB_SPLIT_F64 2          // 255 registers: no spills for 17 joints (3 CTAs / 168 registers spill and measured equal to the solo kernel)
#endif
    constexpr int MINB = (sizeof(T) == 4 ? VFK_MINB_SPLIT_F32 : VFK_MINB_SPLIT_F64) * (128 / kBlock);
    auto kern = vfk_split_kernel<T, N, L, MINB>;
    static PlanCache plans;
    int per_sm = 0;
    VFK_CUDA(h, plans.get(kern, h->device, smem, &per_sm));
    if (per_sm < 1) return fail(h, VFK_ERR_CUDA, "split kernel does not fit an SM with %zu bytes of shared memory", smem);
    const int64_t units = (n + 31) / 32 * L;
    const int64_t want = (units + kBlock / 32 - 1) / (kBlock / 32);
    const int64_t cap = (int64_t)h->sm_count * per_sm;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    kern<<<grid, kBlock, smem, st>>>(c, a, maps);
    VFK_CUDA(h, cudaGetLastError());
    return 1;
}

template <typename T, int N, class PAT>
static int dispatch_feat(vfk_ctx* h, const KConst<T>& c, const vfk_buffers* b, int64_t n, int n_obst, int k_cycles,
                         cudaStream_t st, const vfk_io* io) {
    const bool ext = b->obst_ext && n_obst > 0;
    // Small FP64 batches cannot fill the GPU with one thread per instance: switch to the cooperative latency shape.
    // Measured (scripts/latency.py, one robot, 1000 fused cycles): FP64 M = 32: 6.4 vs 7.1 us per cycle, M = 256: 23.6 vs
    // 36.3 us.  In FP32 the repeller is three MUFU instructions and the shape only adds ring overhead (3.7 vs 2.7 us),
    // so it is neither selected nor instantiated there.
    if constexpr (sizeof(T) == 8) {
        const bool coop = n <= kCoopMaxInstances && n_obst >= 2 * kChunk && !getenv("VFK_NO_COOP");
        if (coop) return ext ? launch_cycle<T, N, PAT, true, false, kChunk>(h, c, b, n, n_obst, k_cycles, st, io)
                             : launch_cycle<T, N, PAT, false, false, kChunk>(h, c, b, n, n_obst, k_cycles, st, io);
    }
    if (ext) return launch_cycle<T, N, PAT, true, false>(h, c, b, n, n_obst, k_cycles, st, io);
    // Two lanes per instance (vfk_split.cuh): default for the FP64 17-joint kernel, opt-in elsewhere.
    if constexpr (kSplitLanes<T, N, PAT> > 0) {
        if (n_obst > 0 && split_ok<T>(h, N, c, b, io)) return launch_split<T, N, kSplitLanes<T, N, PAT>>(h, c, b, n, n_obst, k_cycles, st);
    }
    if (is_lean<T>(c, b) && h->chain.n_joints == N) {
        if constexpr (sizeof(T) == 4 && N <= 7) {
            if (k_cycles == 1) return launch_cycle<T, N, PAT, false, true, 1, true>(h, c, b, n, n_obst, k_cycles, st, io);
        }
        return launch_cycle<T, N, PAT, false, true>(h, c, b, n, n_obst, k_cycles, st, io);
    }
    return launch_cycle<T, N, PAT, false, false>(h, c, b, n, n_obst, k_cycles, st, io);
}

template <typename T, int N>
static int dispatch_ext(vfk_ctx* h, const KConst<T>& c, const vfk_buffers* b, int64_t n, int n_obst, int k_cycles,
                        cudaStream_t st, const vfk_io* io) {
    if constexpr (N == 7) {
        if (h->pattern == 1) return dispatch_feat<T, N, LwrPattern>(h, c, b, n, n_obst, k_cycles, st, io);
    }
    if constexpr (N >= 10) {
        if (h->pattern == 2) return dispatch_feat<T, N, DhPattern>(h, c, b, n, n_obst, k_cycles, st, io);
    }
    return dispatch_feat<T, N, GenericPattern>(h, c, b, n, n_obst, k_cycles, st, io);
}


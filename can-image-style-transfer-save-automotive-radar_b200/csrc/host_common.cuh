// Host-side helpers shared by the plan, the L-BFGS driver and the per-op entry points:
// error reporting across the C ABI, TMA tensor-map construction, kernel launchers.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ist_b200.h"
#include "conv_igemm.cuh"
#include "conv_halo.cuh"
#include "elementwise.cuh"
#include "gram.cuh"
#include "conv_first_tc.cuh"
#include "conv_first_fwd_tc.cuh"

namespace ist {

inline std::string& last_error() {
    static thread_local std::string e;
    return e;
}
inline int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    last_error() = buf;
    return code;
}
#define IST_CUDA(call)                                                                                   \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            return ::ist::fail(IST_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)
#define IST_TRY(call)                  \
    do {                               \
        int rc__ = (call);             \
        if (rc__ != IST_OK) return rc__; \
    } while (0)

// ----------------------------------------------------------------------------------------------------------
// launch bookkeeping: a process-wide count of kernels this library launched (graph replays count their nodes) and an
// optional per-launch profiler (CUDA events around every launch, eager mode only) used by bench.py for the roofline.
// ----------------------------------------------------------------------------------------------------------
struct ProfRec {
    char name[40];
    double flops, bytes;      // algorithmic work of the launch
    cudaEvent_t e0, e1;
};
struct LaunchBook {
    unsigned long long launches = 0;   // kernels launched or replayed
    unsigned long long captured = 0;   // kernels recorded into the graph being captured
    bool capturing = false;
    bool profiling = false;
    std::vector<ProfRec> recs;
};
inline LaunchBook& book() {
    static LaunchBook b;
    return b;
}
inline void launch_pre(const char* name, double flops, double bytes, cudaStream_t st) {
    LaunchBook& b = book();
    if (b.capturing) b.captured++; else b.launches++;
    if (b.profiling && !b.capturing) {
        ProfRec r;
        memset(&r, 0, sizeof(r));
        strncpy(r.name, name, sizeof(r.name) - 1);
        r.flops = flops; r.bytes = bytes;
        cudaEventCreate(&r.e0);
        cudaEventCreate(&r.e1);
        cudaEventRecord(r.e0, st);
        b.recs.push_back(r);
    }
}
inline void launch_post(cudaStream_t st) {
    LaunchBook& b = book();
    if (b.profiling && !b.capturing && !b.recs.empty()) cudaEventRecord(b.recs.back().e1, st);
}

// Kernel launch through cudaLaunchKernelEx. `pdl` marks the launch as programmatically dependent on the previous kernel of
// the stream: the grid may be scheduled while that kernel drains, so its prologue (barrier init, tensor-memory allocation,
// descriptor prefetch, launch latency) overlaps the predecessor's tail. Every kernel launched this way executes
// griddepcontrol.wait (pdl_wait) before it reads or writes global memory. IST_B200_NO_PDL=1 turns the attribute off.
// kernel classes for IST_B200_PDL_MASK (default 3: measured in profiles/r01_pdl_ab.log, the optimiser kernels lose with it): 1 = tensor-core kernels, 2 = elementwise kernels of the closure,
// 4 = optimiser kernels
enum { PDL_TENSOR = 1, PDL_EW = 2, PDL_OPT = 4 };
inline int pdl_mask() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("IST_B200_NO_PDL");
        const char* m = getenv("IST_B200_PDL_MASK");
        v = (e != nullptr && atoi(e) == 1) ? 0 : (m != nullptr ? atoi(m) : 3);
    }
    return v;
}
// `cluster_x` > 1 launches thread-block clusters of that many CTAs along x (CTA pairs of the conv kernel).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kc(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int pdl, int cluster_x,
                             Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int n = 0;
    if ((pdl & pdl_mask()) != 0) {                       // pdl: 0 / false = plain launch, else the kernel's PDL_* class
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    if (cluster_x > 1) {
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = (unsigned)cluster_x; attr[n].val.clusterDim.y = 1; attr[n].val.clusterDim.z = 1;
        ++n;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int pdl, Args&&... args) {
    return launch_kc(kernel, grid, block, smem, st, pdl, 1, std::forward<Args>(args)...);
}

// Per-device caches: function attributes (cudaFuncSetAttribute) and device properties belong to the CURRENT device, and a
// process may drive several (MODEL.DEVICE = 'cuda:1' while device 0 is current elsewhere; one process per GPU is the normal
// multi-GPU mode, but the reference lets the user pick any device through torch).
constexpr int IST_MAX_DEVICES = 64;
inline int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev >= 0 && dev < IST_MAX_DEVICES) ? dev : 0;
}
// Makes `dev` current for the duration of a C-ABI call on an object bound to that device (plans and optimisers remember the
// device they were created on), and restores the caller's device afterwards.
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (dev >= 0 && dev != prev) switched = (cudaSetDevice(dev) == cudaSuccess);
    }
    ~DeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};
struct DeviceOnce {
    bool done[IST_MAX_DEVICES] = {};
    bool first() {                       // true exactly once per device
        const int d = current_device();
        if (done[d]) return false;
        done[d] = true;
        return true;
    }
};
inline int num_sms() {
    static int n[IST_MAX_DEVICES] = {};
    const int dev = current_device();
    if (n[dev] == 0) {
        cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
        if (n[dev] <= 0) n[dev] = 148;
    }
    return n[dev];
}

// ----------------------------------------------------------------------------------------------------------
// TMA tensor maps (cuTensorMapEncodeTiled through the runtime's driver entry point: no -lcuda needed)
// ----------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}
// 16-bit elements, 128-byte swizzle, zero fill out of bounds. dims/box innermost first; strides in bytes for dims 1..rank-1.
inline int make_tmap(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                     const uint32_t* box) {
    EncodeTiledFn fn = encode_fn();
    if (fn == nullptr) return fail(IST_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint32_t ones[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box,
                    ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(IST_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d (rank %d dims %llu %llu %llu box %u %u %u)", (int)r,
                    rank, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0),
                    box[0], box[1], rank > 2 ? box[2] : 0);
    return IST_OK;
}

// which implicit-GEMM kernel runs the convolutions (IST_B200_CONV): "pair" (default) = conv_halo.cuh as CTA pairs
// (tcgen05 cta_group::2), "halo" = conv_halo.cuh single-CTA, "igemm" = conv_igemm.cuh (first generation)
inline int conv_impl() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("IST_B200_CONV");
        v = (e != nullptr && strcmp(e, "igemm") == 0) ? 0 : (e != nullptr && strcmp(e, "halo") == 0) ? 1 : 2;
    }
    return v;
}
inline int conv_impl_halo() { return conv_impl() >= 1; }
inline int conv_impl_pair() { return conv_impl() == 2; }
// independent workers of a conv launch: CTAs, or CTA pairs
inline int conv_workers() { return conv_impl_pair() ? num_sms() / 2 : num_sms(); }
// timing experiments (results are garbage when loads are skipped): IST_B200_DBG_FLAGS bit 1 (2) skips the A loads, bit 2 (4) the B loads
inline int halo_dbg_flags() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("IST_B200_DBG_FLAGS");
        v = (e != nullptr) ? atoi(e) : 0;     // debug: bit 1 (2) skips the A loads, bit 2 (4) skips the B loads (timing experiments)
    }
    return v;
}
inline void pick_tile(int W, int* TW, int* TH) {
    if (conv_impl_halo()) { *TW = 8; *TH = 16; return; }
    if (W >= 16) { *TW = 16; *TH = 8; } else { *TW = 8; *TH = 16; }
}
// NHWC planes [NB,H,W,C] as (C, W, H, NB): the A operand of the implicit GEMM for a `taps`-tap contraction.
// conv_igemm: box (64, TW, TH, 1) fetched once per tap; conv_halo: box (64, TW+2, TH+2, 1) fetched once per 64-channel
// chunk when taps == 9 (exact (64, TW, TH, 1) box for 1x1 contractions).
inline int map_act(CUtensorMap* m, const uint16_t* base, int NB, int H, int W, int C, int taps) {
    int TW, TH;
    pick_tile(W, &TW, &TH);
    const int halo = (conv_impl_halo() && taps == 9) ? 2 : 0;
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)NB};
    const uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
    const uint32_t box[4] = {64, (uint32_t)(TW + halo), (uint32_t)(TH + halo), 1};
    return make_tmap(m, base, 4, dims, strides, box);
}
// weight planes [G][N][K] as (K, N, G), box (64, n_tile, 1): the B operand (G = tap, or frame for the Gram backward)
inline int map_b(CUtensorMap* m, const uint16_t* base, int G, int N, int K, int n_tile) {
    const uint64_t dims[3] = {(uint64_t)K, (uint64_t)N, (uint64_t)G};
    const uint64_t strides[2] = {(uint64_t)K * 2, (uint64_t)N * K * 2};
    const uint32_t box[3] = {64, (uint32_t)n_tile, 1};
    return make_tmap(m, base, 3, dims, strides, box);
}
// planes viewed as [NB][HW][C] -> (C, HW, NB), box (64, 64, 1): both operands of the Gram SYRK
inline int map_gram(CUtensorMap* m, const uint16_t* base, int NB, int HW, int C) {
    const uint64_t dims[3] = {(uint64_t)C, (uint64_t)HW, (uint64_t)NB};
    const uint64_t strides[2] = {(uint64_t)C * 2, (uint64_t)HW * C * 2};
    const uint32_t box[3] = {64, 64, 1};
    return make_tmap(m, base, 3, dims, strides, box);
}

// ----------------------------------------------------------------------------------------------------------
// launchers
// ----------------------------------------------------------------------------------------------------------
template <int N_TILE>
inline int launch_conv_t(cudaStream_t st, const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi,
                         const CUtensorMap& b_lo, const ConvParams& p) {
    static DeviceOnce attr_once;
    if (attr_once.first()) {
        IST_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<N_TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      ConvCfg<N_TILE>::SMEM_BYTES));
    }
    const int total = p.NB * p.tiles_x * p.tiles_y * p.tiles_n;
    const int grid = total < num_sms() ? total : num_sms();
    const double px = (double)p.NB * p.H * p.W;
    const int planes = p.passes == 3 ? 2 : 1;
    launch_pre(p.taps == 9 ? (p.mode == CONV_FWD ? "conv_igemm_fwd" : "conv_igemm_dgrad") : "conv_igemm_gram_bwd",
               2.0 * px * p.Cout * p.Cin * p.taps,
               planes * 2.0 * (px * p.Cin + (double)p.taps * p.Cin * p.Cout) + px * p.Cout * (p.out_f32 != nullptr ? 4.0 : 4.0), st);
    conv_igemm_kernel<N_TILE><<<grid, 192, ConvCfg<N_TILE>::SMEM_BYTES, st>>>(a_hi, a_lo, b_hi, b_lo, p);
    launch_post(st);
    IST_CUDA(cudaGetLastError());
    return IST_OK;
}
// Gram backward fused into a data-gradient launch (conv_halo only): `chunks` extra k-steps A = feature planes (halo-box maps,
// fp16), B = D matrix planes (fp16, [frame][C][C], scaled by 2^e), accumulated in their own tensor-memory accumulator and added
// to the data-gradient as alpha[frame] * (D * F) in the epilogue.
struct GramFuse {
    const CUtensorMap *f_hi, *f_lo, *d_hi, *d_lo;
    int chunks;
    const float* alpha;     // [NB] device multipliers (gram_dmat_kernel's alpha_out)
};
template <int N_TILE, bool PAIR>
inline int launch_halo_t(cudaStream_t st, const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi,
                         const CUtensorMap& b_lo, const CUtensorMap& o_hi, const CUtensorMap& o_lo, const CUtensorMap& f_hi,
                         const CUtensorMap& f_lo, const CUtensorMap& d_hi, const CUtensorMap& d_lo, const ConvParams& p_in) {
    static DeviceOnce attr_once;
    if (attr_once.first()) {
        IST_CUDA(cudaFuncSetAttribute(conv_halo_kernel<N_TILE, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      HaloCfg<N_TILE, PAIR>::SMEM_BYTES));
    }
    const int total = p_in.NB * p_in.tiles_x * p_in.tiles_y * p_in.tiles_n;
    // The stream-K hand-over spins on flags written by other CTAs of the same grid: every CTA of the launch must be resident
    // at once. One CTA (or CTA pair) per SM is what the shared-memory footprint allows; ask the runtime how many it will
    // really keep resident on this device for this kernel (once per device) and refuse to split when the grid would not fit.
    static int max_resident[IST_MAX_DEVICES] = {};
    {
        const int dev = current_device();
        if (max_resident[dev] == 0) {
            int n = 0;
            cudaLaunchConfig_t qc;
            memset(&qc, 0, sizeof(qc));
            qc.gridDim = dim3(num_sms()); qc.blockDim = dim3(HaloCfg<N_TILE, PAIR>::THREADS);
            qc.dynamicSmemBytes = HaloCfg<N_TILE, PAIR>::SMEM_BYTES;
            cudaLaunchAttribute qa[1];
            qa[0].id = cudaLaunchAttributeClusterDimension;
            qa[0].val.clusterDim.x = PAIR ? 2 : 1; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
            qc.attrs = qa; qc.numAttrs = 1;
            if (cudaOccupancyMaxActiveClusters(&n, conv_halo_kernel<N_TILE, PAIR>, &qc) != cudaSuccess || n <= 0) {
                cudaGetLastError();
                n = -1;                                  // unknown: trust the one-CTA-per-SM design
            }
            max_resident[dev] = n > 0 ? n * (PAIR ? 2 : 1) : -1;
        }
    }
    ConvParams p = p_in;
    auto grid_of = [&]() {
        int groups = conv_workers() / p.sk_cpf;        // frame groups that fit side by side
        if (groups > p.NB) groups = p.NB;
        if (groups < 1) groups = 1;
        return groups * p.sk_cpf * (PAIR ? 2 : 1);
    };
    int grid = grid_of();
    if (p.sk_ws != nullptr && max_resident[current_device()] > 0 && grid > max_resident[current_device()]) {
        // fall back to whole-tile ranges: no CTA waits for another one, so co-residency is not needed (the partition, hence the
        // rounding, then differs from a device where the split runs — results stay deterministic on this device)
        const long long tiles_f = (long long)p.tiles_x * p.tiles_y * p.tiles_n;
        p.sk_ws = nullptr; p.sk_flags = nullptr;
        p.sk_cpf = tiles_f < conv_workers() ? (int)tiles_f : conv_workers();
        grid = grid_of();
    }
    const double px = (double)p.NB * p.H * p.W;
    const int planes = p.passes == 3 ? 2 : 1;
    launch_pre(p.taps == 9 ? (p.mode == CONV_FWD ? "conv_halo_fwd" : "conv_halo_dgrad") : "conv_halo_gram_bwd",
               2.0 * px * p.Cout * (p.Cin * p.taps + 64.0 * p.extra_chunks),
               planes * 2.0 * (px * p.Cin + (double)p.taps * p.Cin * p.Cout) + px * p.Cout * 4.0, st);
    static int dbg_on = -1;
    if (dbg_on < 0) { const char* e = getenv("IST_B200_DBG_TIMES"); dbg_on = (e != nullptr && atoi(e) == 1) ? 1 : 0; }
    if (dbg_on && !book().capturing) {
        // phase stamps per CTA: 0 entry, 1 setup done, 2 first MMA issued, 3 last MMA of the last tile issued,
        // 4 last tile's accumulators in registers, 5 last tile stored, 6 exit
        static long long* dbuf = nullptr;
        if (dbuf == nullptr) IST_CUDA(cudaMalloc(&dbuf, sizeof(long long) * 16 * 1024));
        IST_CUDA(cudaMemsetAsync(dbuf, 0, sizeof(long long) * 16 * 1024, st));
        ConvParams q = p;
        q.dbg_times = dbuf;
        IST_CUDA(launch_kc(conv_halo_kernel<N_TILE, PAIR>, dim3(grid), dim3(HaloCfg<N_TILE, PAIR>::THREADS), HaloCfg<N_TILE, PAIR>::SMEM_BYTES, st, 0, PAIR ? 2 : 1, a_hi, a_lo, b_hi, b_lo, o_hi, o_lo, f_hi, f_lo, d_hi, d_lo, q));
        IST_CUDA(cudaStreamSynchronize(st));
        std::vector<long long> h(16 * (size_t)grid);
        IST_CUDA(cudaMemcpy(h.data(), dbuf, sizeof(long long) * 16 * grid, cudaMemcpyDeviceToHost));
        double sum[7] = {0, 0, 0, 0, 0, 0, 0};
        double ns = 0;
        double wsum[8] = {0, 0, 0, 0, 0, 0, 0, 0};       // barrier wait cycles per role (leader CTAs for the issuers)
        int nlead = 0;
        long long gmin = h[2], gmax_start = h[2], gend = h[3];
        for (int c = 0; c < grid; ++c) {
            if (h[16 * c + 2] < gmin) gmin = h[16 * c + 2];
            if (h[16 * c + 2] > gmax_start) gmax_start = h[16 * c + 2];
            if (h[16 * c + 3] > gend) gend = h[16 * c + 3];
        }
        for (int c = 0; c < grid; ++c) {
            for (int k = 1; k < 7; ++k) if (k != 2 && k != 3) sum[k] += (double)(h[16 * c + k] - h[16 * c]);
            ns += (double)h[16 * c + 7];
            const bool lead = !PAIR || (c % 2 == 0);
            if (lead) ++nlead;
            for (int k = 8; k < 16; ++k) if (lead || k == 13 || k == 14 || k == 15) wsum[k - 8] += (double)h[16 * c + k];
        }
        fprintf(stderr, "[dbg-wait] avg cycles waiting | issuer1: acc-empty %.0f B-full %.0f A-full %.0f | issuer2: operands %.0f cross-empty %.0f | producer: A-empty %.0f B-empty %.0f | epilogue warp 2: acc-full %.0f\n",
                wsum[0] / nlead, wsum[1] / nlead, wsum[2] / nlead, wsum[3] / nlead, wsum[4] / nlead, wsum[5] / grid, wsum[6] / grid, wsum[7] / grid);
        fprintf(stderr, "[dbg] conv %dx%d %d->%d taps %d passes %d promote %d grid %d tiles %d | avg clk since entry: setup %.0f acc_read %.0f stored %.0f exit %.0f | CTA life %.1f us -> SM clock %.0f MHz | first CTA start -> last CTA start %.1f us, first start -> last exit %.1f us\n",
                p.H, p.W, p.Cin, p.Cout, p.taps, p.passes, p.promote, grid, total, sum[1] / grid,
                sum[4] / grid, sum[5] / grid, sum[6] / grid, ns / grid * 1e-3, sum[6] / ns * 1e3, (gmax_start - gmin) * 1e-3, (gend - gmin) * 1e-3);
        launch_post(st);
        return IST_OK;
    }
    IST_CUDA(launch_kc(conv_halo_kernel<N_TILE, PAIR>, dim3(grid), dim3(HaloCfg<N_TILE, PAIR>::THREADS), HaloCfg<N_TILE, PAIR>::SMEM_BYTES, st, p.pdl != 0 ? PDL_TENSOR : 0, PAIR ? 2 : 1, a_hi, a_lo, b_hi, b_lo, o_hi, o_lo, f_hi, f_lo, d_hi, d_lo, p));
    launch_post(st);
    return IST_OK;
}
inline int conv_n_tile(int cout) { return cout >= 128 ? 128 : 64; }
// rows of the weight box one CTA loads per tap: a CTA pair splits the N_TILE output channels of a tile between its two CTAs
inline int conv_b_box(int cout) { return conv_impl_pair() ? conv_n_tile(cout) / 2 : conv_n_tile(cout); }
// k-steps per tensor-core accumulation chain before promotion to fp32 registers (IST_B200_PROMOTE overrides; 1 = most accurate)
inline int promote_steps() {
    static int v = 0;
    if (v == 0) {
        const char* e = getenv("IST_B200_PROMOTE");
        v = (e != nullptr && atoi(e) > 0) ? atoi(e) : 1;
    }
    return v;
}
// Forward convolutions: taps per accumulation chain inside a 64-channel chunk (chains never cross a chunk). Draining a
// 128x128 fp32 accumulator takes ~1000 cycles of tensor-memory read bandwidth, the 12 MMAs of a k-step 768 cycles as a CTA
// pair (1032 single-CTA), so short chains cost more as a pair. Measured (profiles/r01_promote_fwd_ab.log, second table):
// chains of 1 / 2 / 3 / 5 taps give a forward rel-L2 error of 1.5 / 2.0 / 2.7 / 3.7e-7 against fp64 and a 512^2 closure of
// 1.331 / 1.304 / 1.294 / 1.296 ms; 3 taps (three chains per chunk) is the default for CTA pairs, 1 for the single-CTA
// kernel. On the parity points of tests/test_closure_gpu.py the image gradient is as close to fp64 with 3 taps as with 1
// (its error is set by ReLU / pool mask flips, and stays at or below the reference's own fp32-vs-fp64 error).
// IST_B200_PROMOTE_FWD overrides.
inline int promote_steps_fwd() {
    static int v = 0;
    if (v == 0) {
        const char* e = getenv("IST_B200_PROMOTE_FWD");
        const char* g = getenv("IST_B200_PROMOTE");
        v = (e != nullptr && atoi(e) > 0) ? atoi(e) : (g != nullptr && atoi(g) > 0) ? atoi(g) : (conv_impl_pair() ? 3 : 1);
    }
    return v;
}
// The data-gradient is linear in its operands (no ReLU / pool decisions depend on it), so the ~3e-8-per-MMA truncation of a
// longer tensor-core chain only adds ~1e-6 relative error to the gradient: one chain per 64-channel chunk (9 taps = 36 MMAs)
// instead of one per k-step saves most of the promotion drains. IST_B200_PROMOTE_BWD overrides.
inline int promote_steps_bwd() {
    static int v = 0;
    if (v == 0) {
        const char* e = getenv("IST_B200_PROMOTE_BWD");
        v = (e != nullptr && atoi(e) > 0) ? atoi(e) : 9;
    }
    return v;
}
// Fills the tiling fields of p (NB,H,W,Cin,Cout,taps,passes,mode and epilogue pointers must be set) and launches.
// Stream-K workspace of the conv_halo kernel: one fp32 partial tile + one flag per CTA. One instance per plan (kernels of a
// plan run on one stream); the per-op entry points share a process-wide one. IST_B200_NO_STREAMK=1 disables the split.
struct ConvWorkspace {
    float* ws = nullptr;
    int* flags = nullptr;
    int alloc() {
        if (ws != nullptr) return IST_OK;
        const char* e = getenv("IST_B200_NO_STREAMK");
        if (e != nullptr && atoi(e) == 1) return IST_OK;
        const size_t n = (size_t)num_sms();
        cudaError_t r = cudaMalloc(&ws, n * 2 * 128 * 128 * sizeof(float));
        if (r == cudaSuccess) r = cudaMalloc(&flags, n * 2 * sizeof(int));
        if (r == cudaSuccess) r = cudaMemset(flags, 0, n * 2 * sizeof(int));
        if (r != cudaSuccess) return fail(IST_ERR_CUDA, "stream-K workspace allocation failed: %s", cudaGetErrorString(r));
        return IST_OK;
    }
    void release() {
        if (ws != nullptr) cudaFree(ws);
        if (flags != nullptr) cudaFree(flags);
        ws = nullptr; flags = nullptr;
    }
};
inline ConvWorkspace& global_conv_workspace() {
    static ConvWorkspace w[IST_MAX_DEVICES];      // one per device: the buffers live in the device they were allocated on
    return w[current_device()];
}

// o_hi / o_lo: optional tensor maps (map_act(..., taps = 1)) of the output planes; with them the conv_halo forward epilogue
// leaves through shared memory + TMA store.
// fmt: operand format of the k-steps, 0 = fp16 x fp16, 1 = bf16 x bf16 (mixing the two in one MMA is an illegal instruction on sm_100a)
inline uint32_t conv_idesc(int fmt, int nt) {
    return umma_idesc_f16(fmt == 1 ? UMMA_FMT_BF16 : UMMA_FMT_F16, conv_impl_pair() ? 256 : 128, (uint32_t)nt, 0, 0);
}
inline int launch_conv(cudaStream_t st, const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi,
                       const CUtensorMap& b_lo, ConvParams p, int fmt, const CUtensorMap* o_hi = nullptr,
                       const CUtensorMap* o_lo = nullptr, const GramFuse* gf = nullptr, ConvWorkspace* wsp = nullptr) {
    if (p.Cin % 64 != 0 || p.Cout % 64 != 0) return fail(IST_ERR_ARG, "conv_igemm needs Cin, Cout %% 64 == 0 (got %d, %d)", p.Cin, p.Cout);
    pick_tile(p.W, &p.TW, &p.TH);
    p.tiles_x = (p.W + p.TW - 1) / p.TW;
    if (conv_impl_pair()) p.tiles_x = (p.tiles_x + 1) / 2;      // pair tiles: two adjacent tile columns (a phantom one past an odd edge)
    p.tiles_y = (p.H + p.TH - 1) / p.TH;
    const int nt = conv_n_tile(p.Cout);
    p.tiles_n = p.Cout / nt;
    if (p.promote < 1) p.promote = (p.mode == CONV_GRAD) ? promote_steps_bwd() : promote_steps_fwd();
    p.idesc = conv_idesc(fmt, nt);
    p.idesc2 = conv_idesc(0, nt);
    p.extra_chunks = 0;
    if (gf != nullptr && !conv_impl_halo()) return fail(IST_ERR_STATE, "the fused Gram backward needs the conv_halo kernel");
    if (conv_impl_halo()) {
        if (gf != nullptr) {
            if (p.taps != 9) return fail(IST_ERR_ARG, "fused Gram backward rides on a 3x3 data-gradient launch");
            if (p.passes != 3) return fail(IST_ERR_ARG, "fused Gram backward needs the 3-pass split");
            p.extra_chunks = gf->chunks;
            p.alpha2_dev = gf->alpha;
        }
        p.dbg_flags = halo_dbg_flags();
        {
            // resident weight tiles: pair kernel, N = 64, ONE 64-channel chunk, 3x3 taps (conv1_2 forward and data-gradient)
            static int res_env = -1;
            if (res_env < 0) { const char* e = getenv("IST_B200_B_RESIDENT"); res_env = (e != nullptr && atoi(e) == 0) ? 0 : 1; }
            p.b_resident = (res_env && conv_impl_pair() && nt == 64 && p.Cin == 64 && p.taps == 9 && !p.b_frame && p.passes == 3) ? 1 : 0;
        }
        if (wsp == nullptr) wsp = &global_conv_workspace();
        IST_TRY(wsp->alloc());
        // stream-K pays (partial-tile exchange, more segments) only when whole-tile waves leave SMs idle: tiles / SMs far from
        // an integer, and at least two chunk units per tile to split. Decided from ONE frame's geometry, so that the partition
        // (hence the rounding) of a frame does not depend on the batch size.
        {
            const long long tiles_f = (long long)p.tiles_x * p.tiles_y * p.tiles_n;
            const int ch = p.Cin / 64 + p.extra_chunks;
            const int workers = conv_workers();
            const long long waves = (tiles_f + workers - 1) / workers;
            const double eff = (double)tiles_f / (double)(waves * workers);
            // ... and only when the tensor work it saves outweighs the serial partial-tile exchange at the head CTA (one
            // 64 KB partial through L2 per extra CTA of a tile, ~3000 cycles each; a k-step is ~770 cycles): tiny launches
            // such as the 1x1 Gram backward of the deepest layer run faster as whole tiles on a few CTAs
            const double ksteps_tile = (double)(p.Cin / 64) * p.taps + p.extra_chunks;
            const double t_whole = (double)waves * ksteps_tile * 770.0;
            const double share = (double)tiles_f / workers;                     // tiles per worker when split evenly
            const double ways = share >= 1.0 ? 2.0 : ceil(1.0 / share);           // CTAs (pairs) that meet in one tile
            const double t_split = share * ksteps_tile * 770.0 + (ways - 1.0) * 3000.0;
            const bool want = wsp->ws != nullptr && ch >= 2 && eff < 0.93 && tiles_f * ch < (1ll << 30) && t_split < t_whole;
            p.sk_ws = want ? wsp->ws : nullptr;
            p.sk_flags = want ? wsp->flags : nullptr;
            const long long units = want ? tiles_f * ch : tiles_f;
            p.sk_cpf = units < workers ? (int)units : workers;
        }
        p.use_tma_store = (p.out_f32 == nullptr && o_hi != nullptr && o_lo != nullptr) ? 1 : 0;
        const CUtensorMap& oh = p.use_tma_store ? *o_hi : a_hi;
        const CUtensorMap& ol = p.use_tma_store ? *o_lo : a_lo;
        const CUtensorMap& fh = gf != nullptr ? *gf->f_hi : a_hi;
        const CUtensorMap& fl = gf != nullptr ? *gf->f_lo : a_lo;
        const CUtensorMap& dh = gf != nullptr ? *gf->d_hi : b_hi;
        const CUtensorMap& dl = gf != nullptr ? *gf->d_lo : b_lo;
        if (conv_impl_pair())
            return nt == 128 ? launch_halo_t<128, true>(st, a_hi, a_lo, b_hi, b_lo, oh, ol, fh, fl, dh, dl, p)
                             : launch_halo_t<64, true>(st, a_hi, a_lo, b_hi, b_lo, oh, ol, fh, fl, dh, dl, p);
        return nt == 128 ? launch_halo_t<128, false>(st, a_hi, a_lo, b_hi, b_lo, oh, ol, fh, fl, dh, dl, p)
                         : launch_halo_t<64, false>(st, a_hi, a_lo, b_hi, b_lo, oh, ol, fh, fl, dh, dl, p);
    }
    return nt == 128 ? launch_conv_t<128>(st, a_hi, a_lo, b_hi, b_lo, p) : launch_conv_t<64>(st, a_hi, a_lo, b_hi, b_lo, p);
}

inline int launch_conv_first_dgrad(cudaStream_t st, const uint16_t* g_hi, const uint16_t* g_lo, const float* w, float* grad, int NB,
                                   int H, int W, bool pdl = false) {
    static DeviceOnce attr_once;
    if (attr_once.first()) {
        IST_CUDA(cudaFuncSetAttribute(conv_first_dgrad_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, CFD_SMEM));
    }
    const double px = (double)NB * H * W;
    launch_pre("conv_first_dgrad", 2.0 * px * 64 * 27, px * (12.0 + 256.0), st);
    dim3 grid((W + CFD_TX - 1) / CFD_TX, (H + CFD_TY - 1) / CFD_TY, NB);
    IST_CUDA(launch_k(conv_first_dgrad_kernel<64>, grid, dim3(256), CFD_SMEM, st, pdl ? PDL_EW : 0, g_hi, g_lo, w, grad, NB, H, W));
    launch_post(st);
    return IST_OK;
}

// conv1_1 forward on the tensor cores (conv_first_fwd_tc.cuh). IST_B200_CFF=cuda selects the CUDA-core kernel.
inline int& cff_flag() { static int v = -1; return v; }      // -1 = not decided yet (environment), ist_set_option overrides
inline int cff_use_tc() {
    int& v = cff_flag();
    if (v < 0) {
        const char* e = getenv("IST_B200_CFF");
        v = (e != nullptr && strcmp(e, "cuda") == 0) ? 0 : 1;
    }
    return v;
}
inline int launch_conv_first_fwd_tc(cudaStream_t st, const CUtensorMap& o_hi, const CUtensorMap& o_lo, const float* x, const float* w,
                                    const float* bias, int NB, int H, int W, float out_scale, int pdl) {
    static DeviceOnce attr_once;
    if (attr_once.first()) {
        IST_CUDA(cudaFuncSetAttribute(conv_first_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CffTcCfg::SMEM_BYTES));
    }
    CffTcParams p;
    memset(&p, 0, sizeof(p));
    p.NB = NB; p.H = H; p.W = W;
    p.tiles_x = (W + CffTcCfg::TW - 1) / CffTcCfg::TW;
    p.tiles_y = (H + CffTcCfg::TH - 1) / CffTcCfg::TH;
    p.x = x; p.w = w; p.bias = bias; p.out_scale = out_scale;
    p.idesc = umma_idesc_f16(UMMA_FMT_F16, 128, 64, 0, 0);
    const long long total = (long long)NB * p.tiles_x * p.tiles_y;
    const int grid = total < num_sms() ? (int)total : num_sms();
    const double px = (double)NB * H * W;
    launch_pre("conv_first_fwd", 2.0 * px * 64 * 27, px * (12.0 + 256.0), st);
    IST_CUDA(launch_k(conv_first_fwd_tc_kernel, dim3(grid), dim3(CffTcCfg::THREADS), CffTcCfg::SMEM_BYTES, st, pdl, o_hi, o_lo, p));
    launch_post(st);
    return IST_OK;
}
// conv1_1 data-gradient on the tensor cores (conv_first_tc.cuh). IST_B200_CFD=cuda selects the CUDA-core kernel.
inline int& cfd_flag() { static int v = -1; return v; }
inline int cfd_use_tc() {
    int& v = cfd_flag();
    if (v < 0) {
        const char* e = getenv("IST_B200_CFD");
        v = (e != nullptr && strcmp(e, "cuda") == 0) ? 0 : 1;
    }
    return v;
}
inline int launch_conv_first_dgrad_tc(cudaStream_t st, const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi,
                                      const CUtensorMap& b_lo, float* grad, int NB, int H, int W, bool pdl) {
    static DeviceOnce attr_once;
    if (attr_once.first()) {
        IST_CUDA(cudaFuncSetAttribute(conv_first_dgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CfdTcCfg::SMEM_BYTES));
    }
    CfdTcParams p;
    memset(&p, 0, sizeof(p));
    p.NB = NB; p.H = H; p.W = W;
    p.tiles_x = (W + CfdTcCfg::TW - 1) / CfdTcCfg::TW;
    p.tiles_y = (H + CfdTcCfg::TH - 1) / CfdTcCfg::TH;
    p.grad = grad;
    p.idesc = umma_idesc_f16(UMMA_FMT_BF16, 128, CfdTcCfg::N_PAD, 0, 0);
    p.idesc32 = umma_idesc_f16(UMMA_FMT_BF16, 128, 2 * CfdTcCfg::N_PAD, 0, 0);
    const long long total = (long long)NB * p.tiles_x * p.tiles_y;
    const int grid = total < num_sms() ? (int)total : num_sms();
    const double px = (double)NB * H * W;
    launch_pre("conv_first_dgrad", 2.0 * px * 64 * 27, px * (12.0 + 256.0), st);
    IST_CUDA(launch_k(conv_first_dgrad_tc_kernel, dim3(grid), dim3(224), CfdTcCfg::SMEM_BYTES, st, pdl ? PDL_TENSOR : 0, a_hi, a_lo, b_hi, b_lo, p));
    launch_post(st);
    return IST_OK;
}

// The pixel split of a frame's Gram matrix depends on the frame geometry only, never on the batch size: every frame of a batch
// is reduced with the partition (hence the rounding) of its single-frame run.
inline void gram_split_plan(int NB, int HW, int C, int* splits, int* chunks_per_split) {
    (void)NB;
    const int tiles_c = (C + 127) / 128;
    const int tri = tiles_c * (tiles_c + 1) / 2;
    const int total_chunks = (HW + 63) / 64;
    int want = (num_sms() + tri - 1) / tri;     // ~one CTA per SM for one frame; accuracy does not depend on the split (register promotion)
    if (want < 1) want = 1;
    if (want > total_chunks) want = total_chunks;
    if (want > 160) want = 160;
    // ... but a split of very few 64-pixel chunks costs more than it gives: every split writes a full C x C fp32 partial that
    // gram_reduce_kernel reads back (relu4_1 / relu5_1: 15 splits x 1 MB each for 1-4 chunks of tensor work). At least
    // `min_chunks` chunks per split (IST_B200_GRAM_MIN_CHUNKS overrides).
    static int min_chunks = 0;
    if (min_chunks == 0) {
        const char* e = getenv("IST_B200_GRAM_MIN_CHUNKS");
        min_chunks = (e != nullptr && atoi(e) > 0) ? atoi(e) : 1;
    }
    if (want > (total_chunks + min_chunks - 1) / min_chunks) want = (total_chunks + min_chunks - 1) / min_chunks;
    if (want < 1) want = 1;
    const int cps = (total_chunks + want - 1) / want;
    *chunks_per_split = cps;
    *splits = (total_chunks + cps - 1) / cps;
}
struct GramLaunch {                     // one layer of a Gram launch
    const CUtensorMap *m_hi, *m_lo;
    int HW, C, splits, chunks_per_split;
    float* partial;
};
inline int launch_gram_multi(cudaStream_t st, int n, const GramLaunch* layers, int NB, int passes, bool pdl = false) {
    static DeviceOnce attr_once;
    if (attr_once.first()) {
        IST_CUDA(cudaFuncSetAttribute(gram_syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GramCfg::SMEM_BYTES));
    }
    if (n < 1 || n > GRAM_MAX_LAYERS) return fail(IST_ERR_ARG, "gram launch of %d layers (1..%d)", n, GRAM_MAX_LAYERS);
    GramMaps maps;
    GramMulti mp;
    memset(&mp, 0, sizeof(mp));
    mp.n = n;
    int blocks = 0;
    double flops = 0, bytes = 0;
    for (int l = 0; l < n; ++l) {
        const GramLaunch& g = layers[l];
        if (g.C % 64 != 0 || (g.C > 64 && g.C % 128 != 0)) return fail(IST_ERR_ARG, "gram needs C == 64 or C %% 128 == 0 (got %d)", g.C);
        GramParams& p = mp.L[l];
        p.NB = NB; p.HW = g.HW; p.C = g.C;
        p.tiles_c = (g.C + 127) / 128;
        p.n_tile = g.C < 128 ? 64 : 128;
        p.splits = g.splits;
        p.chunks_per_split = g.chunks_per_split;
        p.passes = passes;
        p.promote = promote_steps();
        p.idesc = umma_idesc_f16(UMMA_FMT_F16, 128, p.n_tile, 1, 1);
        p.partial = g.partial;
        maps.hi[l] = *g.m_hi;
        maps.lo[l] = *g.m_lo;
        const int tri = p.tiles_c * (p.tiles_c + 1) / 2;
        blocks += g.splits * tri * NB;
        mp.blk_end[l] = blocks;
        flops += 2.0 * NB * (double)g.HW * g.C * g.C;
        bytes += 4.0 * NB * (double)g.HW * g.C + 4.0 * NB * g.splits * (double)g.C * g.C;
    }
    for (int l = n; l < GRAM_MAX_LAYERS; ++l) { maps.hi[l] = maps.hi[0]; maps.lo[l] = maps.lo[0]; mp.blk_end[l] = blocks; }
    launch_pre("gram_syrk", flops, bytes, st);
    IST_CUDA(launch_k(gram_syrk_kernel, dim3(blocks), dim3(192), GramCfg::SMEM_BYTES, st, pdl ? PDL_TENSOR : 0, maps, mp));
    launch_post(st);
    return IST_OK;
}
inline int launch_gram(cudaStream_t st, const CUtensorMap& m_hi, const CUtensorMap& m_lo, int NB, int HW, int C,
                       int splits, int chunks_per_split, float* partial, int passes, bool pdl = false) {
    GramLaunch g = {&m_hi, &m_lo, HW, C, splits, chunks_per_split, partial};
    return launch_gram_multi(st, 1, &g, NB, passes, pdl);
}

inline int ew_grid(size_t work_items, int block) {
    size_t g = (work_items + block - 1) / block;
    const size_t cap = (size_t)num_sms() * 32;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// launch an elementwise / reduction kernel with bookkeeping: IST_EW("name", bytes, stream, kernel<<<...>>>(...));
#define IST_EW(name, bytes, st, ...)                 \
    do {                                             \
        ::ist::launch_pre(name, 0.0, (double)(bytes), st); \
        __VA_ARGS__;                                 \
        ::ist::launch_post(st);                      \
        IST_CUDA(cudaGetLastError());                \
    } while (0)

// same bookkeeping, launched through launch_k (programmatic dependent launch when `pdl`):
// IST_EWK("name", bytes, stream, pdl, kernel, grid, block, smem, args...)
#define IST_EWK(name, bytes, st, pdl, kernel, grid, block, smem, ...)                                   \
    do {                                                                                                \
        ::ist::launch_pre(name, 0.0, (double)(bytes), st);                                              \
        IST_CUDA(::ist::launch_k(kernel, dim3(grid), dim3(block), smem, st, pdl, __VA_ARGS__));         \
        ::ist::launch_post(st);                                                                         \
    } while (0)

struct DevMem {
    std::vector<void*> ptrs;
    size_t bytes = 0;
    ~DevMem() { release(); }
    void release() {
        for (void* p : ptrs) cudaFree(p);
        ptrs.clear();
        bytes = 0;
    }
    template <typename T>
    int alloc(T** out, size_t count) {
        void* p = nullptr;
        const size_t b = ((count * sizeof(T) + 255) / 256) * 256 + 256;
        cudaError_t e = cudaMalloc(&p, b);
        if (e != cudaSuccess) return fail(IST_ERR_CUDA, "cudaMalloc(%zu) failed: %s", b, cudaGetErrorString(e));
        ptrs.push_back(p);
        bytes += b;
        *out = reinterpret_cast<T*>(p);
        return IST_OK;
    }
};

}  // namespace ist

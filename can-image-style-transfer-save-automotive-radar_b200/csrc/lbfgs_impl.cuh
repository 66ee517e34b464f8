// Device-side L-BFGS with the semantics of torch.optim.LBFGS(lr=1, max_iter=20, max_eval=25, tolerance_grad=1e-7,
// tolerance_change=1e-9, history_size=100, line_search_fn=None) as the reference uses it
// (IST/model/engine/utils.py:24,43; torch/optim/lbfgs.py:333-537 of the torch 2.11 the reference path runs on here).
//
// Why it exists: on a B200 the stock optimiser's ~4m+15 tiny kernels and ~2m+5 host syncs per iteration cost more than
// the whole closure. Here one optimizer.step() (20 closure evaluations + 20 updates) is one CUDA graph
// and one host sync.
//
// Same algorithm, rearranged arithmetic. The two-loop recursion needs the dot products s_i.q and y_i.r of vectors that
// change inside the loops; expanding q = -g - sum al_j y_j and r = H q + sum c_j s_j turns them into combinations of
//   sg_i = s_i.g, yg_i = y_i.g, SY_ij = s_i.y_j, YY_ij = y_i.y_j
// so an iteration is: (1) ONE pass over the history computing all new dot products (lbfgs_dots_kernel, HBM-bound),
// (2) the O(m^2) scalar recursion in fp64 on one CTA (lbfgs_solve_kernel, also evaluates every break condition of the
// reference and keeps the (y,s) ring, ro, H_diag, t, n_iter state), (3) ONE pass forming d = cg*g + sum cy_j y_j + cs_j s_j,
// pushing the new (y,s) pair, saving prev_flat_grad and applying x += t*d (lbfgs_update_kernel). Break conditions set a
// per-frame `active` flag on the device; later kernels of the step become no-ops for that frame, exactly like `break`.
// Rounding differs from torch (fp64 dot accumulation, different summation order); the step rule, the ys > 1e-10 gate,
// H_diag, history eviction, evaluation counting and every tolerance test are the reference's.
// Frames of a batch are independent problems with independent optimiser state.
#pragma once
#include "host_common.cuh"

namespace ist {
int plan_batch(const ist_plan* P);
int plan_device(const ist_plan* P);
int plan_image_elems(const ist_plan* P);
int plan_n_losses(const ist_plan* P);
void plan_set_pdl_first(ist_plan* P, bool on);

constexpr int LB_MAXH = 112;          // history_size must be < LB_MAXH (shared-memory budget of the solve kernel)
constexpr int LB_LDG = LB_MAXH + 1;   // leading dimension of the dot-product matrices (odd: bank-conflict-free column walks in shared memory)
constexpr int LB_WT = 512;            // elements per warp-tile (16 per lane)
constexpr int LB_NSTAT = 8;           // ys, yy, s.g, y.g, g.g, |g|_1, max|g|, (unused)
constexpr int LB_PART = LB_MAXH * 5 + LB_NSTAT;

struct LbFrame {
    int n_iter, func_evals, hist_len, head;
    int active, current_evals, apply, accepted, new_slot, nread;
    double H_diag, t, prev_loss, orig_loss, loss, gtd;
    double ro[LB_MAXH];
    // outputs of the solve for the update pass
    float cg, t_f, t_prev_f;                // cg != 0 marks "an iteration was computed" (lbfgs_update_kernel)
    double cg_d, cy_new, cs_new;            // coefficients of d in float64: d = cg*g + cy_new*y_new + cs_new*s_new + sum ...
    int read_slot[LB_MAXH];
    double read_cy[LB_MAXH], read_cs[LB_MAXH];
};

struct LbParams {
    int NB, n, m;                 // frames, elements per frame, history size
    int nblk;                     // CTAs per frame in the update pass
    int nblk_dots;                // CTAs per frame in the dot-product pass
    float* x;                     // [NB][n] (caller's image, fp32 NCHW)
    const float* g;               // [NB][n] gradient of the last closure
    const float* losses;          // [NB][loss_stride], total at index loss_total
    int loss_stride, loss_total;
    float* prev_g;                // [NB][n]
    float* d;                     // [NB][n]
    float* S;                     // [m][NB][n]
    float* Y;                     // [m][NB][n]
    double* part;                 // [NB][nblk][LB_PART]
    double* tot;                  // [NB][LB_PART] fixed-order sums (max for the |g| entry) of `part`
    float* dmax_part;             // [NB][nblk]
    double* SY;                   // [NB][LB_MAXH][LB_LDG]  s_i.y_j by physical slot
    double* YY;                   // [NB][LB_MAXH][LB_LDG]
    LbFrame* frames;              // [NB]
    int it;                       // 1-based iteration index inside this step()
    int max_iter, max_eval;
    double lr, tol_grad, tol_change;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// float64 shuffle tree: the per-lane partial sums (16 fp32 products each) are added exactly from here on. With an fp32 tree the
// dot products of a 512-element tile carried ~3e-7 of sum|a_i b_i|; near-orthogonal history vectors amplify that through the
// two-loop recursion (tools/lbfgs_rounding_study.py: direction error vs float64 2.6e-5 -> 3e-6 together with the float64
// accumulation of d in lbfgs_update_kernel; torch's own fp32 arithmetic sits at 9e-6).
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// 16 elements per lane of a 512-element warp tile: element (k, lane, j) = base + k*128 + lane*4 + j
__device__ __forceinline__ void lb_load16(const float* __restrict__ p, size_t base, int n_left, int lane, bool vec, float (&v)[16]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int off = k * 128 + lane * 4;
        if (vec && off + 3 < n_left) {
            const float4 t = *reinterpret_cast<const float4*>(p + base + off);
            v[4 * k] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) v[4 * k + j] = (off + j < n_left) ? p[base + off + j] : 0.f;
        }
    }
}
__device__ __forceinline__ void lb_store16(float* __restrict__ p, size_t base, int n_left, int lane, bool vec, const float (&v)[16]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int off = k * 128 + lane * 4;
        if (vec && off + 3 < n_left) {
            *reinterpret_cast<float4*>(p + base + off) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (off + j < n_left) p[base + off + j] = v[4 * k + j];
        }
    }
}

// 8 elements per lane of a 256-element warp tile (update pass): element (k, lane, j) = base + k*128 + lane*4 + j
constexpr int LB_UT = 256;
__device__ __forceinline__ void lb_load8(const float* __restrict__ p, size_t base, int n_left, int lane, bool vec, float (&v)[8]) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int off = k * 128 + lane * 4;
        if (vec && off + 3 < n_left) {
            const float4 t = *reinterpret_cast<const float4*>(p + base + off);
            v[4 * k] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) v[4 * k + j] = (off + j < n_left) ? p[base + off + j] : 0.f;
        }
    }
}
__device__ __forceinline__ void lb_store8(float* __restrict__ p, size_t base, int n_left, int lane, bool vec, const float (&v)[8]) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int off = k * 128 + lane * 4;
        if (vec && off + 3 < n_left) {
            *reinterpret_cast<float4*>(p + base + off) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (off + j < n_left) p[base + off + j] = v[4 * k + j];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// pass 1: all dot products of this iteration. grid (nblk_dots, NB), 12 warps; each warp owns whole warp-tiles, so there
// is no block-level synchronisation inside the history loop.
// ---------------------------------------------------------------------------------------------------------------
// The history tiles (2 KB of S_i and 2 KB of Y_i per warp and pair) arrive through a per-warp shared-memory ring filled by 1-D
// bulk copies (cp.async.bulk + mbarrier): LB_DOTS_DEPTH pairs = 12 KB per warp are in flight whatever the register budget,
// ~18 MB over the GPU, which is what HBM needs at its latency (register-staged loads kept 4 KB per warp in flight and ran at
// 4.7 TB/s). One CTA per SM, 12 warps, every warp owns whole warp-tiles.
constexpr int LB_DOTS_WARPS = 12, LB_DOTS_DEPTH = 3;
constexpr int LB_DOTS_RING = LB_DOTS_WARPS * LB_DOTS_DEPTH * 2 * LB_WT * 4;                 // 147456
constexpr int LB_DOTS_BARS = 512;                                                             // 36 mbarriers
constexpr int LB_DOTS_SMEM = LB_DOTS_RING + LB_DOTS_BARS + LB_DOTS_WARPS * LB_PART * 8 + 1024;  // + per-warp fp64 sums
__global__ void __launch_bounds__(LB_DOTS_WARPS * 32, 1)
lbfgs_dots_kernel(const LbParams P) {
    extern __shared__ uint8_t lb_ring_raw[];
    const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const LbFrame& F = P.frames[b];
    // dynamic shared memory: [ring of every warp][mbarriers][per-warp fp64 accumulators]
    const uint32_t ring_base = (smem_u32(lb_ring_raw) + 1023u) & ~1023u;
    double (*wacc)[LB_PART] = reinterpret_cast<double (*)[LB_PART]>(lb_ring_raw + (ring_base - smem_u32(lb_ring_raw)) + LB_DOTS_RING + LB_DOTS_BARS);
    for (int i = lane; i < LB_PART; i += 32) wacc[warp][i] = 0.0;
    wacc[warp][5 * LB_MAXH + 6] = 0.0;
    // ring of this warp: LB_DOTS_DEPTH stages of [S tile 2 KB][Y tile 2 KB], one mbarrier per stage
    const uint32_t my_ring = ring_base + (uint32_t)warp * LB_DOTS_DEPTH * 2 * LB_WT * 4;
    const uint32_t my_bars = ring_base + LB_DOTS_RING + (uint32_t)warp * LB_DOTS_DEPTH * 8;
    const float* ring_f = reinterpret_cast<const float*>(lb_ring_raw + (my_ring - smem_u32(lb_ring_raw)));
    if (lane == 0) {
        for (int s = 0; s < LB_DOTS_DEPTH; ++s) mbar_init(my_bars + 8u * s, 1);
        fence_barrier_init();
    }
    __syncwarp();
    pdl_trigger();
    pdl_wait();          // the frame state, the gradient and the history are written by the previous kernels of the stream
    const bool skip = (P.it > 1 && !F.active);
    const int n = P.n;
    const size_t fo = (size_t)b * n;
    const bool vec = ((n & 3) == 0);
    const bool has_prev = F.n_iter >= 1;
    const float t = (float)F.t;
    const int hist = F.hist_len, head = F.head, m = P.m;
    const int ntiles = (n + LB_WT - 1) / LB_WT;
    float gmax = 0.f;
    uint32_t used = 0;                   // stage uses of this warp so far (ring position = used % DEPTH, parity = used / DEPTH)
    if (!skip) {
        for (int tile = blockIdx.x * LB_DOTS_WARPS + warp; tile < ntiles; tile += gridDim.x * LB_DOTS_WARPS) {
            const size_t base = (size_t)tile * LB_WT;
            const int n_left = n - (int)base;
            const uint32_t tile_bytes = (uint32_t)((n_left < LB_WT ? n_left : LB_WT) * 4);
            auto slot_of = [&](int i) { int s = head + i; return s >= m ? s - m : s; };
            auto issue = [&](int i, uint32_t use) {          // lane 0: fetch pair i into ring stage use % DEPTH
                const uint32_t st = use % LB_DOTS_DEPTH;
                const size_t ho = ((size_t)slot_of(i) * P.NB + b) * (size_t)n + base;
                const uint32_t dst = my_ring + st * (2 * LB_WT * 4), bar = my_bars + 8u * st;
                mbar_arrive_expect_tx(bar, 2 * tile_bytes);
                bulk_load_1d(dst, P.S + ho, tile_bytes, bar);
                bulk_load_1d(dst + LB_WT * 4, P.Y + ho, tile_bytes, bar);
            };
            if (vec && lane == 0)
                for (int i = 0; i < LB_DOTS_DEPTH && i < hist; ++i) issue(i, used + (uint32_t)i);
            float gv[16], yv[16], sv[16];
            lb_load16(P.g + fo, base, n_left, lane, vec, gv);
            if (has_prev) {
                lb_load16(P.prev_g + fo, base, n_left, lane, vec, yv);
                lb_load16(P.d + fo, base, n_left, lane, vec, sv);
            }
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f, a5 = 0.f;
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                if (has_prev) { yv[e] = gv[e] - yv[e]; sv[e] = t * sv[e]; } else { yv[e] = 0.f; sv[e] = 0.f; }
                a0 = fmaf(yv[e], sv[e], a0);
                a1 = fmaf(yv[e], yv[e], a1);
                a2 = fmaf(sv[e], gv[e], a2);
                a3 = fmaf(yv[e], gv[e], a3);
                a4 = fmaf(gv[e], gv[e], a4);
                a5 += fabsf(gv[e]);
                gmax = fmaxf(gmax, fabsf(gv[e]));
            }
            const double b0 = warp_sum_d((double)a0), b1 = warp_sum_d((double)a1), b2 = warp_sum_d((double)a2);
            const double b3 = warp_sum_d((double)a3), b4 = warp_sum_d((double)a4), b5 = warp_sum_d((double)a5);
            if (lane == 0) {
                double* st = &wacc[warp][5 * LB_MAXH];
                st[0] += b0; st[1] += b1; st[2] += b2; st[3] += b3; st[4] += b4; st[5] += b5;
            }
            for (int i = 0; i < hist; ++i) {
                const int slot = slot_of(i);
                float s_i[16], y_i[16];
                if (vec) {
                    const uint32_t use = used + (uint32_t)i, st = use % LB_DOTS_DEPTH;
                    mbar_wait(my_bars + 8u * st, (use / LB_DOTS_DEPTH) & 1u);
                    const float* sp = ring_f + st * (2 * LB_WT);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int off = k * 128 + lane * 4;
                        const bool in = off + 3 < n_left;          // n % 4 == 0: a float4 is inside or outside as a whole
                        const float4 a = in ? *reinterpret_cast<const float4*>(sp + off) : make_float4(0.f, 0.f, 0.f, 0.f);
                        const float4 c = in ? *reinterpret_cast<const float4*>(sp + LB_WT + off) : make_float4(0.f, 0.f, 0.f, 0.f);
                        s_i[4 * k] = a.x; s_i[4 * k + 1] = a.y; s_i[4 * k + 2] = a.z; s_i[4 * k + 3] = a.w;
                        y_i[4 * k] = c.x; y_i[4 * k + 1] = c.y; y_i[4 * k + 2] = c.z; y_i[4 * k + 3] = c.w;
                    }
                    __syncwarp();                                    // every lane has read the stage
                    if (lane == 0 && i + LB_DOTS_DEPTH < hist) {
                        fence_proxy_async_smem();                    // generic reads before the async-proxy refill
                        issue(i + LB_DOTS_DEPTH, use + LB_DOTS_DEPTH);
                    }
                } else {
                    const size_t ho = ((size_t)slot * P.NB + b) * (size_t)n;
                    lb_load16(P.S + ho, base, n_left, lane, vec, s_i);
                    lb_load16(P.Y + ho, base, n_left, lane, vec, y_i);
                }
                float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f, c4 = 0.f;
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    c0 = fmaf(s_i[e], yv[e], c0);   // s_i . y_new
                    c1 = fmaf(y_i[e], sv[e], c1);   // y_i . s_new
                    c2 = fmaf(y_i[e], yv[e], c2);   // y_i . y_new
                    c3 = fmaf(s_i[e], gv[e], c3);   // s_i . g
                    c4 = fmaf(y_i[e], gv[e], c4);   // y_i . g
                }
                const double e0 = warp_sum_d((double)c0), e1 = warp_sum_d((double)c1), e2 = warp_sum_d((double)c2);
                const double e3 = warp_sum_d((double)c3), e4 = warp_sum_d((double)c4);
                if (lane == 0) {
                    double* a = &wacc[warp][slot * 5];
                    a[0] += e0; a[1] += e1; a[2] += e2; a[3] += e3; a[4] += e4;
                }
            }
            if (vec) used += (uint32_t)hist;
        }
        gmax = warp_max(gmax);
        if (lane == 0) wacc[warp][5 * LB_MAXH + 6] = gmax;
    }
    __syncthreads();
    double* out = P.part + ((size_t)b * P.nblk_dots + blockIdx.x) * LB_PART;
    for (int i = threadIdx.x; i < LB_PART; i += blockDim.x) {
        double s = 0.0;
        if (i == 5 * LB_MAXH + 6) {
            for (int w = 0; w < LB_DOTS_WARPS; ++w) s = fmax(s, wacc[w][i]);
        } else {
            for (int w = 0; w < LB_DOTS_WARPS; ++w) s += wacc[w][i];
        }
        out[i] = s;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// pass 1b: fixed-order reduction of the per-CTA partials, one warp per output (grid (ceil(LB_PART / 8), NB), 256 threads):
// lane l sums entries l, l + 32, ... in order, then a shuffle tree — the same order on every run.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
lbfgs_reduce_kernel(const LbParams P) {
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    pdl_trigger();
    pdl_wait();
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        // clears the "iteration computed" marker and remembers the step size of the direction that led to the gradient being
        // processed (the solve replaces F.t with the step size of the new direction)
        LbFrame& F = P.frames[b];
        F.cg = 0.f;
        F.t_prev_f = (float)F.t;
    }
    if (i >= LB_PART) return;
    const double* p = P.part + (size_t)b * P.nblk_dots * LB_PART + i;
    const bool is_max = (i == 5 * LB_MAXH + 6);
    double s = 0.0;
    for (int k = lane; k < P.nblk_dots; k += 32) {
        const double v = p[(size_t)k * LB_PART];
        s = is_max ? fmax(s, v) : s + v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double t = __shfl_xor_sync(0xffffffffu, s, o);
        s = is_max ? fmax(s, t) : s + t;
    }
    if (lane == 0) P.tot[(size_t)b * LB_PART + i] = s;
}

// ---------------------------------------------------------------------------------------------------------------
// pass 2: scalar logic of one iteration of LBFGS.step for one frame (one CTA of LB_SOLVE_THREADS threads).
// ---------------------------------------------------------------------------------------------------------------
constexpr int LB_SOLVE_THREADS = 512;
// dynamic shared memory of the solve: physical-slot copies of SY and YY, rows 0 .. m-1 (the ring only uses slots < m)
// (an even number of rows keeps both copies 16-byte aligned and sized, as cp.async.bulk needs)
inline __host__ __device__ int lb_rows(int m) { return (m + 1) & ~1; }
inline size_t lb_solve_smem(int m) { return 2 * (size_t)lb_rows(m) * LB_LDG * sizeof(double) + 16; }

__global__ void __launch_bounds__(LB_SOLVE_THREADS)
lbfgs_solve_kernel(const LbParams P) {
    extern __shared__ __align__(16) double sm[];
    constexpr int LB_LD = LB_LDG;
    double* SYs = sm;                          // [m][LB_LDG] by PHYSICAL slot: SYs[p][q] = s_p . y_q
    double* YYs = sm + (size_t)lb_rows(P.m) * LB_LD;    // [m][LB_LDG]
    __shared__ __align__(8) unsigned long long copy_bar;
    __shared__ double tot[LB_PART];
    __shared__ double al_s[LB_MAXH], c_s[LB_MAXH], sg_s[LB_MAXH], yg_s[LB_MAXH];
    __shared__ double red[LB_MAXH];
    __shared__ int sh_go, sh_h;
    const int b = blockIdx.x, tid = threadIdx.x, m = P.m;
    pdl_trigger();
    pdl_wait();
    LbFrame& F = P.frames[b];
    double* SYg = P.SY + (size_t)b * LB_MAXH * LB_LDG;
    double* YYg = P.YY + (size_t)b * LB_MAXH * LB_LDG;
    // The dot-product matrices of the previous iterations come in as two 1-D bulk copies (m * 113 doubles each, ~90 KB at
    // m = 100) that overlap the scalar logic below. Gathering them element by element in logical order (the previous form)
    // was a chain of ~20 dependent L2 round trips per thread: 30 of the solve's 39 us at full history.
    const uint32_t mat_bytes = (uint32_t)((size_t)lb_rows(m) * LB_LDG * sizeof(double));
    const uint32_t bar = smem_u32(&copy_bar);
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
        mbar_arrive_expect_tx(bar, 2 * mat_bytes);
        bulk_load_1d(smem_u32(SYs), SYg, mat_bytes, bar);
        bulk_load_1d(smem_u32(YYs), YYg, mat_bytes, bar);
    }

    for (int i = tid; i < LB_PART; i += blockDim.x) tot[i] = P.tot[(size_t)b * LB_PART + i];
    __syncthreads();
    const double* st = &tot[5 * LB_MAXH];
    const double ys = st[0], yy = st[1], sg_new = st[2], yg_new = st[3], gg = st[4], g1 = st[5], gmax = st[6];

    // max |d| of the previous update pass: per-CTA maxima reduced by the whole CTA (thread 0 walking up to 296 values alone was a
    // chain of L2 round trips, most of the solve's 39 us)
    __shared__ float sh_dm[LB_SOLVE_THREADS / 32];
    {
        float dmv = 0.f;
        for (int k = tid; k < P.nblk; k += blockDim.x) dmv = fmaxf(dmv, P.dmax_part[(size_t)b * P.nblk + k]);
        dmv = warp_max(dmv);
        if ((tid & 31) == 0) sh_dm[tid >> 5] = dmv;
    }
    __syncthreads();
    // The frame state lives in global memory: thread 0 reads every scalar it needs up front (independent loads, one L2 round
    // trip), decides on registers and writes the changed fields back once — a read-modify-write per field interleaved with the
    // control flow was a chain of ~20 dependent L2 round trips.
    __shared__ int sh_head, sh_accepted, sh_new_slot, sh_n_iter;
    __shared__ double sh_H, sh_loss;
    if (tid == 0) {
        const double loss = (double)P.losses[(size_t)b * P.loss_stride + P.loss_total];
        int n_iter = F.n_iter, func_evals = F.func_evals, hist_len = F.hist_len, head = F.head, active = F.active;
        int current_evals = F.current_evals;
        double H_diag = F.H_diag;
        const double t_prev = F.t, prev_loss = F.prev_loss;
        int go = 1, accepted = 0, new_slot = 0;
        const bool evaluated = (P.it == 1) || active != 0;      // this closure evaluation counts for the frame
        if (P.it == 1) {                        // lbfgs.py:364-374: first closure of step()
            active = 1;
            current_evals = 1;
            func_evals += 1;
            F.orig_loss = loss;
            if (gmax <= P.tol_grad) { active = 0; go = 0; }
        } else if (active) {                    // lbfgs.py:493-523: the checks that follow the re-evaluation
            current_evals += 1;
            func_evals += 1;
            float dm = 0.f;
            for (int k = 0; k < LB_SOLVE_THREADS / 32; ++k) dm = fmaxf(dm, sh_dm[k]);
            if (current_evals >= P.max_eval) go = 0;
            else if (gmax <= P.tol_grad) go = 0;
            else if ((double)dm * fabs(t_prev) <= P.tol_change) go = 0;
            else if (fabs(loss - prev_loss) < P.tol_change) go = 0;
            if (!go) active = 0;
        } else {
            go = 0;
        }
        if (go) {
            n_iter += 1;
            if (n_iter == 1) {                  // lbfgs.py:396-401
                hist_len = 0; head = 0; H_diag = 1.0;
            } else if (ys > 1e-10) {            // lbfgs.py:404-420
                if (hist_len == m) { new_slot = head; head = (head + 1) % m; }
                else { new_slot = (head + hist_len) % m; hist_len += 1; }
                accepted = 1;
                F.ro[new_slot] = 1.0 / ys;
                H_diag = ys / yy;
            }
        }
        F.active = active; F.current_evals = current_evals; F.func_evals = func_evals;
        if (evaluated) F.loss = loss;
        F.apply = 0; F.accepted = accepted; F.nread = 0; F.new_slot = new_slot;
        F.n_iter = n_iter; F.hist_len = hist_len; F.head = head; F.H_diag = H_diag;
        sh_go = go; sh_h = hist_len; sh_head = head; sh_accepted = accepted; sh_new_slot = new_slot; sh_n_iter = n_iter;
        sh_H = H_diag; sh_loss = loss;
    }
    __syncthreads();
    mbar_wait(bar, 0);                          // the copies have landed (every thread waits: the CTA may not exit before them)
    if (!sh_go) return;
    const int h = sh_h, head = sh_head, accepted = sh_accepted, new_slot = sh_new_slot;
    auto phys = [&](int i) { int s = head + i; return s >= m ? s - m : s; };

    // fold the new pair into the physical-slot matrices (row / column new_slot), in shared memory for this iteration and in
    // global memory for the next ones
    if (accepted) {
        for (int i = tid; i < h; i += blockDim.x) {
            const int p = phys(i);
            if (p == new_slot) continue;
            const double sy_in = tot[p * 5 + 0], sy_ni = tot[p * 5 + 1], yy_in = tot[p * 5 + 2];
            SYs[p * LB_LD + new_slot] = sy_in;                          // s_i . y_new
            SYs[new_slot * LB_LD + p] = sy_ni;                          // s_new . y_i
            YYs[p * LB_LD + new_slot] = yy_in;
            YYs[new_slot * LB_LD + p] = yy_in;
            SYg[(size_t)p * LB_LDG + new_slot] = sy_in;
            SYg[(size_t)new_slot * LB_LDG + p] = sy_ni;
            YYg[(size_t)p * LB_LDG + new_slot] = yy_in;
            YYg[(size_t)new_slot * LB_LDG + p] = yy_in;
        }
        if (tid == 0) {
            SYs[new_slot * LB_LD + new_slot] = ys;
            YYs[new_slot * LB_LD + new_slot] = yy;
            SYg[(size_t)new_slot * LB_LDG + new_slot] = ys;
            YYg[(size_t)new_slot * LB_LDG + new_slot] = yy;
        }
    }
    for (int i = tid; i < h; i += blockDim.x) {
        const int p = phys(i);
        const bool is_new = accepted && p == new_slot;
        sg_s[i] = is_new ? sg_new : tot[p * 5 + 3];
        yg_s[i] = is_new ? yg_new : tot[p * 5 + 4];
    }
    __syncthreads();

    // The two sequential loops of the recursion are triangular solves with R = triu(SY) (logical order, diagonal SY_ii = 1 / ro_i):
    //   first loop  (lbfgs.py:431-436, newest to oldest):  R al = -sg                      (back substitution)
    //   second loop (lbfgs.py:440-443, oldest to newest):  c_i = al_i - ro_i (v_i + sum_{j<i} c_j SY_ji)   (forward substitution)
    // They run blocked on the first 4 warps (h <= 100 < 128, thread k owns unknown k): a 32 x 32 diagonal block is solved inside
    // ONE warp with shuffles (no barrier per unknown), the off-diagonal contributions of a finished block are a small dense
    // update by all 128 threads behind one named barrier per block. One barrier per unknown (the previous form) cost ~370 cycles
    // x 200 unknowns = 39 us per iteration at full history.
    const double H = sh_H;
    if (tid < 128) {
        const int lane = tid & 31, wq = tid >> 5;
        const int nblocks = (h + 31) >> 5;
        const int pk = tid < h ? phys(tid) : 0;                // physical slot (matrix row / column) of this thread's unknown
        const double ro_t = tid < h ? F.ro[pk] : 0.0;
        // FP64 instructions have a long dependent-issue latency here (~50 cycles measured through this kernel), so the chains are
        // kept short: inside a diagonal block the unknown of lane l is carried as z_l = ro_l * (rhs_l - u_l), updated with ONE
        // DFMA per finished unknown (row pre-scaled by ro, off the chain), so a step is shuffle + DFMA; the dense updates
        // between blocks and the y.r product run as four independent partial sums.
        double rhs = tid < h ? -sg_s[tid] : 0.0;              // -sg_k - sum over finished (later) blocks of al_j * SY_kj
        for (int bb = nblocks - 1; bb >= 0; --bb) {
            if (wq == bb) {
                // the pre-scaled row of this lane's unknown comes into registers first (32 independent loads), so that a step of
                // the sequential part is shuffle + DFMA only; unknowns beyond h are zeros that change nothing
                const int lim = (h - bb * 32) < 32 ? (h - bb * 32) : 32;
                double srow[32];
#pragma unroll
                for (int l = 0; l < 32; ++l) srow[l] = (lane < l && l < lim) ? -ro_t * SYs[pk * LB_LD + phys(bb * 32 + l)] : 0.0;
                double z = ro_t * rhs, mine = 0.0;
#pragma unroll
                for (int l = 31; l >= 0; --l) {
                    const double al_l = __shfl_sync(0xffffffffu, z, l);
                    if (lane == l) mine = al_l;
                    z = fma(al_l, srow[l], z);
                }
                if (tid < h) al_s[tid] = mine;
            }
            named_bar_sync(1, 128);
            if (tid < bb * 32) {
                const int j0 = bb * 32, j1 = (bb * 32 + 32) < h ? (bb * 32 + 32) : h;
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                int j = j0;
                for (; j + 4 <= j1; j += 4) {
                    a0 = fma(al_s[j], SYs[pk * LB_LD + phys(j)], a0);
                    a1 = fma(al_s[j + 1], SYs[pk * LB_LD + phys(j + 1)], a1);
                    a2 = fma(al_s[j + 2], SYs[pk * LB_LD + phys(j + 2)], a2);
                    a3 = fma(al_s[j + 3], SYs[pk * LB_LD + phys(j + 3)], a3);
                }
                for (; j < j1; ++j) a0 = fma(al_s[j], SYs[pk * LB_LD + phys(j)], a0);
                rhs -= (a0 + a1) + (a2 + a3);
            }
        }
        // r = H*q: y_k . r before the second loop
        double v = 0.0;
        if (tid < h) {
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            int j = 0;
            for (; j + 4 <= h; j += 4) {
                a0 = fma(al_s[j], YYs[pk * LB_LD + phys(j)], a0);
                a1 = fma(al_s[j + 1], YYs[pk * LB_LD + phys(j + 1)], a1);
                a2 = fma(al_s[j + 2], YYs[pk * LB_LD + phys(j + 2)], a2);
                a3 = fma(al_s[j + 3], YYs[pk * LB_LD + phys(j + 3)], a3);
            }
            for (; j < h; ++j) a0 = fma(al_s[j], YYs[pk * LB_LD + phys(j)], a0);
            v = H * (-yg_s[tid] - ((a0 + a1) + (a2 + a3)));
        }
        const double al_me = tid < h ? al_s[tid] : 0.0;
        double w = 0.0;                                         // sum over finished (earlier) blocks of c_j * SY_jk
        for (int bb = 0; bb < nblocks; ++bb) {
            if (wq == bb) {
                const int lim = (h - bb * 32) < 32 ? (h - bb * 32) : 32;
                double scol[32];
#pragma unroll
                for (int l = 0; l < 32; ++l) scol[l] = (lane > l && l < lim && tid < h) ? -ro_t * SYs[phys(bb * 32 + l) * LB_LD + pk] : 0.0;
                double q = al_me - ro_t * (v + w), mine = 0.0;   // c_k = q_k once every earlier unknown of the block is folded in
#pragma unroll
                for (int l = 0; l < 32; ++l) {
                    const double c_l = __shfl_sync(0xffffffffu, q, l);
                    if (lane == l) mine = c_l;
                    q = fma(c_l, scol[l], q);
                }
                if (tid < h) c_s[tid] = mine;
            }
            named_bar_sync(1, 128);
            if (tid >= bb * 32 + 32 && tid < h) {
                const int j0 = bb * 32;
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    a0 = fma(c_s[j0 + j], SYs[phys(j0 + j) * LB_LD + pk], a0);
                    a1 = fma(c_s[j0 + j + 1], SYs[phys(j0 + j + 1) * LB_LD + pk], a1);
                    a2 = fma(c_s[j0 + j + 2], SYs[phys(j0 + j + 2) * LB_LD + pk], a2);
                    a3 = fma(c_s[j0 + j + 3], SYs[phys(j0 + j + 3) * LB_LD + pk], a3);
                }
                w += (a0 + a1) + (a2 + a3);
            }
        }
    }
    __syncthreads();
    // gtd = g.d with d = -H g - sum H al_j y_j + sum c_j s_j
    if (tid < LB_MAXH) red[tid] = (tid < h) ? (-H * al_s[tid] * yg_s[tid] + c_s[tid] * sg_s[tid]) : 0.0;
    // coefficients of the update pass: an accepted pair is always the newest one, i.e. logical index h - 1
    const int nread = accepted ? h - 1 : h;
    if (tid < nread) {
        F.read_slot[tid] = phys(tid);
        F.read_cy[tid] = -H * al_s[tid];
        F.read_cs[tid] = c_s[tid];
    }
    __syncthreads();
    if (tid < 32) {
        // g.d: lane l adds red[l], red[l + 32], ... in order, then a fixed shuffle tree (deterministic)
        double part = 0.0;
        for (int k = tid; k < h; k += 32) part += red[k];
        part = warp_sum_d(part);
        if (tid == 0) red[0] = part;
    }
    __syncthreads();
    if (tid == 0) {
        const double gtd = -H * gg + red[0];
        F.gtd = gtd;
        F.prev_loss = sh_loss;                                           // lbfgs.py:449
        double t;
        if (sh_n_iter == 1) {                                            // lbfgs.py:454-457
            const float inv = 1.0f / (float)g1;
            t = (double)(inv < 1.0f ? inv : 1.0f) * P.lr;
        } else {
            t = P.lr;
        }
        F.t = t;
        F.t_f = (float)t;
        F.cg = (float)(-H);
        F.cg_d = -H;
        F.cy_new = accepted ? -H * al_s[h - 1] : 0.0;
        F.cs_new = accepted ? c_s[h - 1] : 0.0;
        F.nread = nread;
        if (gtd > -P.tol_change) { F.active = 0; F.apply = 0; }          // lbfgs.py:462-464 (break before the update)
        else F.apply = 1;
        // the direction, prev_flat_grad and the history are saved either way (lbfgs.py:525-535)
    }
}

// ---------------------------------------------------------------------------------------------------------------
// pass 3: d <- cg*g + cy_new*y_new + cs_new*s_new + sum_i cy_i*Y_i + cs_i*S_i; push (y_new, s_new); prev_g <- g;
// x += t*d when the solve allowed it; per-CTA max|d|.
// ---------------------------------------------------------------------------------------------------------------
// d is a sum of up to 2m + 1 vectors with large cancellations. With fp32 coefficients and plain fp32 accumulation its distance
// from float64 arithmetic was 3e-5 ... 1e-4 (relative L2; torch's own fp32 two-loop recursion: ~1e-5), which feeds back through
// s = t * d into the history. The sum is therefore formed as a compensated dot product in fp32 (Ogita / Rump / Oishi "Dot2":
// error-free product by FMA, error-free TwoSum, coefficients as fp32 hi + lo pairs) — the accuracy of a float64 accumulation
// without float64 instructions (an F2F.F64.F32 + DFMA version made this HBM-bound pass 2x slower: 108 -> 203 us per iteration,
// profiles/r02_lbfgs_ab.log). d is rounded to fp32 once, like torch's d.
// (Blackwell's packed FFMA2 / FADD2 / FMUL2 forms of the same step were slower still: 208 us, profiles/r02_lbfgs_ab.log.)
__device__ __forceinline__ void lb_dot2_step(float& hi, float& lo, float ch, float cl, float v) {
    const float p = __fmul_rn(ch, v);
    float pe = __fmaf_rn(ch, v, -p);                    // exact error of the product
    pe = __fmaf_rn(cl, v, pe);                          // low part of the coefficient
    const float s = __fadd_rn(hi, p);
    const float bb = __fsub_rn(s, hi);
    const float err = __fadd_rn(__fsub_rn(hi, __fsub_rn(s, bb)), __fsub_rn(p, bb));      // exact error of the sum
    lo = __fadd_rn(lo, __fadd_rn(err, pe));
    hi = s;
}
__device__ __forceinline__ void lb_split(double c, float& h, float& l) {
    h = (float)c;
    l = (float)(c - (double)h);
}

__global__ void __launch_bounds__(256, 3)
lbfgs_update_kernel(const LbParams P) {
    __shared__ float wmax[8];
    __shared__ int s_slot[LB_MAXH];
    __shared__ float s_cyh[LB_MAXH], s_cyl[LB_MAXH], s_csh[LB_MAXH], s_csl[LB_MAXH];
    const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const LbFrame& F = P.frames[b];
    pdl_trigger();
    pdl_wait();
    for (int i = threadIdx.x; i < F.nread; i += blockDim.x) {
        s_slot[i] = F.read_slot[i];
        lb_split(F.read_cy[i], s_cyh[i], s_cyl[i]);
        lb_split(F.read_cs[i], s_csh[i], s_csl[i]);
    }
    __syncthreads();
    float dmax = 0.f;
    // `F.cg != 0` marks "an iteration was computed": the solve writes cg = -H_diag (never 0) and the host clears it before
    // each solve; a frame that stopped earlier keeps cg == 0 and is left untouched.
    if (F.cg != 0.f) {
        const int n = P.n;
        const size_t fo = (size_t)b * n;
        const bool vec = ((n & 3) == 0);
        const bool has_prev = F.n_iter >= 2;      // a previous direction exists (n_iter was already incremented)
        const float t_old_new = F.t_f;
        const int ntiles = (n + LB_UT - 1) / LB_UT;      // 256-element warp tiles: twice the warps of a 512-element tiling,
                                                         // half the registers per thread — the compensated sum is latency-bound otherwise
        float cgh, cgl, cynh, cynl, csnh, csnl;
        lb_split(F.cg_d, cgh, cgl);
        lb_split(F.cy_new, cynh, cynl);
        lb_split(F.cs_new, csnh, csnl);
        const int nread = F.nread, accepted = F.accepted, new_slot = F.new_slot, apply = F.apply;
        for (int tile = blockIdx.x * 8 + warp; tile < ntiles; tile += gridDim.x * 8) {
            const size_t base = (size_t)tile * LB_UT;
            const int n_left = n - (int)base;
            float gv[8], hi[8], lo[8];
            lb_load8(P.g + fo, base, n_left, lane, vec, gv);
#pragma unroll
            for (int e = 0; e < 8; ++e) { hi[e] = 0.f; lo[e] = 0.f; lb_dot2_step(hi[e], lo[e], cgh, cgl, gv[e]); }
            if (has_prev) {
                float yv[8], sv[8];
                lb_load8(P.prev_g + fo, base, n_left, lane, vec, yv);
                lb_load8(P.d + fo, base, n_left, lane, vec, sv);
#pragma unroll
                for (int e = 0; e < 8; ++e) { yv[e] = gv[e] - yv[e]; }
                if (accepted) {
                    // s_new = t_prev * d_old (lbfgs.py:404): t_prev is the step size saved by lbfgs_reduce_kernel before the
                    // solve replaced F.t with the step size of the new direction
                    const float tp = F.t_prev_f;
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        sv[e] = tp * sv[e];
                        lb_dot2_step(hi[e], lo[e], csnh, csnl, sv[e]);
                        lb_dot2_step(hi[e], lo[e], cynh, cynl, yv[e]);
                    }
                    const size_t ho = ((size_t)new_slot * P.NB + b) * (size_t)n;
                    lb_store8(P.S + ho, base, n_left, lane, vec, sv);
                    lb_store8(P.Y + ho, base, n_left, lane, vec, yv);
                }
            }
            // the loads of pair i + 1 are issued before the (long, dependent) compensated sums of pair i
            float s_n[8], y_n[8];
            if (nread > 0) {
                const size_t h0 = ((size_t)s_slot[0] * P.NB + b) * (size_t)n;
                lb_load8(P.S + h0, base, n_left, lane, vec, s_n);
                lb_load8(P.Y + h0, base, n_left, lane, vec, y_n);
            }
            for (int i = 0; i < nread; ++i) {
                float s_i[8], y_i[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) { s_i[e] = s_n[e]; y_i[e] = y_n[e]; }
                if (i + 1 < nread) {
                    const size_t ho = ((size_t)s_slot[i + 1] * P.NB + b) * (size_t)n;
                    lb_load8(P.S + ho, base, n_left, lane, vec, s_n);
                    lb_load8(P.Y + ho, base, n_left, lane, vec, y_n);
                }
                const float cyh = s_cyh[i], cyl = s_cyl[i], csh = s_csh[i], csl = s_csl[i];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    lb_dot2_step(hi[e], lo[e], csh, csl, s_i[e]);
                    lb_dot2_step(hi[e], lo[e], cyh, cyl, y_i[e]);
                }
            }
            float dv[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) { dv[e] = __fadd_rn(hi[e], lo[e]); dmax = fmaxf(dmax, fabsf(dv[e])); }
            lb_store8(P.d + fo, base, n_left, lane, vec, dv);
            lb_store8(P.prev_g + fo, base, n_left, lane, vec, gv);
            if (apply) {
                float xv[8];
                lb_load8(P.x + fo, base, n_left, lane, vec, xv);
#pragma unroll
                for (int e = 0; e < 8; ++e) xv[e] = fmaf(t_old_new, dv[e], xv[e]);
                lb_store8(P.x + fo, base, n_left, lane, vec, xv);
            }
        }
    }
    dmax = warp_max(dmax);
    if (lane == 0) wmax[warp] = dmax;
    __syncthreads();
    if (threadIdx.x == 0) {
        float mx = 0.f;
        for (int w = 0; w < 8; ++w) mx = fmaxf(mx, wmax[w]);
        if (F.cg != 0.f) P.dmax_part[(size_t)b * P.nblk + blockIdx.x] = mx;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Diagnostics (tests only; never part of a graph unless ist_lbfgs_set_trace was called before the first step).
// Trace: one record per closure evaluation and frame — the evaluation point x, the gradient g the closure returned, the
// direction d the iteration computed from it and the scalar state after the solve. The records let a test feed exactly
// these gradients to a float64 restatement of torch/optim/lbfgs.py and compare every direction, step size, H_diag, the
// accepted / rejected curvature pairs and the exits ("teacher forcing": no chaotic amplification of rounding).
// ---------------------------------------------------------------------------------------------------------------
constexpr int LB_TRACE_SCALARS = 16;
struct LbTrace {
    float *x = nullptr, *g = nullptr, *d = nullptr;      // [cap][NB][n]
    double* sc = nullptr;                                // [cap][NB][LB_TRACE_SCALARS]
    int* count = nullptr;                                // device counter of complete records
    int cap = 0;
};
__global__ void lbfgs_trace_pre_kernel(const LbParams P, const LbTrace T) {
    const int e = *T.count, b = blockIdx.y;
    if (e >= T.cap) return;
    const size_t src = (size_t)b * P.n, dst = ((size_t)e * P.NB + b) * (size_t)P.n;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < (size_t)P.n; i += (size_t)gridDim.x * blockDim.x) {
        T.x[dst + i] = P.x[src + i];
        T.g[dst + i] = P.g[src + i];
    }
}
__global__ void lbfgs_trace_post_kernel(const LbParams P, const LbTrace T) {
    const int e = *T.count, b = blockIdx.y;
    if (e >= T.cap) return;
    const size_t src = (size_t)b * P.n, dst = ((size_t)e * P.NB + b) * (size_t)P.n;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < (size_t)P.n; i += (size_t)gridDim.x * blockDim.x)
        T.d[dst + i] = P.d[src + i];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const LbFrame& F = P.frames[b];
        const double* st = P.tot + (size_t)b * LB_PART + 5 * LB_MAXH;
        double* o = T.sc + ((size_t)e * P.NB + b) * LB_TRACE_SCALARS;
        o[0] = (double)P.losses[(size_t)b * P.loss_stride + P.loss_total];
        o[1] = F.cg != 0.f ? 1.0 : 0.0;       // an iteration (direction) was computed from this evaluation
        o[2] = F.active; o[3] = F.n_iter; o[4] = F.hist_len; o[5] = F.head; o[6] = F.accepted;
        o[7] = F.H_diag; o[8] = F.t; o[9] = F.gtd; o[10] = st[0]; o[11] = st[1]; o[12] = F.apply;
        o[13] = F.func_evals; o[14] = F.current_evals; o[15] = F.new_slot;
    }
}
__global__ void lbfgs_trace_bump_kernel(const LbTrace T) {
    if (*T.count < T.cap) *T.count += 1;
}

// Test objective (ist_lbfgs_create_test): per frame f(x) = sum_i 0.5 a_i (x_i - b_i)^2 + c_i cos(x_i), separable, with
// negative curvature where c_i cos(x_i) > a_i (rejected curvature pairs) and a closed-form float64 oracle. One CTA per
// frame, fixed-order reduction (run-to-run identical like every other reduction of this library).
struct LbTestObjective {
    const float *a = nullptr, *b = nullptr, *c = nullptr;      // [NB][n]
};
__global__ void __launch_bounds__(1024) lbfgs_test_objective_kernel(const LbTestObjective T, const float* __restrict__ x,
                                                                     float* __restrict__ g, float* __restrict__ losses, int n) {
    __shared__ double red[1024];
    const int b = blockIdx.x;
    const size_t fo = (size_t)b * n;
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float xi = x[fo + i], r = xi - T.b[fo + i], ai = T.a[fo + i], ci = T.c[fo + i];
        g[fo + i] = ai * r - ci * sinf(xi);
        acc += 0.5 * (double)ai * (double)r * (double)r + (double)ci * (double)cosf(xi);
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) losses[b] = (float)red[0];
}

}  // namespace ist

struct ist_lbfgs {
    ist_plan* plan = nullptr;
    int device = -1;
    ist::LbParams P;
    ist::DevMem mem;
    float* g = nullptr;
    float* losses = nullptr;
    int n_losses = 0;
    cudaStream_t cap_stream = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
    float* x_own = nullptr;          // optimiser-owned copy of the image: the graph is captured once on it, callers' x is copied in / out
    bool use_graph = true;
    unsigned long long graph_kernels = 0;
    std::vector<ist::LbFrame> host_frames;
    ist::LbTrace trace;                 // diagnostics (ist_lbfgs_set_trace)
    ist::LbTestObjective test_obj;      // closure of ist_lbfgs_create_test (plan == nullptr)
};

namespace ist {

inline int lbfgs_enqueue_step(ist_lbfgs* O, float* x, cudaStream_t st) {
    LbParams P = O->P;
    P.x = x;
    for (int it = 1; it <= P.max_iter; ++it) {
        if (O->plan != nullptr) {
            plan_set_pdl_first(O->plan, it > 1 && O->trace.cap == 0);      // the closure then follows lbfgs_update_kernel on the same stream
            const int rc_closure = ist_plan_loss_and_grad(O->plan, x, O->g, O->losses, (void*)st);
            plan_set_pdl_first(O->plan, false);
            IST_TRY(rc_closure);
        } else {
            IST_EW("lbfgs_test_objective", 16.0 * P.NB * P.n, st,
                   lbfgs_test_objective_kernel<<<P.NB, 1024, 0, st>>>(O->test_obj, x, O->g, O->losses, P.n));
        }
        P.it = it;
        const int tgrid = (P.n + 1023) / 1024 < 64 ? (P.n + 1023) / 1024 : 64;
        if (O->trace.cap > 0)
            IST_EW("lbfgs_trace", 0.0, st, lbfgs_trace_pre_kernel<<<dim3(tgrid, P.NB), 256, 0, st>>>(P, O->trace));
        const double vb = 4.0 * P.NB * (double)P.n;
        IST_EWK("lbfgs_dots", vb * (3 + 2.0 * P.m), st, PDL_OPT, lbfgs_dots_kernel, dim3(P.nblk_dots, P.NB), LB_DOTS_WARPS * 32, LB_DOTS_SMEM, P);
        IST_EWK("lbfgs_reduce", 8.0 * P.NB * P.nblk_dots * LB_PART, st, PDL_OPT, lbfgs_reduce_kernel, dim3((LB_PART + 7) / 8, P.NB), 256, 0, P);
        IST_EWK("lbfgs_solve", 16.0 * P.m * P.m, st, PDL_OPT, lbfgs_solve_kernel, P.NB, LB_SOLVE_THREADS, lb_solve_smem(P.m), P);
        IST_EWK("lbfgs_update", vb * (9 + 2.0 * P.m), st, PDL_OPT, lbfgs_update_kernel, dim3(P.nblk, P.NB), 256, 0, P);
        if (O->trace.cap > 0) {
            IST_EW("lbfgs_trace", 0.0, st, lbfgs_trace_post_kernel<<<dim3(tgrid, P.NB), 256, 0, st>>>(P, O->trace));
            IST_EW("lbfgs_trace", 0.0, st, lbfgs_trace_bump_kernel<<<1, 1, 0, st>>>(O->trace));
        }
    }
    return IST_OK;
}

}  // namespace ist

extern "C" {

static int lbfgs_create_common(ist_lbfgs** out, ist_plan* plan, int batch, int n, int n_losses, int history_size, int max_iter,
                               int max_eval, float lr, double tolerance_grad, double tolerance_change) {
    using namespace ist;
    if (history_size < 1 || history_size > LB_MAXH - 1) return fail(IST_ERR_ARG, "history_size must be in [1, %d]", LB_MAXH - 1);
    if (max_iter < 1) return fail(IST_ERR_ARG, "max_iter must be >= 1");
    DeviceGuard dg(plan != nullptr ? plan_device(plan) : -1);
    ist_lbfgs* O = new ist_lbfgs();
    O->plan = plan;
    O->device = current_device();
    LbParams& P = O->P;
    memset(&P, 0, sizeof(P));
    P.NB = batch;
    P.n = n;
    P.m = history_size;
    const int ntiles = (P.n + LB_WT - 1) / LB_WT;
    const int utiles = (P.n + LB_UT - 1) / LB_UT;
    int nblk = (utiles + 7) / 8;
    if (nblk > 4 * num_sms()) nblk = 4 * num_sms();
    if (nblk < 1) nblk = 1;
    P.nblk = nblk;
    int nblk_dots = (ntiles + LB_DOTS_WARPS - 1) / LB_DOTS_WARPS;        // one CTA per SM (its ring takes the shared memory)
    if (nblk_dots > num_sms()) nblk_dots = num_sms();
    if (nblk_dots < 1) nblk_dots = 1;
    P.nblk_dots = nblk_dots;
    P.max_iter = max_iter;
    P.max_eval = max_eval > 0 ? max_eval : max_iter * 5 / 4;
    P.lr = lr; P.tol_grad = tolerance_grad; P.tol_change = tolerance_change;
    O->n_losses = n_losses;
    P.loss_stride = O->n_losses + 1;
    P.loss_total = O->n_losses;
    const size_t vn = (size_t)P.NB * P.n;
    int rc = IST_OK;
    if (rc == IST_OK) rc = O->mem.alloc(&O->g, vn);
    if (rc == IST_OK) rc = O->mem.alloc(&O->x_own, vn);
    if (rc == IST_OK) rc = O->mem.alloc(&O->losses, (size_t)P.NB * P.loss_stride);
    if (rc == IST_OK) rc = O->mem.alloc(&P.prev_g, vn);
    if (rc == IST_OK) rc = O->mem.alloc(&P.d, vn);
    if (rc == IST_OK) rc = O->mem.alloc(&P.S, vn * P.m);
    if (rc == IST_OK) rc = O->mem.alloc(&P.Y, vn * P.m);
    if (rc == IST_OK) rc = O->mem.alloc(&P.part, (size_t)P.NB * P.nblk_dots * LB_PART);
    if (rc == IST_OK) rc = O->mem.alloc(&P.tot, (size_t)P.NB * LB_PART);
    if (rc == IST_OK) rc = O->mem.alloc(&P.dmax_part, (size_t)P.NB * P.nblk);
    if (rc == IST_OK) rc = O->mem.alloc(&P.SY, (size_t)P.NB * LB_MAXH * LB_LDG);
    if (rc == IST_OK) rc = O->mem.alloc(&P.YY, (size_t)P.NB * LB_MAXH * LB_LDG);
    if (rc == IST_OK) rc = O->mem.alloc(&P.frames, (size_t)P.NB);
    if (rc == IST_OK) rc = O->mem.alloc(&O->trace.count, (size_t)1);
    if (rc != IST_OK) { delete O; return rc; }
    P.g = O->g;
    P.losses = O->losses;
    cudaMemset(P.frames, 0, sizeof(LbFrame) * P.NB);
    cudaMemset(P.dmax_part, 0, sizeof(float) * P.NB * P.nblk);
    cudaMemset(P.SY, 0, sizeof(double) * P.NB * LB_MAXH * LB_LDG);
    cudaMemset(P.YY, 0, sizeof(double) * P.NB * LB_MAXH * LB_LDG);
    cudaMemset(P.d, 0, sizeof(float) * vn);
    cudaMemset(P.prev_g, 0, sizeof(float) * vn);
    cudaMemset(O->trace.count, 0, sizeof(int));
    cudaError_t e = cudaFuncSetAttribute(lbfgs_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)lb_solve_smem(history_size));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(lbfgs_dots_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LB_DOTS_SMEM);
    if (e != cudaSuccess) { delete O; return fail(IST_ERR_CUDA, "cudaFuncSetAttribute(lbfgs_solve): %s", cudaGetErrorString(e)); }
    const char* ng = getenv("IST_B200_NO_GRAPH");
    O->use_graph = !(ng != nullptr && atoi(ng) == 1);
    O->host_frames.resize(P.NB);
    *out = O;
    return IST_OK;
}

int ist_lbfgs_create(ist_lbfgs** out, ist_plan* plan, int history_size, int max_iter, int max_eval, float lr,
                     double tolerance_grad, double tolerance_change) {
    using namespace ist;
    if (out == nullptr || plan == nullptr) return fail(IST_ERR_ARG, "ist_lbfgs_create: null argument");
    if (plan_n_losses(plan) < 1) return fail(IST_ERR_STATE, "configure the plan's losses before creating the optimiser");
    return lbfgs_create_common(out, plan, plan_batch(plan), plan_image_elems(plan), plan_n_losses(plan), history_size, max_iter,
                               max_eval, lr, tolerance_grad, tolerance_change);
}

int ist_lbfgs_create_test(ist_lbfgs** out, int batch, int n, const float* a_dev, const float* b_dev, const float* c_dev,
                          int history_size, int max_iter, int max_eval, float lr, double tolerance_grad, double tolerance_change) {
    using namespace ist;
    if (out == nullptr || a_dev == nullptr || b_dev == nullptr || c_dev == nullptr || batch < 1 || n < 1)
        return fail(IST_ERR_ARG, "ist_lbfgs_create_test: bad argument");
    IST_TRY(ist_device_check());
    IST_TRY(lbfgs_create_common(out, nullptr, batch, n, 0, history_size, max_iter, max_eval, lr, tolerance_grad, tolerance_change));
    (*out)->test_obj.a = a_dev; (*out)->test_obj.b = b_dev; (*out)->test_obj.c = c_dev;
    return IST_OK;
}

int ist_lbfgs_set_trace(ist_lbfgs* O, float* x_dev, float* g_dev, float* d_dev, double* scalars_dev, int capacity) {
    using namespace ist;
    if (O == nullptr || x_dev == nullptr || g_dev == nullptr || d_dev == nullptr || scalars_dev == nullptr || capacity < 1)
        return fail(IST_ERR_ARG, "ist_lbfgs_set_trace: bad argument");
    DeviceGuard dg(O->device);
    if (O->graph_exec != nullptr) return fail(IST_ERR_STATE, "ist_lbfgs_set_trace must be called before the first step (the step graph is already captured)");
    O->trace.x = x_dev; O->trace.g = g_dev; O->trace.d = d_dev; O->trace.sc = scalars_dev; O->trace.cap = capacity;
    IST_CUDA(cudaMemset(O->trace.count, 0, sizeof(int)));
    return IST_OK;
}

int ist_lbfgs_trace_count(ist_lbfgs* O, int* count_host) {
    using namespace ist;
    if (O == nullptr || count_host == nullptr) return fail(IST_ERR_ARG, "ist_lbfgs_trace_count: null argument");
    DeviceGuard dg(O->device);
    IST_CUDA(cudaMemcpy(count_host, O->trace.count, sizeof(int), cudaMemcpyDeviceToHost));
    return IST_OK;
}

int ist_lbfgs_frame_state(ist_lbfgs* O, int frame, int* func_evals, int* n_iter, int* hist_len, int* active, int* step_evals) {
    using namespace ist;
    if (O == nullptr || frame < 0 || frame >= O->P.NB) return fail(IST_ERR_ARG, "ist_lbfgs_frame_state: bad argument");
    const LbFrame& F = O->host_frames[frame];          // copied back by the last ist_lbfgs_step
    if (func_evals != nullptr) *func_evals = F.func_evals;
    if (n_iter != nullptr) *n_iter = F.n_iter;
    if (hist_len != nullptr) *hist_len = F.hist_len;
    if (active != nullptr) *active = F.active;
    if (step_evals != nullptr) *step_evals = F.current_evals;
    return IST_OK;
}

int ist_lbfgs_destroy(ist_lbfgs* O) {
    if (O == nullptr) return IST_OK;
    ist::DeviceGuard dg(O->device);
    if (O->graph_exec != nullptr) cudaGraphExecDestroy(O->graph_exec);
    if (O->cap_stream != nullptr) cudaStreamDestroy(O->cap_stream);
    delete O;
    return IST_OK;
}

int ist_lbfgs_reset(ist_lbfgs* O, void* stream) {
    using namespace ist;
    if (O == nullptr) return fail(IST_ERR_ARG, "ist_lbfgs_reset: null argument");
    DeviceGuard dg(O->device);
    cudaStream_t st = (cudaStream_t)stream;
    const LbParams& P = O->P;
    // a fresh torch.optim.LBFGS([x]) (utils.py:24): empty state; the history buffers need no clearing (hist_len = 0)
    IST_CUDA(cudaMemsetAsync(P.frames, 0, sizeof(LbFrame) * P.NB, st));
    IST_CUDA(cudaMemsetAsync(P.dmax_part, 0, sizeof(float) * P.NB * P.nblk, st));
    IST_CUDA(cudaMemsetAsync(O->trace.count, 0, sizeof(int), st));
    return IST_OK;
}

int ist_lbfgs_step(ist_lbfgs* O, float* x_dev, int* evals_out, float* loss_out, void* stream) {
    using namespace ist;
    if (O == nullptr || x_dev == nullptr) return fail(IST_ERR_ARG, "ist_lbfgs_step: null argument");
    DeviceGuard dg(O->device);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t xbytes = sizeof(float) * (size_t)O->P.NB * O->P.n;
    if (O->use_graph) {
        if (O->graph_exec == nullptr) {
            if (O->cap_stream == nullptr) IST_CUDA(cudaStreamCreateWithFlags(&O->cap_stream, cudaStreamNonBlocking));
            IST_CUDA(cudaStreamSynchronize(st));
            IST_CUDA(cudaStreamBeginCapture(O->cap_stream, cudaStreamCaptureModeThreadLocal));
            book().capturing = true;
            book().captured = 0;
            int rc = lbfgs_enqueue_step(O, O->x_own, O->cap_stream);
            book().capturing = false;
            O->graph_kernels = book().captured;
            cudaGraph_t graph = nullptr;
            cudaError_t e = cudaStreamEndCapture(O->cap_stream, &graph);
            if (rc != IST_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (e != cudaSuccess) return fail(IST_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
            e = cudaGraphInstantiate(&O->graph_exec, graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess) return fail(IST_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
        }
        IST_CUDA(cudaMemcpyAsync(O->x_own, x_dev, xbytes, cudaMemcpyDeviceToDevice, st));
        IST_CUDA(cudaGraphLaunch(O->graph_exec, st));
        IST_CUDA(cudaMemcpyAsync(x_dev, O->x_own, xbytes, cudaMemcpyDeviceToDevice, st));
        book().launches += O->graph_kernels;
    } else {
        IST_TRY(lbfgs_enqueue_step(O, x_dev, st));
    }
    IST_CUDA(cudaMemcpyAsync(O->host_frames.data(), O->P.frames, sizeof(LbFrame) * O->P.NB, cudaMemcpyDeviceToHost, st));
    IST_CUDA(cudaStreamSynchronize(st));
    int evals = 0;
    for (const LbFrame& F : O->host_frames) evals = F.current_evals > evals ? F.current_evals : evals;
    if (evals_out != nullptr) *evals_out += evals;
    if (loss_out != nullptr) *loss_out = (float)O->host_frames[0].orig_loss;
    return IST_OK;
}

int ist_lbfgs_last_losses(ist_lbfgs* O, float* losses_host) {
    using namespace ist;
    if (O == nullptr || losses_host == nullptr) return fail(IST_ERR_ARG, "ist_lbfgs_last_losses: null argument");
    DeviceGuard dg(O->device);
    IST_CUDA(cudaMemcpy(losses_host, O->losses, sizeof(float) * O->P.NB * O->P.loss_stride, cudaMemcpyDeviceToHost));
    return IST_OK;
}

}  // extern "C"

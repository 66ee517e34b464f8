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
constexpr int LB_WT = 512;            // elements per warp-tile (16 per lane)
constexpr int LB_NSTAT = 8;           // ys, yy, s.g, y.g, g.g, |g|_1, max|g|, (unused)
constexpr int LB_PART = LB_MAXH * 5 + LB_NSTAT;

struct LbFrame {
    int n_iter, func_evals, hist_len, head;
    int active, current_evals, apply, accepted, new_slot, nread;
    double H_diag, t, prev_loss, orig_loss, loss, gtd;
    double ro[LB_MAXH];
    // outputs of the solve for the update pass
    float cg, cy_new, cs_new, t_f, t_prev_f;
    int read_slot[LB_MAXH];
    float read_cy[LB_MAXH], read_cs[LB_MAXH];
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
    double* SY;                   // [NB][LB_MAXH][LB_MAXH]  s_i.y_j by physical slot
    double* YY;                   // [NB][LB_MAXH][LB_MAXH]
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
            a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3); a4 = warp_sum(a4); a5 = warp_sum(a5);
            if (lane == 0) {
                double* st = &wacc[warp][5 * LB_MAXH];
                st[0] += a0; st[1] += a1; st[2] += a2; st[3] += a3; st[4] += a4; st[5] += a5;
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
                c0 = warp_sum(c0); c1 = warp_sum(c1); c2 = warp_sum(c2); c3 = warp_sum(c3); c4 = warp_sum(c4);
                if (lane == 0) {
                    double* a = &wacc[warp][slot * 5];
                    a[0] += c0; a[1] += c1; a[2] += c2; a[3] += c3; a[4] += c4;
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
inline __host__ __device__ int lb_ld(int m) { return m + 1; }   // padded leading dimension of the smem matrices
inline size_t lb_solve_smem(int m) { return 2 * (size_t)m * lb_ld(m) * sizeof(double); }

__global__ void __launch_bounds__(LB_SOLVE_THREADS)
lbfgs_solve_kernel(const LbParams P) {
    extern __shared__ double sm[];
    const int LB_LD = lb_ld(P.m);
    double* SYs = sm;                          // [m][LB_LD] logical order: SYs[i][j] = s_i . y_j
    double* YYs = sm + (size_t)P.m * LB_LD;    // [m][LB_LD]
    __shared__ double tot[LB_PART];
    __shared__ double al_s[LB_MAXH], c_s[LB_MAXH], sg_s[LB_MAXH], yg_s[LB_MAXH];
    __shared__ double red[LB_MAXH];
    __shared__ int sh_go, sh_h;
    const int b = blockIdx.x, tid = threadIdx.x, m = P.m;
    pdl_trigger();
    pdl_wait();
    LbFrame& F = P.frames[b];
    double* SYg = P.SY + (size_t)b * LB_MAXH * LB_MAXH;
    double* YYg = P.YY + (size_t)b * LB_MAXH * LB_MAXH;

    for (int i = tid; i < LB_PART; i += blockDim.x) tot[i] = P.tot[(size_t)b * LB_PART + i];
    __syncthreads();
    const double* st = &tot[5 * LB_MAXH];
    const double ys = st[0], yy = st[1], sg_new = st[2], yg_new = st[3], gg = st[4], g1 = st[5], gmax = st[6];

    if (tid == 0) {
        const double loss = (double)P.losses[(size_t)b * P.loss_stride + P.loss_total];
        int go = 1;
        if (P.it == 1) {                        // lbfgs.py:364-374: first closure of step()
            F.active = 1;
            F.current_evals = 1;
            F.func_evals += 1;
            F.orig_loss = loss;
            F.loss = loss;
            if (gmax <= P.tol_grad) { F.active = 0; go = 0; }
        } else if (F.active) {                  // lbfgs.py:493-523: the checks that follow the re-evaluation
            F.current_evals += 1;
            F.func_evals += 1;
            F.loss = loss;
            float dm = 0.f;
            for (int k = 0; k < P.nblk; ++k) dm = fmaxf(dm, P.dmax_part[(size_t)b * P.nblk + k]);
            if (F.current_evals >= P.max_eval) go = 0;
            else if (gmax <= P.tol_grad) go = 0;
            else if ((double)dm * fabs(F.t) <= P.tol_change) go = 0;
            else if (fabs(loss - F.prev_loss) < P.tol_change) go = 0;
            if (!go) F.active = 0;
        } else {
            go = 0;
        }
        F.apply = 0;
        F.accepted = 0;
        F.nread = 0;
        if (go) {
            F.n_iter += 1;
            if (F.n_iter == 1) {                // lbfgs.py:396-401
                F.hist_len = 0; F.head = 0; F.H_diag = 1.0;
            } else if (ys > 1e-10) {            // lbfgs.py:404-420
                int slot;
                if (F.hist_len == m) { slot = F.head; F.head = (F.head + 1) % m; }
                else { slot = (F.head + F.hist_len) % m; F.hist_len += 1; }
                F.accepted = 1;
                F.new_slot = slot;
                F.ro[slot] = 1.0 / ys;
                F.H_diag = ys / yy;
            }
        }
        sh_go = go;
        sh_h = F.hist_len;
    }
    __syncthreads();
    if (!sh_go) return;
    const int h = sh_h, head = F.head, accepted = F.accepted, new_slot = F.new_slot;
    auto phys = [&](int i) { int s = head + i; return s >= m ? s - m : s; };

    // fold the new pair into the physical-slot matrices (row/column new_slot)
    if (accepted) {
        for (int i = tid; i < h; i += blockDim.x) {
            const int p = phys(i);
            if (p == new_slot) continue;
            SYg[(size_t)p * LB_MAXH + new_slot] = tot[p * 5 + 0];      // s_i . y_new
            SYg[(size_t)new_slot * LB_MAXH + p] = tot[p * 5 + 1];      // s_new . y_i
            YYg[(size_t)p * LB_MAXH + new_slot] = tot[p * 5 + 2];
            YYg[(size_t)new_slot * LB_MAXH + p] = tot[p * 5 + 2];
        }
        if (tid == 0) {
            SYg[(size_t)new_slot * LB_MAXH + new_slot] = ys;
            YYg[(size_t)new_slot * LB_MAXH + new_slot] = yy;
        }
    }
    __syncthreads();
    // logical-order copies in shared memory and the g-dots
    for (int e = tid; e < h * h; e += blockDim.x) {
        const int i = e / h, j = e % h;
        SYs[i * LB_LD + j] = SYg[(size_t)phys(i) * LB_MAXH + phys(j)];
        YYs[i * LB_LD + j] = YYg[(size_t)phys(i) * LB_MAXH + phys(j)];
    }
    for (int i = tid; i < h; i += blockDim.x) {
        const int p = phys(i);
        const bool is_new = accepted && p == new_slot;
        sg_s[i] = is_new ? sg_new : tot[p * 5 + 3];
        yg_s[i] = is_new ? yg_new : tot[p * 5 + 4];
    }
    __syncthreads();

    // The two sequential loops run on the first 4 warps only (h <= 100 < 128), synchronised with a named barrier: a barrier
    // over 4 warps is several times cheaper than one over the 16 warps that loaded the matrices.
    const double H = F.H_diag;
    if (tid < 128) {
        // first loop (lbfgs.py:431-436), newest to oldest: al_i = ro_i * s_i.q with q = -g - sum_{j>i} al_j y_j
        const double ro_t = tid < h ? F.ro[phys(tid)] : 0.0;
        double u = 0.0;                              // thread k: sum_{j processed} al_j * (s_k . y_j)
        for (int i = h - 1; i >= 0; --i) {
            if (tid == i) al_s[i] = ro_t * (-sg_s[i] - u);
            named_bar_sync(1, 128);
            if (tid < i) u += al_s[i] * SYs[tid * LB_LD + i];
        }
        named_bar_sync(1, 128);
        // r = H*q: y_k . r before the second loop
        double v = 0.0, w = 0.0;
        if (tid < h) {
            double acc = -yg_s[tid];
            for (int j = 0; j < h; ++j) acc -= al_s[j] * YYs[tid * LB_LD + j];
            v = H * acc;
        }
        // second loop (lbfgs.py:440-443), oldest to newest: be_i = ro_i * y_i.r ; r += (al_i - be_i) s_i
        for (int i = 0; i < h; ++i) {
            if (tid == i) c_s[i] = al_s[i] - ro_t * (v + w);
            named_bar_sync(1, 128);
            if (tid > i && tid < h) w += c_s[i] * SYs[i * LB_LD + tid];
        }
    }
    __syncthreads();
    // gtd = g.d with d = -H g - sum H al_j y_j + sum c_j s_j
    if (tid < LB_MAXH) red[tid] = (tid < h) ? (-H * al_s[tid] * yg_s[tid] + c_s[tid] * sg_s[tid]) : 0.0;
    // coefficients of the update pass: an accepted pair is always the newest one, i.e. logical index h - 1
    const int nread = accepted ? h - 1 : h;
    if (tid < nread) {
        F.read_slot[tid] = phys(tid);
        F.read_cy[tid] = (float)(-H * al_s[tid]);
        F.read_cs[tid] = (float)c_s[tid];
    }
    __syncthreads();
    if (tid == 0) {
        double gtd = -H * gg;
        for (int k = 0; k < h; ++k) gtd += red[k];
        F.gtd = gtd;
        F.prev_loss = F.loss;                                            // lbfgs.py:449
        double t;
        if (F.n_iter == 1) {                                             // lbfgs.py:454-457
            const float inv = 1.0f / (float)g1;
            t = (double)(inv < 1.0f ? inv : 1.0f) * P.lr;
        } else {
            t = P.lr;
        }
        F.t = t;
        F.t_f = (float)t;
        F.cg = (float)(-H);
        F.cy_new = accepted ? (float)(-H * al_s[h - 1]) : 0.f;
        F.cs_new = accepted ? (float)c_s[h - 1] : 0.f;
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
__global__ void __launch_bounds__(256)
lbfgs_update_kernel(const LbParams P) {
    __shared__ float wmax[8];
    __shared__ int s_slot[LB_MAXH];
    __shared__ float s_cy[LB_MAXH], s_cs[LB_MAXH];
    const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const LbFrame& F = P.frames[b];
    pdl_trigger();
    pdl_wait();
    for (int i = threadIdx.x; i < F.nread; i += blockDim.x) { s_slot[i] = F.read_slot[i]; s_cy[i] = F.read_cy[i]; s_cs[i] = F.read_cs[i]; }
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
        const int ntiles = (n + LB_WT - 1) / LB_WT;
        const float cg = F.cg, cyn = F.cy_new, csn = F.cs_new;
        const int nread = F.nread, accepted = F.accepted, new_slot = F.new_slot, apply = F.apply;
        for (int tile = blockIdx.x * 8 + warp; tile < ntiles; tile += gridDim.x * 8) {
            const size_t base = (size_t)tile * LB_WT;
            const int n_left = n - (int)base;
            float gv[16], acc[16];
            lb_load16(P.g + fo, base, n_left, lane, vec, gv);
#pragma unroll
            for (int e = 0; e < 16; ++e) acc[e] = cg * gv[e];
            if (has_prev) {
                float yv[16], sv[16];
                lb_load16(P.prev_g + fo, base, n_left, lane, vec, yv);
                lb_load16(P.d + fo, base, n_left, lane, vec, sv);
#pragma unroll
                for (int e = 0; e < 16; ++e) { yv[e] = gv[e] - yv[e]; }
                if (accepted) {
                    // s_new = t_prev * d_old (lbfgs.py:404): t_prev is the step size saved by lbfgs_reduce_kernel before the
                    // solve replaced F.t with the step size of the new direction
                    const float tp = F.t_prev_f;
#pragma unroll
                    for (int e = 0; e < 16; ++e) { sv[e] = tp * sv[e]; acc[e] = fmaf(cyn, yv[e], fmaf(csn, sv[e], acc[e])); }
                    const size_t ho = ((size_t)new_slot * P.NB + b) * (size_t)n;
                    lb_store16(P.S + ho, base, n_left, lane, vec, sv);
                    lb_store16(P.Y + ho, base, n_left, lane, vec, yv);
                }
            }
            for (int i = 0; i < nread; ++i) {
                const size_t ho = ((size_t)s_slot[i] * P.NB + b) * (size_t)n;
                float s_i[16], y_i[16];
                lb_load16(P.S + ho, base, n_left, lane, vec, s_i);
                lb_load16(P.Y + ho, base, n_left, lane, vec, y_i);
                const float cy = s_cy[i], cs = s_cs[i];
#pragma unroll
                for (int e = 0; e < 16; ++e) acc[e] = fmaf(cy, y_i[e], fmaf(cs, s_i[e], acc[e]));
            }
            lb_store16(P.d + fo, base, n_left, lane, vec, acc);
            lb_store16(P.prev_g + fo, base, n_left, lane, vec, gv);
#pragma unroll
            for (int e = 0; e < 16; ++e) dmax = fmaxf(dmax, fabsf(acc[e]));
            if (apply) {
                float xv[16];
                lb_load16(P.x + fo, base, n_left, lane, vec, xv);
#pragma unroll
                for (int e = 0; e < 16; ++e) xv[e] = fmaf(t_old_new, acc[e], xv[e]);
                lb_store16(P.x + fo, base, n_left, lane, vec, xv);
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
    int nblk = (ntiles + 7) / 8;
    if (nblk > 2 * num_sms()) nblk = 2 * num_sms();
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
    if (rc == IST_OK) rc = O->mem.alloc(&P.SY, (size_t)P.NB * LB_MAXH * LB_MAXH);
    if (rc == IST_OK) rc = O->mem.alloc(&P.YY, (size_t)P.NB * LB_MAXH * LB_MAXH);
    if (rc == IST_OK) rc = O->mem.alloc(&P.frames, (size_t)P.NB);
    if (rc == IST_OK) rc = O->mem.alloc(&O->trace.count, (size_t)1);
    if (rc != IST_OK) { delete O; return rc; }
    P.g = O->g;
    P.losses = O->losses;
    cudaMemset(P.frames, 0, sizeof(LbFrame) * P.NB);
    cudaMemset(P.dmax_part, 0, sizeof(float) * P.NB * P.nblk);
    cudaMemset(P.SY, 0, sizeof(double) * P.NB * LB_MAXH * LB_MAXH);
    cudaMemset(P.YY, 0, sizeof(double) * P.NB * LB_MAXH * LB_MAXH);
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

// conv3x3 implicit GEMM, second generation: one *halo* activation tile per 64-channel chunk feeds all nine taps.
//
// conv_igemm.cuh fetches a 128-pixel x 64-channel A tile per (tap, channel chunk): 9 fetches of the same pixels shifted
// by one. Measured on B200 that makes every VGG layer L2->SM bandwidth bound (~8.5 TB/s aggregate: conv4_2 moves 604 MB
// for 19.3 GF, 68 us, while its tensor time is ~30 us; profiles/r01_ncu_launches_closure512_v2.csv). Here the producer
// loads ONE (TH+2) x (TW+2) pixel halo box per channel chunk (TMA zero-fills the image border = the conv padding) and the
// nine taps address it through shifted shared-memory matrix descriptors:
//   tile TH x TW = 16 x 8, halo 18 x 10 pixels, one pixel = one 128-byte row (64 channels x 2 B), row r = hy*10 + hx.
//   MMA row m = py*8 + px of tap (ky,kx) is halo row (py+ky)*10 + (px+kx) = [ky*10+kx] + py*10 + px, i.e. exactly the
//   K-major SW128 form "8 consecutive 128-byte rows per group, groups SBO apart" with start = base + (ky*10+kx)*128 and
//   SBO = 10*128 B. TMA wrote the box with the 128-byte swizzle keyed on absolute smem address bits [7,10); the
//   descriptor's base_offset field carries (start >> 7) & 7 so the tensor core applies the same phase although the start
//   is no longer 1024-byte aligned.
// A traffic drops 9 x 32 KB -> 46 KB per chunk (6.3x); with the weight tiles unchanged the kernel moves 1.7x (N=128)
// to 2.3x (N=64) fewer bytes per FLOP. 1x1 contractions (Gram backward) use the same kernel with an exact 16x8 box.
// Accumulation (short hi*hi chains promoted to fp32 registers, cross terms in their own accumulator), warp roles and
// epilogue are those of conv_igemm.cuh.
#pragma once
#include "conv_igemm.cuh"

namespace ist {

// PAIR: the kernel runs as clusters of two CTAs (tcgen05 cta_group::2). The pair works on two horizontally adjacent pixel
// tiles (M = 256: rank r owns tile column 2 * pair_column + r) and the same N_TILE output channels; every CTA loads its own
// halo box and HALF of the weight rows (N_TILE / 2), the leader (rank 0) issues all MMAs, every CTA drains / stores its own
// 128 accumulator rows. See ptx.cuh ("CTA pair") for the measured reason.
template <int N_TILE, bool PAIR = false>
struct HaloCfg {
    static constexpr int TW = 8, TH = 16, PW = TW + 2, PH = TH + 2;
    static constexpr int HALO_BYTES = PW * PH * 128;               // 23040
    static constexpr int EXACT_BYTES = TW * TH * 128;              // 16384
    static constexpr int A_PLANE = 23 * 1024;                      // 23552 >= HALO_BYTES, keeps every plane 1024-aligned
    static constexpr int A_STAGE = 2 * A_PLANE;
    static constexpr int A_STAGES = 2;
    static constexpr int B_ROWS = PAIR ? N_TILE / 2 : N_TILE;      // weight rows this CTA holds per tap
    static constexpr int B_PLANE = B_ROWS * 128;
    static constexpr int B_STAGE = 2 * B_PLANE;
#ifndef IST_B_STAGES_N64_PAIR
#define IST_B_STAGES_N64_PAIR 12
#endif
    // pair / N = 64: 12 x 8 KB weight stages (a tile of the 64 -> 64 layers is 9 taps: the ring then spans a tile boundary, and
    // the issuers' largest wait, B-full at 16 % of the launch with 8 stages, shrinks); the shared memory is there (222 KB in all)
    static constexpr int B_STAGES = PAIR ? ((N_TILE == 128) ? 4 : IST_B_STAGES_N64_PAIR) : ((N_TILE == 128) ? 3 : 4);
    static_assert(B_STAGES <= 12, "barrier block holds 12 weight-stage barrier pairs");
    // tensor-memory accumulators (N_TILE fp32 columns each, 512 columns in all):
    //   N_TILE = 128: [main 0][main 1][main 2 | Gram][cross]      (third main buffer when no Gram k-steps are fused)
    //   N_TILE =  64: [main 0..3][cross 0][cross 1][Gram 0][Gram 1]   (the Gram accumulator follows the cross buffer's parity)
    // The cross buffer is read once per tile, so one is enough when the tiles are long; the main ring depth is what hides
    // the promotion round trip (commit -> drain in both CTAs of a pair -> arrival at the leader) behind the next chains.
    static constexpr int TMEM_COLS = 512;
    static constexpr int CROSS_COL = (N_TILE == 128) ? 3 * N_TILE : 4 * N_TILE;
    static constexpr int GRAM_COL = (N_TILE == 128) ? 2 * N_TILE : 6 * N_TILE;
    static constexpr bool XSINGLE = (N_TILE == 128);
    static constexpr int OUT_PLANE = 128 * 128;                    // output staging: 128 pixels x 64 channels x 2 B per plane
    static constexpr int OUT_BYTES = 2 * OUT_PLANE;                // hi + lo
    // a pair's CTA holds half the weight rows, which leaves room for a second staging buffer: the two 64-channel halves of a
    // 128-channel tile are then converted back to back instead of waiting for the first TMA store to drain the buffer
    static constexpr int OUT_BUFS = (PAIR && N_TILE == 128) ? 2 : 1;
    // Epilogue warps: one group of four (one warp per tensor-memory lane quadrant) holds all N_TILE accumulator columns of
    // its pixel; a pair's CTAs run TWO groups (warps 2-5 and 7-10), each holding half of the columns. With one warp per
    // scheduler the conversion code is issue-latency bound (~8k cycles per 128-channel tile); two warps per scheduler overlap,
    // halve the per-thread work and drop the register pressure from 255 (spills) to < 184.
    static constexpr int EPI_GROUPS = PAIR ? 2 : 1;
    static constexpr int NH = N_TILE / EPI_GROUPS;                 // accumulator columns per epilogue thread
    static constexpr int THREADS = 96 + 128 * EPI_GROUPS;          // producer, two issuers (warps 1 and 6), epilogue groups
    static constexpr int SMEM_BYTES = A_STAGES * A_STAGE + B_STAGES * B_STAGE + OUT_BUFS * OUT_BYTES + 512 + 1024;
};

__device__ __forceinline__ uint64_t umma_desc_sw128_bo(uint32_t smem_addr, uint32_t sbo, uint32_t use_base_offset) {
    uint64_t d = umma_smem_desc_sw128(smem_addr, 0, sbo);
    if (use_base_offset) d |= static_cast<uint64_t>((smem_addr >> 7) & 7u) << 49;     // [49,52) matrix base offset
    return d;
}

// Forward epilogue on registers for 32 consecutive channels starting at channel cb: relu(v + bias) * out_scale split into
// fp16 hi / lo packs (same arithmetic as the CONV_FWD branch of conv_epilogue_32).
__device__ __forceinline__ void conv_epilogue_regs_32(const ConvParams& p, const float (&v)[32], int cb, uint32_t (&hi)[16],
                                                      uint32_t (&lo)[16]) {
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
        const float a = fmaxf(v[j] + __ldg(p.bias + cb + j), 0.f) * p.out_scale;
        const float b = fmaxf(v[j + 1] + __ldg(p.bias + cb + j + 1), 0.f) * p.out_scale;
        const uint32_t h = pack_h2(a, b);
        hi[j >> 1] = h;
        lo[j >> 1] = pack_h2(a - h_lo_f(h), b - h_hi_f(h));
    }
}

// Data-gradient epilogue on registers for 32 consecutive channels of one pixel (element offset `o`): adds the optional
// fp32 addend and content term, applies the ReLU mask, and splits into bf16 hi / lo packs (same arithmetic and order as
// the CONV_GRAD branch of conv_epilogue_32). `valid` = the pixel lies inside the image (no loads otherwise).
__device__ __forceinline__ void conv_grad_regs_32(const ConvParams& p, float (&v)[32], size_t o, bool valid, uint32_t (&hi)[16],
                                                  uint32_t (&lo)[16]) {
    if (valid) {
        if (p.addend != nullptr) {
            const float4* ad = reinterpret_cast<const float4*>(p.addend + o);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 t = __ldg(ad + q);
                v[4 * q] += t.x; v[4 * q + 1] += t.y; v[4 * q + 2] += t.z; v[4 * q + 3] += t.w;
            }
        }
        if (p.f_hi != nullptr) {
            const uint4* fh = reinterpret_cast<const uint4*>(p.f_hi + o);
            const uint4* fl = reinterpret_cast<const uint4*>(p.f_lo + o);
            const uint4* th = reinterpret_cast<const uint4*>(p.t_hi + o);
            const uint4* tl = reinterpret_cast<const uint4*>(p.t_lo + o);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint4 a = __ldg(fh + q), b = __ldg(fl + q), c = __ldg(th + q), d = __ldg(tl + q);
                const uint32_t ua[4] = {a.x, a.y, a.z, a.w}, ub[4] = {b.x, b.y, b.z, b.w};
                const uint32_t uc[4] = {c.x, c.y, c.z, c.w}, ud[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float d0 = (h_lo_f(ua[e]) - h_lo_f(uc[e])) + (h_lo_f(ub[e]) - h_lo_f(ud[e]));
                    const float d1 = (h_hi_f(ua[e]) - h_hi_f(uc[e])) + (h_hi_f(ub[e]) - h_hi_f(ud[e]));
                    v[8 * q + 2 * e] += p.content_coef * d0;
                    v[8 * q + 2 * e + 1] += p.content_coef * d1;
                }
            }
        }
        if (p.mask_hi != nullptr) {
            const uint4* mk = reinterpret_cast<const uint4*>(p.mask_hi + o);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint4 a = __ldg(mk + q);
                const uint32_t ua[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (!(h_lo_f(ua[e]) > 0.f)) v[8 * q + 2 * e] = 0.f;
                    if (!(h_hi_f(ua[e]) > 0.f)) v[8 * q + 2 * e + 1] = 0.f;
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
        const uint32_t h = pack_bf2(v[j], v[j + 1]);
        hi[j >> 1] = h;
        lo[j >> 1] = pack_bf2(v[j] - bf_lo_f(h), v[j + 1] - bf_hi_f(h));
    }
}

template <int N_TILE, bool PAIR>
__global__ void __launch_bounds__((HaloCfg<N_TILE, PAIR>::THREADS), 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                 const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                 const __grid_constant__ CUtensorMap tmO_hi, const __grid_constant__ CUtensorMap tmO_lo,
                 const __grid_constant__ CUtensorMap tmF_hi, const __grid_constant__ CUtensorMap tmF_lo,
                 const __grid_constant__ CUtensorMap tmD_hi, const __grid_constant__ CUtensorMap tmD_lo,
                 const ConvParams p) {
    using Cfg = HaloCfg<N_TILE, PAIR>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_base = smem_base;
    const uint32_t b_base = smem_base + Cfg::A_STAGES * Cfg::A_STAGE;
    const uint32_t o_base = b_base + Cfg::B_STAGES * Cfg::B_STAGE;      // output staging (1024-aligned planes)
    const uint32_t bar_base = o_base + Cfg::OUT_BUFS * Cfg::OUT_BYTES;
    // barriers (8 B each): a_full[3] @0, a_empty[3] @24, b_full[12] @48, b_empty[12] @144, main_full[4] @240,
    // main_empty[4] @272, cross_full[2] @304, cross_empty[2] @320, tmem base address @336
    auto afull = [&](int s) { return bar_base + 8u * s; };
    auto aempty = [&](int s) { return bar_base + 24u + 8u * s; };
    auto bfull = [&](int s) { return bar_base + 48u + 8u * s; };
    auto bempty = [&](int s) { return bar_base + 144u + 8u * s; };
    auto mfull = [&](uint32_t b) { return bar_base + 240u + 8u * b; };
    auto mempty = [&](uint32_t b) { return bar_base + 272u + 8u * b; };
    auto xfull = [&](uint32_t a) { return bar_base + 304u + 8u * a; };
    auto xempty = [&](uint32_t a) { return bar_base + 320u + 8u * a; };
    const uint32_t tmem_slot = bar_base + 336u;
    volatile uint32_t* tmem_slot_gen =
        reinterpret_cast<volatile uint32_t*>(smem_raw + (bar_base - smem_u32(smem_raw)) + 336);
    // CTA pair: rank 0 leads (issues the MMAs, owns the operand-full and accumulator-empty barriers both CTAs signal)
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    // barrier operations that differ between the two modes
    auto commit = [&](uint32_t bar) {                 // arrive (in both CTAs of a pair) when this thread's MMAs have completed
        if constexpr (PAIR) umma_pair_commit(bar, 3); else umma_commit(bar);
    };
    auto mma = [&](uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
        if constexpr (PAIR) umma_pair_f16_lh(d, a_lo, a_hi, b_lo, b_hi, idesc, acc); else umma_f16_lh(d, a_lo, a_hi, b_lo, b_hi, idesc, acc);
    };
    auto arrive_leader = [&](uint32_t bar) {          // accumulator-drained arrivals go to the leader's barrier
        if constexpr (PAIR) mbar_arrive_cluster(mapa_u32(bar, 0)); else mbar_arrive(bar);
    };
    auto arrive_both = [&](uint32_t bar) {            // plain arrival on this barrier in every CTA of the pair
        if constexpr (PAIR) { mbar_arrive_cluster(mapa_u32(bar, 0)); mbar_arrive_cluster(mapa_u32(bar, 1)); } else mbar_arrive(bar);
    };

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    long long* dbg = p.dbg_times != nullptr ? p.dbg_times + 16 * (size_t)blockIdx.x : nullptr;
    // debug only: cycles one thread of a role spent waiting on a barrier, accumulated in dbg[slot] (slots 8-15)
    long long wacc[3] = {0, 0, 0};          // per-thread accumulators, flushed by dbg_flush at the end of a role
    auto timed_wait = [&](uint32_t bar, uint32_t parity, int k) {
        if (dbg == nullptr) { mbar_wait(bar, parity); return; }
        const long long t0 = clock64();
        mbar_wait(bar, parity);
        wacc[k] += clock64() - t0;
    };
    auto dbg_flush = [&](int slot0, int n) {
        if (dbg != nullptr && (threadIdx.x & 31) == 0)
            for (int k = 0; k < n; ++k) dbg[slot0 + k] = wacc[k];
    };
    unsigned long long gt0 = 0;
    if (dbg != nullptr && threadIdx.x == 0) {
        dbg[0] = clock64();
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt0));
        dbg[2] = (long long)gt0;                  // absolute start of this CTA (ns): the host prints the launch stagger
    }

    pdl_trigger();       // the next kernel of the stream may be scheduled as SMs free up (it waits for this grid itself)
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA_hi);
        tma_prefetch_desc(&tmB_hi);
        if (p.passes == 3) {
            tma_prefetch_desc(&tmA_lo);
            tma_prefetch_desc(&tmB_lo);
        }
        // a stage is released by the commits of both issuing warps (main chain + cross terms) when the operands are split
        const uint32_t releasers = (p.passes == 3) ? 2u : 1u;
        for (int s = 0; s < Cfg::A_STAGES; ++s) { mbar_init(afull(s), 1); mbar_init(aempty(s), releasers); }
        for (int s = 0; s < Cfg::B_STAGES; ++s) { mbar_init(bfull(s), 1); mbar_init(bempty(s), releasers); }
        for (uint32_t a = 0; a < 4; ++a) {
            mbar_init(mfull(a), 1);
            mbar_init(mempty(a), 4 * Cfg::EPI_GROUPS * (PAIR ? 2 : 1));   // one arrival per epilogue warp (of both CTAs of a pair)
        }
        for (uint32_t a = 0; a < 2; ++a) {
            mbar_init(xfull(a), 1);
            mbar_init(xempty(a), 4 * Cfg::EPI_GROUPS * (PAIR ? 2 : 1));
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        if constexpr (PAIR) tmem_alloc_pair<Cfg::TMEM_COLS>(tmem_slot); else tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
    }
    tc_fence_before();
    if constexpr (PAIR) cluster_sync_all(); else __syncthreads();      // the peer's barriers are initialised, too
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;
    // everything above (barriers, tensor memory, descriptor prefetch) may overlap the previous kernel's tail; global memory
    // is touched only from here on
    pdl_wait();
    if (dbg != nullptr && threadIdx.x == 0) dbg[1] = clock64();

    const int cchunks = p.Cin >> 6;
    const int taps = p.taps;
    // Fused Gram backward (style layers): after the taps * cchunks k-steps of the convolution, `extra_chunks` more k-steps
    // compute D * F of this layer (A = the layer's feature planes, halo box, centre tap; B = the per-frame matrix
    // D = ((G - A) + (G - A)^T) * 2^e as fp16 planes; instruction descriptor idesc2) into their own tensor-memory accumulator
    // (the second cross buffer, so the cross terms are single-buffered in that case). Operand formats cannot be mixed inside
    // one MMA and the two products carry different scales, so the sum is formed in the epilogue: acc += alpha2[frame] * gram.
    const int xchunks = p.extra_chunks;
    const int promote = p.promote < 1 ? 1 : p.promote;
    const bool split = (p.passes == 3);
    const bool halo = (taps == 9);
    const int planes = split ? 2 : 1;
    // bytes counted on the (leader's) operand-full barriers: both CTAs of a pair load the same amounts
    const uint32_t a_tx = (uint32_t)((PAIR ? 2 : 1) * planes * (halo ? Cfg::HALO_BYTES : Cfg::EXACT_BYTES));
    const uint32_t b_tx = (uint32_t)((PAIR ? 2 : 1) * planes * Cfg::B_PLANE);
    const uint32_t a_sbo = halo ? (uint32_t)(Cfg::PW * 128) : 1024u;
    constexpr bool xsingle = Cfg::XSINGLE;
    const uint32_t nmain = (N_TILE == 128) ? (xchunks > 0 ? 2u : 3u) : 4u;       // main accumulation ring (see HaloCfg)
    // resident-weights mode (see the producer): stages 0 .. 8 hold the nine taps, stages 9 .. B_STAGES-1 are the ring of the fused
    // Gram k-steps. Only the pair / N = 64 kernel has the stages for it; the host sets the flag for Cin == 64, 9 taps.
    constexpr int RES_TAPS = 9;
    constexpr int RES_RING = (Cfg::B_STAGES > RES_TAPS) ? (Cfg::B_STAGES - RES_TAPS) : 1;
    const bool resident = PAIR && (N_TILE == 64) && (Cfg::B_STAGES > RES_TAPS) && p.b_resident != 0;

    // Work distribution ("stream-K" over 64-channel chunks), frame by frame. A tile is CH = cchunks + xchunks chunk units; the
    // Gf = tiles_per_frame * CH units of ONE frame are cut into `cpf` equal contiguous ranges (cpf = CTAs per frame, chosen by
    // the host from the frame geometry only), so that every SM gets the same amount of tensor work although the tile counts of
    // the VGG layers (128, 256, 512, ...) never divide by 148. gridDim.x / cpf frame groups work on different frames at the
    // same time. A CTA walks its range segment by segment; a segment that starts inside a tile (cbeg > 0; only the first
    // segment of a range can) leaves its fp32 partial tile in p.sk_ws and raises a flag; the CTA that holds the head of a tile
    // (cbeg == 0) adds the partials of the following CTAs in CTA order and runs the epilogue. The partition, hence the
    // rounding, of a frame is the same whatever the batch size; all sums have a fixed order (deterministic).
    // Partial producers wait only for their own previous-but-one partial to be consumed, heads wait only for producers with
    // a higher CTA index, and all CTAs are co-resident (grid <= number of SMs, one CTA per SM): the waits cannot deadlock.
    // Without a workspace the ranges are rounded to whole tiles (classic persistent tile loop).
    const int CH = cchunks + xchunks;
    const int tiles_xy = p.tiles_y * p.tiles_x;
    const int tiles_f = tiles_xy * p.tiles_n;          // tiles of one frame
    // a CTA pair is ONE worker: tiles are pair tiles (p.tiles_x counts pair columns), ranges are per pair, and the partial
    // tiles of rank r are exchanged between the rank-r CTAs of the pairs
    const int cpf = p.sk_cpf;                          // workers per frame
    const int wid = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int nworkers = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int nfg = nworkers / cpf;                    // frame groups
    const int fgroup = wid / cpf, lid = wid - fgroup * cpf;
    int g_begin, g_end;                                // the host guarantees tiles_f * CH < 2^31
    if (p.sk_ws != nullptr) {
        const long long G = (long long)tiles_f * CH;
        g_begin = (int)(G * lid / cpf);
        g_end = (int)(G * (lid + 1) / cpf);
    } else {
        g_begin = (int)((long long)tiles_f * lid / cpf) * CH;
        g_end = (int)((long long)tiles_f * (lid + 1) / cpf) * CH;
    }
    if (wid >= nfg * cpf) g_end = g_begin;             // surplus workers (grid not a multiple of cpf) have no work
#define IST_FOR_SEGMENTS(fr, fcount, tile, cbeg, cend)                                                          \
    for (int fr = fgroup, fcount = 0; fr < p.NB; fr += nfg, ++fcount)                                            \
        for (int g_ = g_begin, len_ = 0; g_ < g_end; g_ += len_)                                                 \
            if (const int tile = g_ / CH; true)                                                                  \
                if (const int cbeg = g_ - tile * CH; true)                                                       \
                    if (const int cend = (cbeg + (g_end - g_) < CH) ? (cbeg + (g_end - g_)) : CH; (len_ = cend - cbeg, true))

    if (warp == 0) {
        // ------------------------------------------------ TMA producer ------------------------------------------------
        if (lane == 0) {
            int as = 0, bs = 0;
            uint32_t aph = 0, bph = 0;
            bool b_loaded = false;
            int xs = 0;
            uint32_t xph = 0;
            IST_FOR_SEGMENTS(fr, fcount, tile, cbeg, cend) {
                (void)fcount;
                const int tn = tile / tiles_xy;
                const int tm = tile - tn * tiles_xy;
                const int ty = tm / p.tiles_x;
                const int tx = tm - ty * p.tiles_x;
                const int x0 = (PAIR ? 2 * tx + (int)rank : tx) * Cfg::TW - (halo ? 1 : 0), y0 = ty * Cfg::TH - (halo ? 1 : 0);
                const int n0 = tn * N_TILE + (int)rank * Cfg::B_ROWS;        // this CTA's half of the weight rows
                for (int cc = cbeg; cc < cend; ++cc) {
                    const bool ex = cc >= cchunks;
                    const int c64 = (ex ? cc - cchunks : cc) * 64;
                    timed_wait(aempty(as), aph ^ 1u, 0);
                    const uint32_t sA = a_base + as * Cfg::A_STAGE;
                    if constexpr (PAIR) {
                        // the leader announces the bytes of both CTAs; each CTA's loads complete on the leader's barrier
                        const uint32_t lbar = mapa_u32(afull(as), 0);
                        if (p.dbg_flags & 2) { if (rank == 0) mbar_arrive(afull(as)); } else {
                        if (rank == 0) mbar_arrive_expect_tx(afull(as), a_tx);
                        tma_load_4d_pair(sA, ex ? &tmF_hi : &tmA_hi, lbar, c64, x0, y0, fr);
                        if (split) tma_load_4d_pair(sA + Cfg::A_PLANE, ex ? &tmF_lo : &tmA_lo, lbar, c64, x0, y0, fr);
                        }
                    } else if (p.dbg_flags & 2) { mbar_arrive(afull(as)); } else {
                    mbar_arrive_expect_tx(afull(as), a_tx);
                    tma_load_4d(sA, ex ? &tmF_hi : &tmA_hi, afull(as), c64, x0, y0, fr);
                    if (split) tma_load_4d(sA + Cfg::A_PLANE, ex ? &tmF_lo : &tmA_lo, afull(as), c64, x0, y0, fr);
                    }
                    if (++as == Cfg::A_STAGES) { as = 0; aph ^= 1u; }
                    // Resident weights (64 -> 64 layers: ONE 64-channel chunk, so the nine weight tiles of a tile are the same for
                    // every tile of the launch and fit in nine 8 KB stages): loaded once, never released. Reloading them per tile
                    // was 72 KB of shared-memory writes per tile in the kernel that is shared-memory-bandwidth bound, and the
                    // issuers' largest wait. The fused Gram k-step (per-frame D matrix) keeps a ring in the stages behind them.
                    if (resident) {
                        if (!ex) {
                            if (!b_loaded) {
                                for (int tap = 0; tap < 9; ++tap) {
                                    const uint32_t sB = b_base + tap * Cfg::B_STAGE;
                                    const uint32_t lbar = mapa_u32(bfull(tap), 0);
                                    if (rank == 0) mbar_arrive_expect_tx(bfull(tap), b_tx);
                                    tma_load_3d_pair(sB, &tmB_hi, lbar, c64, n0, tap);
                                    if (split) tma_load_3d_pair(sB + Cfg::B_PLANE, &tmB_lo, lbar, c64, n0, tap);
                                }
                                b_loaded = true;
                            }
                        } else {
                            const int st = RES_TAPS + xs;
                            timed_wait(bempty(st), xph ^ 1u, 1);
                            const uint32_t sB = b_base + st * Cfg::B_STAGE;
                            const uint32_t lbar = mapa_u32(bfull(st), 0);
                            if (rank == 0) mbar_arrive_expect_tx(bfull(st), b_tx);
                            tma_load_3d_pair(sB, &tmD_hi, lbar, c64, n0, fr);
                            if (split) tma_load_3d_pair(sB + Cfg::B_PLANE, &tmD_lo, lbar, c64, n0, fr);
                            if (++xs == RES_RING) { xs = 0; xph ^= 1u; }
                        }
                        continue;
                    }
                    const int ntap = ex ? 1 : taps;
                    for (int tap = 0; tap < ntap; ++tap) {
                        const int bz = (p.b_frame || ex) ? fr : tap;
                        timed_wait(bempty(bs), bph ^ 1u, 1);
                        const uint32_t sB = b_base + bs * Cfg::B_STAGE;
                        if constexpr (PAIR) {
                            const uint32_t lbar = mapa_u32(bfull(bs), 0);
                            if (p.dbg_flags & 4) { if (rank == 0) mbar_arrive(bfull(bs)); } else {
                            if (rank == 0) mbar_arrive_expect_tx(bfull(bs), b_tx);
                            tma_load_3d_pair(sB, ex ? &tmD_hi : &tmB_hi, lbar, c64, n0, bz);
                            if (split) tma_load_3d_pair(sB + Cfg::B_PLANE, ex ? &tmD_lo : &tmB_lo, lbar, c64, n0, bz);
                            }
                        } else if (p.dbg_flags & 4) { mbar_arrive(bfull(bs)); } else {
                        mbar_arrive_expect_tx(bfull(bs), b_tx);
                        tma_load_3d(sB, ex ? &tmD_hi : &tmB_hi, bfull(bs), c64, n0, bz);
                        if (split) tma_load_3d(sB + Cfg::B_PLANE, ex ? &tmD_lo : &tmB_lo, bfull(bs), c64, n0, bz);
                        }
                        if (++bs == Cfg::B_STAGES) { bs = 0; bph ^= 1u; }
                    }
                }
            }
            dbg_flush(13, 2);
        }
    } else if (warp == 1) {
      if (rank == 0) {
        // ------------------------------------------ MMA issuer 1: hi*hi chains ------------------------------------------
        // Two warps issue MMAs (this one the short hi*hi chains, warp 6 the hi*lo + lo*hi cross terms): a single issuing
        // thread sustains about one tcgen05.mma per ~107 cycles (measured, tools/probes/umma_probe.cu), above the 64-cycle
        // execution time of a 128x128x16 MMA; two streams that write different accumulators reach 86.
        // The whole warp walks the (warp-uniform) loop; one elected lane issues. Descriptors are a constant high word plus
        // a low word (start address >> 4) that only receives small adds.
        const uint32_t a_hi_w = ((a_sbo >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
        const uint32_t b_hi_w = (1024u >> 4) | (1u << 14) | (2u << 29);
        const uint32_t idesc = p.idesc;
        const int kyn = halo ? 3 : 1;
        int as = 0, bs = 0;
        uint32_t aph = 0, bph = 0;
        uint32_t mb = 0, mph = 0;                  // main ring position and phase
        int xs1 = 0;                               // resident mode: ring of the fused Gram k-steps' D tiles
        uint32_t xph1 = 0;
        bool first = true;
        IST_FOR_SEGMENTS(fr, fcount, tile, cbeg, cend) {
            (void)tile; (void)fr; (void)fcount;
            const int mend = cend < cchunks ? cend : cchunks;          // main chunks of the segment: [cbeg, mend)
            int kit = 0;
            for (int cc = cbeg; cc < mend; ++cc) {
                timed_wait(afull(as), aph, 2);
                const uint32_t a_lo_stage = (a_base + as * Cfg::A_STAGE) >> 4;
                for (int ky = 0; ky < kyn; ++ky) {
                    for (int kx = 0; kx < kyn; ++kx, ++kit) {
                        const bool last_tap = (ky == kyn - 1 && kx == kyn - 1);
                        // chains never cross a 64-channel chunk: every tile of a layer sees the same chain partition wherever
                        // the stream-K ranges are cut (translation-invariant rounding)
                        const int in_chain = (ky * kyn + kx) % promote;
                        const bool chain_end = (in_chain == promote - 1) || last_tap;
                        if (in_chain == 0) timed_wait(mempty(mb), mph ^ 1u, 0);
                        const int bstage = resident ? (ky * kyn + kx) : bs;
                        timed_wait(bfull(bstage), resident ? 0u : bph, 1);      // resident tiles: phase 0 completes once, for good
                        tc_fence_after();
                        if (elect_one()) {
                            const uint32_t d_main = tmem_base + mb * N_TILE;
                            const uint32_t a_lo = a_lo_stage + (uint32_t)((ky * Cfg::PW + kx) * 8);
                            const uint32_t b_lo = (b_base + bstage * Cfg::B_STAGE) >> 4;
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4)
                                mma(d_main, a_lo + 2 * k4, a_hi_w, b_lo + 2 * k4, b_hi_w, idesc, (in_chain | k4) != 0 ? 1u : 0u);
                            if (chain_end) commit(mfull(mb));
                            if (!resident) commit(bempty(bs));
                            if (last_tap) commit(aempty(as));
                        }
                        __syncwarp();
                        first = false;
                        if (chain_end && ++mb == nmain) { mb = 0; mph ^= 1u; }
                        if (!resident && ++bs == Cfg::B_STAGES) { bs = 0; bph ^= 1u; }
                    }
                }
                if (++as == Cfg::A_STAGES) { as = 0; aph ^= 1u; }
            }
            // fused Gram k-steps are issued by warp 6 alone; this warp only keeps the stage rings in step and contributes
            // its share of the release arrivals
            for (int cc = (cbeg > cchunks ? cbeg : cchunks); cc < cend; ++cc) {
                const int bstage = resident ? RES_TAPS + xs1 : bs;
                mbar_wait(afull(as), aph);
                mbar_wait(bfull(bstage), resident ? xph1 : bph);
                if (elect_one()) {
                    arrive_both(bempty(bstage));
                    arrive_both(aempty(as));
                }
                __syncwarp();
                if (resident) { if (++xs1 == RES_RING) { xs1 = 0; xph1 ^= 1u; } }
                else if (++bs == Cfg::B_STAGES) { bs = 0; bph ^= 1u; }
                if (++as == Cfg::A_STAGES) { as = 0; aph ^= 1u; }
            }
        }
        dbg_flush(8, 3);
      }
    } else if (warp == 6) {
        // ------------------------------ MMA issuer 2: hi*lo + lo*hi cross terms, fused Gram k-steps ----------------------
        if (split && rank == 0) {
            const uint32_t a_hi_w = ((a_sbo >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
            const uint32_t b_hi_w = (1024u >> 4) | (1u << 14) | (2u << 29);
            const uint32_t idesc = p.idesc;
            const uint32_t idesc2 = p.idesc2;
            const int kyn = halo ? 3 : 1;
            int as = 0, bs = 0;
            uint32_t aph = 0, bph = 0;
            int xs2 = 0;
            uint32_t xph2 = 0;
            uint32_t scount = 0;
            IST_FOR_SEGMENTS(fr, fcount, tile, cbeg, cend) {
                (void)tile; (void)fr; (void)fcount;
                const uint32_t xa = xsingle ? 0u : (scount & 1u);
                const uint32_t xph = xsingle ? (scount & 1u) : ((scount >> 1) & 1u);
                const uint32_t d_cross = tmem_base + (uint32_t)Cfg::CROSS_COL + xa * N_TILE;
                const uint32_t d_gram = tmem_base + (uint32_t)Cfg::GRAM_COL + xa * N_TILE;
                timed_wait(xempty(xa), xph ^ 1u, 1);
                tc_fence_after();
                int kit = 0, xc = 0;
                for (int cc = cbeg; cc < cend; ++cc) {
                    const bool ex = cc >= cchunks;
                    const bool last_chunk = (cc == cend - 1);
                    timed_wait(afull(as), aph, 0);
                    const uint32_t a_lo_stage = (a_base + as * Cfg::A_STAGE) >> 4;
                    if (!ex) {
                        for (int ky = 0; ky < kyn; ++ky) {
                            for (int kx = 0; kx < kyn; ++kx, ++kit) {
                                const bool last_tap = (ky == kyn - 1 && kx == kyn - 1);
                                const int bstage = resident ? (ky * kyn + kx) : bs;
                                timed_wait(bfull(bstage), resident ? 0u : bph, 0);
                                tc_fence_after();
                                if (elect_one()) {
                                    const uint32_t a_lo = a_lo_stage + (uint32_t)((ky * Cfg::PW + kx) * 8);
                                    const uint32_t b_lo = (b_base + bstage * Cfg::B_STAGE) >> 4;
#pragma unroll
                                    for (int k4 = 0; k4 < 4; ++k4) {
                                        mma(d_cross, a_lo + 2 * k4, a_hi_w, b_lo + (Cfg::B_PLANE >> 4) + 2 * k4, b_hi_w, idesc,
                                            (kit | k4) != 0 ? 1u : 0u);
                                        mma(d_cross, a_lo + (Cfg::A_PLANE >> 4) + 2 * k4, a_hi_w, b_lo + 2 * k4, b_hi_w, idesc, 1u);
                                    }
                                    if (!resident) commit(bempty(bs));
                                    if (last_tap) commit(aempty(as));
                                    if (last_tap && last_chunk) commit(xfull(xa));
                                }
                                __syncwarp();
                                if (!resident && ++bs == Cfg::B_STAGES) { bs = 0; bph ^= 1u; }
                            }
                        }
                    } else {
                        const int bstage = resident ? RES_TAPS + xs2 : bs;
                        mbar_wait(bfull(bstage), resident ? xph2 : bph);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint32_t a_lo = a_lo_stage + (uint32_t)((Cfg::PW + 1) * 8);   // centre tap of the halo box
                            const uint32_t b_lo = (b_base + bstage * Cfg::B_STAGE) >> 4;
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4) {
                                mma(d_gram, a_lo + 2 * k4, a_hi_w, b_lo + 2 * k4, b_hi_w, idesc2, (xc | k4) != 0 ? 1u : 0u);
                                mma(d_gram, a_lo + 2 * k4, a_hi_w, b_lo + (Cfg::B_PLANE >> 4) + 2 * k4, b_hi_w, idesc2, 1u);
                                mma(d_gram, a_lo + (Cfg::A_PLANE >> 4) + 2 * k4, a_hi_w, b_lo + 2 * k4, b_hi_w, idesc2, 1u);
                            }
                            commit(bempty(bstage));
                            commit(aempty(as));
                            if (last_chunk) commit(xfull(xa));
                        }
                        __syncwarp();
                        ++xc;
                        if (resident) { if (++xs2 == RES_RING) { xs2 = 0; xph2 ^= 1u; } }
                        else if (++bs == Cfg::B_STAGES) { bs = 0; bph ^= 1u; }
                    }
                    if (++as == Cfg::A_STAGES) { as = 0; aph ^= 1u; }
                }
                ++scount;
            }
            dbg_flush(11, 2);
        }
    } else {
        // ------------------------------------------- promotion + epilogue ---------------------------------------------
        constexpr int EG = Cfg::EPI_GROUPS, NH = Cfg::NH;
        const int quad = warp & 3;                       // tensor-memory lane quadrant this warp may read
        const int grp = (EG == 2 && warp >= 7) ? 1 : 0;  // column half of this warp's group
        const int col0 = grp * NH;                       // first accumulator column of this thread
        const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)col0;
        const int m = quad * 32 + lane;                  // accumulator row == pixel inside the tile
        // named barriers: all epilogue threads (partial-tile hand-over), and the threads that fill one staging buffer
        auto bar_all = [&]() { if constexpr (EG == 1) named_bar_sync(1, 128); else named_bar_sync(3, 256); };
        auto bar_half = [&]() {
            if constexpr (EG == 1) named_bar_sync(1, 128);
            else if constexpr (N_TILE == 128) named_bar_sync(1 + grp, 128);
            else named_bar_sync(3, 256);
        };
        // the thread that issues the TMA store of a staging buffer (one per group when each group owns a 64-channel half)
        const bool storer = (EG == 2 && N_TILE == 128) ? (lane == 0 && (warp == 2 || warp == 7)) : (threadIdx.x == 64);
        uint32_t mb = 0, mph = 0, scount = 0;
        IST_FOR_SEGMENTS(fr, fcount, tile, cbeg, cend) {
            const int mend = cend < cchunks ? cend : cchunks;
            const int kit_seg = (mend > cbeg ? mend - cbeg : 0) * taps;
            const int nchains = (mend > cbeg ? mend - cbeg : 0) * ((taps + promote - 1) / promote);      // per chunk: ceil(taps / promote)
            const bool has_main = kit_seg > 0, has_extra = cend > cchunks;
            float acc[NH];
#pragma unroll
            for (int j = 0; j < NH; ++j) acc[j] = 0.f;
            const int slot = fcount & 1;                  // partial tiles are double-buffered over the frames a CTA walks
            // Head of a tile whose remaining chunks belong to the following workers of this frame group: their partial tiles
            // (published at the START of their ranges) are added in worker order. The 64 KB read through L2 happens `nmain`
            // chains before the end of this segment: the remaining chains fit in the accumulator ring, so the tensor core
            // finishes the segment while the partials are fetched (and while a late partial is waited for). The position of
            // the additions in the sum depends on the geometry only: results stay run-to-run identical.
            bool partials_done = !(p.sk_ws != nullptr && cbeg == 0 && cend < CH);
            auto add_partials = [&]() {
                int covered = cend;
                for (int ol = lid + 1; covered < CH; ++ol) {
                    const long long Gt = (long long)tiles_f * CH;
                    const int ob = (int)(Gt * ol / cpf), oe = (int)(Gt * (ol + 1) / cpf);
                    const int olen = (oe - ob) < (CH - covered) ? (oe - ob) : (CH - covered);
                    const int oc = PAIR ? 2 * (fgroup * cpf + ol) + (int)rank : fgroup * cpf + ol;      // same-rank CTA of worker ol
                    int* flag = p.sk_flags + 2 * oc + slot;
                    if (threadIdx.x == 64) {
                        int v = 0;
                        uint32_t spins = 0;
                        do {
                            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
                            if (++spins > (1u << 26)) __trap();
                        } while (v == 0);
                    }
                    bar_all();
                    const float* ws = p.sk_ws + (size_t)(2 * oc + slot) * (128 * N_TILE) + (size_t)col0 * 128;
#pragma unroll
                    for (int j = 0; j < NH; ++j) acc[j] += __ldcg(ws + j * 128 + m);
                    bar_all();
                    if (threadIdx.x == 64) {
                        asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flag), "r"(0) : "memory");
                    }
                    covered += olen;
                }
                partials_done = true;
            };
            const int gather_at = nchains > (int)nmain ? nchains - (int)nmain : 0;
            for (int ch = 0; ch < nchains; ++ch) {
                if (!partials_done && ch == gather_at) add_partials();
                timed_wait(mfull(mb), mph, 0);
                tc_fence_after();
#pragma unroll
                for (int c0 = 0; c0 < NH; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld_32x32(lane_base + mb * N_TILE + c0, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[c0 + j] += __uint_as_float(r[j]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) arrive_leader(mempty(mb));      // tcgen05.wait::ld is warp-wide: every lane's reads are done
                if (++mb == nmain) { mb = 0; mph ^= 1u; }
            }
            const int tn = tile / tiles_xy;
            const int tm = tile - tn * tiles_xy;
            const int ty = tm / p.tiles_x;
            const int tx = PAIR ? 2 * (tm - ty * p.tiles_x) + (int)rank : tm - ty * p.tiles_x;      // this CTA's pixel-tile column
            const int n0 = tn * N_TILE;
            if (split) {
                const uint32_t xa = xsingle ? 0u : (scount & 1u);
                mbar_wait(xfull(xa), xsingle ? (scount & 1u) : ((scount >> 1) & 1u));
                tc_fence_after();
                if (has_main) {
#pragma unroll
                    for (int c0 = 0; c0 < NH; c0 += 32) {
                        uint32_t r[32];
                        tmem_ld_32x32(lane_base + (uint32_t)Cfg::CROSS_COL + xa * N_TILE + c0, r);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) acc[c0 + j] += __uint_as_float(r[j]);
                    }
                }
                if (has_extra) {
                    // fused Gram term: acc (true units, alpha == 1 for data-gradients) += alpha2[frame] * (D * F)
                    const float a2 = __ldg(p.alpha2_dev + fr);
#pragma unroll
                    for (int c0 = 0; c0 < NH; c0 += 32) {
                        uint32_t r[32];
                        tmem_ld_32x32(lane_base + (uint32_t)Cfg::GRAM_COL + xa * N_TILE + c0, r);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) acc[c0 + j] = fmaf(a2, __uint_as_float(r[j]), acc[c0 + j]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) arrive_leader(xempty(xa));
            }
            ++scount;
            if (dbg != nullptr && threadIdx.x == 64) dbg[4] = clock64();
            if (cbeg != 0) {
                // partial tile of a segment that starts inside a tile: hand it to the CTA that holds the tile's head
                int* flag = p.sk_flags + 2 * blockIdx.x + slot;
                if (threadIdx.x == 64) {                  // the partial of two frames ago must have been consumed
                    int v = 1;
                    uint32_t spins = 0;
                    do {
                        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
                        if (++spins > (1u << 26)) __trap();
                    } while (v != 0);
                }
                bar_all();
                float* ws = p.sk_ws + (size_t)(2 * blockIdx.x + slot) * (128 * N_TILE) + (size_t)col0 * 128;
#pragma unroll
                for (int j = 0; j < NH; ++j) ws[j * 128 + m] = acc[j];
                __threadfence();
                bar_all();
                if (threadIdx.x == 64) {
                    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flag), "r"(1) : "memory");
                }
                continue;
            }
            if (!partials_done) add_partials();            // segments without main chains
            float alpha = p.alpha;
            if (p.alpha_dev != nullptr) alpha *= __ldg(p.alpha_dev + (size_t)p.alpha_stride * fr);
            const int gx = tx * Cfg::TW + (m % Cfg::TW), gy = ty * Cfg::TH + (m / Cfg::TW);
            const bool gvalid = (gx < p.W) && (gy < p.H);
            const size_t gpix = ((size_t)fr * p.H + (gvalid ? gy : 0)) * p.W + (gvalid ? gx : 0);
            if (p.use_tma_store) {
                // Plane outputs leave through shared memory and a TMA store: a "thread = pixel" register tile written
                // directly to NHWC memory touches 32 different lines per warp instruction. Each thread writes its pixel's
                // 64 channels (one 128-byte row, 16-byte chunks XOR-swizzled by the row index exactly as the tensor map's
                // SWIZZLE_128B expects) and one elected thread stores the 16x8-pixel box; the image border is clipped by TMA.
                // Staging buffers hold 64 channels (one 128-byte row per pixel). This thread's NH columns start at tile channel
                // col0: one group fills both halves of a 128-channel tile in turn; with two groups each group fills its own half
                // (N_TILE = 128, own buffer, own TMA store) or its own 32 channels of the single half (N_TILE = 64).
#pragma unroll
                for (int hh = 0; hh < (NH + 63) / 64; ++hh) {
                    const int h0 = (col0 & ~63) + 64 * hh;                  // 64-channel half this round belongs to
                    const uint32_t o_buf = o_base + (Cfg::OUT_BUFS == 2 ? (uint32_t)(h0 >> 6) * Cfg::OUT_BYTES : 0u);
                    if (storer) tma_store_wait_read();                     // the box last staged in this buffer has left it
                    bar_half();
#pragma unroll
                    for (int cb = 0; cb < (NH < 64 ? NH : 64); cb += 32) {
                        const int c0 = (col0 & 63) + cb;                    // channel offset inside the 64-channel half
                        const int ja = 64 * hh + cb;                        // accumulator index of the block's first column
                        float v[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = acc[ja + j] * alpha;
                        uint32_t hi[16], lo[16];
                        if (p.mode == CONV_FWD) conv_epilogue_regs_32(p, v, n0 + h0 + c0, hi, lo);
                        else conv_grad_regs_32(p, v, gpix * (size_t)p.Cout + n0 + h0 + c0, gvalid, hi, lo);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const uint32_t chunk = (uint32_t)(((c0 >> 3) + q) ^ (m & 7));
                            const uint32_t a = o_buf + (uint32_t)m * 128u + chunk * 16u;
                            st_shared_v4(a, hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
                            st_shared_v4(a + Cfg::OUT_PLANE, lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
                        }
                    }
                    fence_proxy_async_smem();
                    bar_half();
                    if (storer) {
                        tma_store_4d(&tmO_hi, o_buf, n0 + h0, tx * Cfg::TW, ty * Cfg::TH, fr);
                        tma_store_4d(&tmO_lo, o_buf + Cfg::OUT_PLANE, n0 + h0, tx * Cfg::TW, ty * Cfg::TH, fr);
                        tma_store_commit();
                    }
                }
            } else {
                if (gvalid) {
                    const size_t obase = gpix * (size_t)p.Cout + n0 + col0;
#pragma unroll
                    for (int c0 = 0; c0 < NH; c0 += 32) {
                        float v[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = acc[c0 + j] * alpha;
                        conv_epilogue_32(p, v, obase + c0, n0 + col0 + c0);
                    }
                }
            }
            if (dbg != nullptr && threadIdx.x == 64) dbg[5] = clock64();
        }
        // the staging buffers must have been read before the CTA may exit; the global writes themselves complete with the grid
        if (p.use_tma_store && storer) tma_store_wait_read();
        if (warp == 2) dbg_flush(15, 1);
    }
#undef IST_FOR_SEGMENTS

    tc_fence_before();
    if constexpr (PAIR) cluster_sync_all(); else __syncthreads();      // the peer's MMAs read this CTA's shared memory
    if (dbg != nullptr && threadIdx.x == 0) {
        dbg[6] = clock64();
        unsigned long long gt1;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt1));
        dbg[7] = (long long)(gt1 - gt0);          // nanoseconds: with dbg[6] - dbg[0] gives the SM clock actually running
        dbg[3] = (long long)gt1;                  // absolute end of this CTA (ns)
    }
    if (warp == 1) {
        tc_fence_after();
        if constexpr (PAIR) tmem_dealloc_pair<Cfg::TMEM_COLS>(tmem_base); else tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
    }
}

}  // namespace ist

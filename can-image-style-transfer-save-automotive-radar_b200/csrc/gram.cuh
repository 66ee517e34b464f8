// Gram matrix partial products G_part[split] = F^T F over a pixel range, as a tcgen05 SYRK.
//
// Replaces torch.bmm(F, F.transpose(1, 2)) of the reference (IST/model/meta_arch/gram_matrix.py:9).
// F lives in HBM as NHWC 16-bit planes (hi, lo): element (pixel n, channel c) at n*C + c. The contraction
// index is the pixel, so both operands are *MN-major* for the tensor core (channel contiguous): a TMA box
// of [64 pixels][64 channels] with 128-byte swizzle is exactly one column of MN-major SW128 atoms
// (8 pixel rows of 128 B each), SBO = 1024 between 8-pixel groups, LBO = 8192 between 64-channel groups.
// Only tiles with n_tile >= m_tile are computed (SYRK); off-diagonal tiles are also stored mirrored. Diagonal tiles reuse
// the A tile as B (no second load). Split over pixels gives >= 1 CTA per SM even for C = 64; partial
// results go to a [split] array that gram_finalize sums in a fixed order (deterministic, no atomics).
// 3-pass split (hi*hi + hi*lo + lo*hi) and the short-chain / register-promotion accumulation scheme of
// conv_igemm.cuh (the tensor core truncates when it adds into its accumulator; Gram sums are all-positive, so a
// long chain would be biased low by ~3e-8 per MMA).
#pragma once
#include "ptx.cuh"

namespace ist {

struct GramParams {
    int NB, HW, C;
    int tiles_c;           // ceil(C / 128)
    int n_tile;            // min(C, 128)
    int splits;            // pixel splits per frame
    int chunks_per_split;  // 64-pixel chunks per split
    int passes;
    int promote;           // 64-pixel k-steps per main accumulation chain
    uint32_t idesc;        // M = 128, N = n_tile, both MN-major
    float* partial;        // [NB][splits][C][C]
};

struct GramCfg {
    static constexpr int T_BYTES = 128 * 128;            // [64 pixels][128 channels] x 2 B, as two 8 KB channel groups
    static constexpr int STAGE_BYTES = 4 * T_BYTES;      // largest stage: A_hi, A_lo, B_hi, B_lo of an off-diagonal 128-wide tile
    static constexpr int STAGES = 3;                     // stages of that size; smaller stages get more (see RING_BYTES)
    static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
    static constexpr int MAX_STAGES = 8;                 // barrier slots
    static constexpr int TMEM_COLS = 512;                // main[2] @0,128; cross @256
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 + 1024;
};

// Several layers per launch: the style layers whose features are ready at the same time (relu1_1 ... relu4_1 while conv5_1 is
// still to run) share ONE grid — blocks [blk_end[l-1], blk_end[l]) work on layer l, longest CTAs (the shallow layers) first —
// so that the fixed cost of a launch (a wave of ~150 CTAs that live 9-17 us each, then an idle tail) is paid once and the
// tails of one layer fill with CTAs of the next.
constexpr int GRAM_MAX_LAYERS = 4;
struct GramMulti {
    GramParams L[GRAM_MAX_LAYERS];
    int blk_end[GRAM_MAX_LAYERS];
    int n;
};
struct alignas(64) GramMaps {
    CUtensorMap hi[GRAM_MAX_LAYERS];
    CUtensorMap lo[GRAM_MAX_LAYERS];
};

__global__ void __launch_bounds__(192, 1)
gram_syrk_kernel(const __grid_constant__ GramMaps maps, const __grid_constant__ GramMulti mp) {
    using Cfg = GramCfg;
    int layer = 0;
    while (layer + 1 < mp.n && (int)blockIdx.x >= mp.blk_end[layer]) ++layer;
    const GramParams& p = mp.L[layer];
    const CUtensorMap* tm_hi_p = &maps.hi[layer];
    const CUtensorMap* tm_lo_p = &maps.lo[layer];
    const int local = (int)blockIdx.x - (layer > 0 ? mp.blk_end[layer - 1] : 0);
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + Cfg::RING_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 64u + 8u * s; };
    auto mfull_bar = [&](uint32_t b) { return bar_base + 128u + 8u * b; };
    auto mempty_bar = [&](uint32_t b) { return bar_base + 144u + 8u * b; };
    const uint32_t xfull_bar = bar_base + 160u;
    const uint32_t tmem_slot = bar_base + 192u;
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(
        smem_raw + (smem_base - smem_u32(smem_raw)) + Cfg::RING_BYTES + 192);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // block decode inside the layer: split fastest, then the triangular tile id, then the frame
    const int tri = p.tiles_c * (p.tiles_c + 1) / 2;
    const int split = local % p.splits;
    const int tile_id = (local / p.splits) % tri;
    const int fr = local / (p.splits * tri);
    int mt = 0, nt = 0;
    {
        int t = tile_id;
        for (int i = 0; i < p.tiles_c; ++i) {
            const int row = p.tiles_c - i;
            if (t < row) { mt = i; nt = i + t; break; }
            t -= row;
        }
    }
    const bool diag = (mt == nt);
    const int total_chunks = (p.HW + 63) >> 6;
    const int c_begin = split * p.chunks_per_split;
    int c_end = c_begin + p.chunks_per_split;
    if (c_end > total_chunks) c_end = total_chunks;
    const int kiters = c_end > c_begin ? c_end - c_begin : 0;
    const int groups_a = (p.C >= 128) ? 2 : 1;           // 64-channel groups actually loaded per operand
    const int groups_b = p.n_tile >> 6;
    const int promote = p.promote < 1 ? 1 : p.promote;
    // Stage layout [A_hi][A_lo][B_hi][B_lo], every plane `plane_bytes`; diagonal tiles load no B, C = 64 has 8 KB planes. The
    // 192 KB ring is cut into as many stages as fit (up to 8): a C = 64 / C = 128 layer is a pure HBM stream (16 / 32 KB per
    // 64-pixel chunk), and with three stages in flight it ran at the DRAM latency (2.0-2.6 TB/s) instead of the bandwidth.
    const uint32_t plane_bytes = (uint32_t)groups_a * 8192u;
    const uint32_t stage_bytes = (uint32_t)((p.passes == 3) ? 2 : 1) * (diag ? 1u : 2u) * plane_bytes;
    int nstages = (int)(Cfg::RING_BYTES / stage_bytes);
    if (nstages > Cfg::MAX_STAGES) nstages = Cfg::MAX_STAGES;
    const uint32_t b_off = ((p.passes == 3) ? 2u : 1u) * plane_bytes;      // B planes follow the A planes

    pdl_trigger();
    if (threadIdx.x == 0) {
        tma_prefetch_desc(tm_hi_p);
        if (p.passes == 3) tma_prefetch_desc(tm_lo_p);
        for (int s = 0; s < Cfg::MAX_STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (uint32_t b = 0; b < 2; ++b) {
            mbar_init(mfull_bar(b), 1);
            mbar_init(mempty_bar(b), 128);
        }
        mbar_init(xfull_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) { tmem_alloc<Cfg::TMEM_COLS>(tmem_slot); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;
    pdl_wait();          // setup above overlaps the previous kernel's tail; the feature planes are read from here on

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const int planes = (p.passes == 3) ? 2 : 1;
            const uint32_t tx = (uint32_t)(planes * (groups_a + (diag ? 0 : groups_b)) * 8192);
            for (int kit = 0; kit < kiters; ++kit) {
                const int pix = (c_begin + kit) * 64;
                mbar_wait(empty_bar(stage), phase ^ 1u);
                const uint32_t s0 = smem_base + stage * stage_bytes;
                mbar_arrive_expect_tx(full_bar(stage), tx);
                for (int pl = 0; pl < planes; ++pl) {
                    const CUtensorMap* tm = pl == 0 ? tm_hi_p : tm_lo_p;
                    for (int g = 0; g < groups_a; ++g)
                        tma_load_3d(s0 + pl * plane_bytes + g * 8192, tm, full_bar(stage), mt * 128 + g * 64, pix, fr);
                    if (!diag)
                        for (int g = 0; g < groups_b; ++g)
                            tma_load_3d(s0 + b_off + pl * plane_bytes + g * 8192, tm, full_bar(stage),
                                        nt * 128 + g * 64, pix, fr);
                }
                if (++stage == nstages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        int stage = 0;
        uint32_t phase = 0;
        uint32_t mcount = 0;
        const bool split3 = (p.passes == 3);
        for (int kit = 0; kit < kiters; ++kit) {
            const int in_chain = kit % promote;
            const uint32_t mb = mcount & 1u;
            if (in_chain == 0) {
                mbar_wait(mempty_bar(mb), ((mcount >> 1) & 1u) ^ 1u);
                tc_fence_after();
            }
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            if (elect_one()) {
                // MN-major SW128 descriptors: low word = start >> 4 | (LBO = 8192 B) >> 4 << 16, constant high word
                const uint32_t sA = smem_base + stage * stage_bytes;
                const uint32_t sB = diag ? sA : sA + b_off;
                const uint32_t a_lo = (sA >> 4) | ((8192u >> 4) << 16);
                const uint32_t b_lo = (sB >> 4) | ((8192u >> 4) << 16);
                const uint32_t hi_w = (1024u >> 4) | (1u << 14) | (2u << 29);
                const uint32_t d_main = tmem_base + mb * 128u;
                const uint32_t d_cross = tmem_base + 256u;
                const uint32_t idesc = p.idesc;
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4)            // 16 pixels per MMA = two 8-row groups = 2048 B
                    umma_f16_lh(d_main, a_lo + k4 * 128, hi_w, b_lo + k4 * 128, hi_w, idesc, (in_chain | k4) != 0 ? 1u : 0u);
                if (in_chain == promote - 1 || kit == kiters - 1) umma_commit(mfull_bar(mb));
                if (split3) {
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                        umma_f16_lh(d_cross, a_lo + k4 * 128, hi_w, b_lo + (plane_bytes >> 4) + k4 * 128, hi_w, idesc,
                                    (kit | k4) != 0 ? 1u : 0u);
                        umma_f16_lh(d_cross, a_lo + (plane_bytes >> 4) + k4 * 128, hi_w, b_lo + k4 * 128, hi_w, idesc, 1u);
                    }
                }
                umma_commit(empty_bar(stage));
                if (split3 && kit == kiters - 1) umma_commit(xfull_bar);
            }
            __syncwarp();
            if (in_chain == promote - 1 || kit == kiters - 1) ++mcount;
            if (++stage == nstages) { stage = 0; phase ^= 1u; }
        }
    } else {
        const int quad = warp & 3;
        const int c1 = mt * 128 + quad * 32 + lane;
        const bool valid = (quad * 32 + lane) < (p.C >= 128 ? 128 : 64) && c1 < p.C;
        float* dst = p.partial + (((size_t)fr * p.splits + split) * p.C + (valid ? c1 : 0)) * (size_t)p.C + nt * 128;
        const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
        const int nchains = (kiters + promote - 1) / promote;
        float acc[128];
#pragma unroll
        for (int j = 0; j < 128; ++j) acc[j] = 0.f;
        for (int ch = 0; ch < nchains; ++ch) {
            const uint32_t mb = (uint32_t)ch & 1u;
            mbar_wait(mfull_bar(mb), ((uint32_t)ch >> 1) & 1u);
            tc_fence_after();
#pragma unroll
            for (int c0 = 0; c0 < 128; c0 += 32) {
                if (c0 < p.n_tile) {
                    uint32_t r[32];
                    tmem_ld_32x32(lane_base + mb * 128u + c0, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[c0 + j] += __uint_as_float(r[j]);
                }
            }
            tc_fence_before();
            mbar_arrive(mempty_bar(mb));
        }
        if (p.passes == 3 && kiters > 0) {
            mbar_wait(xfull_bar, 0);
            tc_fence_after();
#pragma unroll
            for (int c0 = 0; c0 < 128; c0 += 32) {
                if (c0 < p.n_tile) {
                    uint32_t r[32];
                    tmem_ld_32x32(lane_base + 256u + c0, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[c0 + j] += __uint_as_float(r[j]);
                }
            }
        }
        if (valid) {
#pragma unroll
            for (int c0 = 0; c0 < 128; c0 += 32) {
                if (c0 < p.n_tile) {
                    float4* d = reinterpret_cast<float4*>(dst + c0);
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        d[q] = make_float4(acc[c0 + 4 * q], acc[c0 + 4 * q + 1], acc[c0 + 4 * q + 2], acc[c0 + 4 * q + 3]);
                }
            }
            if (!diag) {
                // mirrored tile G[c2][c1] = G[c1][c2]: for a fixed column j the 32 lanes of a warp hold 32 consecutive c1,
                // so these scalar stores coalesce; gram_reduce then reads every element at its own (coalesced) address.
                float* mir = p.partial + (((size_t)fr * p.splits + split) * p.C + (size_t)nt * 128) * (size_t)p.C + c1;
#pragma unroll
                for (int j = 0; j < 128; ++j) mir[(size_t)j * p.C] = acc[j];
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
    }
}

}  // namespace ist

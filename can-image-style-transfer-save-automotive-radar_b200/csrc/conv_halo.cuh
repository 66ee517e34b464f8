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

template <int N_TILE>
struct HaloCfg {
    static constexpr int TW = 8, TH = 16, PW = TW + 2, PH = TH + 2;
    static constexpr int HALO_BYTES = PW * PH * 128;               // 23040
    static constexpr int EXACT_BYTES = TW * TH * 128;              // 16384
    static constexpr int A_PLANE = 23 * 1024;                      // 23552 >= HALO_BYTES, keeps every plane 1024-aligned
    static constexpr int A_STAGE = 2 * A_PLANE;
    static constexpr int A_STAGES = (N_TILE == 128) ? 2 : 3;
    static constexpr int B_PLANE = N_TILE * 128;
    static constexpr int B_STAGE = 2 * B_PLANE;
    static constexpr int B_STAGES = 4;
    static constexpr int TMEM_COLS = (4 * N_TILE < 32) ? 32 : 4 * N_TILE;
    static constexpr int SMEM_BYTES = A_STAGES * A_STAGE + B_STAGES * B_STAGE + 256 + 1024;
};

__device__ __forceinline__ uint64_t umma_desc_sw128_bo(uint32_t smem_addr, uint32_t sbo, uint32_t use_base_offset) {
    uint64_t d = umma_smem_desc_sw128(smem_addr, 0, sbo);
    if (use_base_offset) d |= static_cast<uint64_t>((smem_addr >> 7) & 7u) << 49;     // [49,52) matrix base offset
    return d;
}

template <int N_TILE>
__global__ void __launch_bounds__(192, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                 const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                 const ConvParams p) {
    using Cfg = HaloCfg<N_TILE>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_base = smem_base;
    const uint32_t b_base = smem_base + Cfg::A_STAGES * Cfg::A_STAGE;
    const uint32_t bar_base = b_base + Cfg::B_STAGES * Cfg::B_STAGE;
    // barriers (8 B each): a_full[3] @0, a_empty[3] @24, b_full[4] @48, b_empty[4] @80, main_full[2] @112,
    // main_empty[2] @128, cross_full[2] @144, cross_empty[2] @160, tmem base address @192
    auto afull = [&](int s) { return bar_base + 8u * s; };
    auto aempty = [&](int s) { return bar_base + 24u + 8u * s; };
    auto bfull = [&](int s) { return bar_base + 48u + 8u * s; };
    auto bempty = [&](int s) { return bar_base + 80u + 8u * s; };
    auto mfull = [&](uint32_t b) { return bar_base + 112u + 8u * b; };
    auto mempty = [&](uint32_t b) { return bar_base + 128u + 8u * b; };
    auto xfull = [&](uint32_t a) { return bar_base + 144u + 8u * a; };
    auto xempty = [&](uint32_t a) { return bar_base + 160u + 8u * a; };
    const uint32_t tmem_slot = bar_base + 192u;
    volatile uint32_t* tmem_slot_gen =
        reinterpret_cast<volatile uint32_t*>(smem_raw + (bar_base - smem_u32(smem_raw)) + 192);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA_hi);
        tma_prefetch_desc(&tmB_hi);
        if (p.passes == 3) {
            tma_prefetch_desc(&tmA_lo);
            tma_prefetch_desc(&tmB_lo);
        }
        for (int s = 0; s < Cfg::A_STAGES; ++s) { mbar_init(afull(s), 1); mbar_init(aempty(s), 1); }
        for (int s = 0; s < Cfg::B_STAGES; ++s) { mbar_init(bfull(s), 1); mbar_init(bempty(s), 1); }
        for (uint32_t a = 0; a < 2; ++a) {
            mbar_init(mfull(a), 1);
            mbar_init(mempty(a), 128);
            mbar_init(xfull(a), 1);
            mbar_init(xempty(a), 128);
        }
        fence_barrier_init();
    }
    if (warp == 1) { tmem_alloc<Cfg::TMEM_COLS>(tmem_slot); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    const int tiles_m = p.NB * p.tiles_y * p.tiles_x;
    const int total_tiles = tiles_m * p.tiles_n;
    const int cchunks = p.Cin >> 6;
    const int taps = p.taps;
    const int kiters = taps * cchunks;
    const int promote = p.promote < 1 ? 1 : p.promote;
    const bool split = (p.passes == 3);
    const bool halo = (taps == 9);
    const int planes = split ? 2 : 1;
    const uint32_t a_tx = (uint32_t)(planes * (halo ? Cfg::HALO_BYTES : Cfg::EXACT_BYTES));
    const uint32_t b_tx = (uint32_t)(planes * Cfg::B_PLANE);
    const uint32_t a_sbo = halo ? (uint32_t)(Cfg::PW * 128) : 1024u;
    const uint32_t use_bo = (uint32_t)p.desc_base_offset;

    if (warp == 0) {
        // ------------------------------------------------ TMA producer ------------------------------------------------
        if (lane == 0) {
            int as = 0, bs = 0;
            uint32_t aph = 0, bph = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int tm = tile % tiles_m;
                const int tn = tile / tiles_m;
                const int tx = tm % p.tiles_x;
                const int ty = (tm / p.tiles_x) % p.tiles_y;
                const int fr = tm / (p.tiles_x * p.tiles_y);
                const int x0 = tx * Cfg::TW - (halo ? 1 : 0), y0 = ty * Cfg::TH - (halo ? 1 : 0), n0 = tn * N_TILE;
                for (int cc = 0; cc < cchunks; ++cc) {
                    mbar_wait(aempty(as), aph ^ 1u);
                    const uint32_t sA = a_base + as * Cfg::A_STAGE;
                    mbar_arrive_expect_tx(afull(as), a_tx);
                    tma_load_4d(sA, &tmA_hi, afull(as), cc * 64, x0, y0, fr);
                    if (split) tma_load_4d(sA + Cfg::A_PLANE, &tmA_lo, afull(as), cc * 64, x0, y0, fr);
                    if (++as == Cfg::A_STAGES) { as = 0; aph ^= 1u; }
                    for (int tap = 0; tap < taps; ++tap) {
                        const int bz = p.b_frame ? fr : tap;
                        mbar_wait(bempty(bs), bph ^ 1u);
                        const uint32_t sB = b_base + bs * Cfg::B_STAGE;
                        mbar_arrive_expect_tx(bfull(bs), b_tx);
                        tma_load_3d(sB, &tmB_hi, bfull(bs), cc * 64, n0, bz);
                        if (split) tma_load_3d(sB + Cfg::B_PLANE, &tmB_lo, bfull(bs), cc * 64, n0, bz);
                        if (++bs == Cfg::B_STAGES) { bs = 0; bph ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer --------------------------------------------------
        int as = 0, bs = 0;
        uint32_t aph = 0, bph = 0;
        uint32_t mcount = 0, tcount = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
            const uint32_t xa = tcount & 1u;
            const uint32_t d_cross = tmem_base + (uint32_t)(2 * N_TILE) + xa * N_TILE;
            if (split) {
                mbar_wait(xempty(xa), ((tcount >> 1) & 1u) ^ 1u);
                tc_fence_after();
            }
            int kit = 0;
            for (int cc = 0; cc < cchunks; ++cc) {
                mbar_wait(afull(as), aph);
                const uint32_t sA = a_base + as * Cfg::A_STAGE;
                for (int tap = 0; tap < taps; ++tap, ++kit) {
                    const int in_chain = kit % promote;
                    const uint32_t mb = mcount & 1u;
                    if (in_chain == 0) {
                        mbar_wait(mempty(mb), ((mcount >> 1) & 1u) ^ 1u);
                    }
                    mbar_wait(bfull(bs), bph);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint32_t d_main = tmem_base + mb * N_TILE;
                        const uint32_t sB = b_base + bs * Cfg::B_STAGE;
                        uint32_t a_off = 0;
                        if (halo) a_off = (uint32_t)(((tap / 3) * Cfg::PW + (tap % 3)) * 128);
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4) {
                            const uint64_t a_hi = umma_desc_sw128_bo(sA + a_off + k4 * 32, a_sbo, use_bo);
                            const uint64_t b_hi = umma_smem_desc_sw128(sB + k4 * 32, 0, 1024);
                            umma_f16(d_main, a_hi, b_hi, p.idesc, (in_chain | k4) != 0 ? 1u : 0u);
                        }
                        const bool chain_end = (in_chain == promote - 1) || (kit == kiters - 1);
                        if (chain_end) umma_commit(mfull(mb));
                        if (split) {
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4) {
                                const uint64_t a_hi = umma_desc_sw128_bo(sA + a_off + k4 * 32, a_sbo, use_bo);
                                const uint64_t a_lo = umma_desc_sw128_bo(sA + Cfg::A_PLANE + a_off + k4 * 32, a_sbo, use_bo);
                                const uint64_t b_hi = umma_smem_desc_sw128(sB + k4 * 32, 0, 1024);
                                const uint64_t b_lo = umma_smem_desc_sw128(sB + Cfg::B_PLANE + k4 * 32, 0, 1024);
                                umma_f16(d_cross, a_hi, b_lo, p.idesc, (kit | k4) != 0 ? 1u : 0u);
                                umma_f16(d_cross, a_lo, b_hi, p.idesc, 1u);
                            }
                        }
                        umma_commit(bempty(bs));
                        if (tap == taps - 1) umma_commit(aempty(as));
                        if (split && kit == kiters - 1) umma_commit(xfull(xa));
                    }
                    __syncwarp();
                    if (in_chain == promote - 1 || kit == kiters - 1) ++mcount;
                    if (++bs == Cfg::B_STAGES) { bs = 0; bph ^= 1u; }
                }
                if (++as == Cfg::A_STAGES) { as = 0; aph ^= 1u; }
            }
        }
    } else {
        // ------------------------------------------- promotion + epilogue ---------------------------------------------
        const int quad = warp & 3;
        const int m = quad * 32 + lane;
        const int px = m % Cfg::TW, py = m / Cfg::TW;
        const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
        const int nchains = (kiters + promote - 1) / promote;
        uint32_t mcount = 0, tcount = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
            float acc[N_TILE];
#pragma unroll
            for (int j = 0; j < N_TILE; ++j) acc[j] = 0.f;
            for (int ch = 0; ch < nchains; ++ch, ++mcount) {
                const uint32_t mb = mcount & 1u;
                mbar_wait(mfull(mb), (mcount >> 1) & 1u);
                tc_fence_after();
#pragma unroll
                for (int c0 = 0; c0 < N_TILE; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld_32x32(lane_base + mb * N_TILE + c0, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[c0 + j] += __uint_as_float(r[j]);
                }
                tc_fence_before();
                mbar_arrive(mempty(mb));
            }
            if (split) {
                const uint32_t xa = tcount & 1u;
                mbar_wait(xfull(xa), (tcount >> 1) & 1u);
                tc_fence_after();
#pragma unroll
                for (int c0 = 0; c0 < N_TILE; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld_32x32(lane_base + (uint32_t)(2 * N_TILE) + xa * N_TILE + c0, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[c0 + j] += __uint_as_float(r[j]);
                }
                tc_fence_before();
                mbar_arrive(xempty(xa));
            }
            const int tm = tile % tiles_m;
            const int tn = tile / tiles_m;
            const int tx = tm % p.tiles_x;
            const int ty = (tm / p.tiles_x) % p.tiles_y;
            const int fr = tm / (p.tiles_x * p.tiles_y);
            const int x = tx * Cfg::TW + px, y = ty * Cfg::TH + py, n0 = tn * N_TILE;
            if ((x < p.W) && (y < p.H)) {
                const size_t pix = ((size_t)fr * p.H + y) * p.W + x;
                const size_t obase = pix * (size_t)p.Cout + n0;
                float alpha = p.alpha;
                if (p.alpha_dev != nullptr) alpha *= __ldg(p.alpha_dev + (size_t)p.alpha_stride * fr);
#pragma unroll
                for (int c0 = 0; c0 < N_TILE; c0 += 32) {
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = acc[c0 + j] * alpha;
                    conv_epilogue_32(p, v, obase + c0, n0 + c0);
                }
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

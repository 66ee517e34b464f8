// First conv (conv1_1: 3 -> 64 channels, K = 27) + bias + ReLU on the tensor cores.
//
// Reference: vgg.py:52 for conv1_1. The CUDA-core kernel (conv_first_fwd_kernel) is shared-memory bound (42 us at 512^2 for a
// 67 MB output). Here the 27 patch values of a pixel are one K = 32 row of an M = 128 (16 x 8 pixels), N = 64 tcgen05.mma:
//   * "builder" warps (thread = pixel) read the 3 x 18 x 10 input halo of the tile (fp32, the L-BFGS vector itself) from shared
//     memory, split every patch value into fp16 hi + lo and write their row of the A operand (K-major, 128-byte rows, the
//     16-byte chunks XOR-swizzled by the row index as SWIZZLE_128B expects); the halo of the next tile is fetched meanwhile;
//   * one warp issues 2 (k16) x 3 (hi*hi, hi*lo, lo*hi) MMAs per tile: hi*hi into one accumulator, the cross terms into a
//     second one (weights: fp16 hi / lo of 256 * w, resident in shared memory);
//   * two groups of four epilogue warps (32 channels each) add the accumulators, apply bias / ReLU / plane scale, split into
//     fp16 hi / lo and leave through a swizzled staging tile and a TMA store, as in conv_halo.cuh.
// Persistent CTAs, one per SM; tiles are independent (no split), accumulators double-buffered in tensor memory.
#pragma once
#include "conv_halo.cuh"

namespace ist {

struct CffTcParams {
    int NB, H, W, tiles_x, tiles_y;
    const float* x;         // fp32 NCHW [NB, 3, H, W]
    const float* w;         // fp32 [64][3][3][3]
    const float* bias;      // [64]
    float out_scale;        // plane scale of the activations
    uint32_t idesc;         // M = 128, N = 64, fp16 x fp16, both K-major
};

struct CffTcCfg {
    static constexpr int TW = 8, TH = 16, PW = TW + 2, PH = TH + 2;
    static constexpr int A_PLANE = 128 * 128;                      // 128 rows of 128 B (K = 32 used)
    static constexpr int A_STAGE = 2 * A_PLANE;                    // hi + lo
    static constexpr int A_STAGES = 2;
    static constexpr int B_PLANE = 64 * 128;
    static constexpr int OUT_PLANE = 128 * 128;
    static constexpr int SX_FLOATS = 3 * PH * PW;                  // 540
    static constexpr int SX_BYTES = 2304;                          // >= 540 * 4, multiple of 256
    static constexpr int W_SCALE_LOG2 = 8;                         // weights enter the MMA as fp16(256 * w)
    static constexpr int TMEM_COLS = 256;                          // 2 tiles x (main 64 + cross 64)
    static constexpr int THREADS = 13 * 32;                        // issuer, 4 builder warps, 8 epilogue warps
    static constexpr int OUT_BUFS = 2;                             // staging tiles in rotation: a tile is converted while the previous one's TMA store drains
    static constexpr int SMEM_BYTES = A_STAGES * A_STAGE + 2 * B_PLANE + OUT_BUFS * 2 * OUT_PLANE + SX_BYTES + 256 + 1024;
};

__global__ void __launch_bounds__(CffTcCfg::THREADS, 1)
conv_first_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmO_hi, const __grid_constant__ CUtensorMap tmO_lo,
                         const CffTcParams p) {
    using Cfg = CffTcCfg;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t a_base = smem_base;
    const uint32_t b_base = a_base + Cfg::A_STAGES * Cfg::A_STAGE;
    const uint32_t o_base = b_base + 2 * Cfg::B_PLANE;
    const uint32_t sx_base = o_base + Cfg::OUT_BUFS * 2 * Cfg::OUT_PLANE;
    const uint32_t bar_base = sx_base + Cfg::SX_BYTES;
    float* sx = reinterpret_cast<float*>(smem_al + (sx_base - smem_base));
    auto afull = [&](int s) { return bar_base + 8u * s; };
    auto aempty = [&](int s) { return bar_base + 16u + 8u * s; };
    auto accfull = [&](uint32_t b) { return bar_base + 32u + 8u * b; };
    auto accempty = [&](uint32_t b) { return bar_base + 48u + 8u * b; };
    const uint32_t tmem_slot = bar_base + 64u;
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_al + (bar_base - smem_base) + 64);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    pdl_trigger();
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmO_hi); tma_prefetch_desc(&tmO_lo);
        for (int s = 0; s < Cfg::A_STAGES; ++s) { mbar_init(afull(s), 1); mbar_init(aempty(s), 1); }
        for (uint32_t b = 0; b < 2; ++b) { mbar_init(accfull(b), 1); mbar_init(accempty(b), 8); }
        fence_barrier_init();
    }
    if (warp == 0) { tmem_alloc<Cfg::TMEM_COLS>(tmem_slot); }
    // weights (constants): B operand rows n = co, k = ci * 9 + tap (k >= 27 zero), fp16 hi / lo of 256 * w, swizzled rows.
    // Every (row, 16-byte chunk) of the first four logical chunks is written by one thread; the A stages' unused logical chunks
    // are never read (K = 32), so nothing else needs clearing.
    for (int i = threadIdx.x; i < 64 * 4; i += blockDim.x) {
        const int n = i >> 2, c = i & 3;
        uint32_t hw[4], lw[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int k0 = c * 8 + 2 * e;
            const float v0 = k0 < 27 ? __ldg(p.w + n * 27 + k0) * (float)(1 << Cfg::W_SCALE_LOG2) : 0.f;
            const float v1 = k0 + 1 < 27 ? __ldg(p.w + n * 27 + k0 + 1) * (float)(1 << Cfg::W_SCALE_LOG2) : 0.f;
            const uint32_t h = pack_h2(v0, v1);
            hw[e] = h;
            lw[e] = pack_h2(v0 - h_lo_f(h), v1 - h_hi_f(h));
        }
        const uint32_t a = b_base + (uint32_t)n * 128u + (uint32_t)((c ^ (n & 7)) * 16);
        st_shared_v4(a, hw[0], hw[1], hw[2], hw[3]);
        st_shared_v4(a + Cfg::B_PLANE, lw[0], lw[1], lw[2], lw[3]);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;
    pdl_wait();

    const int tiles_f = p.tiles_x * p.tiles_y;
    const int total = p.NB * tiles_f;
    const size_t HW = (size_t)p.H * p.W;

    if (warp == 0) {
        // ------------------------------------------------ MMA issuer ------------------------------------------------
        const uint32_t hi_w = (1024u >> 4) | (1u << 14) | (2u << 29);      // K-major SW128, SBO = 1024 (8 rows of 128 B)
        const uint32_t idesc = p.idesc;
        uint32_t cnt = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x, ++cnt) {
            const uint32_t s = cnt & 1u, ab = cnt & 1u, ph = (cnt >> 1) & 1u;
            mbar_wait(accempty(ab), ph ^ 1u);
            mbar_wait(afull(s), ph);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t d_main = tmem_base + ab * 128u, d_cross = d_main + 64u;
                const uint32_t a_lo = (a_base + s * Cfg::A_STAGE) >> 4, b_lo = b_base >> 4;
#pragma unroll
                for (int k2 = 0; k2 < 2; ++k2) {
                    umma_f16_lh(d_main, a_lo + 2 * k2, hi_w, b_lo + 2 * k2, hi_w, idesc, k2 != 0 ? 1u : 0u);
                    umma_f16_lh(d_cross, a_lo + 2 * k2, hi_w, b_lo + (Cfg::B_PLANE >> 4) + 2 * k2, hi_w, idesc, k2 != 0 ? 1u : 0u);
                    umma_f16_lh(d_cross, a_lo + (Cfg::A_PLANE >> 4) + 2 * k2, hi_w, b_lo + 2 * k2, hi_w, idesc, 1u);
                }
                umma_commit(aempty(s));
                umma_commit(accfull(ab));
            }
            __syncwarp();
        }
    } else if (warp <= 4) {
        // ------------------------------------- builders: thread = pixel of the tile -------------------------------------
        const int bt = threadIdx.x - 32;                  // 0..127
        const int m = bt, r = m >> 3, c = m & 7;
        // halo element ids this thread fetches: bt, bt + 128, ... < 540
        auto halo_fetch = [&](int t, float (&h)[5]) {
            const int fr = t / tiles_f, tm = t - fr * tiles_f, ty = tm / p.tiles_x, tx = tm - ty * p.tiles_x;
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                const int id = bt + 128 * q;
                float v = 0.f;
                if (id < Cfg::SX_FLOATS) {
                    const int ci = id / (Cfg::PH * Cfg::PW), rr = id - ci * (Cfg::PH * Cfg::PW), hy = rr / Cfg::PW, hx = rr - hy * Cfg::PW;
                    const int gy = ty * Cfg::TH - 1 + hy, gx = tx * Cfg::TW - 1 + hx;
                    if (gy >= 0 && gy < p.H && gx >= 0 && gx < p.W) v = __ldg(p.x + ((size_t)fr * 3 + ci) * HW + (size_t)gy * p.W + gx);
                }
                h[q] = v;
            }
        };
        auto halo_store = [&](const float (&h)[5]) {
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                const int id = bt + 128 * q;
                if (id < Cfg::SX_FLOATS) sx[id] = h[q];
            }
        };
        float hreg[5];
        if ((int)blockIdx.x < total) { halo_fetch(blockIdx.x, hreg); halo_store(hreg); }
        named_bar_sync(1, 128);
        uint32_t cnt = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x, ++cnt) {
            const uint32_t s = cnt & 1u, ph = (cnt >> 1) & 1u;
            const int tn = t + gridDim.x;
            if (tn < total) halo_fetch(tn, hreg);                    // in flight while this tile's rows are built
            mbar_wait(aempty(s), ph ^ 1u);
            // the pixel's 27 patch values, k = ci * 9 + ky * 3 + kx, as fp16 hi / lo pairs
            uint32_t hw[16], lw[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                float v[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int k = 2 * e + u;
                    if (k < 27) {
                        const int ci = k / 9, tap = k - ci * 9, ky = tap / 3, kx = tap - ky * 3;
                        v[u] = sx[(ci * Cfg::PH + r + ky) * Cfg::PW + c + kx];
                    } else {
                        v[u] = 0.f;
                    }
                }
                const uint32_t h = pack_h2(v[0], v[1]);
                hw[e] = h;
                lw[e] = pack_h2(v[0] - h_lo_f(h), v[1] - h_hi_f(h));
            }
            const uint32_t row = a_base + s * Cfg::A_STAGE + (uint32_t)m * 128u;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t a = row + (uint32_t)((q ^ (m & 7)) * 16);
                st_shared_v4(a, hw[4 * q], hw[4 * q + 1], hw[4 * q + 2], hw[4 * q + 3]);
                st_shared_v4(a + Cfg::A_PLANE, lw[4 * q], lw[4 * q + 1], lw[4 * q + 2], lw[4 * q + 3]);
            }
            fence_proxy_async_smem();                                // rows visible to the tensor core (async proxy)
            named_bar_sync(1, 128);                                  // all rows written, all reads of sx done
            if (bt == 0) mbar_arrive(afull(s));
            if (tn < total) halo_store(hreg);
            named_bar_sync(1, 128);                                  // next halo in place
        }
    } else {
        // --------------------------------------------- epilogue: two groups of four warps ---------------------------------------------
        const int ew = warp - 5;                          // 0..7
        const int grp = ew >> 2;                          // channels [32 * grp, 32 * grp + 32)
        const int quad = warp & 3;
        const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(grp * 32);
        const int m = quad * 32 + lane;
        const bool storer = (warp == 5 && lane == 0);
        ConvParams cp;                                    // the fields conv_epilogue_regs_32 reads
        cp.bias = p.bias;
        cp.out_scale = p.out_scale;
        const float alpha = 1.f / (float)(1 << Cfg::W_SCALE_LOG2);
        uint32_t cnt = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x, ++cnt) {
            const int fr = t / tiles_f, tm = t - fr * tiles_f, ty = tm / p.tiles_x, tx = tm - ty * p.tiles_x;
            const uint32_t ab = cnt & 1u;
            mbar_wait(accfull(ab), (cnt >> 1) & 1u);
            tc_fence_after();
            uint32_t rm[32], rc[32];
            tmem_ld_32x32(lane_base + ab * 128u, rm);
            tmem_ld_32x32(lane_base + ab * 128u + 64u, rc);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(accempty(ab));
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = (__uint_as_float(rm[j]) + __uint_as_float(rc[j])) * alpha;
            uint32_t hi[16], lo[16];
            conv_epilogue_regs_32(cp, v, grp * 32, hi, lo);
            const uint32_t o_buf = o_base + (cnt & 1u) * (2u * Cfg::OUT_PLANE);
            if (storer) tma_store_wait_read_1();                     // the tile staged two tiles ago has left this buffer
            named_bar_sync(2, 256);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t chunk = (uint32_t)((grp * 4 + q) ^ (m & 7));
                const uint32_t a = o_buf + (uint32_t)m * 128u + chunk * 16u;
                st_shared_v4(a, hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
                st_shared_v4(a + Cfg::OUT_PLANE, lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
            }
            fence_proxy_async_smem();
            named_bar_sync(2, 256);
            if (storer) {
                tma_store_4d(&tmO_hi, o_buf, 0, tx * Cfg::TW, ty * Cfg::TH, fr);
                tma_store_4d(&tmO_lo, o_buf + Cfg::OUT_PLANE, 0, tx * Cfg::TW, ty * Cfg::TH, fr);
                tma_store_commit();
            }
        }
        if (storer) tma_store_wait_read();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
    }
}

}  // namespace ist

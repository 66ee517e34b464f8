// Data-gradient of the first conv (conv1_1, 3 input channels) on the tensor cores.
//
// dX[ci, y, x] = sum_{co, ky, kx} W[co, ci, ky, kx] * dY[co, y - ky + 1, x - kx + 1]   (autograd of vgg.py:52 for conv1_1; the
// result is x.grad of IST/model/engine/utils.py:36). As an implicit GEMM: M = pixels (tiles of 16 x 8), K = 64 channels x 9
// taps, N = 3 padded to 16. The CUDA-core kernel (conv_first_dgrad_kernel) is shared-memory / L1 bound at ~1 TB/s (72 us at
// 512^2 for 70 MB); here the dY halo tile of conv_halo.cuh feeds all nine taps of an M = 128, N = 16 tcgen05.mma, whose cost is
// the 4 KB of A it reads per instruction. bf16 hi/lo split operands and the three products of every other data-gradient, in
// TWO instructions per k-slice: the hi and lo weight planes of a tap are adjacent in shared memory (32 rows), so
// dY_hi x [W_hi ; W_lo] is one N = 32 MMA (columns 0-15 = hi*hi, 16-31 = hi*lo) and dY_lo x W_hi one N = 16 MMA — the A
// operand is read twice instead of three times.
// Warp roles (224 threads, persistent): warp 0 TMA producer (weights once: 9 taps x 2 planes x 2 KB; dY ring of 3 halo tiles),
// warps 1 and 6 MMA issuers (taps 0-4 and 5-8, each with its own pair of accumulators: a single issuing thread sustains one
// small SS-mode MMA per ~50 cycles and was the bound at 45 us), warps 2-5 read the accumulators (double-buffered over tiles),
// add them and store the fp32 NCHW image gradient.
#pragma once
#include "conv_halo.cuh"

namespace ist {

struct CfdTcParams {
    int NB, H, W, tiles_x, tiles_y;
    float* grad;            // fp32 NCHW [NB, 3, H, W]
    uint32_t idesc;         // M = 128, N = 16, bf16 x bf16, both K-major
    uint32_t idesc32;       // same with N = 32 (hi and lo weight planes in one instruction)
};

struct CfdTcCfg {
    static constexpr int TW = 8, TH = 16, PW = TW + 2, PH = TH + 2;
    static constexpr int HALO_BYTES = PW * PH * 128;               // 23040
    static constexpr int A_PLANE = 23 * 1024;
    static constexpr int A_STAGE = 2 * A_PLANE;
    static constexpr int A_STAGES = 3;
    static constexpr int N_PAD = 16;
    static constexpr int B_TAP = N_PAD * 128;                      // one tap, one plane: 16 rows of 64 channels
    static constexpr int B_BYTES = 9 * 2 * B_TAP;                  // 36864
    static constexpr int TMEM_COLS = 256;                          // 2 tiles x 2 issuers x (hi*[hi;lo] 32 + lo*hi 16, padded to 64)
    static constexpr int SMEM_BYTES = A_STAGES * A_STAGE + B_BYTES + 256 + 1024;
};

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

__global__ void __launch_bounds__(224, 1)
conv_first_dgrad_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                           const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                           const CfdTcParams p) {
    using Cfg = CfdTcCfg;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_base = smem_base;
    const uint32_t b_base = smem_base + Cfg::A_STAGES * Cfg::A_STAGE;
    const uint32_t bar_base = b_base + Cfg::B_BYTES;
    auto afull = [&](int s) { return bar_base + 8u * s; };
    auto aempty = [&](int s) { return bar_base + 24u + 8u * s; };
    const uint32_t bfull = bar_base + 48u;
    auto accfull = [&](uint32_t b) { return bar_base + 56u + 8u * b; };
    auto accempty = [&](uint32_t b) { return bar_base + 72u + 8u * b; };
    const uint32_t tmem_slot = bar_base + 96u;
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (bar_base - smem_u32(smem_raw)) + 96);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    pdl_trigger();
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA_hi); tma_prefetch_desc(&tmA_lo);
        tma_prefetch_desc(&tmB_hi); tma_prefetch_desc(&tmB_lo);
        for (int s = 0; s < Cfg::A_STAGES; ++s) { mbar_init(afull(s), 1); mbar_init(aempty(s), 2); }   // both issuers release
        mbar_init(bfull, 1);
        for (uint32_t b = 0; b < 2; ++b) { mbar_init(accfull(b), 2); mbar_init(accempty(b), 4); }
        fence_barrier_init();
        // the weights are constants: they may be fetched while the previous kernel drains
        mbar_arrive_expect_tx(bfull, (uint32_t)Cfg::B_BYTES);
        for (int tap = 0; tap < 9; ++tap) {
            tma_load_3d(b_base + (uint32_t)(tap * 2) * Cfg::B_TAP, &tmB_hi, bfull, 0, 0, tap);
            tma_load_3d(b_base + (uint32_t)(tap * 2 + 1) * Cfg::B_TAP, &tmB_lo, bfull, 0, 0, tap);
        }
    }
    if (warp == 1) { tmem_alloc<Cfg::TMEM_COLS>(tmem_slot); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;
    pdl_wait();

    const int tiles_f = p.tiles_x * p.tiles_y;
    const int total = p.NB * tiles_f;

    if (warp == 0) {
        if (lane == 0) {
            int as = 0;
            uint32_t aph = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
                const int fr = t / tiles_f, tm = t - fr * tiles_f, ty = tm / p.tiles_x, tx = tm - ty * p.tiles_x;
                mbar_wait(aempty(as), aph ^ 1u);
                const uint32_t sA = a_base + as * Cfg::A_STAGE;
                mbar_arrive_expect_tx(afull(as), 2u * Cfg::HALO_BYTES);
                tma_load_4d(sA, &tmA_hi, afull(as), 0, tx * Cfg::TW - 1, ty * Cfg::TH - 1, fr);
                tma_load_4d(sA + Cfg::A_PLANE, &tmA_lo, afull(as), 0, tx * Cfg::TW - 1, ty * Cfg::TH - 1, fr);
                if (++as == Cfg::A_STAGES) { as = 0; aph ^= 1u; }
            }
        }
    } else if (warp == 1 || warp == 6) {
        const int issuer = warp == 1 ? 0 : 1;
        const int tap_begin = issuer == 0 ? 0 : 5, tap_end = issuer == 0 ? 5 : 9;
        const uint32_t a_hi_w = (((uint32_t)(Cfg::PW * 128) >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
        const uint32_t b_hi_w = (1024u >> 4) | (1u << 14) | (2u << 29);
        const uint32_t idesc = p.idesc, idesc32 = p.idesc32;
        int as = 0;
        uint32_t aph = 0, cnt = 0;
        mbar_wait(bfull, 0);
        for (int t = blockIdx.x; t < total; t += gridDim.x, ++cnt) {
            const uint32_t ab = cnt & 1u;
            mbar_wait(accempty(ab), ((cnt >> 1) & 1u) ^ 1u);
            mbar_wait(afull(as), aph);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t d_main = tmem_base + ab * 128u + (uint32_t)issuer * 64u, d_cross = d_main + 32u;
                const uint32_t a_stage = (a_base + as * Cfg::A_STAGE) >> 4;
                for (int tap = tap_begin; tap < tap_end; ++tap) {
                    const uint32_t a_lo = a_stage + (uint32_t)(((tap / 3) * Cfg::PW + (tap % 3)) * 8);
                    const uint32_t b_lo = (b_base + (uint32_t)(tap * 2) * Cfg::B_TAP) >> 4;
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                        const uint32_t acc = ((tap - tap_begin) | k4) != 0 ? 1u : 0u;
                        umma_f16_lh(d_main, a_lo + 2 * k4, a_hi_w, b_lo + 2 * k4, b_hi_w, idesc32, acc);             // hi * [hi ; lo]
                        umma_f16_lh(d_cross, a_lo + (Cfg::A_PLANE >> 4) + 2 * k4, a_hi_w, b_lo + 2 * k4, b_hi_w, idesc, acc);   // lo * hi
                    }
                }
                umma_commit(aempty(as));
                umma_commit(accfull(ab));
            }
            __syncwarp();
            if (++as == Cfg::A_STAGES) { as = 0; aph ^= 1u; }
        }
    } else {
        const int quad = warp & 3;
        const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
        const int m = quad * 32 + lane;
        const size_t HW = (size_t)p.H * p.W;
        uint32_t cnt = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x, ++cnt) {
            const int fr = t / tiles_f, tm = t - fr * tiles_f, ty = tm / p.tiles_x, tx = tm - ty * p.tiles_x;
            const uint32_t ab = cnt & 1u;
            mbar_wait(accfull(ab), (cnt >> 1) & 1u);
            tc_fence_after();
            // per issuer: columns [0,16) hi*hi, [16,32) hi*lo, [32,48) lo*hi; only the first three of each 16 are real channels
            uint32_t rm[16], rc[16], rx[16], rm2[16], rc2[16], rx2[16];
            tmem_ld_32x16(lane_base + ab * 128u, rm);
            tmem_ld_32x16(lane_base + ab * 128u + 16u, rc);
            tmem_ld_32x16(lane_base + ab * 128u + 32u, rx);
            tmem_ld_32x16(lane_base + ab * 128u + 64u, rm2);
            tmem_ld_32x16(lane_base + ab * 128u + 80u, rc2);
            tmem_ld_32x16(lane_base + ab * 128u + 96u, rx2);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(accempty(ab));
            const int gx = tx * Cfg::TW + (m % Cfg::TW), gy = ty * Cfg::TH + (m / Cfg::TW);
            if (gx < p.W && gy < p.H) {
                float* dst = p.grad + (size_t)fr * 3 * HW + (size_t)gy * p.W + gx;
#pragma unroll
                for (int ci = 0; ci < 3; ++ci)          // fixed order: (main + main') + ((hi*lo + lo*hi) + (hi*lo + lo*hi)')
                    dst[ci * HW] = (__uint_as_float(rm[ci]) + __uint_as_float(rm2[ci])) +
                                   ((__uint_as_float(rc[ci]) + __uint_as_float(rx[ci])) + (__uint_as_float(rc2[ci]) + __uint_as_float(rx2[ci])));
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

// weights of conv1_1 as the B operand of its data-gradient: [tap'][16 rows: ci, zero padded][64: co] bf16 hi / lo with
// tap' = 8 - tap (the halo formulation reads dY at y + ty - 1, so tap' = (ty, tx) pairs with W[.., 2 - ty, 2 - tx])
__global__ void cfd_tc_weight_repack_kernel(const float* __restrict__ w /*[64][3][3][3]*/, uint16_t* __restrict__ d_hi,
                                            uint16_t* __restrict__ d_lo) {
    const int total = 9 * 16 * 64;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int co = i & 63, n = (i >> 6) & 15, tp = i >> 10;
        float v = 0.f;
        if (n < 3) v = w[(co * 3 + n) * 9 + (8 - tp)];
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
        d_hi[i] = __bfloat16_as_ushort(h);
        d_lo[i] = __bfloat16_as_ushort(l);
    }
}

}  // namespace ist

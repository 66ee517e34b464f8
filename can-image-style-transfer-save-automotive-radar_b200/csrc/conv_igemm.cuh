// conv3x3 (pad 1, stride 1) and 1x1 contractions as a TMA-fed tcgen05 implicit GEMM.
//
// Replaces, on the IST hot path, cuDNN conv forward (reference IST/model/meta_arch/vgg.py:52),
// cuDNN conv backward-data reached through loss.backward() (IST/model/engine/utils.py:36) and the two
// bmm of the Gram backward (autograd of IST/model/meta_arch/gram_matrix.py:9).
//
// GEMM view: D[pixel, cout] = sum_{tap, cin} X[pixel + off(tap), cin] * Wt[tap][cout][cin]
//   M = 128 output pixels (a TH x TW rectangle of one frame), N = N_TILE output channels,
//   K = taps * Cin walked as (tap, 64-channel chunk).
// Operands are NHWC 16-bit *planes*: every fp32 tensor is carried as hi + lo (two fp16 or two bf16
// arrays); each K step issues hi*hi, hi*lo and lo*hi ("3-pass split", ~22 operand mantissa bits;
// the forward must be fp32-accurate or ReLU/pool masks flip).
// Zero padding comes from TMA out-of-bounds fill: the A box is fetched at (x0 + kx - 1, y0 + ky - 1).
// Every output pixel sees the same K order, so equal patches give bit-equal outputs.
//
// Accumulation. The tensor core adds into its fp32 accumulator with truncation, so one long accumulation
// chain drifts by ~3e-8 per MMA (measured: rel. error 2e-6 at K=576 growing to 1.3e-5 at K=4608, see
// profiles/r01_selftest_v1_single_accumulator.log). Therefore:
//   * hi*hi products go to a *short* chain (p.promote k-steps = 4*promote MMAs) in one of two "main" TMEM
//     buffers; the epilogue warps drain each finished chain into fp32 registers (round-to-nearest adds)
//     while the tensor core fills the other buffer;
//   * hi*lo + lo*hi (2^-11 smaller) run as one long chain in a separate "cross" TMEM buffer, whose
//     truncation error is 2^-11 smaller as well, and are added once at the end of the tile.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + single-thread MMA issuer,
// warps 2..5 = promotion + epilogue (TMEM -> registers -> global). Persistent CTAs walk tiles round-robin.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "ptx.cuh"

namespace ist {

enum ConvMode : int { CONV_FWD = 0, CONV_GRAD = 1 };

struct ConvParams {
    int NB, H, W;        // frames, spatial size (input == output)
    int Cin, Cout;       // contraction channels, output channels (GEMM K/taps and N)
    int taps;            // 9 (3x3) or 1 (1x1)
    int TW, TH;          // pixel tile, TW * TH == 128
    int tiles_x, tiles_y, tiles_n;
    int passes;          // 3 = hi*hi + hi*lo + lo*hi, 1 = hi*hi only
    int promote;         // k-steps (of 64 channels) per main accumulation chain, >= 1
    int b_frame;         // 1: third B coordinate is the frame (per-frame B, Gram backward), 0: the tap
    int b_resident;      // conv_halo pair kernel, Cin == 64, 9 taps: the nine weight tiles stay in shared memory for the whole launch
    int extra_chunks;      // conv_halo: fused Gram-backward k-steps appended to the convolution (0 = none)
    uint32_t idesc2;       // instruction descriptor of those k-steps (fp16 features x fp16 D matrix)
    const float* alpha2_dev;   // [NB] multiplier of the fused Gram accumulator
    float* sk_ws;          // conv_halo stream-K: [gridDim.x][2][128 * N_TILE] fp32 partial tiles (nullptr = whole-tile ranges)
    int* sk_flags;         // [gridDim.x][2] partial-ready flags, zero between launches
    int sk_cpf;            // CTAs that share one frame (gridDim.x / sk_cpf frame groups run side by side)
    int use_tma_store;     // conv_halo, CONV_FWD: the fp16 planes leave through shared memory + TMA store (tmO_hi / tmO_lo)
    int dbg_flags;         // timing experiments only: 2 = skip the A loads, 4 = skip the B loads (results are then garbage)
    long long* dbg_times;  // optional [gridDim.x][8] clock64 stamps of the kernel phases (IST_B200_DBG_TIMES=1)
    int pdl;               // host-side hint: launch as programmatically dependent on the previous kernel of the stream
    uint32_t idesc;
    int mode;
    float alpha;             // host multiplier on the accumulator
    const float* alpha_dev;  // optional device multiplier (per frame: alpha_dev[alpha_stride * frame])
    int alpha_stride;
    // CONV_FWD: v = relu(acc*alpha + bias); planes <- split_fp16(v * out_scale)
    const float* bias;
    float out_scale;
    // outputs: FWD fp16 planes; GRAD bf16 planes (when out_f32 == nullptr) or one fp32 NHWC array
    uint16_t* out_hi;
    uint16_t* out_lo;
    float* out_f32;
    // CONV_GRAD extras: v = acc*alpha (+ addend) (+ content_coef * (F - T)); v = mask > 0 ? v : 0
    const float* addend;        // fp32 NHWC, same shape as the output
    const uint16_t* mask_hi;    // fp16 hi plane of the ReLU output the gradient flows into
    const uint16_t* f_hi;       // content term: current feature planes (fp16, scaled)
    const uint16_t* f_lo;
    const uint16_t* t_hi;       // content target planes (same scaling)
    const uint16_t* t_lo;
    float content_coef;         // 2*w/(C*H*W) / plane scale
};

template <int N_TILE>
struct ConvCfg {
    static constexpr int A_BYTES = 128 * 128;        // 128 pixels x 64 ch x 2 B
    static constexpr int B_BYTES = N_TILE * 128;     // N_TILE couts x 64 ch x 2 B
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int STAGES = (N_TILE == 128 ? 3 : 4);
    static constexpr int TMEM_COLS = 4 * N_TILE;     // main[2] + cross[2]
    static constexpr int BAR_BYTES = 256;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;   // + alignment slack
};

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_bf2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float h_lo_f(uint32_t u) { return __half2float(__ushort_as_half((unsigned short)(u & 0xFFFF))); }
__device__ __forceinline__ float h_hi_f(uint32_t u) { return __half2float(__ushort_as_half((unsigned short)(u >> 16))); }
__device__ __forceinline__ float bf_lo_f(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi_f(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }

// Epilogue for 32 consecutive output channels of one pixel; `o` = element offset of the first channel, `cb` = its
// channel index (for the bias). v[] holds acc * alpha on entry.
__device__ __forceinline__ void conv_epilogue_32(const ConvParams& p, float (&v)[32], size_t o, int cb) {
    if (p.mode == CONV_FWD) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
            const float a = fmaxf(v[j] + __ldg(p.bias + cb + j), 0.f) * p.out_scale;
            const float b = fmaxf(v[j + 1] + __ldg(p.bias + cb + j + 1), 0.f) * p.out_scale;
            const uint32_t h = pack_h2(a, b);
            hi[j >> 1] = h;
            lo[j >> 1] = pack_h2(a - h_lo_f(h), b - h_hi_f(h));
        }
        uint4* dh = reinterpret_cast<uint4*>(p.out_hi + o);
        uint4* dl = reinterpret_cast<uint4*>(p.out_lo + o);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            dh[q] = make_uint4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
            dl[q] = make_uint4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
        }
        return;
    }
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
    if (p.out_f32 != nullptr) {
        float4* d = reinterpret_cast<float4*>(p.out_f32 + o);
#pragma unroll
        for (int q = 0; q < 8; ++q) d[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    } else {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
            const uint32_t h = pack_bf2(v[j], v[j + 1]);
            hi[j >> 1] = h;
            lo[j >> 1] = pack_bf2(v[j] - bf_lo_f(h), v[j + 1] - bf_hi_f(h));
        }
        uint4* dh = reinterpret_cast<uint4*>(p.out_hi + o);
        uint4* dl = reinterpret_cast<uint4*>(p.out_lo + o);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            dh[q] = make_uint4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
            dl[q] = make_uint4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
        }
    }
}

template <int N_TILE>
__global__ void __launch_bounds__(192, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                  const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                  const ConvParams p) {
    using Cfg = ConvCfg<N_TILE>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
    // barrier map (8 B each): full[s] @0, empty[s] @64, main_full[2] @128, main_empty[2] @144,
    // cross_full[2] @160, cross_empty[2] @176, tmem base address @192
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 64u + 8u * s; };
    auto mfull_bar = [&](uint32_t b) { return bar_base + 128u + 8u * b; };
    auto mempty_bar = [&](uint32_t b) { return bar_base + 144u + 8u * b; };
    auto xfull_bar = [&](uint32_t a) { return bar_base + 160u + 8u * a; };
    auto xempty_bar = [&](uint32_t a) { return bar_base + 176u + 8u * a; };
    const uint32_t tmem_slot = bar_base + 192u;
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(
        smem_raw + (smem_base - smem_u32(smem_raw)) + STAGES * Cfg::STAGE_BYTES + 192);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA_hi);
        tma_prefetch_desc(&tmB_hi);
        if (p.passes == 3) {
            tma_prefetch_desc(&tmA_lo);
            tma_prefetch_desc(&tmB_lo);
        }
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (uint32_t a = 0; a < 2; ++a) {
            mbar_init(mfull_bar(a), 1);
            mbar_init(mempty_bar(a), 128);
            mbar_init(xfull_bar(a), 1);
            mbar_init(xempty_bar(a), 128);
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
    const int kiters = p.taps * cchunks;
    const int promote = p.promote < 1 ? 1 : p.promote;
    const bool split = (p.passes == 3);
    const uint32_t stage_tx = split ? (uint32_t)Cfg::STAGE_BYTES : (uint32_t)(Cfg::A_BYTES + Cfg::B_BYTES);

    if (warp == 0) {
        // ------------------------------------------------ TMA producer ------------------------------------------------
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int tm = tile % tiles_m;
                const int tn = tile / tiles_m;
                const int tx = tm % p.tiles_x;
                const int ty = (tm / p.tiles_x) % p.tiles_y;
                const int fr = tm / (p.tiles_x * p.tiles_y);
                const int x0 = tx * p.TW, y0 = ty * p.TH, n0 = tn * N_TILE;
                for (int kit = 0; kit < kiters; ++kit) {
                    const int tap = kit / cchunks;
                    const int cc = kit - tap * cchunks;
                    int dx = 0, dy = 0;
                    if (p.taps == 9) {
                        dy = tap / 3 - 1;
                        dx = tap % 3 - 1;
                    }
                    const int bz = p.b_frame ? fr : tap;
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t sA = smem_base + stage * Cfg::STAGE_BYTES;
                    const uint32_t sB = sA + 2 * Cfg::A_BYTES;
                    mbar_arrive_expect_tx(full_bar(stage), stage_tx);
                    tma_load_4d(sA, &tmA_hi, full_bar(stage), cc * 64, x0 + dx, y0 + dy, fr);
                    tma_load_3d(sB, &tmB_hi, full_bar(stage), cc * 64, n0, bz);
                    if (split) {
                        tma_load_4d(sA + Cfg::A_BYTES, &tmA_lo, full_bar(stage), cc * 64, x0 + dx, y0 + dy, fr);
                        tma_load_3d(sB + Cfg::B_BYTES, &tmB_lo, full_bar(stage), cc * 64, n0, bz);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer --------------------------------------------------
        int stage = 0;
        uint32_t phase = 0;
        uint32_t mcount = 0;     // main chains issued so far (ring of 2 buffers)
        uint32_t tcount = 0;     // tiles issued so far (ring of 2 cross buffers)
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
            const uint32_t xa = tcount & 1u;
            const uint32_t d_cross = tmem_base + (uint32_t)(2 * N_TILE) + xa * N_TILE;
            if (split) {
                mbar_wait(xempty_bar(xa), ((tcount >> 1) & 1u) ^ 1u);
                tc_fence_after();
            }
            for (int kit = 0; kit < kiters; ++kit) {
                const int in_chain = kit % promote;
                const uint32_t mb = mcount & 1u;
                if (in_chain == 0) {
                    mbar_wait(mempty_bar(mb), ((mcount >> 1) & 1u) ^ 1u);
                    tc_fence_after();
                }
                mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t d_main = tmem_base + mb * N_TILE;
                    const uint32_t sA = smem_base + stage * Cfg::STAGE_BYTES;
                    const uint32_t sB = sA + 2 * Cfg::A_BYTES;
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                        const uint64_t a_hi = umma_smem_desc_sw128(sA + k4 * 32, 0, 1024);
                        const uint64_t b_hi = umma_smem_desc_sw128(sB + k4 * 32, 0, 1024);
                        umma_f16(d_main, a_hi, b_hi, p.idesc, (in_chain | k4) != 0 ? 1u : 0u);
                    }
                    const bool chain_end = (in_chain == promote - 1) || (kit == kiters - 1);
                    if (chain_end) umma_commit(mfull_bar(mb));
                    if (split) {
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4) {
                            const uint64_t a_hi = umma_smem_desc_sw128(sA + k4 * 32, 0, 1024);
                            const uint64_t b_hi = umma_smem_desc_sw128(sB + k4 * 32, 0, 1024);
                            const uint64_t a_lo = umma_smem_desc_sw128(sA + Cfg::A_BYTES + k4 * 32, 0, 1024);
                            const uint64_t b_lo = umma_smem_desc_sw128(sB + Cfg::B_BYTES + k4 * 32, 0, 1024);
                            umma_f16(d_cross, a_hi, b_lo, p.idesc, (kit | k4) != 0 ? 1u : 0u);
                            umma_f16(d_cross, a_lo, b_hi, p.idesc, 1u);
                        }
                    }
                    umma_commit(empty_bar(stage));
                    if (split && kit == kiters - 1) umma_commit(xfull_bar(xa));
                }
                __syncwarp();
                if (in_chain == promote - 1 || kit == kiters - 1) ++mcount;
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
        }
    } else {
        // ------------------------------------------- promotion + epilogue ---------------------------------------------
        const int quad = warp & 3;              // TMEM lane quadrant this warp may read
        const int m = quad * 32 + lane;         // accumulator row == pixel inside the tile
        const int px = m % p.TW, py = m / p.TW;
        const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
        const int nchains = (kiters + promote - 1) / promote;
        uint32_t mcount = 0, tcount = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
            float acc[N_TILE];
#pragma unroll
            for (int j = 0; j < N_TILE; ++j) acc[j] = 0.f;
            for (int ch = 0; ch < nchains; ++ch, ++mcount) {
                const uint32_t mb = mcount & 1u;
                mbar_wait(mfull_bar(mb), (mcount >> 1) & 1u);
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
                mbar_arrive(mempty_bar(mb));
            }
            if (split) {
                const uint32_t xa = tcount & 1u;
                mbar_wait(xfull_bar(xa), (tcount >> 1) & 1u);
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
                mbar_arrive(xempty_bar(xa));
            }
            // final epilogue from registers
            const int tm = tile % tiles_m;
            const int tn = tile / tiles_m;
            const int tx = tm % p.tiles_x;
            const int ty = (tm / p.tiles_x) % p.tiles_y;
            const int fr = tm / (p.tiles_x * p.tiles_y);
            const int x = tx * p.TW + px, y = ty * p.TH + py, n0 = tn * N_TILE;
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

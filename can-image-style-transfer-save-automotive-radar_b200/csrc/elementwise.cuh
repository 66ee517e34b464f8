// HBM-bound kernels of the Gatys closure: layout converts, the Cin=3 first conv (forward and data-gradient),
// 2x2 max-pool forward / scatter-backward with the reference's tie rule, content MSE, Gram finalisation.
// All tensors are NHWC 16-bit planes (hi, lo) unless a name says nchw / f32. Each thread moves 16-byte vectors
// along the channel axis (fully coalesced); reductions are two-level with a fixed order (no float atomics), so
// results are run-to-run bit-identical (SURVEY 8d "Determinism").
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "conv_igemm.cuh"

namespace ist {

// hi/lo pair of packed 16-bit values -> two fp32 sums (exact: hi + lo has <= 22 significant bits)
template <bool BF>
__device__ __forceinline__ void unpack_sum(uint32_t h, uint32_t l, float& a, float& b) {
    if (BF) {
        a = bf_lo_f(h) + bf_lo_f(l);
        b = bf_hi_f(h) + bf_hi_f(l);
    } else {
        a = h_lo_f(h) + h_lo_f(l);
        b = h_hi_f(h) + h_hi_f(l);
    }
}
template <bool BF>
__device__ __forceinline__ void split_pack(float a, float b, uint32_t& h, uint32_t& l) {
    if (BF) {
        h = pack_bf2(a, b);
        l = pack_bf2(a - bf_lo_f(h), b - bf_hi_f(h));
    } else {
        h = pack_h2(a, b);
        l = pack_h2(a - h_lo_f(h), b - h_hi_f(h));
    }
}

// Fixed-order block reduction (sum or max) over 256 threads; result valid in thread 0.
template <bool MAX>
__device__ __forceinline__ float block_reduce_256(float v, float* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float t = __shfl_down_sync(0xffffffffu, v, o);
        v = MAX ? fmaxf(v, t) : v + t;
    }
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        v = threadIdx.x < 8 ? sh[threadIdx.x] : (MAX ? 0.f : 0.f);
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            float t = __shfl_down_sync(0xffffffffu, v, o);
            v = MAX ? fmaxf(v, t) : v + t;
        }
    }
    __syncthreads();
    return v;
}

// ------------------------------------------------------------------------------------------------------------
// Layout converts (used at the API boundary: reference tensors are fp32 NCHW, vgg.py:44-58)
// ------------------------------------------------------------------------------------------------------------
template <bool BF>
__global__ void nchw_to_planes_kernel(const float* __restrict__ src, uint16_t* __restrict__ hi,
                                      uint16_t* __restrict__ lo, int NB, int C, int HW, float scale) {
    // one thread per (frame, pixel, channel pair); channel fastest so plane stores coalesce
    const size_t total = (size_t)NB * HW * (C >> 1);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int cp = (int)(i % (C >> 1));
        const size_t np = i / (C >> 1);
        const int pix = (int)(np % HW);
        const int n = (int)(np / HW);
        const float a = src[((size_t)n * C + 2 * cp) * HW + pix] * scale;
        const float b = src[((size_t)n * C + 2 * cp + 1) * HW + pix] * scale;
        uint32_t h, l;
        split_pack<BF>(a, b, h, l);
        reinterpret_cast<uint32_t*>(hi)[i] = h;
        reinterpret_cast<uint32_t*>(lo)[i] = l;
    }
}
template <bool BF>
__global__ void planes_to_nchw_kernel(const uint16_t* __restrict__ hi, const uint16_t* __restrict__ lo,
                                      float* __restrict__ dst, int NB, int C, int HW, float inv_scale) {
    // one thread per (frame, channel pair, pixel); pixel fastest so NCHW stores coalesce
    const size_t total = (size_t)NB * HW * (C >> 1);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int pix = (int)(i % HW);
        const size_t nc = i / HW;
        const int cp = (int)(nc % (C >> 1));
        const int n = (int)(nc / (C >> 1));
        const size_t s = ((size_t)n * HW + pix) * (C >> 1) + cp;
        float a, b;
        unpack_sum<BF>(reinterpret_cast<const uint32_t*>(hi)[s], reinterpret_cast<const uint32_t*>(lo)[s], a, b);
        dst[((size_t)n * C + 2 * cp) * HW + pix] = a * inv_scale;
        dst[((size_t)n * C + 2 * cp + 1) * HW + pix] = b * inv_scale;
    }
}
// argmax bytes of a pool layer, NHWC [NB,HW,C] -> NCHW [NB,C,HW] (ist_plan_get_pool_index)
__global__ void u8_nhwc_to_nchw_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int NB, int C, int HW) {
    const size_t total = (size_t)NB * HW * C;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int pix = (int)(i % HW);
        const size_t nc = i / HW;
        const int c = (int)(nc % C);
        const int n = (int)(nc / C);
        dst[i] = src[((size_t)n * HW + pix) * C + c];
    }
}
__global__ void nchw_to_nhwc_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, int NB, int C, int HW) {
    const size_t total = (size_t)NB * HW * C;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const size_t np = i / C;
        const int pix = (int)(np % HW);
        const int n = (int)(np / HW);
        dst[i] = src[((size_t)n * C + c) * HW + pix];
    }
}
__global__ void nhwc_to_nchw_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, int NB, int C, int HW) {
    const size_t total = (size_t)NB * HW * C;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int pix = (int)(i % HW);
        const size_t nc = i / HW;
        const int c = (int)(nc % C);
        const int n = (int)(nc / C);
        dst[i] = src[((size_t)n * HW + pix) * C + c];
    }
}

// ------------------------------------------------------------------------------------------------------------
// Weight repack: OIHW fp32 -> [tap][Cout][Cin] planes (forward, fp16, scaled) and [tap'][Cin][Cout] planes
// (data-gradient, bf16, taps flipped): dX = conv(dY, flip(W)^T).
// ------------------------------------------------------------------------------------------------------------
__global__ void weight_repack_kernel(const float* __restrict__ w, int Cout, int Cin, float scale_fwd,
                                     uint16_t* __restrict__ f_hi, uint16_t* __restrict__ f_lo,
                                     uint16_t* __restrict__ d_hi, uint16_t* __restrict__ d_lo) {
    const size_t total = (size_t)Cout * Cin * 9;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int tap = (int)(i % 9);
        const size_t oc = i / 9;
        const int ci = (int)(oc % Cin);
        const int co = (int)(oc / Cin);
        const float v = w[i];
        {
            const float s = v * scale_fwd;
            const __half h = __float2half_rn(s);
            const __half l = __float2half_rn(s - __half2float(h));
            const size_t o = ((size_t)tap * Cout + co) * Cin + ci;
            f_hi[o] = __half_as_ushort(h);
            f_lo[o] = __half_as_ushort(l);
        }
        {
            const __nv_bfloat16 h = __float2bfloat16_rn(v);
            const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
            const size_t o = ((size_t)(8 - tap) * Cin + ci) * Cout + co;
            d_hi[o] = __bfloat16_as_ushort(h);
            d_lo[o] = __bfloat16_as_ushort(l);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// First conv (Cin = 3, K = 27): HBM-bound, exact fp32 on CUDA cores. x is the reference's fp32 NCHW image
// (the L-BFGS vector); output = relu planes. Reference: vgg.py:52 for conv1_1.
// ------------------------------------------------------------------------------------------------------------
// Register-tiled: a block owns a 32 x 8 pixel tile and stages its 34 x 10 x 3 input halo in shared memory; four threads share
// a group of 4 consecutive pixels, thread q producing channels [8q, 8q+8) and [32+8q, 32+8q+8) for all four pixels (64
// accumulators). Per (ci, ky) a thread reads 6 inputs and, per kx, 4 broadcast float4 of weights for 64 FMAs; the four
// 16-byte stores of a pixel group form one contiguous 64-byte run per plane.
constexpr int CFF_TX = 32, CFF_TY = 8, CFF_HX = 36 /* 34 padded */, CFF_HY = 10;

template <int COUT>
__global__ void __launch_bounds__(256)
conv_first_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w /*[COUT][3][3][3]*/,
                      const float* __restrict__ bias, uint16_t* __restrict__ out_hi, uint16_t* __restrict__ out_lo,
                      int NB, int H, int W, float out_scale) {
    static_assert(COUT == 64, "channel split below is written for 64 output channels");
    __shared__ __align__(16) float ws[27][COUT];
    __shared__ __align__(16) float sx[3][CFF_HY][CFF_HX];
    __shared__ float bs[COUT];
    const int tid = threadIdx.x;
    for (int i = tid; i < 27 * COUT; i += blockDim.x) {
        const int co = i / 27, t = i % 27;      // w index = co*27 + (ci*9 + ky*3 + kx)
        ws[t][co] = w[i];
    }
    for (int i = tid; i < COUT; i += blockDim.x) bs[i] = bias[i];
    pdl_trigger();
    pdl_wait();          // the weights above are constants; the image is written by the previous kernel of the stream
    const int n = blockIdx.z;
    const int x0 = blockIdx.x * CFF_TX, y0 = blockIdx.y * CFF_TY;
    const size_t HW = (size_t)H * W;
    for (int i = tid; i < 3 * CFF_HY * 34; i += blockDim.x) {
        const int ci = i / (CFF_HY * 34), r = i % (CFF_HY * 34), hy = r / 34, hx = r % 34;
        const int gy = y0 - 1 + hy, gx = x0 - 1 + hx;
        sx[ci][hy][hx] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? __ldg(x + ((size_t)n * 3 + ci) * HW + (size_t)gy * W + gx) : 0.f;
    }
    __syncthreads();
    const int q = tid & 3, pq = tid >> 2;
    const int lx0 = (pq & 7) * 4, ly = pq >> 3;
    float acc[4][16];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 16; ++c) acc[i][c] = 0.f;
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            // output (ly, lx) reads input halo row ly + ky, columns lx + kx  (halo origin = tile origin - 1)
            const float* row = &sx[ci][ly + ky][lx0];
            const float4 d0 = *reinterpret_cast<const float4*>(row);
            const float2 d1 = *reinterpret_cast<const float2*>(row + 4);
            const float d[6] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y};
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int t = ci * 9 + ky * 3 + kx;
                float wv[16];
#pragma unroll
                for (int g = 0; g < 2; ++g)
#pragma unroll
                    for (int c = 0; c < 8; c += 4) {
                        const float4 w4 = *reinterpret_cast<const float4*>(&ws[t][g * 32 + q * 8 + c]);
                        wv[g * 8 + c] = w4.x; wv[g * 8 + c + 1] = w4.y; wv[g * 8 + c + 2] = w4.z; wv[g * 8 + c + 3] = w4.w;
                    }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int c = 0; c < 16; ++c) acc[i][c] = fmaf(d[i + kx], wv[c], acc[i][c]);
            }
        }
    }
    const int y = y0 + ly;
    if (y >= H) return;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int xx = x0 + lx0 + i;
        if (xx >= W) continue;
        const size_t pix = ((size_t)n * H + y) * W + xx;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const int cb = g * 32 + q * 8;
            uint32_t h[4], l[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float a = fmaxf(acc[i][g * 8 + 2 * e] + bs[cb + 2 * e], 0.f) * out_scale;
                const float b = fmaxf(acc[i][g * 8 + 2 * e + 1] + bs[cb + 2 * e + 1], 0.f) * out_scale;
                split_pack<false>(a, b, h[e], l[e]);
            }
            *reinterpret_cast<uint4*>(out_hi + pix * COUT + cb) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(out_lo + pix * COUT + cb) = make_uint4(l[0], l[1], l[2], l[3]);
        }
    }
}

// Data-gradient of the first conv: dX[ci,y,x] = sum_{co,ky,kx} W[co,ci,ky,kx] * dY[co, y-ky+1, x-kx+1].
// dY = bf16 planes already masked by relu1_1; output = fp32 NCHW image gradient (x.grad of utils.py:36).
// N = 3 output channels is no tensor-core shape, so this is a register-tiled CUDA-core kernel: a block owns a 32 x 16
// pixel tile, stages the (34 x 18)-pixel halo of dY for 16 channels at a time in shared memory as fp32 (hi + lo summed
// once instead of once per tap), and every thread produces 4 consecutive pixels x 3 channels, so that per (co, ky) it
// issues 2 shared loads of dY and 3 broadcast loads of weights for 36 FMAs. The 256 threads of a block form two halves that
// take 8 of the 16 staged channels each (twice the warps per block for the same shared memory); the halves' sums are combined
// through shared memory at the end, half 1 added to half 0 — a fixed order.
constexpr int CFD_TX = 32, CFD_TY = 16, CFD_HX = 36 /* 34 padded to a multiple of 4 */, CFD_HY = 18, CFD_CO = 16;
constexpr int CFD_SMEM = (CFD_CO * CFD_HY * CFD_HX + 64 * 9 * 4) * 4;

template <int COUT>
__global__ void __launch_bounds__(256)
conv_first_dgrad_kernel(const uint16_t* __restrict__ g_hi, const uint16_t* __restrict__ g_lo,
                        const float* __restrict__ w /*[COUT][3][3][3]*/, float* __restrict__ grad, int NB, int H, int W) {
    extern __shared__ __align__(16) float cfd_smem[];
    float* sd = cfd_smem;                                   // [CFD_CO][CFD_HY][CFD_HX]
    float4* ws = reinterpret_cast<float4*>(cfd_smem + CFD_CO * CFD_HY * CFD_HX);   // [COUT][9] (ci in .x .y .z)
    const int tid = threadIdx.x;
    for (int i = tid; i < COUT * 9; i += blockDim.x) {
        const int co = i / 9, tap = i % 9;
        ws[i] = make_float4(w[(co * 3 + 0) * 9 + tap], w[(co * 3 + 1) * 9 + tap], w[(co * 3 + 2) * 9 + tap], 0.f);
    }
    pdl_trigger();
    pdl_wait();          // weights are constants; the gradient planes come from the previous kernel
    const int n = blockIdx.z;
    const int x0 = blockIdx.x * CFD_TX, y0 = blockIdx.y * CFD_TY;
    const int half = tid >> 7, t7 = tid & 127;
    const int lx0 = (t7 & 7) * 4, ly = t7 >> 3;
    float acc[4][3];
#pragma unroll
    for (int i = 0; i < 4; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; acc[i][2] = 0.f; }
    for (int c0 = 0; c0 < COUT; c0 += CFD_CO) {
        __syncthreads();
        for (int hp = tid; hp < CFD_HY * 34; hp += blockDim.x) {
            const int hy = hp / 34, hx = hp % 34;
            const int gy = y0 - 1 + hy, gx = x0 - 1 + hx;
            float v[CFD_CO];
            if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
                const size_t o = ((((size_t)n * H + gy) * W + gx) * COUT + c0) / 8;
                const uint4* ph = reinterpret_cast<const uint4*>(g_hi) + o;
                const uint4* pl = reinterpret_cast<const uint4*>(g_lo) + o;
#pragma unroll
                for (int q = 0; q < CFD_CO / 8; ++q) {
                    const uint4 h = __ldg(ph + q), l = __ldg(pl + q);
                    const uint32_t uh[4] = {h.x, h.y, h.z, h.w}, ul[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) unpack_sum<true>(uh[e], ul[e], v[q * 8 + 2 * e], v[q * 8 + 2 * e + 1]);
                }
            } else {
#pragma unroll
                for (int c = 0; c < CFD_CO; ++c) v[c] = 0.f;
            }
#pragma unroll
            for (int c = 0; c < CFD_CO; ++c) sd[(c * CFD_HY + hy) * CFD_HX + hx] = v[c];
        }
        __syncthreads();
#pragma unroll 2
        for (int c = half * (CFD_CO / 2); c < (half + 1) * (CFD_CO / 2); ++c) {
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                // halo row of dY that tap ky of output row ly reads: (y - ky + 1) - (y0 - 1) = ly - ky + 2
                const float* row = sd + (c * CFD_HY + (ly - ky + 2)) * CFD_HX + lx0;
                const float4 d0 = *reinterpret_cast<const float4*>(row);
                const float2 d1 = *reinterpret_cast<const float2*>(row + 4);
                const float d[6] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y};
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const float4 wv = ws[(c0 + c) * 9 + ky * 3 + kx];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float dv = d[i - kx + 2];         // halo column (x - kx + 1) - (x0 - 1) - lx0
                        acc[i][0] = fmaf(dv, wv.x, acc[i][0]);
                        acc[i][1] = fmaf(dv, wv.y, acc[i][1]);
                        acc[i][2] = fmaf(dv, wv.z, acc[i][2]);
                    }
                }
            }
        }
    }
    __syncthreads();                                   // the dY stage is free: reuse it to combine the two channel halves
    if (half == 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) sd[(i * 3 + ci) * 128 + t7] = acc[i][ci];
    }
    __syncthreads();
    if (half == 1) return;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) acc[i][ci] += sd[(i * 3 + ci) * 128 + t7];
    const int y = y0 + ly, x = x0 + lx0;
    if (y >= H || x >= W) return;
    const size_t HW = (size_t)H * W;
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
        float* dst = grad + ((size_t)n * 3 + ci) * HW + (size_t)y * W + x;
        if (x + 3 < W && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
            *reinterpret_cast<float4*>(dst) = make_float4(acc[0][ci], acc[1][ci], acc[2][ci], acc[3][ci]);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (x + i < W) dst[i] = acc[i][ci];
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// MaxPool2d(2,2) forward (vgg.py:21-22,54): floor(H/2) x floor(W/2); first maximum in row-major window order
// wins (strict >), matching ATen. One thread = one output pixel x 8 channels.
// ------------------------------------------------------------------------------------------------------------
// `idx` (optional, one byte per pooled element): position 0..3 of the first maximum in row-major window order, or 4 when the
// maximum is not positive (the ReLU mask then stops the gradient). grad_route_kernel routes from it instead of re-reading
// the four pre-pool activations.
__global__ void maxpool_fwd_kernel(const uint16_t* __restrict__ in_hi, const uint16_t* __restrict__ in_lo,
                                   uint16_t* __restrict__ out_hi, uint16_t* __restrict__ out_lo, int NB, int H, int W,
                                   int C, uint8_t* __restrict__ idx) {
    const int Ho = H >> 1, Wo = W >> 1, C8 = C >> 3;
    const size_t total = (size_t)NB * Ho * Wo * C8;
    pdl_trigger();
    pdl_wait();
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % C8);
        size_t r = i / C8;
        const int xo = (int)(r % Wo);
        r /= Wo;
        const int yo = (int)(r % Ho);
        const int n = (int)(r / Ho);
        uint32_t bh[4], bl[4];
        float bv[8];
        uint32_t bk[2] = {0u, 0u};          // argmax position per channel, one byte each (channels 0-3, 4-7)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const size_t o = ((((size_t)n * H + 2 * yo + (k >> 1)) * W + 2 * xo + (k & 1)) * C) / 8 + c8;
            const uint4 h = __ldg(reinterpret_cast<const uint4*>(in_hi) + o);
            const uint4 l = __ldg(reinterpret_cast<const uint4*>(in_lo) + o);
            const uint32_t uh[4] = {h.x, h.y, h.z, h.w}, ul[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float a, b;
                unpack_sum<false>(uh[e], ul[e], a, b);
                if (k == 0) {
                    bv[2 * e] = a; bv[2 * e + 1] = b; bh[e] = uh[e]; bl[e] = ul[e];
                } else {
                    if (a > bv[2 * e]) {
                        bv[2 * e] = a;
                        bh[e] = (bh[e] & 0xFFFF0000u) | (uh[e] & 0xFFFFu);
                        bl[e] = (bl[e] & 0xFFFF0000u) | (ul[e] & 0xFFFFu);
                        const int sh = 8 * ((2 * e) & 3);
                        bk[e >> 1] = (bk[e >> 1] & ~(0xFFu << sh)) | ((uint32_t)k << sh);
                    }
                    if (b > bv[2 * e + 1]) {
                        bv[2 * e + 1] = b;
                        bh[e] = (bh[e] & 0xFFFFu) | (uh[e] & 0xFFFF0000u);
                        bl[e] = (bl[e] & 0xFFFFu) | (ul[e] & 0xFFFF0000u);
                        const int sh = 8 * ((2 * e + 1) & 3);
                        bk[e >> 1] = (bk[e >> 1] & ~(0xFFu << sh)) | ((uint32_t)k << sh);
                    }
                }
            }
        }
        reinterpret_cast<uint4*>(out_hi)[i] = make_uint4(bh[0], bh[1], bh[2], bh[3]);
        reinterpret_cast<uint4*>(out_lo)[i] = make_uint4(bl[0], bl[1], bl[2], bl[3]);
        if (idx != nullptr) {
#pragma unroll
            for (int c = 0; c < 8; ++c)
                if (!(bv[c] > 0.f)) bk[c >> 2] = (bk[c >> 2] & ~(0xFFu << (8 * (c & 3)))) | (4u << (8 * (c & 3)));
            reinterpret_cast<uint2*>(idx)[i] = make_uint2(bk[0], bk[1]);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Gradient routing at a ReLU output F (pre-pool resolution H x W):
//   v(pos) = [pos == first-argmax of its 2x2 window ? g_pool(window) : 0]  (if g_pool != nullptr)
//          + addend(pos) (if any) + content_coef * (F - T)(pos) (if any);  then v = F > 0 ? v : 0 (if mask)
// Output: bf16 planes (the dY the next dgrad consumes) or fp32 NHWC (when a style seed is still to be added).
// Covers autograd of max_pool2d + relu (threshold_backward) and the content MSE seed (main.py:36-37).
// One thread = one 2x2 window (ceil grid so odd edges are written too) x 4 channels.
// ------------------------------------------------------------------------------------------------------------
struct RouteParams {
    int NB, H, W, C;
    const float* g_pool;        // fp32 NHWC [NB, H/2, W/2, C] or nullptr
    const uint8_t* idx;         // maxpool_fwd_kernel's argmax bytes [NB, H/2, W/2, C] or nullptr; when given (only with
                                // g_pool, mask, plane output and no addend / content term) the activations are not read
    const uint16_t* f_hi;       // fp16 planes of F (argmax + mask + content term)
    const uint16_t* f_lo;
    const float* addend;        // fp32 NHWC [NB,H,W,C] or nullptr
    const uint16_t* t_hi;       // content target planes or nullptr
    const uint16_t* t_lo;
    float content_coef;
    int apply_mask;
    uint16_t* out_hi;           // bf16 planes, used when out_f32 == nullptr
    uint16_t* out_lo;
    float* out_f32;
};

// One thread = one 2x2 window x 4 channels (8-byte plane accesses, 16-byte fp32 accesses): half the registers of an
// 8-channel thread, so four 256-thread blocks fit per SM and enough loads are in flight to stream at HBM speed.
__global__ void __launch_bounds__(256, 4) grad_route_kernel(const RouteParams p) {
    const int Hc = (p.H + 1) >> 1, Wc = (p.W + 1) >> 1, C4 = p.C >> 2;
    const int Ho = p.H >> 1, Wo = p.W >> 1;
    const size_t total = (size_t)p.NB * Hc * Wc * C4;
    pdl_trigger();
    pdl_wait();
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % C4);
        size_t r = i / C4;
        const int xo = (int)(r % Wc);
        r /= Wc;
        const int yo = (int)(r % Hc);
        const int n = (int)(r / Hc);
        const bool window = (p.g_pool != nullptr) && (yo < Ho) && (xo < Wo);
        float g[4] = {0.f, 0.f, 0.f, 0.f};
        if (window) {
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.g_pool + ((((size_t)n * Ho + yo) * Wo + xo) * p.C) + c4 * 4));
            g[0] = g0.x; g[1] = g0.y; g[2] = g0.z; g[3] = g0.w;
        }
        float fv[4][4];
        bool inb[4];
        size_t o4[4];
        if (p.idx != nullptr) {
            // routing from the forward pass's argmax bytes: position k of the window receives g where idx == k (4 = masked)
            uint32_t code = 0x04040404u;
            if (window) code = __ldg(reinterpret_cast<const uint32_t*>(p.idx + ((((size_t)n * Ho + yo) * Wo + xo) * p.C)) + c4);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int yy = 2 * yo + (k >> 1), xx = 2 * xo + (k & 1);
                if (!((yy < p.H) && (xx < p.W))) continue;
                const size_t o = ((((size_t)n * p.H + yy) * p.W + xx) * p.C) / 4 + c4;
                float v[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) v[e] = (((code >> (8 * e)) & 0xFFu) == (uint32_t)k) ? g[e] : 0.f;
                uint32_t h0, l0, h1, l1;
                split_pack<true>(v[0], v[1], h0, l0);
                split_pack<true>(v[2], v[3], h1, l1);
                reinterpret_cast<uint2*>(p.out_hi)[o] = make_uint2(h0, h1);
                reinterpret_cast<uint2*>(p.out_lo)[o] = make_uint2(l0, l1);
            }
            continue;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int yy = 2 * yo + (k >> 1), xx = 2 * xo + (k & 1);
            inb[k] = (yy < p.H) && (xx < p.W);
            o4[k] = ((((size_t)n * p.H + (inb[k] ? yy : 0)) * p.W + (inb[k] ? xx : 0)) * p.C) / 4 + c4;
            if (inb[k]) {
                const uint2 h = __ldg(reinterpret_cast<const uint2*>(p.f_hi) + o4[k]);
                const uint2 l = __ldg(reinterpret_cast<const uint2*>(p.f_lo) + o4[k]);
                unpack_sum<false>(h.x, l.x, fv[k][0], fv[k][1]);
                unpack_sum<false>(h.y, l.y, fv[k][2], fv[k][3]);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) fv[k][e] = 0.f;
            }
        }
        int arg[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            int a = 0;
            float best = fv[0][e];
#pragma unroll
            for (int k = 1; k < 4; ++k)
                if (fv[k][e] > best) { best = fv[k][e]; a = k; }
            arg[e] = a;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (!inb[k]) continue;
            float v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = (window && arg[e] == k) ? g[e] : 0.f;
            if (p.addend != nullptr) {
                const float4 a0 = __ldg(reinterpret_cast<const float4*>(p.addend) + o4[k]);
                v[0] += a0.x; v[1] += a0.y; v[2] += a0.z; v[3] += a0.w;
            }
            if (p.t_hi != nullptr) {
                const uint2 fh = __ldg(reinterpret_cast<const uint2*>(p.f_hi) + o4[k]);
                const uint2 fl = __ldg(reinterpret_cast<const uint2*>(p.f_lo) + o4[k]);
                const uint2 th = __ldg(reinterpret_cast<const uint2*>(p.t_hi) + o4[k]);
                const uint2 tl = __ldg(reinterpret_cast<const uint2*>(p.t_lo) + o4[k]);
                v[0] += p.content_coef * ((h_lo_f(fh.x) - h_lo_f(th.x)) + (h_lo_f(fl.x) - h_lo_f(tl.x)));
                v[1] += p.content_coef * ((h_hi_f(fh.x) - h_hi_f(th.x)) + (h_hi_f(fl.x) - h_hi_f(tl.x)));
                v[2] += p.content_coef * ((h_lo_f(fh.y) - h_lo_f(th.y)) + (h_lo_f(fl.y) - h_lo_f(tl.y)));
                v[3] += p.content_coef * ((h_hi_f(fh.y) - h_hi_f(th.y)) + (h_hi_f(fl.y) - h_hi_f(tl.y)));
            }
            if (p.apply_mask) {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (!(fv[k][e] > 0.f)) v[e] = 0.f;
            }
            if (p.out_f32 != nullptr) {
                reinterpret_cast<float4*>(p.out_f32)[o4[k]] = make_float4(v[0], v[1], v[2], v[3]);
            } else {
                uint32_t h0, l0, h1, l1;
                split_pack<true>(v[0], v[1], h0, l0);
                split_pack<true>(v[2], v[3], h1, l1);
                reinterpret_cast<uint2*>(p.out_hi)[o4[k]] = make_uint2(h0, h1);
                reinterpret_cast<uint2*>(p.out_lo)[o4[k]] = make_uint2(l0, l1);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Content loss partial sums: sum (F - T)^2 in plane units; one block = one fixed slice; fixed-order tree.
// Reference: nn.MSELoss() on relu4_2 (main.py:36-37, utils.py:32).
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
content_partial_kernel(const uint16_t* __restrict__ f_hi, const uint16_t* __restrict__ f_lo,
                       const uint16_t* __restrict__ t_hi, const uint16_t* __restrict__ t_lo, size_t n8_per_frame,
                       float* __restrict__ partial /*[NB][gridDim.x]*/) {
    __shared__ float sh[8];
    const int fr = blockIdx.y;
    const uint4* a = reinterpret_cast<const uint4*>(f_hi) + (size_t)fr * n8_per_frame;
    const uint4* b = reinterpret_cast<const uint4*>(f_lo) + (size_t)fr * n8_per_frame;
    const uint4* c = reinterpret_cast<const uint4*>(t_hi) + (size_t)fr * n8_per_frame;
    const uint4* d = reinterpret_cast<const uint4*>(t_lo) + (size_t)fr * n8_per_frame;
    float s = 0.f;
    pdl_trigger();
    pdl_wait();
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n8_per_frame; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 va = __ldg(a + i), vb = __ldg(b + i), vc = __ldg(c + i), vd = __ldg(d + i);
        const uint32_t ua[4] = {va.x, va.y, va.z, va.w}, ub[4] = {vb.x, vb.y, vb.z, vb.w};
        const uint32_t uc[4] = {vc.x, vc.y, vc.z, vc.w}, ud[4] = {vd.x, vd.y, vd.z, vd.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float d0 = (h_lo_f(ua[e]) - h_lo_f(uc[e])) + (h_lo_f(ub[e]) - h_lo_f(ud[e]));
            const float d1 = (h_hi_f(ua[e]) - h_hi_f(uc[e])) + (h_hi_f(ub[e]) - h_hi_f(ud[e]));
            s = fmaf(d0, d0, s);
            s = fmaf(d1, d1, s);
        }
    }
    s = block_reduce_256<false>(s, sh);
    if (threadIdx.x == 0) partial[(size_t)fr * gridDim.x + blockIdx.x] = s;
}

// ------------------------------------------------------------------------------------------------------------
// Gram finalisation for up to 8 style layers in one launch (blockIdx.y = layer, blockIdx.z = frame).
//   G = (sum_splits partial) * g_scale               (GramMatrix.forward: bmm then div_(h*w), gram_matrix.py:9-10)
//   target mode: g_out <- G.   loss mode: diff <- G - A; per-block sum(diff^2) and max|diff|.
// ------------------------------------------------------------------------------------------------------------
struct GramLayer {
    const float* partial;   // [NB][splits][C][C] (the SYRK kernel also stores the mirrored tiles)
    const float* target;    // [C][C] shared by all frames (loss mode)
    float* g_out;           // [NB][C][C] (target mode) or nullptr
    float* diff;            // [NB][C][C]
    uint16_t* d_hi;         // [NB][C][C] fp16 planes of (diff + diff^T) * 2^e
    uint16_t* d_lo;
    float* blk_sum;         // [NB][GRAM_FIN_BLOCKS]
    float* blk_max;
    float* alpha_out;       // [NB] multiplier the Gram-backward GEMM applies to its accumulator
    float* loss_out;        // &losses[frame * loss_stride + slot], weighted
    int C, splits;
    float g_scale;          // 1 / (H*W * s_act^2)
    float weight;           // loss weight (STYLE_WEIGHTS[k], defaults.py:68)
    float bwd_coef;         // 2*w / (C^2 * H*W * s_act)
};
constexpr int GRAM_FIN_BLOCKS = 256;     // blocks per (layer, frame) of gram_reduce
struct GramFinalizeParams {
    GramLayer L[8];
    int n_layers, NB, loss_stride;
};

__global__ void __launch_bounds__(256)
gram_reduce_kernel(const GramFinalizeParams p) {
    __shared__ float sh[8];
    const GramLayer& L = p.L[blockIdx.y];
    const int fr = blockIdx.z, C = L.C;
    const size_t CC = (size_t)C * C;
    float s = 0.f, mx = 0.f;
    pdl_trigger();
    pdl_wait();
    // Every element sums its split partials in the fixed order sp = 0, 1, 2, ... . The loop is latency-bound (a thread owns
    // few elements): two elements x four splits = eight independent loads are kept in flight.
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const float* pbase = L.partial + (size_t)fr * L.splits * CC;
    if (L.splits >= 32) {
        // Narrow layers (C = 64 / 128: one split per SM, few elements): four lanes share an element, each adds a contiguous
        // quarter of the splits in order, a fixed two-step shuffle tree adds the quarters. A single thread per element would
        // walk splits / 16 dependent rounds of loads through L2.
        const int lane = threadIdx.x & 31, part = lane & 3, q = lane >> 2;
        const int per = (L.splits + 3) >> 2;
        const int sp0 = part * per, sp1 = (sp0 + per < L.splits) ? sp0 + per : L.splits;
        const size_t nquads = stride >> 2;
        for (size_t eb = ((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 2) - q; eb < CC; eb += nquads) {
            const size_t e = eb + q;
            const bool ok = e < CC;
            float g = 0.f;
            if (ok) {
                const float* p0 = pbase + e;
                int sp = sp0;
                for (; sp + 16 <= sp1; sp += 16) {
                    float a[16];
#pragma unroll
                    for (int u = 0; u < 16; ++u) a[u] = __ldg(p0 + (size_t)(sp + u) * CC);
#pragma unroll
                    for (int u = 0; u < 16; ++u) g += a[u];
                }
                for (; sp + 4 <= sp1; sp += 4) {
                    const float a0 = __ldg(p0 + (size_t)sp * CC), a1 = __ldg(p0 + (size_t)(sp + 1) * CC);
                    const float a2 = __ldg(p0 + (size_t)(sp + 2) * CC), a3 = __ldg(p0 + (size_t)(sp + 3) * CC);
                    g += a0; g += a1; g += a2; g += a3;
                }
                for (; sp < sp1; ++sp) g += __ldg(p0 + (size_t)sp * CC);
            }
            g += __shfl_xor_sync(0xffffffffu, g, 1);
            g += __shfl_xor_sync(0xffffffffu, g, 2);
            if (ok && part == 0) {
                g *= L.g_scale;
                if (L.g_out != nullptr) {
                    L.g_out[(size_t)fr * CC + e] = g;
                } else {
                    const float d = g - __ldg(L.target + e);
                    L.diff[(size_t)fr * CC + e] = d;
                    s = fmaf(d, d, s);
                    mx = fmaxf(mx, fabsf(d));
                }
            }
        }
    } else
    for (size_t e0 = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e0 < CC; e0 += 2 * stride) {
        const size_t e1 = e0 + stride;
        const bool two = e1 < CC;
        const float* p0 = pbase + e0;
        const float* p1 = pbase + (two ? e1 : e0);
        float g0 = 0.f, g1 = 0.f;
        int sp = 0;
        // layers with many splits (relu1_1: one split per SM) would otherwise walk a chain of splits / 4 dependent load rounds
        for (; sp + 16 <= L.splits; sp += 16) {
            float a[16], b[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) { a[u] = __ldg(p0 + (size_t)(sp + u) * CC); b[u] = __ldg(p1 + (size_t)(sp + u) * CC); }
#pragma unroll
            for (int u = 0; u < 16; ++u) { g0 += a[u]; g1 += b[u]; }
        }
        // remaining splits (fewer than 16): batches of 8 predicated loads (adding 0.f leaves the sum unchanged)
        for (; sp < L.splits; sp += 8) {
            float a[8], b[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const bool in = sp + u < L.splits;
                a[u] = in ? __ldg(p0 + (size_t)(sp + u) * CC) : 0.f;
                b[u] = in ? __ldg(p1 + (size_t)(sp + u) * CC) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) { g0 += a[u]; g1 += b[u]; }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (u == 1 && !two) break;
            const size_t e = u == 0 ? e0 : e1;
            const float g = (u == 0 ? g0 : g1) * L.g_scale;
            if (L.g_out != nullptr) {
                L.g_out[(size_t)fr * CC + e] = g;
            } else {
                const float d = g - __ldg(L.target + e);
                L.diff[(size_t)fr * CC + e] = d;
                s = fmaf(d, d, s);
                mx = fmaxf(mx, fabsf(d));
            }
        }
    }
    if (L.g_out == nullptr) {
        s = block_reduce_256<false>(s, sh);
        mx = block_reduce_256<true>(mx, sh);
        if (threadIdx.x == 0) {
            L.blk_sum[(size_t)fr * GRAM_FIN_BLOCKS + blockIdx.x] = s;
            L.blk_max[(size_t)fr * GRAM_FIN_BLOCKS + blockIdx.x] = mx;
        }
    }
}

// D = (diff + diff^T) * 2^e as fp16 planes (B operand of the Gram-backward GEMM: dF = coef * D * F,
// autograd of bmm + div_ + MSELoss), e chosen so |D| < 2^15; alpha_out = coef / 2^e.
__global__ void __launch_bounds__(256)
gram_dmat_kernel(const GramFinalizeParams p) {
    const GramLayer& L = p.L[blockIdx.y];
    const int fr = blockIdx.z, C = L.C;
    const size_t CC = (size_t)C * C;
    float mx = 0.f;
    pdl_trigger();
    pdl_wait();
    // every warp reduces the block maxima on its own (lane-strided loads + shuffles; max is order-independent)
    static_assert(GRAM_FIN_BLOCKS % 32 == 0, "lane-strided reductions below");
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int b = 0; b < GRAM_FIN_BLOCKS; b += 32) mx = fmaxf(mx, L.blk_max[(size_t)fr * GRAM_FIN_BLOCKS + b + lane]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    int e2 = 0;
    if (mx > 0.f && isfinite(mx)) e2 = 13 - ilogbf(mx);
    const float sc = ldexpf(1.f, e2);
    if (blockIdx.x == 0 && threadIdx.x < 32) {
        // sum of the block sums in a fixed order: lane l adds blocks [l * K, (l + 1) * K) in sequence, then a fixed shuffle tree
        constexpr int K = GRAM_FIN_BLOCKS / 32;
        double s = 0.0;
#pragma unroll
        for (int b = 0; b < K; ++b) s += (double)L.blk_sum[(size_t)fr * GRAM_FIN_BLOCKS + lane * K + b];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) {
        const float mean = (float)(s / (double)CC);
        *(L.loss_out + (size_t)fr * p.loss_stride) = L.weight * mean;
        L.alpha_out[fr] = L.bwd_coef * ldexpf(1.f, -e2);
      }
    }
    const float* df = L.diff + (size_t)fr * CC;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < CC; e += (size_t)gridDim.x * blockDim.x) {
        const int c1 = (int)(e / C), c2 = (int)(e % C);
        const float v = (df[e] + df[(size_t)c2 * C + c1]) * sc;
        const __half h = __float2half_rn(v);
        const __half l = __float2half_rn(v - __half2float(h));
        L.d_hi[(size_t)fr * CC + e] = __half_as_ushort(h);
        L.d_lo[(size_t)fr * CC + e] = __half_as_ushort(l);
    }
}

// total = sum_k losses[k] in list order (python sum(layer_losses), utils.py:32-35); content losses come from partials.
struct LossTotalParams {
    float* losses;           // [NB][loss_stride]
    int loss_stride, n_losses, NB;
    const float* c_partial[4];   // per content slot: [NB][c_blocks]
    int c_slot[4];
    float c_scale[4];        // weight / (C*H*W * s_act^2)
    int n_content, c_blocks;
};
// One warp per frame (launch <<<NB, 32>>>): lane l adds the partials [l * K, (l + 1) * K) in sequence, a fixed shuffle tree adds
// the lanes (deterministic order, no serial chain of dependent loads on the path between the forward and the backward pass).
__global__ void __launch_bounds__(32) loss_total_kernel(const LossTotalParams p) {
    const int fr = blockIdx.x, lane = threadIdx.x;
    pdl_trigger();
    pdl_wait();
    if (fr >= p.NB) return;
    float* L = p.losses + (size_t)fr * p.loss_stride;
    for (int k = 0; k < p.n_content; ++k) {
        const int K = (p.c_blocks + 31) / 32;
        double s = 0.0;
        for (int b = lane * K; b < (lane + 1) * K && b < p.c_blocks; ++b) s += (double)p.c_partial[k][(size_t)fr * p.c_blocks + b];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) L[p.c_slot[k]] = (float)(s * (double)p.c_scale[k]);
    }
    __syncwarp();
    if (lane == 0) {
        float t = 0.f;
        for (int k = 0; k < p.n_losses; ++k) t += L[k];
        L[p.n_losses] = t;
    }
}

}  // namespace ist

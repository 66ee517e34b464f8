// Image pre/post-processing on the device: the arithmetic of ImageTransform.preparation /
// post_preparation (IST/data/image_transform.py:8-31) and of the coarse-to-fine hand-off
// (IST/model/engine/hr_transfer_style.py:21-27), bit for bit:
//   post:   x * float(1/255) - float(-mean) ; BGR->RGB ; clamp [0,1] ; * 255 ; truncate to uint8     (HWC RGB)
//   resize: PIL Image.resize(BILINEAR) on 8-bit pixels = Pillow's two-pass fixed-point resampling (22 fractional bits,
//           horizontal pass first, 8-bit intermediate), coefficients computed on the host in double like Pillow does
//   prep:   (u8 / 255 - float(mean)) * 255 ; RGB->BGR                                                (CHW fp32)
// All three are HBM-bound byte kernels (<= 12 bytes moved per pixel); they keep a frame on the device between stages.
#pragma once
#include <map>
#include <mutex>
#include <tuple>

#include "host_common.cuh"

namespace ist {

constexpr int RS_PREC = 32 - 8 - 2;   // Pillow's PRECISION_BITS for 8-bit resampling

struct Mean3 { float m[3]; };

// x [NB,3,H,W] fp32 (BGR planes) -> rgb [NB,H,W,3] uint8. One thread = 4 consecutive pixels of one frame.
__global__ void __launch_bounds__(256) image_post_kernel(const float* __restrict__ x, uint8_t* __restrict__ rgb, int NB,
                                                          int HW, Mean3 negmean) {
    const float inv255 = (float)(1.0 / 255);
    const size_t quads = (size_t)(HW + 3) / 4;
    const size_t total = (size_t)NB * quads;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int n = (int)(i / quads);
        const int p0 = (int)(i % quads) * 4;
        const float* xf = x + (size_t)n * 3 * HW;
        uint8_t* o = rgb + ((size_t)n * HW + p0) * 3;
        const int cnt = HW - p0 < 4 ? HW - p0 : 4;
        uint8_t bytes[12];
#pragma unroll
        for (int c = 0; c < 3; ++c) {          // c = output channel (RGB); source plane 2 - c (BGR)
            const float* src = xf + (size_t)(2 - c) * HW + p0;
            float v[4];
            if (cnt == 4 && (HW & 3) == 0) {
                const float4 q = *reinterpret_cast<const float4*>(src);
                v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = j < cnt ? src[j] : 0.f;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float t = __fsub_rn(__fmul_rn(v[j], inv255), negmean.m[2 - c]);
                t = t > 1.f ? 1.f : t;
                t = t < 0.f ? 0.f : t;
                bytes[j * 3 + c] = (uint8_t)(int)__fmul_rn(t, 255.f);
            }
        }
        if (cnt == 4 && (HW & 3) == 0) {       // 12 bytes, 4-byte aligned because p0 % 4 == 0 and HW % 4 == 0
            uint32_t* o32 = reinterpret_cast<uint32_t*>(o);
#pragma unroll
            for (int k = 0; k < 3; ++k)
                o32[k] = (uint32_t)bytes[4 * k] | ((uint32_t)bytes[4 * k + 1] << 8) | ((uint32_t)bytes[4 * k + 2] << 16) |
                         ((uint32_t)bytes[4 * k + 3] << 24);
        } else {
            for (int k = 0; k < cnt * 3; ++k) o[k] = bytes[k];
        }
    }
}

// rgb [NB,H,W,3] uint8 -> x [NB,3,H,W] fp32 (BGR planes). One thread = 4 consecutive pixels.
__global__ void __launch_bounds__(256) image_prep_kernel(const uint8_t* __restrict__ rgb, float* __restrict__ x, int NB,
                                                          int HW, Mean3 mean) {
    const size_t quads = (size_t)(HW + 3) / 4;
    const size_t total = (size_t)NB * quads;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int n = (int)(i / quads);
        const int p0 = (int)(i % quads) * 4;
        const uint8_t* s = rgb + ((size_t)n * HW + p0) * 3;
        float* xf = x + (size_t)n * 3 * HW;
        const int cnt = HW - p0 < 4 ? HW - p0 : 4;
        const bool vec = cnt == 4 && (HW & 3) == 0;
        uint8_t bytes[12];
        if (vec) {
            const uint32_t* s32 = reinterpret_cast<const uint32_t*>(s);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const uint32_t w = s32[k];
                bytes[4 * k] = w & 255; bytes[4 * k + 1] = (w >> 8) & 255; bytes[4 * k + 2] = (w >> 16) & 255; bytes[4 * k + 3] = w >> 24;
            }
        } else {
            for (int k = 0; k < 12; ++k) bytes[k] = k < cnt * 3 ? s[k] : 0;
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {          // c = destination plane (BGR); source channel 2 - c
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                v[j] = __fmul_rn(__fsub_rn(__fdiv_rn((float)bytes[j * 3 + (2 - c)], 255.f), mean.m[c]), 255.f);
            float* dst = xf + (size_t)c * HW + p0;
            if (vec) *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
            else for (int j = 0; j < cnt; ++j) dst[j] = v[j];
        }
    }
}

// One resampling pass over bytes. The image is [outer][n_in][inner] bytes (inner = 3 for the horizontal pass over pixels,
// W*3 for the vertical pass over rows); output [outer][n_out][inner]. One thread per output byte.
__global__ void __launch_bounds__(256) image_resample_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                                                              size_t outer, int n_in, int n_out, int inner,
                                                              const int* __restrict__ bounds, const int* __restrict__ kk, int ksize) {
    const size_t total = outer * (size_t)n_out * inner;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % inner);
        const size_t r = i / inner;
        const int xx = (int)(r % n_out);
        const size_t o = r / n_out;
        const int x0 = bounds[2 * xx], cnt = bounds[2 * xx + 1];
        const int* k = kk + (size_t)xx * ksize;
        const uint8_t* src = in + (o * n_in + x0) * (size_t)inner + c;
        int ss = 1 << (RS_PREC - 1);
        for (int j = 0; j < cnt; ++j) ss += (int)src[(size_t)j * inner] * k[j];
        ss >>= RS_PREC;
        out[i] = (uint8_t)(ss < 0 ? 0 : (ss > 255 ? 255 : ss));
    }
}

// ---- host side: Pillow's precompute_coeffs + normalize_coeffs_8bpc for the bilinear filter, cached per (device, in, out) ----
struct ResampleTable {
    int ksize = 0;
    int* bounds = nullptr;   // device [out][2]: first input index, tap count
    int* kk = nullptr;       // device [out][ksize]: fixed-point weights
};

inline void resample_coeffs_host(int in_size, int out_size, int* ksize_out, std::vector<int>& bounds, std::vector<int>& kk) {
    const double scale = (double)in_size / (double)out_size;
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = 1.0 * filterscale;
    const int ksize = (int)ceil(support) * 2 + 1;
    bounds.assign((size_t)out_size * 2, 0);
    kk.assign((size_t)out_size * ksize, 0);
    std::vector<double> w((size_t)ksize);
    const double ss = 1.0 / filterscale;
    for (int xx = 0; xx < out_size; ++xx) {
        const double center = 0.0 + (xx + 0.5) * scale;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        double ww = 0.0;
        for (int x = 0; x < xmax; ++x) {
            double a = (x + xmin - center + 0.5) * ss;
            if (a < 0.0) a = -a;
            const double v = a < 1.0 ? 1.0 - a : 0.0;
            w[x] = v;
            ww += v;
        }
        for (int x = 0; x < xmax; ++x) {
            double v = w[x];
            if (ww != 0.0) v /= ww;
            kk[(size_t)xx * ksize + x] = v < 0 ? (int)(-0.5 + v * (1 << RS_PREC)) : (int)(0.5 + v * (1 << RS_PREC));
        }
        bounds[2 * xx] = xmin;
        bounds[2 * xx + 1] = xmax;
    }
    *ksize_out = ksize;
}

inline int resample_table(int in_size, int out_size, ResampleTable* out) {
    static std::mutex mu;
    static std::map<std::tuple<int, int, int>, ResampleTable> cache;
    int dev = 0;
    IST_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    auto key = std::make_tuple(dev, in_size, out_size);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return IST_OK; }
    std::vector<int> b, k;
    ResampleTable t;
    resample_coeffs_host(in_size, out_size, &t.ksize, b, k);
    IST_CUDA(cudaMalloc(&t.bounds, b.size() * sizeof(int)));
    IST_CUDA(cudaMalloc(&t.kk, k.size() * sizeof(int)));
    IST_CUDA(cudaMemcpy(t.bounds, b.data(), b.size() * sizeof(int), cudaMemcpyHostToDevice));
    IST_CUDA(cudaMemcpy(t.kk, k.data(), k.size() * sizeof(int), cudaMemcpyHostToDevice));
    cache[key] = t;
    *out = t;
    return IST_OK;
}

inline int launch_image_post(cudaStream_t st, const float* x, uint8_t* rgb, int NB, int H, int W, const double* mean_bgr) {
    Mean3 nm;
    for (int c = 0; c < 3; ++c) nm.m[c] = (float)((-1) * mean_bgr[c]);    // torch.as_tensor([(-1)*m], dtype=float32)
    const size_t quads = (size_t)NB * (((size_t)H * W + 3) / 4);
    IST_EW("image_post", (double)NB * H * W * 15.0, st,
           image_post_kernel<<<ew_grid(quads, 256), 256, 0, st>>>(x, rgb, NB, H * W, nm));
    return IST_OK;
}
inline int launch_image_prep(cudaStream_t st, const uint8_t* rgb, float* x, int NB, int H, int W, const double* mean_bgr) {
    Mean3 m;
    for (int c = 0; c < 3; ++c) m.m[c] = (float)mean_bgr[c];
    const size_t quads = (size_t)NB * (((size_t)H * W + 3) / 4);
    IST_EW("image_prep", (double)NB * H * W * 15.0, st,
           image_prep_kernel<<<ew_grid(quads, 256), 256, 0, st>>>(rgb, x, NB, H * W, m));
    return IST_OK;
}
// in [NB,Hin,Win,3] -> out [NB,Hout,Wout,3]; tmp [NB,Hin,Wout,3] is needed only when both axes change
inline int launch_image_resize(cudaStream_t st, const uint8_t* in, uint8_t* out, uint8_t* tmp, int NB, int Hin, int Win,
                               int Hout, int Wout) {
    const bool need_h = Win != Wout, need_v = Hin != Hout;
    if (!need_h && !need_v) {
        IST_CUDA(cudaMemcpyAsync(out, in, (size_t)NB * Hin * Win * 3, cudaMemcpyDeviceToDevice, st));
        return IST_OK;
    }
    if (need_h && need_v && tmp == nullptr) return fail(IST_ERR_ARG, "resize of both axes needs the [batch,Hin,Wout,3] scratch image");
    const uint8_t* src = in;
    if (need_h) {
        ResampleTable t;
        IST_TRY(resample_table(Win, Wout, &t));
        uint8_t* dst = need_v ? tmp : out;
        const size_t total = (size_t)NB * Hin * Wout * 3;
        IST_EW("image_resample_h", (double)total * (1.0 + t.ksize * 0.5), st,
               image_resample_kernel<<<ew_grid(total, 256), 256, 0, st>>>(src, dst, (size_t)NB * Hin, Win, Wout, 3, t.bounds, t.kk, t.ksize));
        src = dst;
    }
    if (need_v) {
        ResampleTable t;
        IST_TRY(resample_table(Hin, Hout, &t));
        const size_t total = (size_t)NB * Hout * Wout * 3;
        IST_EW("image_resample_v", (double)total * (1.0 + t.ksize * 0.5), st,
               image_resample_kernel<<<ew_grid(total, 256), 256, 0, st>>>(src, out, (size_t)NB, Hin, Hout, Wout * 3, t.bounds, t.kk, t.ksize));
    }
    return IST_OK;
}

}  // namespace ist

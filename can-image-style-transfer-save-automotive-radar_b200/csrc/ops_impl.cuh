// Per-op entry points of the C ABI (fp32 NCHW in/out) for unit parity against the reference modules:
// nn.Conv2d+F.relu and MaxPool2d (IST/model/meta_arch/vgg.py:13-22,52-54), GramMatrix (gram_matrix.py:6-11),
// GramMSELoss (gram_mse_loss.py:6-8), nn.MSELoss (main.py:36-37) and their autograd backward.
// They run exactly the kernels the plan runs; only the fp32 <-> plane converts and temporaries are extra.
// Temporaries are allocated and freed inside each call (these are test entry points, not the hot path).
#include "host_common.cuh"

using namespace ist;

namespace {
constexpr float kS = 0.25f;   // same activation-plane scale as the plan

int dev_ok() { return ist_device_check(); }

struct Tmp : DevMem {
    cudaStream_t st;
    explicit Tmp(cudaStream_t s) : st(s) {}
    ~Tmp() { cudaStreamSynchronize(st); }
};

int to_planes(cudaStream_t st, const float* src, uint16_t* hi, uint16_t* lo, int NB, int C, int HW, float scale, bool bf) {
    const size_t items = (size_t)NB * HW * (C / 2);
    if (bf) nchw_to_planes_kernel<true><<<ew_grid(items, 256), 256, 0, st>>>(src, hi, lo, NB, C, HW, scale);
    else nchw_to_planes_kernel<false><<<ew_grid(items, 256), 256, 0, st>>>(src, hi, lo, NB, C, HW, scale);
    IST_CUDA(cudaGetLastError());
    return IST_OK;
}
int from_planes(cudaStream_t st, const uint16_t* hi, const uint16_t* lo, float* dst, int NB, int C, int HW, float inv, bool bf) {
    const size_t items = (size_t)NB * HW * (C / 2);
    if (bf) planes_to_nchw_kernel<true><<<ew_grid(items, 256), 256, 0, st>>>(hi, lo, dst, NB, C, HW, inv);
    else planes_to_nchw_kernel<false><<<ew_grid(items, 256), 256, 0, st>>>(hi, lo, dst, NB, C, HW, inv);
    IST_CUDA(cudaGetLastError());
    return IST_OK;
}
int weight_scale(cudaStream_t st, const float* w_dev, size_t n, float* scale) {
    std::vector<float> h(n);
    IST_CUDA(cudaMemcpyAsync(h.data(), w_dev, n * sizeof(float), cudaMemcpyDeviceToHost, st));
    IST_CUDA(cudaStreamSynchronize(st));
    float mx = 0.f;
    for (float v : h) { const float a = v < 0 ? -v : v; if (a > mx) mx = a; }
    *scale = mx > 0.f ? ldexpf(1.f, 13 - ilogbf(mx)) : 1.f;
    return IST_OK;
}
}  // namespace

extern "C" {

int ist_op_conv3x3_relu_fwd(const float* x, const float* w, const float* b, float* y, int NB, int cin, int cout, int H,
                            int W, int apply_relu, void* stream) {
    IST_TRY(dev_ok());
    if (!apply_relu) return fail(IST_ERR_ARG, "the B200 conv kernels fuse bias+ReLU (vgg.py:52 always applies F.relu)");
    cudaStream_t st = (cudaStream_t)stream;
    Tmp t(st);
    const int HW = H * W;
    uint16_t *oh, *ol;
    IST_TRY(t.alloc(&oh, (size_t)NB * HW * cout));
    IST_TRY(t.alloc(&ol, (size_t)NB * HW * cout));
    if (cin == 3) {
        if (cout != 64) return fail(IST_ERR_ARG, "first-layer kernel is built for cout == 64");
        // the dispatch of the plan (plan_impl.cuh run_forward): tensor-core kernel unless IST_B200_CFF=cuda
        if (cff_use_tc() && conv_impl_halo()) {
            CUtensorMap o_hi, o_lo;
            IST_TRY(map_act(&o_hi, oh, NB, H, W, cout, 1));
            IST_TRY(map_act(&o_lo, ol, NB, H, W, cout, 1));
            IST_TRY(launch_conv_first_fwd_tc(st, o_hi, o_lo, x, w, b, NB, H, W, kS, 0));
        } else {
            conv_first_fwd_kernel<64><<<dim3((W + CFF_TX - 1) / CFF_TX, (H + CFF_TY - 1) / CFF_TY, NB), 256, 0, st>>>(x, w, b, oh, ol, NB, H, W, kS);
            IST_CUDA(cudaGetLastError());
        }
    } else {
        uint16_t *ih, *il, *fh, *fl, *dh, *dl;
        const size_t wn = (size_t)cout * cin * 9;
        IST_TRY(t.alloc(&ih, (size_t)NB * HW * cin));
        IST_TRY(t.alloc(&il, (size_t)NB * HW * cin));
        IST_TRY(t.alloc(&fh, wn)); IST_TRY(t.alloc(&fl, wn)); IST_TRY(t.alloc(&dh, wn)); IST_TRY(t.alloc(&dl, wn));
        float ws;
        IST_TRY(weight_scale(st, w, wn, &ws));
        IST_TRY(to_planes(st, x, ih, il, NB, cin, HW, kS, false));
        weight_repack_kernel<<<ew_grid(wn, 256), 256, 0, st>>>(w, cout, cin, ws, fh, fl, dh, dl);
        IST_CUDA(cudaGetLastError());
        CUtensorMap a_hi, a_lo, b_hi, b_lo, o_hi, o_lo;
        IST_TRY(map_act(&a_hi, ih, NB, H, W, cin, 9));
        IST_TRY(map_act(&a_lo, il, NB, H, W, cin, 9));
        IST_TRY(map_act(&o_hi, oh, NB, H, W, cout, 1));
        IST_TRY(map_act(&o_lo, ol, NB, H, W, cout, 1));
        IST_TRY(map_b(&b_hi, fh, 9, cout, cin, conv_b_box(cout)));
        IST_TRY(map_b(&b_lo, fl, 9, cout, cin, conv_b_box(cout)));
        ConvParams p;
        memset(&p, 0, sizeof(p));
        p.NB = NB; p.H = H; p.W = W; p.Cin = cin; p.Cout = cout; p.taps = 9; p.passes = 3; p.mode = CONV_FWD;
        p.alpha = 1.f / (kS * ws); p.bias = b; p.out_scale = kS; p.out_hi = oh; p.out_lo = ol;
        IST_TRY(launch_conv(st, a_hi, a_lo, b_hi, b_lo, p, 0, &o_hi, &o_lo));
    }
    IST_TRY(from_planes(st, oh, ol, y, NB, cout, HW, 1.f / kS, false));
    return IST_OK;
}

int ist_op_conv3x3_dgrad(const float* dy, const float* w, float* dx, int NB, int cin, int cout, int H, int W, int passes,
                         void* stream) {
    IST_TRY(dev_ok());
    cudaStream_t st = (cudaStream_t)stream;
    Tmp t(st);
    const int HW = H * W;
    uint16_t *gh, *gl;
    IST_TRY(t.alloc(&gh, (size_t)NB * HW * cout));
    IST_TRY(t.alloc(&gl, (size_t)NB * HW * cout));
    IST_TRY(to_planes(st, dy, gh, gl, NB, cout, HW, 1.f, true));
    if (cin == 3) {
        if (cout != 64) return fail(IST_ERR_ARG, "first-layer kernel is built for cout == 64");
        // the dispatch of the plan (plan_impl.cuh run_backward): tensor-core kernel unless IST_B200_CFD=cuda
        if (cfd_use_tc() && conv_impl_halo()) {
            uint16_t *wh, *wl;
            IST_TRY(t.alloc(&wh, (size_t)9 * CfdTcCfg::N_PAD * 64));
            IST_TRY(t.alloc(&wl, (size_t)9 * CfdTcCfg::N_PAD * 64));
            cfd_tc_weight_repack_kernel<<<36, 256, 0, st>>>(w, wh, wl);
            IST_CUDA(cudaGetLastError());
            CUtensorMap g_hi, g_lo, b_hi, b_lo;
            IST_TRY(map_act(&g_hi, gh, NB, H, W, cout, 9));
            IST_TRY(map_act(&g_lo, gl, NB, H, W, cout, 9));
            IST_TRY(map_b(&b_hi, wh, 9, CfdTcCfg::N_PAD, 64, CfdTcCfg::N_PAD));
            IST_TRY(map_b(&b_lo, wl, 9, CfdTcCfg::N_PAD, 64, CfdTcCfg::N_PAD));
            return launch_conv_first_dgrad_tc(st, g_hi, g_lo, b_hi, b_lo, dx, NB, H, W, false);
        }
        return launch_conv_first_dgrad(st, gh, gl, w, dx, NB, H, W);
    }
    uint16_t *fh, *fl, *dh, *dl;
    float* o32;
    const size_t wn = (size_t)cout * cin * 9;
    IST_TRY(t.alloc(&fh, wn)); IST_TRY(t.alloc(&fl, wn)); IST_TRY(t.alloc(&dh, wn)); IST_TRY(t.alloc(&dl, wn));
    IST_TRY(t.alloc(&o32, (size_t)NB * HW * cin));
    weight_repack_kernel<<<ew_grid(wn, 256), 256, 0, st>>>(w, cout, cin, 1.f, fh, fl, dh, dl);
    IST_CUDA(cudaGetLastError());
    CUtensorMap a_hi, a_lo, b_hi, b_lo;
    IST_TRY(map_act(&a_hi, gh, NB, H, W, cout, 9));
    IST_TRY(map_act(&a_lo, gl, NB, H, W, cout, 9));
    IST_TRY(map_b(&b_hi, dh, 9, cin, cout, conv_b_box(cin)));
    IST_TRY(map_b(&b_lo, dl, 9, cin, cout, conv_b_box(cin)));
    ConvParams p;
    memset(&p, 0, sizeof(p));
    p.NB = NB; p.H = H; p.W = W; p.Cin = cout; p.Cout = cin; p.taps = 9; p.passes = passes == 1 ? 1 : 3; p.mode = CONV_GRAD;
    p.alpha = 1.f; p.out_f32 = o32;
    IST_TRY(launch_conv(st, a_hi, a_lo, b_hi, b_lo, p, 1));
    nhwc_to_nchw_f32_kernel<<<ew_grid((size_t)NB * HW * cin, 256), 256, 0, st>>>(o32, dx, NB, cin, HW);
    IST_CUDA(cudaGetLastError());
    return IST_OK;
}

int ist_op_maxpool2x2_fwd(const float* x, float* y, int NB, int C, int H, int W, void* stream) {
    IST_TRY(dev_ok());
    if (C % 8 != 0) return fail(IST_ERR_ARG, "C must be a multiple of 8");
    cudaStream_t st = (cudaStream_t)stream;
    Tmp t(st);
    const int Ho = H / 2, Wo = W / 2;
    uint16_t *ih, *il, *oh, *ol;
    IST_TRY(t.alloc(&ih, (size_t)NB * H * W * C)); IST_TRY(t.alloc(&il, (size_t)NB * H * W * C));
    IST_TRY(t.alloc(&oh, (size_t)NB * Ho * Wo * C)); IST_TRY(t.alloc(&ol, (size_t)NB * Ho * Wo * C));
    IST_TRY(to_planes(st, x, ih, il, NB, C, H * W, kS, false));
    maxpool_fwd_kernel<<<ew_grid((size_t)NB * Ho * Wo * (C / 8), 256), 256, 0, st>>>(ih, il, oh, ol, NB, H, W, C, nullptr);
    IST_CUDA(cudaGetLastError());
    IST_TRY(from_planes(st, oh, ol, y, NB, C, Ho * Wo, 1.f / kS, false));
    return IST_OK;
}

static int route_op(const float* feat, const float* g_pool_nchw, const float* addend_nchw, float* dx, int NB, int C, int H,
                    int W, int mask, cudaStream_t st) {
    Tmp t(st);
    uint16_t *fh, *fl;
    float *gp = nullptr, *ad = nullptr, *o32;
    IST_TRY(t.alloc(&fh, (size_t)NB * H * W * C)); IST_TRY(t.alloc(&fl, (size_t)NB * H * W * C));
    IST_TRY(t.alloc(&o32, (size_t)NB * H * W * C));
    IST_TRY(to_planes(st, feat, fh, fl, NB, C, H * W, kS, false));
    if (g_pool_nchw != nullptr) {
        const int HWo = (H / 2) * (W / 2);
        IST_TRY(t.alloc(&gp, (size_t)NB * HWo * C));
        nchw_to_nhwc_f32_kernel<<<ew_grid((size_t)NB * HWo * C, 256), 256, 0, st>>>(g_pool_nchw, gp, NB, C, HWo);
        IST_CUDA(cudaGetLastError());
    }
    if (addend_nchw != nullptr) {
        IST_TRY(t.alloc(&ad, (size_t)NB * H * W * C));
        nchw_to_nhwc_f32_kernel<<<ew_grid((size_t)NB * H * W * C, 256), 256, 0, st>>>(addend_nchw, ad, NB, C, H * W);
        IST_CUDA(cudaGetLastError());
    }
    RouteParams r;
    memset(&r, 0, sizeof(r));
    r.NB = NB; r.H = H; r.W = W; r.C = C; r.g_pool = gp; r.f_hi = fh; r.f_lo = fl; r.addend = ad; r.apply_mask = mask;
    r.out_f32 = o32;
    const size_t items = (size_t)NB * ((H + 1) / 2) * ((W + 1) / 2) * (C / 4);
    grad_route_kernel<<<ew_grid(items, 256), 256, 0, st>>>(r);
    IST_CUDA(cudaGetLastError());
    nhwc_to_nchw_f32_kernel<<<ew_grid((size_t)NB * H * W * C, 256), 256, 0, st>>>(o32, dx, NB, C, H * W);
    IST_CUDA(cudaGetLastError());
    return IST_OK;
}

int ist_op_maxpool2x2_bwd(const float* x, const float* dy, float* dx, int NB, int C, int H, int W, void* stream) {
    IST_TRY(dev_ok());
    if (C % 8 != 0) return fail(IST_ERR_ARG, "C must be a multiple of 8");
    return route_op(x, dy, nullptr, dx, NB, C, H, W, 0, (cudaStream_t)stream);
}

int ist_op_relu_bwd(const float* yv, const float* dy, float* dx, int NB, int C, int H, int W, void* stream) {
    IST_TRY(dev_ok());
    if (C % 8 != 0) return fail(IST_ERR_ARG, "C must be a multiple of 8");
    return route_op(yv, nullptr, dy, dx, NB, C, H, W, 1, (cudaStream_t)stream);
}

// |dG| block maxima for the scaling step of gram_dmat_kernel when dG comes from the caller
__global__ void __launch_bounds__(256) absmax_blocks_kernel(const float* __restrict__ v, size_t n_per_frame, float* __restrict__ blk_max,
                                                            float* __restrict__ blk_sum) {
    __shared__ float sh[8];
    const int fr = blockIdx.y;
    float mx = 0.f;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_per_frame; i += (size_t)gridDim.x * blockDim.x)
        mx = fmaxf(mx, fabsf(v[(size_t)fr * n_per_frame + i]));
    mx = block_reduce_256<true>(mx, sh);
    if (threadIdx.x == 0) {
        blk_max[(size_t)fr * GRAM_FIN_BLOCKS + blockIdx.x] = mx;
        blk_sum[(size_t)fr * GRAM_FIN_BLOCKS + blockIdx.x] = 0.f;
    }
}

// target != nullptr: GramMSE loss + gradient; g_out != nullptr: Gram only; dg != nullptr: dx = (dG + dG^T) F / (H*W)
static int gram_common(const float* x, const float* target, float weight, float* g_out, float* loss_out, float* dx, int NB,
                       int C, int H, int W, cudaStream_t st, const float* dg = nullptr) {
    Tmp t(st);
    const int HW = H * W;
    const size_t CC = (size_t)C * C;
    uint16_t *fh, *fl, *dh, *dl;
    float *partial, *diff, *bs, *bm, *alpha, *o32;
    int splits, cps;
    gram_split_plan(NB, HW, C, &splits, &cps);
    IST_TRY(t.alloc(&fh, (size_t)NB * HW * C)); IST_TRY(t.alloc(&fl, (size_t)NB * HW * C));
    IST_TRY(t.alloc(&partial, (size_t)NB * splits * CC));
    IST_CUDA(cudaMemsetAsync(partial, 0, (size_t)NB * splits * CC * sizeof(float), st));
    IST_TRY(t.alloc(&diff, (size_t)NB * CC));
    IST_TRY(t.alloc(&bs, (size_t)NB * GRAM_FIN_BLOCKS)); IST_TRY(t.alloc(&bm, (size_t)NB * GRAM_FIN_BLOCKS));
    IST_TRY(t.alloc(&alpha, (size_t)NB));
    IST_TRY(t.alloc(&dh, (size_t)NB * CC)); IST_TRY(t.alloc(&dl, (size_t)NB * CC));
    IST_TRY(to_planes(st, x, fh, fl, NB, C, HW, kS, false));
    CUtensorMap g_hi, g_lo;
    IST_TRY(map_gram(&g_hi, fh, NB, HW, C));
    IST_TRY(map_gram(&g_lo, fl, NB, HW, C));
    if (dg == nullptr) IST_TRY(launch_gram(st, g_hi, g_lo, NB, HW, C, splits, cps, partial, 3));
    float* dummy_loss = nullptr;
    if (dg != nullptr) IST_TRY(t.alloc(&dummy_loss, (size_t)NB));
    GramFinalizeParams gp;
    memset(&gp, 0, sizeof(gp));
    gp.NB = NB; gp.n_layers = 1; gp.loss_stride = 1;
    GramLayer& L = gp.L[0];
    L.partial = partial; L.target = target; L.g_out = g_out; L.diff = diff; L.d_hi = dh; L.d_lo = dl;
    L.blk_sum = bs; L.blk_max = bm; L.alpha_out = alpha; L.loss_out = loss_out; L.C = C; L.splits = splits;
    L.g_scale = (float)(1.0 / ((double)HW * kS * kS));
    L.weight = weight;
    L.bwd_coef = (float)(2.0 * weight / ((double)C * C * HW * kS));
    dim3 grid(GRAM_FIN_BLOCKS, 1, NB);
    if (dg == nullptr) {
        gram_reduce_kernel<<<grid, 256, 0, st>>>(gp);
        IST_CUDA(cudaGetLastError());
        if (g_out != nullptr) return IST_OK;
    } else {
        IST_CUDA(cudaMemcpyAsync(diff, dg, (size_t)NB * CC * sizeof(float), cudaMemcpyDeviceToDevice, st));
        absmax_blocks_kernel<<<dim3(GRAM_FIN_BLOCKS, NB), 256, 0, st>>>(diff, CC, bm, bs);
        IST_CUDA(cudaGetLastError());
        gp.L[0].loss_out = dummy_loss;
        gp.L[0].weight = 0.f;
        gp.L[0].bwd_coef = (float)(1.0 / ((double)HW * kS));
    }
    gram_dmat_kernel<<<grid, 256, 0, st>>>(gp);
    IST_CUDA(cudaGetLastError());
    IST_TRY(t.alloc(&o32, (size_t)NB * HW * C));
    CUtensorMap a_hi, a_lo, b_hi, b_lo;
    IST_TRY(map_act(&a_hi, fh, NB, H, W, C, 1));
    IST_TRY(map_act(&a_lo, fl, NB, H, W, C, 1));
    IST_TRY(map_b(&b_hi, dh, NB, C, C, conv_b_box(C)));
    IST_TRY(map_b(&b_lo, dl, NB, C, C, conv_b_box(C)));
    ConvParams p;
    memset(&p, 0, sizeof(p));
    p.NB = NB; p.H = H; p.W = W; p.Cin = C; p.Cout = C; p.taps = 1; p.b_frame = 1; p.passes = 3; p.mode = CONV_GRAD;
    p.alpha = 1.f; p.alpha_dev = alpha; p.alpha_stride = 1; p.out_f32 = o32;
    IST_TRY(launch_conv(st, a_hi, a_lo, b_hi, b_lo, p, 0));
    nhwc_to_nchw_f32_kernel<<<ew_grid((size_t)NB * HW * C, 256), 256, 0, st>>>(o32, dx, NB, C, HW);
    IST_CUDA(cudaGetLastError());
    return IST_OK;
}

int ist_op_gram(const float* x, float* g, int NB, int C, int H, int W, void* stream) {
    IST_TRY(dev_ok());
    if (x == nullptr || g == nullptr) return fail(IST_ERR_ARG, "ist_op_gram: null argument");
    return gram_common(x, nullptr, 0.f, g, nullptr, nullptr, NB, C, H, W, (cudaStream_t)stream);
}

int ist_op_gram_bwd(const float* x, const float* dg, float* dx, int NB, int C, int H, int W, void* stream) {
    IST_TRY(dev_ok());
    if (x == nullptr || dg == nullptr || dx == nullptr) return fail(IST_ERR_ARG, "ist_op_gram_bwd: null argument");
    return gram_common(x, nullptr, 0.f, nullptr, nullptr, dx, NB, C, H, W, (cudaStream_t)stream, dg);
}

int ist_op_gram_mse(const float* x, const float* target, float weight, float* loss, float* dx, int NB, int C, int H, int W,
                    void* stream) {
    IST_TRY(dev_ok());
    if (x == nullptr || target == nullptr || loss == nullptr || dx == nullptr) return fail(IST_ERR_ARG, "ist_op_gram_mse: null argument");
    return gram_common(x, target, weight, nullptr, loss, dx, NB, C, H, W, (cudaStream_t)stream);
}

int ist_op_mse(const float* x, const float* tg, float weight, float* loss, float* dx, int NB, int C, int H, int W, void* stream) {
    IST_TRY(dev_ok());
    if (C % 8 != 0) return fail(IST_ERR_ARG, "C must be a multiple of 8");
    cudaStream_t st = (cudaStream_t)stream;
    Tmp t(st);
    const int HW = H * W;
    const size_t n = (size_t)NB * HW * C;
    uint16_t *fh, *fl, *th, *tl;
    float *part, *o32;
    IST_TRY(t.alloc(&fh, n)); IST_TRY(t.alloc(&fl, n)); IST_TRY(t.alloc(&th, n)); IST_TRY(t.alloc(&tl, n));
    IST_TRY(t.alloc(&part, (size_t)NB * 64)); IST_TRY(t.alloc(&o32, n));
    IST_TRY(to_planes(st, x, fh, fl, NB, C, HW, kS, false));
    IST_TRY(to_planes(st, tg, th, tl, NB, C, HW, kS, false));
    dim3 grid(64, NB);
    content_partial_kernel<<<grid, 256, 0, st>>>(fh, fl, th, tl, (size_t)HW * C / 8, part);
    IST_CUDA(cudaGetLastError());
    LossTotalParams lt;
    memset(&lt, 0, sizeof(lt));
    lt.losses = loss; lt.loss_stride = 2; lt.n_losses = 1; lt.NB = NB; lt.c_blocks = 64; lt.n_content = 1;
    lt.c_partial[0] = part; lt.c_slot[0] = 0;
    lt.c_scale[0] = (float)((double)weight / ((double)C * HW * kS * kS));
    loss_total_kernel<<<NB, 32, 0, st>>>(lt);
    IST_CUDA(cudaGetLastError());
    RouteParams r;
    memset(&r, 0, sizeof(r));
    r.NB = NB; r.H = H; r.W = W; r.C = C; r.f_hi = fh; r.f_lo = fl; r.t_hi = th; r.t_lo = tl;
    r.content_coef = (float)(2.0 * weight / ((double)C * HW * kS));
    r.apply_mask = 0; r.out_f32 = o32;
    const size_t items = (size_t)NB * ((H + 1) / 2) * ((W + 1) / 2) * (C / 4);
    grad_route_kernel<<<ew_grid(items, 256), 256, 0, st>>>(r);
    IST_CUDA(cudaGetLastError());
    nhwc_to_nchw_f32_kernel<<<ew_grid(n, 256), 256, 0, st>>>(o32, dx, NB, C, HW);
    IST_CUDA(cudaGetLastError());
    return IST_OK;
}

}  // extern "C"

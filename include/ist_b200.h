/* ist_b200.h — C ABI of the B200-native Gatys style-transfer hot path.
 *
 * The reference (DJNing/Can-Image-Style-Transfer-Save-Automotive-Radar, IST/) has no FFI: its boundary is the
 * Python plugin API  build_model -> meta_arch.VGG(cfg, pool) / GramMatrix / GramMSELoss / optimize().
 * Each entry point below cites the reference interface it stands in for (paths relative to the reference
 * root). The Python package `can-image-style-transfer-save-automotive-radar_b200` binds these with ctypes and
 * mirrors the reference's classes on top.
 *
 * Conventions
 *   - every pointer named *_dev is a CUDA device pointer owned by the caller (PyTorch allocator) and only
 *     borrowed for the call; fp32, contiguous, NCHW for images/features (the reference layout), [C,C] for Grams;
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream); all work is
 *     enqueued on it, nothing synchronises unless stated;
 *   - return value 0 = ok, non-zero = error (never throws across the ABI); ist_last_error() gives the text;
 *   - there is no CPU fallback: without an sm_100 device every compute entry point returns IST_ERR_DEVICE;
 *   - a plan is bound to one device and must not be used from two threads at once.
 */
#ifndef IST_B200_H_
#define IST_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IST_OK 0
#define IST_ERR_ARG 1
#define IST_ERR_CUDA 2
#define IST_ERR_DEVICE 3
#define IST_ERR_STATE 4

#define IST_LAYER_CONV3X3_RELU 0 /* nn.Conv2d(k=3,pad=1) + F.relu, IST/model/meta_arch/vgg.py:13-16,52 */
#define IST_LAYER_MAXPOOL2X2 1   /* nn.MaxPool2d(2,2),            IST/model/meta_arch/vgg.py:21-22,54 */

typedef struct ist_plan ist_plan;

typedef struct ist_layer_desc {
    int kind;     /* IST_LAYER_* */
    int cin;      /* conv only */
    int cout;     /* conv only */
} ist_layer_desc;

/* library / device ------------------------------------------------------------------------------------- */
const char* ist_last_error(void);
int ist_version(void);
/* 0 when the current CUDA device can run the kernels (compute capability 10.x), IST_ERR_DEVICE otherwise. */
int ist_device_check(void);

/* run-time kernel selection for the conv1_1 kernels (same effect as the IST_B200_CFF / IST_B200_CFD environment variables,
 * but switchable inside one process so that the parity tests cover both variants): "first_conv_fwd_tc",
 * "first_conv_dgrad_tc": 1 = tensor-core kernel (default), 0 = CUDA-core kernel. Applies to plans and per-op entry points.
 * "overlap": 1 = the loss partials of the shallow style layers run on a side stream next to the deepest conv (default),
 * 0 = everything on the caller's stream (per-launch profiling: a kernel's event time is then its own). */
int ist_set_option(const char* name, int value);

/* kernels launched by this library in this process so far (kernels inside a replayed CUDA graph are counted per replay) */
unsigned long long ist_launch_count(void);
/* per-launch profiling for benchmarks: between begin and end every kernel launched eagerly (not through a graph) is
 * bracketed by CUDA events on its stream. end() synchronises the device and returns up to max_records records:
 * names [n][40] chars, algorithmic flops / bytes per launch, elapsed milliseconds. */
int ist_profile_begin(void);
int ist_profile_end(int max_records, char* names, double* flops, double* bytes, float* ms, int* n_out);

/* plan: VGG.__init__ (IST/model/meta_arch/vgg.py:6-42) for one (batch, H, W) ------------------------------ */
/* layers[] follows cfg.MODEL.VGG.FORWARD_SEQ (IST/config/defaults.py:48-54) truncated at the deepest layer the
 * caller will ever request. The first layer must be a conv with cin == 3; all other convs need cin, cout % 64 == 0. */
int ist_plan_create(ist_plan** out, int n_layers, const ist_layer_desc* layers, int batch, int H, int W);
int ist_plan_destroy(ist_plan* plan);
/* vgg.load_state_dict (IST/main.py:30): weight OIHW [cout,cin,3,3], bias [cout]; copied and repacked, the caller
 * keeps ownership. conv_index counts conv layers only (0 = conv1_1). */
int ist_plan_set_weights(ist_plan* plan, int conv_index, const float* w_dev, const float* b_dev, void* stream);
/* bytes of device memory held by the plan */
size_t ist_plan_bytes(const ist_plan* plan);

/* forward: VGG.forward(input, out_keys) (IST/model/meta_arch/vgg.py:44-58) ------------------------------- */
/* runs layers [0, upto_layer] on x (fp32 NCHW [batch,3,H,W]); features stay inside the plan */
int ist_plan_forward(ist_plan* plan, const float* x_dev, int upto_layer, void* stream);
/* copy the output of `layer` (after ReLU / after pool) out as fp32 NCHW [batch,C,h,w] */
int ist_plan_get_feature(ist_plan* plan, int layer, float* out_dev, void* stream);
int ist_plan_feature_shape(const ist_plan* plan, int layer, int* C, int* h, int* w);
/* the decisions nn.MaxPool2d(2,2) took in the last forward (IST/model/meta_arch/vgg.py:54): for the pool layer `layer`,
 * one byte per pooled element as NCHW [batch,C,h,w]: window position 0..3 (row-major, first maximum) that max_pool2d's
 * backward routes the gradient to, or 4 where the pooled value is not positive (the ReLU below stops the gradient).
 * Together with ist_plan_get_feature (> 0 = ReLU sign map) this exposes every discontinuous decision of the forward. */
int ist_plan_get_pool_index(ist_plan* plan, int layer, uint8_t* out_dev, void* stream);
/* GramMatrix.forward (IST/model/meta_arch/gram_matrix.py:6-11) of the current features of `layer`: out [batch,C,C] */
int ist_plan_gram(ist_plan* plan, int layer, float* out_dev, void* stream);

/* loss configuration: StyleTransfer(vgg, loss_layers, loss_functions, loss_weights) (IST/main.py:35-43) ---- */
/* style layers use GramMSELoss (IST/model/meta_arch/gram_mse_loss.py:6-8), content layers nn.MSELoss
 * (IST/main.py:36-37); loss k of the output vector is style[0..n_style) then content[0..n_content). */
int ist_plan_set_loss(ist_plan* plan, int n_style, const int* style_layers, const float* style_weights,
                      int n_content, const int* content_layers, const float* content_weights);
/* targets (IST/model/engine/utils.py:19-20), cached in the plan until replaced */
int ist_plan_set_style_target(ist_plan* plan, int style_slot, const float* gram_dev /*[C,C]*/, void* stream);
/* content target of slot := current features of that layer (run ist_plan_forward on the content image first) */
int ist_plan_capture_content_target(ist_plan* plan, int content_slot, void* stream);

/* closure: IST/model/engine/utils.py:29-41 — forward, weighted layer losses, total, d(total)/dx -------------- */
/* x, grad: fp32 NCHW [batch,3,H,W]; losses: [batch, n_style+n_content+1], last entry = sum in list order.
 * Frames of a batch are independent problems (each is the reference's b = 1 case). */
int ist_plan_loss_and_grad(ist_plan* plan, const float* x_dev, float* grad_dev, float* losses_dev, void* stream);

/* generic backward for VGG.forward used under autograd with arbitrary downstream losses:
 * seeds[i] = dL/d(output of layers[i]) as fp32 NCHW; requires a preceding ist_plan_forward on the same x. */
int ist_plan_backward(ist_plan* plan, int n_seeds, const int* layers, const float* const* seeds_dev,
                      float* grad_dev, void* stream);

/* L-BFGS: torch.optim.LBFGS([x]) with the defaults the reference uses (IST/model/engine/utils.py:24,43) --- */
typedef struct ist_lbfgs ist_lbfgs;
int ist_lbfgs_create(ist_lbfgs** out, ist_plan* plan, int history_size, int max_iter, int max_eval, float lr,
                     double tolerance_grad, double tolerance_change);
int ist_lbfgs_destroy(ist_lbfgs* opt);
/* back to the state of a freshly constructed optimiser (a new `optim.LBFGS([x])` per frame, IST/model/engine/utils.py:24)
 * without re-allocating the history or re-capturing the CUDA graph */
int ist_lbfgs_reset(ist_lbfgs* opt, void* stream);
/* one optimizer.step(closure) on x (updated in place); evals_out += closure evaluations performed;
 * loss_out (host) = loss of the first closure call of this step, as LBFGS.step returns it. Synchronises once. */
int ist_lbfgs_step(ist_lbfgs* opt, float* x_dev, int* evals_out, float* loss_out, void* stream);
/* per-frame losses of the most recent closure evaluation: [batch, n_losses+1] */
int ist_lbfgs_last_losses(ist_lbfgs* opt, float* losses_host);

/* per-frame optimiser state after the last ist_lbfgs_step (frames of a batch stop independently): state['func_evals'],
 * state['n_iter'], len(old_dirs), whether the frame was still iterating when the step ended, closure evaluations this step */
int ist_lbfgs_frame_state(ist_lbfgs* opt, int frame, int* func_evals, int* n_iter, int* hist_len, int* active, int* step_evals);

/* diagnostics for the parity tests of the optimiser (torch/optim/lbfgs.py:333-537 as used by IST/model/engine/utils.py:24,43) */
/* Record every closure evaluation: x_dev/g_dev/d_dev [capacity][batch][n] receive the evaluation point, the gradient the
 * closure returned and the direction computed from it; scalars_dev [capacity][batch][16] doubles = loss, direction computed
 * (0/1), still active, n_iter, history length, ring head, pair accepted (ys > 1e-10), H_diag, t, g.d, y.s, y.y, x updated,
 * func_evals, evaluations this step, ring slot of the new pair. Call before the first step. */
int ist_lbfgs_set_trace(ist_lbfgs* opt, float* x_dev, float* g_dev, float* d_dev, double* scalars_dev, int capacity);
int ist_lbfgs_trace_count(ist_lbfgs* opt, int* count_host);
/* the same optimiser on a closed-form separable objective instead of a plan's closure — per frame
 * f(x) = sum_i 0.5 a_i (x_i - b_i)^2 + c_i cos(x_i), a/b/c [batch][n] — to drive it through rejected curvature pairs
 * and every tolerance exit against a float64 restatement of torch's algorithm */
int ist_lbfgs_create_test(ist_lbfgs** out, int batch, int n, const float* a_dev, const float* b_dev, const float* c_dev,
                          int history_size, int max_iter, int max_eval, float lr, double tolerance_grad,
                          double tolerance_change);

/* image pre/post-processing on the device ------------------------------------------------------- */
/* 8-bit images are RGB, HWC, [batch,H,W,3] uint8 on the device; network images fp32 NCHW [batch,3,H,W] in the reference's
 * preprocessed space (BGR, mean-subtracted, x255). mean_bgr = cfg.DATA.IMAGENET_MEAN (IST/config/defaults.py:86) as 3 host
 * doubles. Results are bit-identical to the reference's torchvision / PIL pipeline. */
/* output size of transforms.Scale(size) (IST/data/image_transform.py:9): smaller edge -> size, aspect kept (truncated) */
int ist_image_resize_target(int H, int W, int size, int* Hout, int* Wout);
/* ImageTransform.post_preparation (IST/data/image_transform.py:16-31): x/255 + mean, BGR->RGB, clamp [0,1], ToPILImage */
int ist_image_post_u8(const float* x_dev, uint8_t* rgb_dev, int batch, int H, int W, const double* mean_bgr, void* stream);
/* ImageTransform.preparation after the resize (image_transform.py:10-13): ToTensor, RGB->BGR, Normalize, mul_(255) */
int ist_image_prep_u8(const uint8_t* rgb_dev, float* x_dev, int batch, int H, int W, const double* mean_bgr, void* stream);
/* transforms.Scale on a PIL image (image_transform.py:9) = Image.resize(BILINEAR): Pillow's two-pass 8-bit resampling.
 * tmp_dev: scratch [batch,Hin,Wout,3] bytes, needed only when both axes change (may be NULL otherwise). */
int ist_image_resize_u8(const uint8_t* in_dev, uint8_t* out_dev, uint8_t* tmp_dev, int batch, int Hin, int Win, int Hout,
                        int Wout, void* stream);
/* coarse-to-fine hand-off of the optimised image (IST/model/engine/hr_transfer_style.py:21-27: post_preparation of the
 * low-resolution result, preparation at HRDATA.IMG_SIZE) without leaving the device: x_lo [batch,3,Hin,Win] ->
 * x_hi [batch,3,Hout,Wout]; work_dev = ist_image_handoff_workspace(...) bytes of scratch. */
size_t ist_image_handoff_workspace(int batch, int Hin, int Win, int Hout, int Wout);
int ist_image_handoff(const float* x_lo_dev, float* x_hi_dev, uint8_t* work_dev, size_t work_bytes, int batch, int Hin,
                      int Win, int Hout, int Wout, const double* mean_bgr, void* stream);

/* per-op entry points for unit parity (fp32 NCHW in/out, temporaries allocated inside) ---------------------- */
int ist_op_conv3x3_relu_fwd(const float* x_dev, const float* w_dev, const float* b_dev, float* y_dev, int batch,
                            int cin, int cout, int H, int W, int apply_relu, void* stream);
int ist_op_conv3x3_dgrad(const float* dy_dev, const float* w_dev, float* dx_dev, int batch, int cin, int cout,
                         int H, int W, int passes, void* stream);
int ist_op_maxpool2x2_fwd(const float* x_dev, float* y_dev, int batch, int C, int H, int W, void* stream);
int ist_op_maxpool2x2_bwd(const float* x_dev, const float* dy_dev, float* dx_dev, int batch, int C, int H, int W,
                          void* stream);
int ist_op_relu_bwd(const float* y_dev, const float* dy_dev, float* dx_dev, int batch, int C, int H, int W,
                    void* stream);
int ist_op_gram(const float* x_dev, float* g_dev, int batch, int C, int H, int W, void* stream);
/* backward of GramMatrix for an arbitrary upstream gradient dg [batch,C,C]: dx = (dg + dg^T) F / (H*W) */
int ist_op_gram_bwd(const float* x_dev, const float* dg_dev, float* dx_dev, int batch, int C, int H, int W, void* stream);
/* weight * mean((Gram(x) - target)^2) [loss_dev: batch floats] and its gradient w.r.t. x; target [C,C] */
int ist_op_gram_mse(const float* x_dev, const float* target_dev, float weight, float* loss_dev, float* dx_dev,
                    int batch, int C, int H, int W, void* stream);
/* weight * mean((x - t)^2) per frame [loss_dev: batch x 2 floats, column 0 = loss] and its gradient w.r.t. x */
int ist_op_mse(const float* x_dev, const float* t_dev, float weight, float* loss_dev, float* dx_dev, int batch,
               int C, int H, int W, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IST_B200_H_ */

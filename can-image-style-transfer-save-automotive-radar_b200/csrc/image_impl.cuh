// C-ABI entry points of the device image pipeline (include/ist_b200.h, "image pre/post-processing"): the work of
// ImageTransform.preparation / post_preparation (IST/data/image_transform.py:8-31) and of the coarse-to-fine hand-off
// (IST/model/engine/hr_transfer_style.py:21-27) on device buffers.
#include "image_ops.cuh"

using namespace ist;

extern "C" {

int ist_image_resize_target(int H, int W, int size, int* Hout, int* Wout) {
    if (H <= 0 || W <= 0 || size <= 0 || Hout == nullptr || Wout == nullptr) return fail(IST_ERR_ARG, "resize_target: bad arguments");
    // torchvision Resize(int): smaller edge -> size, other edge int(size * long / short)
    if (W <= H) { *Wout = size; *Hout = (int)((double)size * H / W); }
    else { *Hout = size; *Wout = (int)((double)size * W / H); }
    return IST_OK;
}

int ist_image_post_u8(const float* x_dev, uint8_t* rgb_dev, int batch, int H, int W, const double* mean_bgr, void* stream) {
    IST_TRY(ist_device_check());
    if (x_dev == nullptr || rgb_dev == nullptr || mean_bgr == nullptr || batch <= 0 || H <= 0 || W <= 0)
        return fail(IST_ERR_ARG, "image_post: bad arguments");
    return launch_image_post((cudaStream_t)stream, x_dev, rgb_dev, batch, H, W, mean_bgr);
}

int ist_image_prep_u8(const uint8_t* rgb_dev, float* x_dev, int batch, int H, int W, const double* mean_bgr, void* stream) {
    IST_TRY(ist_device_check());
    if (x_dev == nullptr || rgb_dev == nullptr || mean_bgr == nullptr || batch <= 0 || H <= 0 || W <= 0)
        return fail(IST_ERR_ARG, "image_prep: bad arguments");
    return launch_image_prep((cudaStream_t)stream, rgb_dev, x_dev, batch, H, W, mean_bgr);
}

int ist_image_resize_u8(const uint8_t* in_dev, uint8_t* out_dev, uint8_t* tmp_dev, int batch, int Hin, int Win, int Hout,
                        int Wout, void* stream) {
    IST_TRY(ist_device_check());
    if (in_dev == nullptr || out_dev == nullptr || batch <= 0 || Hin <= 0 || Win <= 0 || Hout <= 0 || Wout <= 0)
        return fail(IST_ERR_ARG, "image_resize: bad arguments");
    return launch_image_resize((cudaStream_t)stream, in_dev, out_dev, tmp_dev, batch, Hin, Win, Hout, Wout);
}

size_t ist_image_handoff_workspace(int batch, int Hin, int Win, int Hout, int Wout) {
    auto up = [](size_t b) { return (b + 255) / 256 * 256; };
    return up((size_t)batch * Hin * Win * 3) + up((size_t)batch * Hin * Wout * 3) + up((size_t)batch * Hout * Wout * 3);
}

int ist_image_handoff(const float* x_lo_dev, float* x_hi_dev, uint8_t* work_dev, size_t work_bytes, int batch, int Hin,
                      int Win, int Hout, int Wout, const double* mean_bgr, void* stream) {
    IST_TRY(ist_device_check());
    if (x_lo_dev == nullptr || x_hi_dev == nullptr || work_dev == nullptr || mean_bgr == nullptr || batch <= 0)
        return fail(IST_ERR_ARG, "image_handoff: bad arguments");
    if (work_bytes < ist_image_handoff_workspace(batch, Hin, Win, Hout, Wout))
        return fail(IST_ERR_ARG, "image_handoff: workspace of %zu bytes is too small", work_bytes);
    auto up = [](size_t b) { return (b + 255) / 256 * 256; };
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t* lo8 = work_dev;
    uint8_t* tmp = lo8 + up((size_t)batch * Hin * Win * 3);
    uint8_t* hi8 = tmp + up((size_t)batch * Hin * Wout * 3);
    IST_TRY(launch_image_post(st, x_lo_dev, lo8, batch, Hin, Win, mean_bgr));
    IST_TRY(launch_image_resize(st, lo8, hi8, tmp, batch, Hin, Win, Hout, Wout));
    return launch_image_prep(st, hi8, x_hi_dev, batch, Hout, Wout, mean_bgr);
}

}  // extern "C"

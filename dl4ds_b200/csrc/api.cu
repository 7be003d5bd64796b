// C-ABI glue: error text, device probe, argument validation and dispatch for the convolution
// family (include/dl4ds_b200.h).  Kernels live in conv_simt.cu (CUDA-core fp32) and conv_tc.cu
// (tcgen05 tf32 / 3xtf32).
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

namespace dl4ds {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return DL4DS_E_CUDA;
    }
    return DL4DS_OK;
}

}  // namespace dl4ds

using namespace dl4ds;

extern "C" {

const char* dl4ds_last_error(void) { return g_err; }
int dl4ds_version(void) { return 100; }

int64_t dl4ds_tc_launch_count(void) { return g_tc_launches.load(); }

int dl4ds_debug_set_buffer(void* dev_i64) {
    wgrad2_set_debug_buffer(reinterpret_cast<long long*>(dev_i64));
    wgrad3_set_debug_buffer(reinterpret_cast<long long*>(dev_i64));
    halo_set_debug_buffer(reinterpret_cast<long long*>(dev_i64));
    return DL4DS_OK;
}

int dl4ds_device_is_sm100(void) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
    return major == 10 ? 1 : 0;
}

int dl4ds_conv2d_fwd(const float* x, int x_ld, const float* w, const float* bias,
                     const float* res, int res_ld, float* y, int y_ld,
                     int N, int H, int W, int Cin, int Ho, int Wo, int Cout,
                     int KH, int KW, int stride, int up, int pad_t, int pad_l,
                     int wmode, int act, int d2s_r, int beta, int math_mode, void* ws, void* stream) {
    DL4DS_REQUIRE(x && w && y, DL4DS_E_BADARG, "conv2d_fwd: null pointer");
    const int prepacked = (wmode & DL4DS_W_PREPACKED) ? 1 : 0;
    wmode &= ~DL4DS_W_PREPACKED;
    DL4DS_REQUIRE(N > 0 && H > 0 && W > 0 && Cin > 0 && Ho > 0 && Wo > 0 && Cout > 0,
                  DL4DS_E_SHAPE, "conv2d_fwd: non-positive dimension");
    DL4DS_REQUIRE(KH > 0 && KW > 0 && stride > 0 && up > 0, DL4DS_E_SHAPE, "conv2d_fwd: bad kernel/stride");
    DL4DS_REQUIRE(x_ld >= Cin, DL4DS_E_SHAPE, "conv2d_fwd: x_ld < Cin");
    DL4DS_REQUIRE(wmode == DL4DS_W_HWIO || wmode == DL4DS_W_FLIP_T, DL4DS_E_BADARG, "conv2d_fwd: wmode");
    DL4DS_REQUIRE(act >= 0 && act <= 3, DL4DS_E_BADARG, "conv2d_fwd: act");
    DL4DS_REQUIRE((int64_t)N * Ho * Wo < (1ll << 31), DL4DS_E_SHAPE, "conv2d_fwd: too many pixels");
    if (d2s_r > 1) {
        DL4DS_REQUIRE(Cout % (d2s_r * d2s_r) == 0, DL4DS_E_SHAPE, "conv2d_fwd: Cout %% r^2 != 0");
        DL4DS_REQUIRE(res == nullptr && beta == 0, DL4DS_E_BADARG, "conv2d_fwd: d2s excludes res/beta");
        DL4DS_REQUIRE(y_ld >= Cout / (d2s_r * d2s_r), DL4DS_E_SHAPE, "conv2d_fwd: y_ld too small");
    } else {
        d2s_r = 1;
        DL4DS_REQUIRE(y_ld >= Cout, DL4DS_E_SHAPE, "conv2d_fwd: y_ld < Cout");
        DL4DS_REQUIRE(!res || res_ld >= Cout, DL4DS_E_SHAPE, "conv2d_fwd: res_ld < Cout");
    }
    DL4DS_REQUIRE(!(beta && act != DL4DS_ACT_NONE), DL4DS_E_BADARG, "conv2d_fwd: beta needs act NONE");
    ConvArgs a;
    a.x = x; a.w = w; a.bias = bias; a.res = res; a.y = y;
    a.x_ld = x_ld; a.res_ld = res_ld; a.y_ld = y_ld;
    a.N = N; a.H = H; a.W = W; a.Cin = Cin; a.Ho = Ho; a.Wo = Wo; a.Cout = Cout;
    a.KH = KH; a.KW = KW; a.stride = stride; a.up = up; a.pad_t = pad_t; a.pad_l = pad_l;
    a.wmode = wmode; a.act = act; a.d2s_r = d2s_r; a.beta = beta;
    a.M = N * Ho * Wo; a.HoWo = Ho * Wo;
    a.vec = (Cin % 4 == 0) && (x_ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    {   // narrow 3x3 layers (Cin, Cout in {1, 8}): HBM-bound, exact-fp32 direct kernel in every math mode
        int rc = conv2d_fwd_thin_mma(a, math_mode, st);    // 8 -> 8, tensor-core math modes: warp-level mma.sync
        if (rc != DL4DS_E_UNSUPPORTED) return rc;
        rc = conv2d_fwd_thin(a, st);
        if (rc != DL4DS_E_UNSUPPORTED) return rc;
        rc = conv2d_fwd_pointwise(a, st);
        if (rc != DL4DS_E_UNSUPPORTED) return rc;
    }
    if (math_mode != DL4DS_MATH_FP32) {
        int rc = conv2d_fwd_tc(a, math_mode, ws, prepacked, st);
        if (rc != DL4DS_E_UNSUPPORTED) return rc;
    }
    return conv2d_fwd_simt(a, st);
}

int dl4ds_conv2d_dgrad_fused_supported(int N, int H, int W, int Cq, int Cp, int KH, int KW, int math_mode) {
    static const bool disabled = [] { const char* e = getenv("DL4DS_NO_FUSED_DGRAD"); return e && e[0] == '1'; }();
    if (disabled || dl4ds_device_is_sm100() != 1) return 0;
    if (Cq == 8 && Cp == 8)      // the 8-channel HR tail: the warp-level fp16 kernel (thin_mma.cu) or not fused at all
        return conv2d_thin_fused_dgrad_supported(N, H, W, KH, KW, math_mode) ? 1 : 0;
    ConvArgs a = {};
    a.N = N; a.H = H; a.W = W; a.Cin = Cq; a.Ho = H; a.Wo = W; a.Cout = Cp;
    a.KH = KH; a.KW = KW; a.stride = 1; a.up = 1; a.d2s_r = 1;
    return (conv2d_fwd_halo_supported(a, math_mode) ||
            (math_mode == DL4DS_MATH_F16X3 && conv2d_fwd_halo_supported(a, DL4DS_MATH_TF32X3))) ? 1 : 0;
}

int dl4ds_conv2d_dgrad_fused(const float* dq, int dq_ld, const float* w, float* dz, int dz_ld,
                             const float* y_prod, int y_ld, int act, float* dbias,
                             int N, int H, int W, int Cq, int Cp, int KH, int KW, int pad_t, int pad_l,
                             int wmode, int beta, int math_mode, void* ws, void* stream) {
    DL4DS_REQUIRE(dq && w && dz, DL4DS_E_BADARG, "conv2d_dgrad_fused: null pointer");
    DL4DS_REQUIRE(act >= 0 && act <= 3, DL4DS_E_BADARG, "conv2d_dgrad_fused: act");
    DL4DS_REQUIRE(dq_ld >= Cq && dz_ld >= Cp && (!y_prod || y_ld >= Cp), DL4DS_E_SHAPE, "conv2d_dgrad_fused: pitch < channels");
    const int prepacked = (wmode & DL4DS_W_PREPACKED) ? 1 : 0;
    wmode &= ~DL4DS_W_PREPACKED;
    DL4DS_REQUIRE(wmode == DL4DS_W_FLIP_T, DL4DS_E_BADARG, "conv2d_dgrad_fused: wmode must be DL4DS_W_FLIP_T");
    if (!dl4ds_conv2d_dgrad_fused_supported(N, H, W, Cq, Cp, KH, KW, math_mode)) {
        set_error("conv2d_dgrad_fused: shape outside the halo-tile kernel's domain");
        return DL4DS_E_UNSUPPORTED;
    }
    DL4DS_REQUIRE((!y_prod || (y_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(y_prod) & 15) == 0)) &&
                  (!dbias || (reinterpret_cast<uintptr_t>(dbias) & 3) == 0), DL4DS_E_BADARG,
                  "conv2d_dgrad_fused: y_prod must be 16-byte aligned with a pitch multiple of 4");
    ConvArgs a;
    a.x = dq; a.w = w; a.bias = nullptr; a.res = nullptr; a.y = dz;
    a.x_ld = dq_ld; a.res_ld = 0; a.y_ld = dz_ld;
    a.N = N; a.H = H; a.W = W; a.Cin = Cq; a.Ho = H; a.Wo = W; a.Cout = Cp;
    a.KH = KH; a.KW = KW; a.stride = 1; a.up = 1; a.pad_t = pad_t; a.pad_l = pad_l;
    a.wmode = wmode; a.act = DL4DS_ACT_NONE; a.d2s_r = 1; a.beta = beta;
    a.M = N * H * W; a.HoWo = H * W;
    a.vec = 1;
    a.mask_y = y_prod; a.mask_ld = y_ld; a.mask_act = act; a.dbias = dbias;
    const int rc = (Cq == 8 && Cp == 8) ? conv2d_fwd_thin_mma(a, math_mode, reinterpret_cast<cudaStream_t>(stream))
                                        : conv2d_fwd_tc(a, math_mode, ws, prepacked, reinterpret_cast<cudaStream_t>(stream));
    if (rc == DL4DS_E_UNSUPPORTED) set_error("conv2d_dgrad_fused: tensors not aligned for the tensor-core kernel");
    return rc;
}

int64_t dl4ds_conv2d_fwd_workspace_bytes(int N, int H, int W, int Cin, int Ho, int Wo, int Cout,
                                         int KH, int KW, int stride, int up, int d2s_r, int math_mode) {
    if (math_mode == DL4DS_MATH_FP32) return 0;
    ConvArgs a = {};
    a.N = N; a.H = H; a.W = W; a.Cin = Cin; a.Ho = Ho; a.Wo = Wo; a.Cout = Cout;
    a.KH = KH; a.KW = KW; a.stride = stride; a.up = up; a.d2s_r = d2s_r > 1 ? d2s_r : 1;
    return conv2d_fwd_tc_workspace(a, math_mode);
}

int dl4ds_conv2d_pack(const float* w, int wmode, int KH, int KW, int Cin, int Cout, int math_mode,
                      void* ws, void* stream) {
    DL4DS_REQUIRE(w && ws, DL4DS_E_BADARG, "conv2d_pack: null pointer");
    DL4DS_REQUIRE(math_mode == DL4DS_MATH_TF32 || math_mode == DL4DS_MATH_TF32X3 || math_mode == DL4DS_MATH_F16X3,
                  DL4DS_E_BADARG, "conv2d_pack: math_mode must be a tensor-core mode");
    DL4DS_REQUIRE(Cin % 8 == 0 && Cout % 8 == 0, DL4DS_E_SHAPE, "conv2d_pack: channels must be multiples of 8");
    return conv2d_pack_tc(w, wmode & ~DL4DS_W_PREPACKED, KH, KW, Cin, Cout, math_mode, ws,
                          reinterpret_cast<cudaStream_t>(stream));
}

int64_t dl4ds_conv2d_pack_desc(const float* w, int wmode, int KH, int KW, int Cin, int Cout, int math_mode,
                               void* ws, void* desc_out_host) {
    if (!w || !ws || !desc_out_host ||
        (math_mode != DL4DS_MATH_TF32 && math_mode != DL4DS_MATH_TF32X3 && math_mode != DL4DS_MATH_F16X3) || Cin % 8 ||
        Cout % 8) {
        set_error("conv2d_pack_desc: bad argument");
        return DL4DS_E_BADARG;
    }
    return conv2d_pack_desc(w, wmode & ~DL4DS_W_PREPACKED, KH, KW, Cin, Cout, math_mode, ws, desc_out_host);
}

int dl4ds_conv2d_pack_multi(const void* descs_dev, int n, int64_t total_units, void* stream) {
    DL4DS_REQUIRE(descs_dev || n == 0, DL4DS_E_BADARG, "conv2d_pack_multi: null descriptor table");
    return conv2d_pack_multi(descs_dev, n, total_units, reinterpret_cast<cudaStream_t>(stream));
}

int64_t dl4ds_conv2d_wgrad_workspace_bytes(int N, int Hq, int Wq, int Ca, int Cb, int KH, int KW,
                                           int math_mode) {
    if (math_mode == DL4DS_MATH_FP32) return 0;
    return conv2d_wgrad_tc_workspace(N, Hq, Wq, Ca, Cb, KH, KW);
}

int dl4ds_conv2d_wgrad(const float* P, int p_ld, const float* Q, int q_ld, float* dw,
                       int N, int Hp, int Wp, int Ca, int Hq, int Wq, int Cb,
                       int KH, int KW, int stride, int pad_t, int pad_l,
                       void* ws, int math_mode, void* stream) {
    DL4DS_REQUIRE(P && Q && dw, DL4DS_E_BADARG, "conv2d_wgrad: null pointer");
    DL4DS_REQUIRE(N > 0 && Hp > 0 && Wp > 0 && Ca > 0 && Hq > 0 && Wq > 0 && Cb > 0 && KH > 0 &&
                  KW > 0 && stride > 0, DL4DS_E_SHAPE, "conv2d_wgrad: non-positive dimension");
    DL4DS_REQUIRE(p_ld >= Ca && q_ld >= Cb, DL4DS_E_SHAPE, "conv2d_wgrad: pitch < channels");
    WgradArgs a;
    a.P = P; a.Q = Q; a.dw = dw; a.p_ld = p_ld; a.q_ld = q_ld;
    a.N = N; a.Hp = Hp; a.Wp = Wp; a.Ca = Ca; a.Hq = Hq; a.Wq = Wq; a.Cb = Cb;
    a.KH = KH; a.KW = KW; a.stride = stride; a.pad_t = pad_t; a.pad_l = pad_l;
    if (math_mode == DL4DS_MATH_F16X3) math_mode = DL4DS_MATH_TF32X3;       // weight gradients: the 3xTF32 kernels
    a.Mw = KH * KW * Ca;
    a.NQ = (int64_t)N * Hq * Wq;
    a.chunks_per_split = 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    {   // narrow layers (Ca, Cb in {1, 8}, 3x3): exact-fp32 sliding-window kernel in every math mode
        int rc = conv2d_wgrad_thin_mma(a, math_mode, st);   // 8 x 8, tensor-core math modes: warp-level mma.sync
        if (rc != DL4DS_E_UNSUPPORTED) return rc;
        rc = conv2d_wgrad_thin(a, st);
        if (rc != DL4DS_E_UNSUPPORTED) return rc;
        rc = conv2d_wgrad_pointwise(a, st);        // 1x1, Ca*Cb <= 1024: streaming CUDA-core kernel
        if (rc != DL4DS_E_UNSUPPORTED) return rc;
    }
    if (math_mode != DL4DS_MATH_FP32) {
        int rc = conv2d_wgrad_tc3(a, math_mode, st);
        if (rc != DL4DS_E_UNSUPPORTED) return rc;
        rc = conv2d_wgrad_tc2(a, math_mode, st);
        if (rc != DL4DS_E_UNSUPPORTED) return rc;
        rc = conv2d_wgrad_tc(a, ws, math_mode, st);
        if (rc != DL4DS_E_UNSUPPORTED) return rc;
    }
    return conv2d_wgrad_simt(a, st);
}

}  // extern "C"

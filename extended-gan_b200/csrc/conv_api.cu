// C-ABI entry points of the convolution kernels: argument checks and the direct / tcgen05 switch.
#include "common.cuh"

namespace cgat {
int validate_conv(const cgat_conv_desc* d);
int conv_fprop_direct_launch(const cgat_conv_desc*, const void*, const void*, const float*, void*, cudaStream_t);
int conv_dgrad_direct_launch(const cgat_conv_desc*, const void*, const void*, void*, cudaStream_t);
int conv_wgrad_direct_launch(const cgat_conv_desc*, const void*, const void*, float*, float*, cudaStream_t);
int conv_tc_supported(const cgat_conv_desc* d, int which);
size_t conv_tc_workspace(const cgat_conv_desc* d, int which);
int conv_fprop_tc_launch(const cgat_conv_desc*, const void*, const void*, const float*, void*, void*, cudaStream_t);
int conv_dgrad_tc_launch(const cgat_conv_desc*, const void*, const void*, void*, void*, cudaStream_t);
int conv_wgrad_tc_launch(const cgat_conv_desc*, const void*, const void*, float*, float*, void*, cudaStream_t,
                         int* ncta_out = nullptr, int* nt_out = nullptr);
int conv_tc_packed_launch(const cgat_conv_desc*, int dgrad, const void*, const void*, const float*, void*, cudaStream_t);
void set_debug_buffer(long long* p);
int conv_is_pointwise(const cgat_conv_desc* d);
int conv_pointwise_launch(int which, const cgat_conv_desc*, const void*, const void*, void*, const float*, cudaStream_t);
int conv_dbias_launch(const cgat_conv_desc*, const void* dy, float* dbias, cudaStream_t);
int conv_big_supported(const cgat_conv_desc* d, int which);
size_t conv_big_workspace(const cgat_conv_desc* d, int which);
int conv_big_fprop_launch(const cgat_conv_desc*, const void*, const void*, const float*, void*, cudaStream_t);
int conv_big_dgrad_launch(const cgat_conv_desc*, const void*, const void*, void*, void*, cudaStream_t);
int conv_big_wgrad_launch(const cgat_conv_desc*, const void*, const void*, float*, void*, cudaStream_t);
int conv_big_wgrad_small_ok(const cgat_conv_desc* d);
int conv_is_fullwindow(const cgat_conv_desc* d);
int conv_fullwindow_fprop_launch(const cgat_conv_desc*, const void*, const void*, const float*, void*, cudaStream_t);
int conv_dw3x3_served(const cgat_conv_desc* d);
int conv_dw3x3_launch(int which, const cgat_conv_desc*, const void*, const void*, void*, float*, const float*, cudaStream_t);
int conv_wgrad_small_served(const cgat_conv_desc* d);
int conv_wgrad_small_launch(const cgat_conv_desc*, const void*, const void*, float*, float*, cudaStream_t);
int conv_dbias_ws_launch(const cgat_conv_desc*, const void*, float*, void*, cudaStream_t);
size_t conv_dbias_workspace(const cgat_conv_desc* d);
int conv_small_served(const cgat_conv_desc* d, int which);
int conv_small_launch(int which, const cgat_conv_desc*, const void*, const void*, const float*, void*, cudaStream_t);
int conv_gemm_served(const cgat_conv_desc* d);
int conv_gemm_launch(int which, const cgat_conv_desc*, const void*, const void*, void*, const float*, cudaStream_t);
}  // namespace cgat

using namespace cgat;

// developer aid, not part of the public header: registers a device buffer for the kernel timeline
extern "C" void cgat_debug_timeline(long long* device_buffer) { set_debug_buffer(device_buffer); }

static const char* kNames[3] = {"fprop", "dgrad", "wgrad"};

// the streamed-operand kernels (conv_tc_big.cu) take every shape with >= 64 channels on the axes they tile;
// the resident-weight kernels (conv_tc.cu) keep the small-channel convs
// (wgrad also falls to them for small-channel shapes the resident kernel does not serve)
static int use_big(const cgat_conv_desc* d, int which) {
  if (conv_big_supported(d, which)) return 1;
  // ... unless the CUDA-core small-channel wgrad (conv_wgrad_small.cu) serves them better
  return which == 2 && !conv_tc_supported(d, 2) && conv_big_wgrad_small_ok(d) && !conv_wgrad_small_served(d);
}

static int tc_ready(const cgat_conv_desc* d, int which, void* workspace) {
  if (!conv_tc_supported(d, which)) return fail(CGAT_EUNSUPPORTED, "tcgen05 %s does not support this conv shape", kNames[which]);
  if (conv_tc_workspace(d, which) > 0 && !workspace)
    return fail(CGAT_EINVAL, "tcgen05 %s needs a workspace of %zu bytes", kNames[which], conv_tc_workspace(d, which));
  return 0;
}

extern "C" int cgat_conv_tc_supported(const cgat_conv_desc* d, int which) {
  if (validate_conv(d) || which < 0 || which > 2) return 0;
  return use_big(d, which) || conv_tc_supported(d, which);
}

extern "C" int64_t cgat_conv_dbias_workspace_bytes(const cgat_conv_desc* d) {
  return validate_conv(d) ? 0 : (int64_t)conv_dbias_workspace(d);
}

extern "C" int64_t cgat_conv_workspace_bytes(const cgat_conv_desc* d, int which) {
  if (validate_conv(d) || which < 0 || which > 2) return 0;
  if (use_big(d, which)) return (int64_t)conv_big_workspace(d, which);
  return (int64_t)conv_tc_workspace(d, which);
}

extern "C" int cgat_conv2d_fprop(const cgat_conv_desc* d, const void* x, const void* w, const float* bias, void* y,
                                 int impl, void* workspace, void* stream) {
  if (int rc = validate_conv(d)) return rc;
  if (!x || !w || !y) return fail(CGAT_EINVAL, "null x/w/y");
  if (impl == 0 && conv_dw3x3_served(d)) return conv_dw3x3_launch(0, d, x, w, y, nullptr, bias, (cudaStream_t)stream);
  if (impl == 0 && conv_is_fullwindow(d)) return conv_fullwindow_fprop_launch(d, x, w, bias, y, (cudaStream_t)stream);
  if (impl == 0 && conv_is_pointwise(d)) return conv_pointwise_launch(0, d, x, w, y, bias, (cudaStream_t)stream);
  if (impl == 0 && conv_small_served(d, 0)) return conv_small_launch(0, d, x, w, bias, y, (cudaStream_t)stream);
  if (impl == 0 && conv_gemm_served(d)) return conv_gemm_launch(0, d, x, w, y, bias, (cudaStream_t)stream);
  if (impl == 0) return conv_fprop_direct_launch(d, x, w, bias, y, (cudaStream_t)stream);
  if (use_big(d, 0)) return conv_big_fprop_launch(d, x, w, bias, y, (cudaStream_t)stream);
  if (int rc = tc_ready(d, 0, workspace)) return rc;
  return conv_fprop_tc_launch(d, x, w, bias, y, workspace, (cudaStream_t)stream);
}

extern "C" int cgat_conv2d_dgrad(const cgat_conv_desc* d, const void* dy, const void* w, void* dx, int impl,
                                 void* workspace, void* stream) {
  if (int rc = validate_conv(d)) return rc;
  if (!dy || !w || !dx) return fail(CGAT_EINVAL, "null dy/w/dx");
  if (impl == 0 && conv_dw3x3_served(d)) return conv_dw3x3_launch(1, d, dy, w, dx, nullptr, nullptr, (cudaStream_t)stream);
  if (impl == 0 && conv_is_pointwise(d)) return conv_pointwise_launch(1, d, dy, w, dx, nullptr, (cudaStream_t)stream);
  if (impl == 0 && conv_small_served(d, 1)) return conv_small_launch(1, d, dy, w, nullptr, dx, (cudaStream_t)stream);
  if (impl == 0 && conv_gemm_served(d)) return conv_gemm_launch(1, d, dy, w, dx, nullptr, (cudaStream_t)stream);
  if (impl == 0) return conv_dgrad_direct_launch(d, dy, w, dx, (cudaStream_t)stream);
  if (use_big(d, 1)) return conv_big_dgrad_launch(d, dy, w, dx, workspace, (cudaStream_t)stream);
  if (int rc = tc_ready(d, 1, workspace)) return rc;
  return conv_dgrad_tc_launch(d, dy, w, dx, workspace, (cudaStream_t)stream);
}

extern "C" int cgat_conv2d_wgrad(const cgat_conv_desc* d, const void* x, const void* dy, float* dw, float* dbias,
                                 int impl, void* workspace, void* stream) {
  if (int rc = validate_conv(d)) return rc;
  if (!x || !dy || !dw) return fail(CGAT_EINVAL, "null x/dy/dw");
  if (impl == 0 && conv_dw3x3_served(d)) return conv_dw3x3_launch(2, d, x, dy, dw, dbias, nullptr, (cudaStream_t)stream);
  if (impl == 0 && conv_wgrad_small_served(d)) return conv_wgrad_small_launch(d, x, dy, dw, dbias, (cudaStream_t)stream);
  if (impl == 0 && conv_is_pointwise(d)) {
    if (int rc = conv_pointwise_launch(2, d, dy, x, dw, nullptr, (cudaStream_t)stream)) return rc;
    return dbias ? conv_dbias_ws_launch(d, dy, dbias, workspace, (cudaStream_t)stream) : 0;  // (workspace optional here)
  }
  if (impl == 0 && conv_gemm_served(d)) {
    if (int rc = conv_gemm_launch(2, d, dy, x, dw, nullptr, (cudaStream_t)stream)) return rc;
    return dbias ? conv_dbias_ws_launch(d, dy, dbias, workspace, (cudaStream_t)stream) : 0;  // (workspace optional here)
  }
  if (impl == 0) return conv_wgrad_direct_launch(d, x, dy, dw, dbias, (cudaStream_t)stream);
  if (use_big(d, 2)) {
    if (int rc = conv_big_wgrad_launch(d, x, dy, dw, workspace, (cudaStream_t)stream)) return rc;
    return dbias ? conv_dbias_ws_launch(d, dy, dbias, workspace, (cudaStream_t)stream) : 0;
  }
  if (int rc = tc_ready(d, 2, workspace)) return rc;
  return conv_wgrad_tc_launch(d, x, dy, dw, dbias, workspace, (cudaStream_t)stream);
}

// The one-launch conv-GAT stream path (cgat_conv2d_fprop_packed / _wgrad_partial / _dgrad_packed) runs on the
// RESIDENT-weight kernels only: cgat_conv_tc_supported also answers 1 for shapes only the streamed kernels serve.
extern "C" int cgat_conv_stream_supported(const cgat_conv_desc* d, int need_dx) {
  if (validate_conv(d)) return 0;
  return conv_tc_supported(d, 0) && conv_tc_supported(d, 2) && (!need_dx || conv_tc_supported(d, 1));
}

extern "C" int64_t cgat_conv_stream_workspace_bytes(const cgat_conv_desc* d) {
  if (validate_conv(d) || !conv_tc_supported(d, 2)) return 0;
  return (int64_t)conv_tc_workspace(d, 2);
}

extern "C" int cgat_conv2d_fprop_packed(const cgat_conv_desc* d, const void* x, const void* wpack, const float* bias,
                                        void* y, void* stream) {
  if (int rc = validate_conv(d)) return rc;
  if (!x || !wpack || !y) return fail(CGAT_EINVAL, "null x/wpack/y");
  if (!conv_tc_supported(d, 0)) return fail(CGAT_EUNSUPPORTED, "tcgen05 fprop does not support this conv shape");
  return conv_tc_packed_launch(d, 0, x, wpack, bias, y, (cudaStream_t)stream);
}

extern "C" int cgat_conv2d_dgrad_packed(const cgat_conv_desc* d, const void* dy, const void* wpack, void* dx,
                                        void* stream) {
  if (int rc = validate_conv(d)) return rc;
  if (!dy || !wpack || !dx) return fail(CGAT_EINVAL, "null dy/wpack/dx");
  if (!conv_tc_supported(d, 1)) return fail(CGAT_EUNSUPPORTED, "tcgen05 dgrad does not support this conv shape");
  return conv_tc_packed_launch(d, 1, dy, wpack, nullptr, dx, (cudaStream_t)stream);
}

extern "C" int cgat_conv2d_wgrad_partial(const cgat_conv_desc* d, const void* x, const void* dy, void* workspace,
                                         int32_t* ncta_out, int32_t* nt_out, void* stream) {
  if (int rc = validate_conv(d)) return rc;
  if (!x || !dy || !workspace || !ncta_out || !nt_out) return fail(CGAT_EINVAL, "null argument");
  if (int rc = tc_ready(d, 2, workspace)) return rc;
  int ncta = 0, nt = 0;
  const int rc = conv_wgrad_tc_launch(d, x, dy, nullptr, nullptr, workspace, (cudaStream_t)stream, &ncta, &nt);
  *ncta_out = ncta;
  *nt_out = nt;
  return rc;
}

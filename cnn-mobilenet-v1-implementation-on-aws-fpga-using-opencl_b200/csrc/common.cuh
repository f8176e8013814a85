// common.cuh — internal types shared by the sm_100a kernels and the C-ABI layer.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <utility>
#include <vector>

#include "../../include/mnv1.h"

typedef __nv_bfloat16 bf16;

struct mnv1_buf {
  void* d = nullptr;     // device storage
  size_t bytes = 0;
  int n = 0, c = 0, h = 0, w = 0;  // feature map (NHWC on the device); c==0 -> raw bytes
  bool is_u8 = false;
  bool owned = true;
};

struct mnv1_filter {
  mnv1_kind kind;
  int cin = 0, cout = 0;
  mnv1_act act = MNV1_ACT_NONE;
  float* w_f32 = nullptr;   // kernel-native fp32 copy (stem [27][Cout], dw [9][C], pw/fc [Cout][Cin])
  float* w_scaled = nullptr; // depthwise: [9][C] taps pre-multiplied by `scale` (TMA path)
  bf16* w_bf16 = nullptr;   // bf16 [Cout][Cin] (pointwise / fc, bf16 contexts)
  float* scale = nullptr;   // [Cout] or nullptr
  float* shift = nullptr;   // [Cout] or nullptr
  // stem on tensor cores (bf16 contexts): fp16 filter bank x input scale and the folded shift are
  // rebuilt whenever the context's input transform changes (stem_tc.cu)
  std::vector<float> h_w, h_scale, h_shift;
  __half* wq = nullptr;
  float* shift2 = nullptr;
  float h_shift2[32] = {};  // host copy of shift2 (constant-bank epilogue of stem_rows.cu)
  float prep_scale = 0.f, prep_bias = 0.f, p0 = 0.f;
  bool prepared = false;
  // integer contexts (MNV1_U8, int8.cu): s8 filter in the kernel's order (stem: packed [32][7] words in w_q32,
  // depthwise [9][C], pointwise / fc [Cout][Cin]), s32 bias, right shift of the requantisation
  int8_t* w_s8 = nullptr;
  int* w_q32 = nullptr;
  int* bias_i32 = nullptr;
  int rshift = 0;
  std::vector<int> h_q32, h_bias;   // stem: host copies (the integer stem reads them from the constant bank)
  CUtensorMap tmap_b;       // TMA descriptor of w_bf16 (pointwise, bf16 contexts)
  bool has_tmap = false;
  int tmap_bn = 0;          // N-tile the descriptor's box was built for
};

struct Epilogue {
  const float* scale;  // may be nullptr
  const float* shift;  // may be nullptr
  int act;
};

__device__ __forceinline__ float apply_epilogue(float acc, float s, float t, int act) {
  float y = fmaf(acc, s, t);
  if (act != MNV1_ACT_NONE) y = fmaxf(y, 0.0f);
  if (act == MNV1_ACT_RELU6) y = fminf(y, 6.0f);
  return y;
}

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&p);
}
__device__ __forceinline__ float bf16lo_to_f32(uint32_t p) { return __uint_as_float(p << 16); }
__device__ __forceinline__ float bf16hi_to_f32(uint32_t p) { return __uint_as_float(p & 0xffff0000u); }

// Pipeline stamps and experiment switches inside the kernels exist only in the MNV1_TRACE build
// (`make trace` -> libmnv1_trace.so, read by tools/*_trace.py); the shipped library compiles them out.
#ifdef MNV1_TRACE
#define MNV1_DBG(x) (x)
#define MNV1_TRC(x) (x)
#else
#define MNV1_DBG(x) 0
#define MNV1_TRC(x) ((unsigned long long*)nullptr)
#endif

// ---- launchers implemented in the .cu files (all asynchronous on `st`) -------------------
namespace mnv1 {

// Programmatic dependent launch: consecutive kernels of the layer schedule are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization.  Every kernel triggers its dependents at entry
// (pdl_trigger) and waits for its predecessor (pdl_wait: full completion + memory flush) only after its
// own prologue — barrier init, TMEM allocation, tensor-map prefetch, constants — so the launch latency
// and the prologue of layer k+1 hide under the tail of layer k.  MNV1_NO_PDL=1 turns the attribute off.
bool pdl_enabled();
// Environment switches (fall back to the previous kernel of a layer; timing experiments): read ONCE per
// process, by the first mnv1_ctx_create — never from a launch path.
struct Switches { bool no_pdl, no_pair, no_cw, no_stem_rows, fused_head, fused_pair, pp_direct, no_pp_tail, h2d_wc; long rb_mask; };
const Switches& switches();
// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute: remembered per (function,
// current device), thread-safe, so that a second context on another GPU of the same process opts in too.
cudaError_t ensure_dyn_smem(const void* fn, int bytes);
inline bool capturing(cudaStream_t st) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  return cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif

struct StemArgs {
  const uint8_t *r, *g, *b;  // plane base pointers (interleaved: g = r+1, b = r+2)
  int pix_stride;            // 1 planar, 3 interleaved
  long img_stride;           // bytes between images for each pointer
  int n, rows, cols, stride, cout, pad_lo;
  float in_scale, in_bias;
};
cudaError_t launch_stem(mnv1_dtype dt, void* out, const StemArgs& a, const float* w27xC,
                        Epilogue ep, cudaStream_t st);
cudaError_t stem_tc_prepare(const float* w_oihw_host, const float* scale_host, const float* shift_host,
                            float in_scale, float in_bias, __half* wq_dev, float* shift2_dev, float* p0_out,
                            float* shift2_host_out = nullptr);
cudaError_t launch_stem_rows(bf16* out, const StemArgs& a, const __half* wq_dev, const float* scale_host,
                             const float* shift2_host, float p0, int act, int num_sms, cudaStream_t st,
                             std::string* err);
cudaError_t launch_stem_tc(bf16* out, const StemArgs& a, const __half* wq_dev, const float* scale_dev,
                           const float* shift2_dev, float p0, int act, int num_sms, cudaStream_t st,
                           std::string* err);
cudaError_t launch_depthwise(mnv1_dtype dt, void* out, const void* in, const float* w9xC, int n,
                             int rows, int cols, int stride, int c, int pad_lo, Epilogue ep,
                             cudaStream_t st);
// TMA-staged depthwise (bf16); cudaErrorNotSupported = no variant for this shape, nothing launched
cudaError_t launch_depthwise_tma(bf16* out, const bf16* in, const float* w9xC_scaled, const float* shift, int act,
                                 int n, int rows, int cols, int stride, int c, int pad_lo, int num_sms,
                                 cudaStream_t st, std::string* err);
// small many-channel maps (28x28x256 and below): group-ring stencil, depthwise_ring.cu
cudaError_t launch_depthwise_ring(bf16* out, const bf16* in, const float* w9xC_scaled, const float* shift, int act, int n,
                                  int rows, int cols, int stride, int c, int pad_lo, int num_sms, cudaStream_t st,
                                  std::string* err);
// 14x14 / 7x7 maps: register-window stencil, one thread per channel pair (depthwise_cw.cu)
cudaError_t launch_depthwise_cw(bf16* out, const bf16* in, const float* w9xC_scaled, const float* shift, int act, int n,
                                int rows, int cols, int stride, int c, int pad_lo, cudaStream_t st);
// generic SIMT 1x1 conv / FC: out[M][Cout] = in[M][K] * w[Cout][K]^T  (fp32 contexts, FC)
cudaError_t launch_pointwise_simt(mnv1_dtype dt, void* out, const void* in, const float* w_f32,
                                  const bf16* w_bf16, long m, int k, int cout, Epilogue ep,
                                  bool out_f32, cudaStream_t st);
// tcgen05 / TMEM / TMA 1x1 conv (bf16 contexts)
cudaError_t launch_pointwise_tc(bf16* out, const bf16* in, const mnv1_filter* f, long m, int k,
                                int cout, int num_sms, cudaStream_t st, std::string* err);
cudaError_t make_weight_tmap(mnv1_filter* f, std::string* err);
// CTA-pair tcgen05 (cta_group::2) 1x1 conv for Cout % 256 == 0, K >= 256 (pointwise_pair.cu)
cudaError_t launch_pointwise_pair(bf16* out, const bf16* in, const mnv1_filter* f, long m, int k, int cout, int num_sms,
                                  cudaStream_t st, std::string* err);
// fused depthwise -> pointwise block (bf16, fused_rb.cu); cudaErrorNotSupported = no variant, nothing launched
cudaError_t launch_fused_dw_pw(bf16* out, const bf16* in, const mnv1_filter* dw, const mnv1_filter* pw, int n,
                               int rows, int cols, int stride, int pad_lo, int num_sms, cudaStream_t st,
                               std::string* err);
bool fused_dw_pw_supported(const mnv1_filter* dw, const mnv1_filter* pw, int rows, int cols, int stride);
// fused depthwise -> pointwise on CTA pairs with a streamed filter (the 512-channel 14x14 blocks, fused_pair.cu)
cudaError_t launch_fused_pair(bf16* out, const bf16* in, const mnv1_filter* dw, const mnv1_filter* pw, int n, int rows, int cols,
                              int stride, int pad_lo, int num_sms, cudaStream_t st, std::string* err);
bool fused_pair_supported(const mnv1_filter* dw, const mnv1_filter* pw, int rows, int cols, int stride);
cudaError_t launch_pool(mnv1_dtype dt, void* out, const void* in, int n, int hw, int c, bool out_f32,
                        cudaStream_t st);
// Logits gather without a collective kernel: the head kernel also stores every image's logits / top-1 /
// top-1 probability into the gather buffers of up to 8 GPUs (own buffer included) at row `row0 + image`.
// The pointers are peer-mapped (cudaDeviceEnablePeerAccess inside one process, CUDA IPC across processes).
struct HeadGather {
  float* logits[8];
  int* top1[8];
  float* prob[8];
  int n_dst;
  long row0;
};
// head: global average pool -> FC (+bias) -> softmax -> argmax (one cluster kernel on bf16 contexts)
cudaError_t launch_head(mnv1_dtype dt, const void* in, int n, int hw, int c, const mnv1_filter* fc,
                        float* pooled_scratch, float* logits, int* top1, float* top1_prob,
                        int classes, const HeadGather& gather, cudaStream_t st, int* launches);
struct HeadGather;
cudaError_t launch_softmax(const float* logits, int n, int classes, float* prob, int* top1,
                           float* top1_prob, cudaStream_t st, const HeadGather* gather = nullptr);
cudaError_t launch_nchw_to_nhwc(mnv1_dtype dt, void* out_nhwc, const float* in_nchw, int n, int c,
                                int h, int w, cudaStream_t st);
cudaError_t launch_nhwc_to_nchw(mnv1_dtype dt, float* out_nchw, const void* in_nhwc, int n, int c,
                                int h, int w, cudaStream_t st);
cudaError_t launch_synth_images(uint8_t* out, long first_byte, long nbytes, uint64_t seed,
                                cudaStream_t st);
// ---- integer contexts (int8.cu): u8 activations x s8 filters -> s32 -> u8
cudaError_t launch_pointwise_i8(uint8_t* out, const uint8_t* in, const mnv1_filter* f, long m, int k, int cout, int wrap,
                                int num_sms, cudaStream_t st, std::string* err);
cudaError_t launch_depthwise_u8(uint8_t* out, const uint8_t* in, const mnv1_filter* f, int n, int rows, int cols, int stride,
                                int c, int pad_lo, int wrap, cudaStream_t st);
cudaError_t launch_stem_u8(uint8_t* out, const StemArgs& a, const mnv1_filter* f, int wrap, int num_sms, cudaStream_t st,
                           const char** kernel_name = nullptr);
cudaError_t launch_pool_u8(uint8_t* out, const uint8_t* in, int n, int hw, int c, int wrap, cudaStream_t st);
// dir 0: planar host order (u8 or float values) -> NHWC u8; dir 1: NHWC u8 -> planar (u8 or float)
cudaError_t launch_u8_layout(int dir, void* out, const void* in, bool host_is_u8, int n, int c, int hw, cudaStream_t st);
cudaError_t launch_u8_to_f32(float* out, const uint8_t* in, long count, cudaStream_t st);
}  // namespace mnv1

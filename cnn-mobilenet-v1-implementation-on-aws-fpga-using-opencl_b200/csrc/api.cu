// api.cu — the C-ABI of include/mnv1.h: context, buffers, filters, the four kernel entry
// points with kernel.cl's argument order, and the whole-network executor (activation arena +
// CUDA graph) that replaces the 29 copy-pasted layer blocks of MobileNet.c:207-2763.
#include <algorithm>
#include <chrono>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <tuple>

#include "common.cuh"

namespace mnv1 {
int load_weight_file(const char* path, std::vector<float>* weights, std::vector<float>* scale,
                     std::vector<float>* shift, std::string* err);
int save_weight_file_bin(const char* path, const float* weights, const float* scale, const float* shift,
                         std::string* err);
int read_ppm(const char* path, uint8_t* out, int height, int width, std::string* err);
}  // namespace mnv1

static thread_local std::string g_last_error;

const mnv1::Switches& mnv1::switches() {
  static Switches sw;
  static std::once_flag once;
  std::call_once(once, [] {
    sw.no_pdl = getenv("MNV1_NO_PDL") != nullptr;
    sw.no_pair = getenv("MNV1_NO_PAIR") != nullptr;
    sw.no_cw = getenv("MNV1_NO_CW") != nullptr;
    sw.no_stem_rows = getenv("MNV1_NO_STEM_ROWS") != nullptr;
    sw.fused_head = getenv("MNV1_FUSED_HEAD") != nullptr;   // one cluster kernel for pool+FC+softmax: correct, still slower than the three launches
    sw.fused_pair = getenv("MNV1_FUSED_PAIR") != nullptr;   // layers 14-23 as CTA-pair fused blocks inside mnv1_forward*: correct, but
                                                             // stencil-bound at 67 us per block against 45 us for the two kernels
    sw.pp_direct = getenv("MNV1_PP_DIRECT") != nullptr;
    sw.no_pp_tail = getenv("MNV1_NO_PP_TAIL") != nullptr;
    sw.h2d_wc = getenv("MNV1_H2D_WC") != nullptr;       // write-combined source buffers in mnv1_h2d_probe_*
    sw.rb_mask = getenv("MNV1_RB_MASK") ? strtol(getenv("MNV1_RB_MASK"), nullptr, 0) : ~0L;
  });
  return sw;
}
bool mnv1::pdl_enabled() { return !switches().no_pdl; }

cudaError_t mnv1::ensure_dyn_smem(const void* fn, int bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, int> done;   // (function, device) -> bytes granted
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lk(mu);
  auto it = done.find({fn, dev});
  if (it != done.end() && it->second >= bytes) return cudaSuccess;
  e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) done[{fn, dev}] = bytes;
  return e;
}

// Every entry point that takes a context runs on the context's device, whatever device the calling
// thread had current, and leaves the thread's current device as it found it.
struct DeviceGuard {
  int prev = -1, dev;
  explicit DeviceGuard(int d) : dev(d) {
    if (dev < 0) return;                                   // null context: the callee reports it
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DeviceGuard() { if (dev >= 0 && prev >= 0 && prev != dev) cudaSetDevice(prev); }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define GUARD(ctx) DeviceGuard guard_((ctx) ? (ctx)->device : -1)

struct LayerDef { int kind, cin, cout, hin, hout, stride; long w_off, w_cnt, c_off; };

static const LayerDef* layer_defs() {
  static LayerDef L[MNV1_NUM_LAYERS];
  static bool init = false;
  if (!init) {
    // MobileNet.c schedule; L26 is stride 1 (App. B note)
    const int dw_c[13] = {32, 64, 128, 128, 256, 256, 512, 512, 512, 512, 512, 512, 1024};
    const int dw_s[13] = {1, 2, 1, 2, 1, 2, 1, 1, 1, 1, 1, 2, 1};
    const int pw_o[13] = {64, 128, 128, 256, 256, 512, 512, 512, 512, 512, 512, 1024, 1024};
    long w = 0, c = 0;
    int h = 112, k = 0;
    L[k++] = {MNV1_CONVOLUTE, 3, 32, 224, 112, 2, 0, 864, 0};
    w = 864; c = 32;
    for (int b = 0; b < 13; ++b) {
      const int ho = h / dw_s[b];
      L[k++] = {MNV1_DEPTHWISE, dw_c[b], dw_c[b], h, ho, dw_s[b], w, dw_c[b] * 9L, c};
      w += dw_c[b] * 9L; c += dw_c[b]; h = ho;
      L[k++] = {MNV1_POINTWISE, dw_c[b], pw_o[b], h, h, 1, w, (long)dw_c[b] * pw_o[b], c};
      w += (long)dw_c[b] * pw_o[b]; c += pw_o[b];
    }
    L[k++] = {MNV1_POOL, 1024, 1024, 7, 1, 1, w, 0, c};
    L[k++] = {MNV1_FC, 1024, 1000, 1, 1, 1, w, 1024000L, c};
    init = true;
  }
  return L;
}

struct GraphKey {
  const void* img; int n; void* logits; void* top1; void* prob;
  bool operator<(const GraphKey& o) const {
    return std::tie(img, n, logits, top1, prob) < std::tie(o.img, o.n, o.logits, o.top1, o.prob);
  }
};

struct mnv1_ctx {
  int device = 0;
  mnv1_dtype dtype = MNV1_F32;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  mnv1_pad pad = MNV1_PAD_REF;
  float in_scale = 1.f, in_bias = 0.f;
  int num_sms = 148;
  bool timing = true;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool ev_valid = false;
  long launches = 0;
  std::string err;
  const char* last_kernel = "";
  // whole-network state
  mnv1_filter* net[MNV1_NUM_LAYERS] = {};
  bool have_weights = false;
  int plan_batch = 0;
  void* act[2] = {nullptr, nullptr};
  uint8_t* d_images = nullptr;
  float* d_pooled = nullptr;
  float* d_logits = nullptr;
  int* d_top1 = nullptr;
  float* d_prob = nullptr;
  uint8_t* h_images = nullptr;  // pinned staging (slot 0; aliases of slots[0])
  float* h_logits = nullptr;
  int* h_top1 = nullptr;
  float* h_prob = nullptr;
  // kSlots in-flight batches for mnv1_forward_submit / _wait: while batch i computes, the images of
  // batches i+1 and i+2 are copied in (copy stream) and batch i-1's results are copied out (d2h
  // stream).  Three, not two: the upload of a 256-image batch (38.5 MB over PCIe, ~0.77 ms) is
  // almost as long as its kernels (0.86 ms), so with two slots any host wake-up delay stalls the GPU.
  static constexpr int kSlots = 3;
  struct Slot {
    uint8_t* d_images = nullptr; float* d_logits = nullptr; int* d_top1 = nullptr; float* d_prob = nullptr;
    uint8_t* h_images = nullptr; float* h_logits = nullptr; int* h_top1 = nullptr; float* h_prob = nullptr;
    cudaEvent_t ev_h2d = nullptr, ev_compute = nullptr, ev_done = nullptr;
    bool busy = false; int n = 0; long ticket = -1;
    float* u_logits = nullptr; int* u_top1 = nullptr; float* u_prob = nullptr;  // unpinned user outputs
  } slots[kSlots];
  cudaStream_t copy_stream = nullptr, d2h_stream = nullptr;
  long next_ticket = 0;
  long graph_kernels = 30;
  std::map<GraphKey, cudaGraphExec_t> graphs;
  bool use_graph = true;
  bool use_fused = true;   // depthwise->pointwise block fusion inside mnv1_forward*
  int u8_wrap = 0;         // integer contexts: 1 = wrap modulo 256 like kernel.cl's store to unsigned char, 0 = saturate
  // staging for the planar <-> NHWC boundary copies and mnv1_softmax: grown on demand, then reused, so the
  // per-layer loop of the host programs allocates nothing after its first pass
  void* scratch = nullptr;
  size_t scratch_bytes = 0;
  // logits gather of the data-parallel mode (mnv1_gather_*): this rank's gather block, the peers' blocks it
  // stores into, and the HeadGather handed to the head kernel
  int g_world = 0, g_rank = 0, g_rows = 0;
  long g_first = 0; int g_max = 0;   // this rank's window of the gather block (mnv1_gather_set_rows)
  uint8_t* g_block = nullptr;                 // [world*rows][1000] f32 | [world*rows] i32 | [world*rows] f32
  void* g_peer[8] = {};                       // base of rank d's block as mapped into this process / device
  bool g_peer_ipc[8] = {};
  mnv1::HeadGather gather = {};
};

extern "C" {
static void drop_graphs(mnv1_ctx* ctx);
static void gather_release(mnv1_ctx* ctx);
}

static void* scratch(mnv1_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->scratch_bytes) return ctx->scratch;
  cudaStreamSynchronize(ctx->stream);
  cudaFree(ctx->scratch); ctx->scratch = nullptr; ctx->scratch_bytes = 0;
  const size_t want = (bytes + (size_t(1) << 20) - 1) & ~((size_t(1) << 20) - 1);
  if (cudaMalloc(&ctx->scratch, want) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  ctx->scratch_bytes = want;
  return ctx->scratch;
}

static int fail(mnv1_ctx* ctx, int code, const std::string& msg) {
  g_last_error = msg;
  if (ctx) ctx->err = msg;
  return code;
}
static int fail_cuda(mnv1_ctx* ctx, cudaError_t e, const char* what) {
  std::string m = std::string(what) + ": " + cudaGetErrorString(e);
  if (ctx && !ctx->err.empty() && e == cudaErrorInvalidValue) m += " (" + ctx->err + ")";
  return fail(ctx, MNV1_ECUDA, m);
}
#define CK(ctx, call)                                                  \
  do {                                                                 \
    cudaError_t e_ = (call);                                           \
    if (e_ != cudaSuccess) return fail_cuda(ctx, e_, #call);           \
  } while (0)

static size_t elem_size(mnv1_dtype dt) { return dt == MNV1_BF16 ? 2 : dt == MNV1_U8 ? 1 : 4; }
static int pad_lo_for(const mnv1_ctx* ctx, int stride) {
  return (stride == 2 && ctx->pad == MNV1_PAD_TFSAME) ? 0 : 1;
}

struct TimedLaunch {  // brackets one per-layer launch with the event pair (MobileNet.c:303-305)
  mnv1_ctx* c;
  explicit TimedLaunch(mnv1_ctx* ctx) : c(ctx) {
    if (c->timing) cudaEventRecord(c->ev0, c->stream);
  }
  ~TimedLaunch() {
    if (c->timing) { cudaEventRecord(c->ev1, c->stream); c->ev_valid = true; }
  }
};

extern "C" {

const char* mnv1_version(void) { return "mnv1-b200 0.2 (sm_100a)"; }

const char* mnv1_last_error(const mnv1_ctx* ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

int mnv1_ctx_create(int device, mnv1_dtype dtype, mnv1_ctx** out) {
  if (!out || (dtype != MNV1_F32 && dtype != MNV1_BF16 && dtype != MNV1_U8)) return fail(nullptr, MNV1_EINVAL, "bad ctx_create args");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(nullptr, MNV1_ECUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) +
                                         " (this library has no CPU fallback)");
  if (device < 0 || device >= count) return fail(nullptr, MNV1_EINVAL, "device index out of range");
  (void)mnv1::switches();          // environment switches are read here, once per process
  DeviceGuard guard_(device);
  cudaDeviceProp prop;
  CK(nullptr, cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(nullptr, MNV1_EUNSUPPORTED, std::string("device ") + prop.name +
                                                " is not sm_100 (kernels are built for sm_100a only)");
  std::unique_ptr<mnv1_ctx> ctx(new mnv1_ctx);
  ctx->device = device;
  ctx->dtype = dtype;
  ctx->num_sms = prop.multiProcessorCount;
  CK(nullptr, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  CK(nullptr, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  CK(nullptr, cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
  CK(nullptr, cudaEventCreate(&ctx->ev0));
  CK(nullptr, cudaEventCreate(&ctx->ev1));
  *out = ctx.release();
  return MNV1_OK;
}

static void free_plan(mnv1_ctx* ctx) {
  for (auto& kv : ctx->graphs) cudaGraphExecDestroy(kv.second);
  ctx->graphs.clear();
  for (int i = 0; i < 2; ++i) { cudaFree(ctx->act[i]); ctx->act[i] = nullptr; }
  cudaFree(ctx->d_pooled); ctx->d_pooled = nullptr;
  for (auto& sl : ctx->slots) {
    cudaFree(sl.d_images); cudaFree(sl.d_logits); cudaFree(sl.d_top1); cudaFree(sl.d_prob);
    cudaFreeHost(sl.h_images); cudaFreeHost(sl.h_logits); cudaFreeHost(sl.h_top1); cudaFreeHost(sl.h_prob);
    if (sl.ev_h2d) cudaEventDestroy(sl.ev_h2d);
    if (sl.ev_compute) cudaEventDestroy(sl.ev_compute);
    if (sl.ev_done) cudaEventDestroy(sl.ev_done);
    sl = mnv1_ctx::Slot();
  }
  ctx->d_images = nullptr; ctx->d_logits = nullptr; ctx->d_top1 = nullptr; ctx->d_prob = nullptr;
  ctx->h_images = nullptr; ctx->h_logits = nullptr; ctx->h_top1 = nullptr; ctx->h_prob = nullptr;
  ctx->plan_batch = 0;
}

int mnv1_filter_destroy(mnv1_ctx* ctx, mnv1_filter* f);

int mnv1_ctx_destroy(mnv1_ctx* ctx) {
  if (!ctx) return MNV1_OK;
  GUARD(ctx);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
  if (ctx->d2h_stream) cudaStreamSynchronize(ctx->d2h_stream);
  free_plan(ctx);
  gather_release(ctx);
  cudaFree(ctx->scratch); ctx->scratch = nullptr; ctx->scratch_bytes = 0;
  for (auto& f : ctx->net) { if (f) mnv1_filter_destroy(ctx, f); f = nullptr; }
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
  delete ctx;
  return MNV1_OK;
}

int mnv1_ctx_set_stream(mnv1_ctx* ctx, void* s) {
  GUARD(ctx);
  if (!ctx) return fail(nullptr, MNV1_EINVAL, "null ctx");
  if (ctx->own_stream && ctx->stream) { cudaStreamSynchronize(ctx->stream); cudaStreamDestroy(ctx->stream); }
  ctx->stream = (cudaStream_t)s;
  ctx->own_stream = false;
  for (auto& kv : ctx->graphs) cudaGraphExecDestroy(kv.second);
  ctx->graphs.clear();
  return MNV1_OK;
}
int mnv1_ctx_set_pad_mode(mnv1_ctx* ctx, mnv1_pad pad) {
  GUARD(ctx);
  if (!ctx || (pad != MNV1_PAD_REF && pad != MNV1_PAD_TFSAME)) return fail(ctx, MNV1_EINVAL, "bad pad mode");
  ctx->pad = pad;
  for (auto& kv : ctx->graphs) cudaGraphExecDestroy(kv.second);
  ctx->graphs.clear();
  return MNV1_OK;
}
int mnv1_ctx_set_input_transform(mnv1_ctx* ctx, float scale, float bias) {
  GUARD(ctx);
  if (!ctx) return fail(nullptr, MNV1_EINVAL, "null ctx");
  ctx->in_scale = scale; ctx->in_bias = bias;
  for (auto& kv : ctx->graphs) cudaGraphExecDestroy(kv.second);
  ctx->graphs.clear();
  return MNV1_OK;
}
int mnv1_ctx_enable_timing(mnv1_ctx* ctx, int on) {
  GUARD(ctx);
  if (!ctx) return fail(nullptr, MNV1_EINVAL, "null ctx");
  ctx->timing = on != 0;
  return MNV1_OK;
}
int mnv1_sync(mnv1_ctx* ctx) {
  GUARD(ctx);
  if (!ctx) return fail(nullptr, MNV1_EINVAL, "null ctx");
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  return MNV1_OK;
}
int mnv1_last_kernel_ms(mnv1_ctx* ctx, float* ms) {
  GUARD(ctx);
  if (!ctx || !ms) return fail(ctx, MNV1_EINVAL, "null arg");
  if (!ctx->timing || !ctx->ev_valid) return fail(ctx, MNV1_ESTATE, "no timed launch yet");
  CK(ctx, cudaEventSynchronize(ctx->ev1));
  CK(ctx, cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
  return MNV1_OK;
}
long mnv1_launch_count(const mnv1_ctx* ctx) { return ctx ? ctx->launches : 0; }
const char* mnv1_last_kernel_name(const mnv1_ctx* ctx) { return ctx ? ctx->last_kernel : ""; }

// ---------------------------------------------------------------- buffers
int mnv1_malloc(mnv1_ctx* ctx, int n, int c, int h, int w, mnv1_buf** out) {
  GUARD(ctx);
  if (!ctx || !out || n < 0 || c <= 0 || h <= 0 || w <= 0) return fail(ctx, MNV1_EINVAL, "bad malloc shape");
  std::unique_ptr<mnv1_buf> b(new mnv1_buf);
  b->n = n; b->c = c; b->h = h; b->w = w;
  b->bytes = (size_t)n * c * h * w * elem_size(ctx->dtype);
  if (b->bytes) {
    cudaError_t e = cudaMalloc(&b->d, b->bytes);
    if (e != cudaSuccess) return fail(ctx, MNV1_ENOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
  }
  *out = b.release();
  return MNV1_OK;
}
int mnv1_malloc_u8(mnv1_ctx* ctx, size_t bytes, mnv1_buf** out) {
  GUARD(ctx);
  if (!ctx || !out) return fail(ctx, MNV1_EINVAL, "null arg");
  std::unique_ptr<mnv1_buf> b(new mnv1_buf);
  b->bytes = bytes; b->is_u8 = true;
  if (bytes) {
    cudaError_t e = cudaMalloc(&b->d, bytes);
    if (e != cudaSuccess) return fail(ctx, MNV1_ENOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
  }
  *out = b.release();
  return MNV1_OK;
}
int mnv1_free(mnv1_ctx* ctx, mnv1_buf* b) {
  GUARD(ctx);
  if (!b) return MNV1_OK;
  if (ctx) cudaStreamSynchronize(ctx->stream);
  if (b->owned && b->d) cudaFree(b->d);
  delete b;
  return MNV1_OK;
}
void* mnv1_buf_device_ptr(mnv1_buf* b) { return b ? b->d : nullptr; }

int mnv1_upload_u8(mnv1_ctx* ctx, mnv1_buf* b, const uint8_t* host, size_t bytes) {
  GUARD(ctx);
  if (!ctx || !b || !b->is_u8 || bytes > b->bytes || (!host && bytes)) return fail(ctx, MNV1_EINVAL, "bad upload_u8");
  if (bytes) CK(ctx, cudaMemcpyAsync(b->d, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  CK(ctx, cudaStreamSynchronize(ctx->stream));  // CL_TRUE blocking write, MobileNet.c:350
  return MNV1_OK;
}
int mnv1_upload_planar(mnv1_ctx* ctx, mnv1_buf* b, const float* host) {
  GUARD(ctx);
  if (!ctx || !b || b->is_u8 || (!host && b->bytes)) return fail(ctx, MNV1_EINVAL, "bad upload_planar");
  if (!b->bytes) return MNV1_OK;
  const size_t elems = (size_t)b->n * b->c * b->h * b->w;
  float* stage = (float*)scratch(ctx, elems * 4);
  if (!stage) return fail(ctx, MNV1_ENOMEM, "staging cudaMalloc failed");
  cudaError_t e = cudaMemcpyAsync(stage, host, elems * 4, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess)
    e = ctx->dtype == MNV1_U8 ? mnv1::launch_u8_layout(0, b->d, stage, false, b->n, b->c, b->h * b->w, ctx->stream)
                              : mnv1::launch_nchw_to_nhwc(ctx->dtype, b->d, stage, b->n, b->c, b->h, b->w, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  ctx->launches++;
  if (e != cudaSuccess) return fail_cuda(ctx, e, "upload_planar");
  return MNV1_OK;
}
int mnv1_download_planar(mnv1_ctx* ctx, mnv1_buf* b, float* host) {
  GUARD(ctx);
  if (!ctx || !b || b->is_u8 || (!host && b->bytes)) return fail(ctx, MNV1_EINVAL, "bad download_planar");
  if (!b->bytes) return MNV1_OK;
  const size_t elems = (size_t)b->n * b->c * b->h * b->w;
  float* stage = (float*)scratch(ctx, elems * 4);
  if (!stage) return fail(ctx, MNV1_ENOMEM, "staging cudaMalloc failed");
  cudaError_t e = ctx->dtype == MNV1_U8 ? mnv1::launch_u8_layout(1, stage, b->d, false, b->n, b->c, b->h * b->w, ctx->stream)
                                        : mnv1::launch_nhwc_to_nchw(ctx->dtype, stage, b->d, b->n, b->c, b->h, b->w, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(host, stage, elems * 4, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  ctx->launches++;
  if (e != cudaSuccess) return fail_cuda(ctx, e, "download_planar");
  return MNV1_OK;
}

// planar u8 host arrays — the reference's own host layout (`unsigned char* output_image`, MobileNet.c:116-143,
// written / read by clEnqueueWriteBuffer / ReadBuffer :350,:395) — without a detour through fp32
int mnv1_upload_planar_u8(mnv1_ctx* ctx, mnv1_buf* b, const uint8_t* host) {
  GUARD(ctx);
  if (!ctx || !b || b->is_u8 || (!host && b->bytes)) return fail(ctx, MNV1_EINVAL, "bad upload_planar_u8");
  if (ctx->dtype != MNV1_U8) return fail(ctx, MNV1_EUNSUPPORTED, "upload_planar_u8: integer (MNV1_U8) contexts only");
  if (!b->bytes) return MNV1_OK;
  const size_t elems = (size_t)b->n * b->c * b->h * b->w;
  uint8_t* stage = (uint8_t*)scratch(ctx, elems);
  if (!stage) return fail(ctx, MNV1_ENOMEM, "staging cudaMalloc failed");
  cudaError_t e = cudaMemcpyAsync(stage, host, elems, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = mnv1::launch_u8_layout(0, b->d, stage, true, b->n, b->c, b->h * b->w, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  ctx->launches++;
  if (e != cudaSuccess) return fail_cuda(ctx, e, "upload_planar_u8");
  return MNV1_OK;
}
int mnv1_download_planar_u8(mnv1_ctx* ctx, mnv1_buf* b, uint8_t* host) {
  GUARD(ctx);
  if (!ctx || !b || b->is_u8 || (!host && b->bytes)) return fail(ctx, MNV1_EINVAL, "bad download_planar_u8");
  if (ctx->dtype != MNV1_U8) return fail(ctx, MNV1_EUNSUPPORTED, "download_planar_u8: integer (MNV1_U8) contexts only");
  if (!b->bytes) return MNV1_OK;
  const size_t elems = (size_t)b->n * b->c * b->h * b->w;
  uint8_t* stage = (uint8_t*)scratch(ctx, elems);
  if (!stage) return fail(ctx, MNV1_ENOMEM, "staging cudaMalloc failed");
  cudaError_t e = mnv1::launch_u8_layout(1, stage, b->d, true, b->n, b->c, b->h * b->w, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(host, stage, elems, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  ctx->launches++;
  if (e != cudaSuccess) return fail_cuda(ctx, e, "download_planar_u8");
  return MNV1_OK;
}
int mnv1_ctx_set_u8_store(mnv1_ctx* ctx, int wrap) {
  if (!ctx) return fail(nullptr, MNV1_EINVAL, "null ctx");
  ctx->u8_wrap = wrap ? 1 : 0;
  drop_graphs(ctx);
  return MNV1_OK;
}

// ---------------------------------------------------------------- filters
static int upload_vec(mnv1_ctx* ctx, const std::vector<float>& v, float** d) {
  cudaError_t e = cudaMalloc(d, v.size() * 4);
  if (e != cudaSuccess) return fail(ctx, MNV1_ENOMEM, "cudaMalloc(filter) failed");
  CK(ctx, cudaMemcpy(*d, v.data(), v.size() * 4, cudaMemcpyHostToDevice));
  return MNV1_OK;
}
static uint16_t f32_to_bf16_rne(float x) {
  uint32_t u; memcpy(&u, &x, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

// Integer contexts: the filter values must be integers in [-128, 127] (the reference's int8_t buffers,
// MobileNet.c:116,248); `shift` is an integer bias added to the s32 accumulator; `scale`, when given, must be
// the same power of two 2^-s for every channel and becomes the right shift s of the requantisation.
static int create_filter_u8(mnv1_ctx* ctx, mnv1_filter* f, const float* w, const float* scale, const float* shift) {
  const int cin = f->cin, cout = f->cout;
  size_t cnt = 0;
  switch (f->kind) {
    case MNV1_CONVOLUTE: if (cin != 3 || cout != 32) return fail(ctx, MNV1_EUNSUPPORTED, "convolute (u8): 3 -> 32 channels only"); cnt = (size_t)27 * cout; break;
    case MNV1_DEPTHWISE: if (cin != cout || cout % 4) return fail(ctx, MNV1_EINVAL, "depthwise (u8): cin == cout, a multiple of 4"); cnt = (size_t)9 * cout; break;
    case MNV1_POINTWISE: case MNV1_FC: cnt = (size_t)cin * cout; break;
    default: return fail(ctx, MNV1_EINVAL, "filter kind has no weights");
  }
  for (size_t i = 0; i < cnt; ++i)
    if (w[i] != (float)(int)w[i] || w[i] < -128.f || w[i] > 127.f)
      return fail(ctx, MNV1_EINVAL, "integer context: filter values must be integers in [-128, 127]");
  if (scale) {
    int s = 0;
    while (s < 31 && scale[0] != 1.0f / (float)(1u << s)) ++s;
    if (scale[0] != 1.0f / (float)(1u << s)) return fail(ctx, MNV1_EINVAL, "integer context: scale must be 2^-s, 0 <= s <= 31");
    for (int c = 1; c < cout; ++c)
      if (scale[c] != scale[0]) return fail(ctx, MNV1_EINVAL, "integer context: scale must be the same power of two for every channel");
    f->rshift = s;
  }
  std::vector<int8_t> q(cnt);
  std::vector<int> packed;
  if (f->kind == MNV1_CONVOLUTE) {           // [O][plane][ty][tx] -> 7 words of 4 taps per filter in the order an interleaved
    packed.assign((size_t)cout * 7, 0);      // RGB row delivers the window bytes: k = 9 ty + 3 tx + plane (28th tap = 0)
    for (int o = 0; o < cout; ++o)
      for (int pl = 0; pl < 3; ++pl)
        for (int ty = 0; ty < 3; ++ty)
          for (int tx = 0; tx < 3; ++tx) {
            const int k = 9 * ty + 3 * tx + pl;
            packed[(size_t)o * 7 + (k >> 2)] |= (int)((uint32_t)(uint8_t)(int8_t)(int)w[(size_t)o * 27 + pl * 9 + ty * 3 + tx] << (8 * (k & 3)));
          }
  } else if (f->kind == MNV1_DEPTHWISE) {    // [C][3][3] -> [C][3] words (w[ty][0], w[ty][1], w[ty][2], 0): one DP4A per tap row
    packed.assign((size_t)cout * 3, 0);
    for (int c = 0; c < cout; ++c)
      for (int t = 0; t < 9; ++t)
        packed[(size_t)c * 3 + t / 3] |= (int)((uint32_t)(uint8_t)(int8_t)(int)w[(size_t)c * 9 + t] << (8 * (t % 3)));
  } else {
    for (size_t i = 0; i < cnt; ++i) q[i] = (int8_t)(int)w[i];                 // [Cout][Cin]: the K-major B operand as is
  }
  if (!packed.empty()) {
    if (cudaMalloc(&f->w_q32, packed.size() * 4) != cudaSuccess) return fail(ctx, MNV1_ENOMEM, "cudaMalloc(filter) failed");
    CK(ctx, cudaMemcpy(f->w_q32, packed.data(), packed.size() * 4, cudaMemcpyHostToDevice));
    if (f->kind == MNV1_CONVOLUTE) f->h_q32 = packed;
  } else {
    if (cudaMalloc(&f->w_s8, cnt) != cudaSuccess) return fail(ctx, MNV1_ENOMEM, "cudaMalloc(filter) failed");
    CK(ctx, cudaMemcpy(f->w_s8, q.data(), cnt, cudaMemcpyHostToDevice));
  }
  if (shift) {
    std::vector<int> b(cout);
    for (int c = 0; c < cout; ++c) {
      if (shift[c] != (float)(int)shift[c]) return fail(ctx, MNV1_EINVAL, "integer context: shift (bias) must be integers");
      b[c] = (int)shift[c];
    }
    if (cudaMalloc(&f->bias_i32, (size_t)cout * 4) != cudaSuccess) return fail(ctx, MNV1_ENOMEM, "cudaMalloc(filter) failed");
    CK(ctx, cudaMemcpy(f->bias_i32, b.data(), (size_t)cout * 4, cudaMemcpyHostToDevice));
    if (f->kind == MNV1_CONVOLUTE) f->h_bias = b;
  }
  return MNV1_OK;
}

int mnv1_filter_create(mnv1_ctx* ctx, mnv1_kind kind, const float* w, int cin, int cout, const float* scale,
                       const float* shift, mnv1_act act, mnv1_filter** out) {
  GUARD(ctx);
  if (!ctx || !w || !out || cin <= 0 || cout <= 0) return fail(ctx, MNV1_EINVAL, "bad filter_create args");
  std::unique_ptr<mnv1_filter> f(new mnv1_filter);
  f->kind = kind; f->cin = cin; f->cout = cout; f->act = act;
  if (ctx->dtype == MNV1_U8) {
    int rc8 = create_filter_u8(ctx, f.get(), w, scale, shift);
    if (rc8) { mnv1_filter_destroy(ctx, f.release()); return rc8; }
    *out = f.release();
    return MNV1_OK;
  }
  std::vector<float> dev;
  int rc = MNV1_OK;
  switch (kind) {
    case MNV1_CONVOLUTE: {  // [O][3(R,G,B)][3][3] (kernel.cl:15-51 findex order) -> [27][O]
      if (cin != 3) return fail(ctx, MNV1_EINVAL, "convolute filter needs cin == 3");
      dev.resize((size_t)27 * cout);
      for (int o = 0; o < cout; ++o)
        for (int t = 0; t < 27; ++t) dev[(size_t)t * cout + o] = w[(size_t)o * 27 + t];
      if (ctx->dtype == MNV1_BF16 && cout == 32) {  // tensor-core stem: prepared lazily per input transform
        f->h_w.assign(w, w + (size_t)27 * cout);
        if (scale) f->h_scale.assign(scale, scale + cout);
        if (shift) f->h_shift.assign(shift, shift + cout);
        if (cudaMalloc(&f->wq, 32 * 32 * sizeof(__half)) != cudaSuccess || cudaMalloc(&f->shift2, 32 * 4) != cudaSuccess)
          return fail(ctx, MNV1_ENOMEM, "cudaMalloc(stem filter) failed");
      }
      break;
    }
    case MNV1_DEPTHWISE: {  // [C][3][3] (kernel.cl:77) -> [9][C]
      if (cin != cout) return fail(ctx, MNV1_EINVAL, "depthwise filter needs cin == cout");
      dev.resize((size_t)9 * cout);
      for (int c = 0; c < cout; ++c)
        for (int t = 0; t < 9; ++t) dev[(size_t)t * cout + c] = w[(size_t)c * 9 + t];
      if (ctx->dtype == MNV1_BF16) {  // taps with the folded-BN scale multiplied in (depthwise_tma.cu)
        std::vector<float> sc(dev);
        if (scale)
          for (int c = 0; c < cout; ++c)
            for (int t = 0; t < 9; ++t) sc[(size_t)t * cout + c] *= scale[c];
        if ((rc = upload_vec(ctx, sc, &f->w_scaled)) != MNV1_OK) return rc;
      }
      break;
    }
    case MNV1_POINTWISE:
    case MNV1_FC:           // [Cout][Cin] (kernel.cl:106) kept as is: it is the K-major B operand
      dev.assign(w, w + (size_t)cin * cout);
      if (scale) f->h_scale.assign(scale, scale + cout);   // host copies: the fused block kernel takes them by value
      if (shift) f->h_shift.assign(shift, shift + cout);
      break;
    default:
      return fail(ctx, MNV1_EINVAL, "filter kind has no weights");
  }
  if ((rc = upload_vec(ctx, dev, &f->w_f32)) != MNV1_OK) return rc;
  if (scale) { std::vector<float> s(scale, scale + cout); if ((rc = upload_vec(ctx, s, &f->scale)) != MNV1_OK) return rc; }
  if (shift) { std::vector<float> s(shift, shift + cout); if ((rc = upload_vec(ctx, s, &f->shift)) != MNV1_OK) return rc; }
  if (ctx->dtype == MNV1_BF16 && (kind == MNV1_POINTWISE || kind == MNV1_FC)) {
    std::vector<uint16_t> h((size_t)cin * cout);
    for (size_t i = 0; i < h.size(); ++i) h[i] = f32_to_bf16_rne(w[i]);
    cudaError_t e = cudaMalloc(&f->w_bf16, h.size() * 2);
    if (e != cudaSuccess) return fail(ctx, MNV1_ENOMEM, "cudaMalloc(filter bf16) failed");
    CK(ctx, cudaMemcpy(f->w_bf16, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
    if (kind == MNV1_POINTWISE && cout % 64 == 0 && cin % 8 == 0 && cout <= 1024) {
      std::string err;
      cudaError_t te = mnv1::make_weight_tmap(f.get(), &err);
      if (te != cudaSuccess) return fail(ctx, MNV1_ECUDA, "TMA descriptor for pointwise filter: " + err);
    }
  }
  *out = f.release();
  return MNV1_OK;
}
int mnv1_filter_destroy(mnv1_ctx* ctx, mnv1_filter* f) {
  GUARD(ctx);
  if (!f) return MNV1_OK;
  if (ctx) cudaStreamSynchronize(ctx->stream);
  cudaFree(f->w_f32); cudaFree(f->w_scaled); cudaFree(f->w_bf16); cudaFree(f->wq); cudaFree(f->shift2); cudaFree(f->scale); cudaFree(f->shift);
  cudaFree(f->w_s8); cudaFree(f->w_q32); cudaFree(f->bias_i32);
  delete f;
  return MNV1_OK;
}

// ---------------------------------------------------------------- raw launches (device pointers)
// (re)build the stem's fp16 filter bank for the context's current input transform; must run
// outside graph capture (it copies synchronously)
static int prepare_stem(mnv1_ctx* ctx, mnv1_filter* f) {
  if (!f || !f->wq || ctx->in_scale == 0.f) return MNV1_OK;
  if (f->prepared && f->prep_scale == ctx->in_scale && f->prep_bias == ctx->in_bias) return MNV1_OK;
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  CK(ctx, mnv1::stem_tc_prepare(f->h_w.data(), f->h_scale.empty() ? nullptr : f->h_scale.data(),
                                f->h_shift.empty() ? nullptr : f->h_shift.data(), ctx->in_scale, ctx->in_bias, f->wq,
                                f->shift2, &f->p0, f->h_shift2));
  f->prep_scale = ctx->in_scale; f->prep_bias = ctx->in_bias; f->prepared = true;
  return MNV1_OK;
}

static cudaError_t run_stem(mnv1_ctx* ctx, void* out, const uint8_t* r, const uint8_t* g, const uint8_t* b,
                            int pix_stride, long img_stride, const mnv1_filter* f, int n, int rows, int cols,
                            int stride) {
  mnv1::StemArgs a{r, g, b, pix_stride, img_stride, n, rows, cols, stride, f->cout, pad_lo_for(ctx, stride),
                   ctx->in_scale, ctx->in_bias};
  ctx->launches++;
  if (ctx->dtype == MNV1_U8) {
    ctx->last_kernel = "stem_u8_kernel";
    return mnv1::launch_stem_u8((uint8_t*)out, a, f, ctx->u8_wrap, ctx->num_sms, ctx->stream, &ctx->last_kernel);
  }
  if (ctx->dtype == MNV1_BF16 && f->prepared && f->prep_scale == ctx->in_scale && f->prep_bias == ctx->in_bias) {
    ctx->err.clear();
    if (!mnv1::switches().no_stem_rows) {
      cudaError_t er = mnv1::launch_stem_rows((bf16*)out, a, f->wq, f->h_scale.empty() ? nullptr : f->h_scale.data(),
                                              f->h_shift2, f->p0, (int)f->act, ctx->num_sms, ctx->stream, &ctx->err);
      if (er != cudaErrorNotSupported) { ctx->last_kernel = "stem_rows_kernel"; return er; }
    }
    cudaError_t e = mnv1::launch_stem_tc((bf16*)out, a, f->wq, f->scale, f->shift2, f->p0, (int)f->act, ctx->num_sms,
                                         ctx->stream, &ctx->err);
    if (e != cudaErrorNotSupported) { ctx->last_kernel = "stem_tc_kernel"; return e; }
  }
  ctx->last_kernel = "stem_kernel";
  return mnv1::launch_stem(ctx->dtype, out, a, f->w_f32, Epilogue{f->scale, f->shift, (int)f->act}, ctx->stream);
}
static cudaError_t run_depthwise(mnv1_ctx* ctx, void* out, const void* in, const mnv1_filter* f, int n, int rows,
                                 int cols, int stride) {
  ctx->launches++;
  if (ctx->dtype == MNV1_U8) {
    ctx->last_kernel = "depthwise_u8_kernel";
    return mnv1::launch_depthwise_u8((uint8_t*)out, (const uint8_t*)in, f, n, rows, cols, stride, f->cout, pad_lo_for(ctx, stride),
                                     ctx->u8_wrap, ctx->stream);
  }
  if (ctx->dtype == MNV1_BF16 && f->w_scaled) {
    ctx->err.clear();
    if (!mnv1::switches().no_cw) {
      cudaError_t ec = mnv1::launch_depthwise_cw((bf16*)out, (const bf16*)in, f->w_scaled, f->shift, (int)f->act, n, rows,
                                                 cols, stride, f->cout, pad_lo_for(ctx, stride), ctx->stream);
      if (ec != cudaErrorNotSupported) { ctx->last_kernel = "depthwise_cw_kernel"; return ec; }
    }
    cudaError_t er = mnv1::launch_depthwise_ring((bf16*)out, (const bf16*)in, f->w_scaled, f->shift, (int)f->act, n, rows,
                                                 cols, stride, f->cout, pad_lo_for(ctx, stride), ctx->num_sms,
                                                 ctx->stream, &ctx->err);
    if (er != cudaErrorNotSupported) { ctx->last_kernel = "depthwise_ring_kernel"; return er; }
    cudaError_t e = mnv1::launch_depthwise_tma((bf16*)out, (const bf16*)in, f->w_scaled, f->shift, (int)f->act, n, rows,
                                               cols, stride, f->cout, pad_lo_for(ctx, stride), ctx->num_sms,
                                               ctx->stream, &ctx->err);
    if (e != cudaErrorNotSupported) { ctx->last_kernel = "depthwise_tma_kernel"; return e; }
  }
  ctx->last_kernel = "depthwise_kernel";
  return mnv1::launch_depthwise(ctx->dtype, out, in, f->w_f32, n, rows, cols, stride, f->cout,
                                pad_lo_for(ctx, stride), Epilogue{f->scale, f->shift, (int)f->act}, ctx->stream);
}
static cudaError_t run_pointwise(mnv1_ctx* ctx, void* out, const void* in, const mnv1_filter* f, long m,
                                 bool force_simt) {
  ctx->launches++;
  if (ctx->dtype == MNV1_U8) {
    ctx->last_kernel = "pointwise_i8_kernel";
    ctx->err.clear();
    return mnv1::launch_pointwise_i8((uint8_t*)out, (const uint8_t*)in, f, m, f->cin, f->cout, ctx->u8_wrap, ctx->num_sms, ctx->stream, &ctx->err);
  }
  if (ctx->dtype == MNV1_BF16 && f->has_tmap && !force_simt) {
    ctx->err.clear();
    cudaError_t ep = mnv1::launch_pointwise_pair((bf16*)out, (const bf16*)in, f, m, f->cin, f->cout, ctx->num_sms,
                                                 ctx->stream, &ctx->err);
    if (ep != cudaErrorNotSupported) { ctx->last_kernel = "pointwise_pair_kernel"; return ep; }
    ctx->last_kernel = "pointwise_tc_kernel";
    return mnv1::launch_pointwise_tc((bf16*)out, (const bf16*)in, f, m, f->cin, f->cout, ctx->num_sms,
                                     ctx->stream, &ctx->err);
  }
  ctx->last_kernel = "pointwise_simt_kernel";
  return mnv1::launch_pointwise_simt(ctx->dtype, out, in, f->w_f32, f->w_bf16, m, f->cin, f->cout,
                                     Epilogue{f->scale, f->shift, (int)f->act}, false, ctx->stream);
}

// ---------------------------------------------------------------- the four kernels
// depthwise -> pointwise as one kernel: the CTA-pair kernel with a streamed filter for the 512-channel blocks, the
// resident-filter kernel for the blocks whose filter fits in shared memory; cudaErrorNotSupported = no fused variant
static cudaError_t run_fused_block(mnv1_ctx* ctx, void* out, const void* in, const mnv1_filter* dw, const mnv1_filter* pw, int n,
                                   int rows, int cols, int stride, bool allow_pair) {
  ctx->err.clear();
  cudaError_t e = cudaErrorNotSupported;
  if (allow_pair)
    e = mnv1::launch_fused_pair((bf16*)out, (const bf16*)in, dw, pw, n, rows, cols, stride, pad_lo_for(ctx, stride), ctx->num_sms,
                                ctx->stream, &ctx->err);
  if (e != cudaErrorNotSupported) { ctx->launches++; ctx->last_kernel = "fused_pair_kernel"; return e; }
  e = mnv1::launch_fused_dw_pw((bf16*)out, (const bf16*)in, dw, pw, n, rows, cols, stride, pad_lo_for(ctx, stride), ctx->num_sms,
                               ctx->stream, &ctx->err);
  if (e != cudaErrorNotSupported) { ctx->launches++; ctx->last_kernel = "fused_dw_pw_kernel"; }
  return e;
}

static int check_fmap(mnv1_ctx* ctx, const mnv1_buf* b, int c, int h, int w, const char* what) {
  if (!b || b->is_u8 || b->c != c || b->h != h || b->w != w)
    return fail(ctx, MNV1_EINVAL, std::string(what) + ": buffer shape does not match the kernel arguments");
  return MNV1_OK;
}

static int convolute_common(mnv1_ctx* ctx, mnv1_buf* out, const uint8_t* r, const uint8_t* g, const uint8_t* b,
                            int pix_stride, size_t avail_bytes, const mnv1_filter* f, int rows, int cols,
                            int filtersize, int stride, int op_size) {
  if (!ctx || !out || !f) return fail(ctx, MNV1_EINVAL, "convolute: null argument");
  if (f->kind != MNV1_CONVOLUTE || f->cout != op_size) return fail(ctx, MNV1_EINVAL, "convolute: filter mismatch");
  if (filtersize != 3) return fail(ctx, MNV1_EUNSUPPORTED, "convolute: only 3x3 (K = 3, MobileNet.c:15)");
  if (op_size != 32) return fail(ctx, MNV1_EUNSUPPORTED, "convolute: op_size must be 32 (FILTER_SIZE_L1)");
  if ((stride != 1 && stride != 2) || rows % stride || cols % stride)
    return fail(ctx, MNV1_EUNSUPPORTED, "convolute: stride must be 1 or 2 and divide rows/cols");
  int rc = check_fmap(ctx, out, op_size, rows / stride, cols / stride, "convolute(out)");
  if (rc) return rc;
  const size_t plane = (size_t)rows * cols * pix_stride;
  if (avail_bytes < plane * out->n) return fail(ctx, MNV1_EINVAL, "convolute: image buffer too small for the batch");
  if (ctx->dtype == MNV1_U8 && (ctx->in_scale != 1.f || ctx->in_bias != 0.f))
    return fail(ctx, MNV1_EUNSUPPORTED, "convolute: integer contexts read the raw u8 pixels (input transform must be 1, 0)");
  rc = prepare_stem(ctx, const_cast<mnv1_filter*>(f));
  if (rc) return rc;
  TimedLaunch tl(ctx);
  CK(ctx, run_stem(ctx, out->d, r, g, b, pix_stride, (long)plane, f, out->n, rows, cols, stride));
  return MNV1_OK;
}

int mnv1_convolute(mnv1_ctx* ctx, mnv1_buf* out, const mnv1_buf* in_r, const mnv1_buf* in_g, const mnv1_buf* in_b,
                   const mnv1_filter* f, int rows, int cols, int filtersize, int stride, int op_size) {
  GUARD(ctx);
  if (!in_r || !in_g || !in_b || !in_r->is_u8 || !in_g->is_u8 || !in_b->is_u8)
    return fail(ctx, MNV1_EINVAL, "convolute: r/g/b must be u8 buffers");
  size_t avail = in_r->bytes < in_g->bytes ? in_r->bytes : in_g->bytes;
  if (in_b->bytes < avail) avail = in_b->bytes;
  return convolute_common(ctx, out, (const uint8_t*)in_r->d, (const uint8_t*)in_g->d, (const uint8_t*)in_b->d, 1,
                          avail, f, rows, cols, filtersize, stride, op_size);
}
int mnv1_convolute_rgb(mnv1_ctx* ctx, mnv1_buf* out, const mnv1_buf* in_rgb, const mnv1_filter* f, int rows,
                       int cols, int filtersize, int stride, int op_size) {
  GUARD(ctx);
  if (!in_rgb || !in_rgb->is_u8) return fail(ctx, MNV1_EINVAL, "convolute_rgb: image must be a u8 buffer");
  const uint8_t* p = (const uint8_t*)in_rgb->d;
  return convolute_common(ctx, out, p, p + 1, p + 2, 3, in_rgb->bytes, f, rows, cols, filtersize, stride, op_size);
}

int mnv1_depthwise(mnv1_ctx* ctx, mnv1_buf* out, const mnv1_buf* in, const mnv1_filter* f, int rows, int cols,
                   int filtersize, int stride, int op_size) {
  GUARD(ctx);
  if (!ctx || !out || !in || !f) return fail(ctx, MNV1_EINVAL, "depthwise: null argument");
  if (f->kind != MNV1_DEPTHWISE || f->cout != op_size) return fail(ctx, MNV1_EINVAL, "depthwise: filter mismatch");
  if (filtersize != 3) return fail(ctx, MNV1_EUNSUPPORTED, "depthwise: only 3x3 (K = 3, MobileNet.c:15)");
  if ((stride != 1 && stride != 2) || rows % stride || cols % stride)
    return fail(ctx, MNV1_EUNSUPPORTED, "depthwise: stride must be 1 or 2 and divide rows/cols");
  const int vec = ctx->dtype == MNV1_BF16 ? 8 : 4;   // u8: one 32-bit word
  if (op_size % vec) return fail(ctx, MNV1_EUNSUPPORTED, "depthwise: op_size must be a multiple of the 128-bit channel vector");
  int rc = check_fmap(ctx, in, op_size, rows, cols, "depthwise(in)");
  if (rc) return rc;
  rc = check_fmap(ctx, out, op_size, rows / stride, cols / stride, "depthwise(out)");
  if (rc) return rc;
  if (in->n != out->n) return fail(ctx, MNV1_EINVAL, "depthwise: batch mismatch");
  TimedLaunch tl(ctx);
  CK(ctx, run_depthwise(ctx, out->d, in->d, f, in->n, rows, cols, stride));
  return MNV1_OK;
}

static int pointwise_impl(mnv1_ctx* ctx, mnv1_buf* out, const mnv1_buf* in, const mnv1_filter* f, int rows,
                          int cols, int filtersize, int op_size, bool force_simt) {
  if (!ctx || !out || !in || !f) return fail(ctx, MNV1_EINVAL, "pointwise: null argument");
  if ((f->kind != MNV1_POINTWISE && f->kind != MNV1_FC) || f->cout != op_size || f->cin != filtersize)
    return fail(ctx, MNV1_EINVAL, "pointwise: filter mismatch (filtersize is the number of input planes, Cin)");
  int rc = check_fmap(ctx, in, filtersize, rows, cols, "pointwise(in)");
  if (rc) return rc;
  rc = check_fmap(ctx, out, op_size, rows, cols, "pointwise(out)");
  if (rc) return rc;
  if (in->n != out->n) return fail(ctx, MNV1_EINVAL, "pointwise: batch mismatch");
  TimedLaunch tl(ctx);
  CK(ctx, run_pointwise(ctx, out->d, in->d, f, (long)in->n * rows * cols, force_simt));
  return MNV1_OK;
}
int mnv1_pointwise(mnv1_ctx* ctx, mnv1_buf* out, const mnv1_buf* in, const mnv1_filter* f, int rows, int cols,
                   int filtersize, int op_size) {
  GUARD(ctx);
  return pointwise_impl(ctx, out, in, f, rows, cols, filtersize, op_size, false);
}
// the CUDA-core GEMM on any context (what fp32 contexts always run); used to cross-check the
// tcgen05 kernel on the device
int mnv1_pointwise_simt(mnv1_ctx* ctx, mnv1_buf* out, const mnv1_buf* in, const mnv1_filter* f, int rows, int cols,
                        int filtersize, int op_size) {
  GUARD(ctx);
  return pointwise_impl(ctx, out, in, f, rows, cols, filtersize, op_size, true);
}

int mnv1_pool(mnv1_ctx* ctx, mnv1_buf* out, const mnv1_buf* in, int rows, int cols, int filtersize, int op_size) {
  GUARD(ctx);
  if (!ctx || !out || !in) return fail(ctx, MNV1_EINVAL, "pool: null argument");
  if (rows != filtersize || cols != filtersize)
    return fail(ctx, MNV1_EUNSUPPORTED, "pool: global average only (rows = cols = filtersize, kernel.cl:126)");
  if (op_size % 4) return fail(ctx, MNV1_EUNSUPPORTED, "pool: op_size must be a multiple of 4");
  int rc = check_fmap(ctx, in, op_size, rows, cols, "pool(in)");
  if (rc) return rc;
  rc = check_fmap(ctx, out, op_size, 1, 1, "pool(out)");
  if (rc) return rc;
  if (in->n != out->n) return fail(ctx, MNV1_EINVAL, "pool: batch mismatch");
  TimedLaunch tl(ctx);
  ctx->launches++; ctx->last_kernel = "pool_kernel";
  if (ctx->dtype == MNV1_U8) {
    ctx->last_kernel = "pool_u8_kernel";
    CK(ctx, mnv1::launch_pool_u8((uint8_t*)out->d, (const uint8_t*)in->d, in->n, rows * cols, op_size, ctx->u8_wrap, ctx->stream));
    return MNV1_OK;
  }
  CK(ctx, mnv1::launch_pool(ctx->dtype, out->d, in->d, in->n, rows * cols, op_size, false, ctx->stream));
  return MNV1_OK;
}

__global__ void widen_logits_kernel(float* out, const bf16* in, long n) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __bfloat162float(in[i]);
}

int mnv1_softmax(mnv1_ctx* ctx, const mnv1_buf* logits, int classes, float* prob, int* top1, float* top1_prob) {
  GUARD(ctx);
  if (!ctx || !logits || logits->is_u8 || logits->c != classes || logits->h != 1 || logits->w != 1)
    return fail(ctx, MNV1_EINVAL, "softmax: logits must be [n][classes][1][1]");
  const int n = logits->n;
  if (n == 0) return MNV1_OK;
  // one carve-up of the context's scratch: [logits fp32 (bf16 contexts)] [prob] [top1] [top1_prob]
  const long cnt = (long)n * classes;
  const size_t b_logits = ctx->dtype != MNV1_F32 ? (size_t)cnt * 4 : 0, b_prob = prob ? (size_t)cnt * 4 : 0;
  uint8_t* base = (uint8_t*)scratch(ctx, b_logits + b_prob + (size_t)n * 8 + 64);
  if (!base) return fail(ctx, MNV1_ENOMEM, "softmax: staging cudaMalloc failed");
  float* d_logits = ctx->dtype != MNV1_F32 ? (float*)base : (float*)logits->d;
  float* d_prob = prob ? (float*)(base + b_logits) : nullptr;
  int* d_top1 = (int*)(base + b_logits + b_prob);
  float* d_p1 = (float*)(base + b_logits + b_prob + (size_t)n * 4);
  cudaError_t e = cudaSuccess;
  if (ctx->dtype == MNV1_BF16) {
    widen_logits_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, ctx->stream>>>(d_logits, (const bf16*)logits->d, cnt);
    ctx->launches++;
    e = cudaGetLastError();
  } else if (ctx->dtype == MNV1_U8) {     // the reference's softmax runs over the u8 logits (MobileNet.c:2769-2781)
    e = mnv1::launch_u8_to_f32(d_logits, (const uint8_t*)logits->d, cnt, ctx->stream);
    ctx->launches++;
  }
  if (e == cudaSuccess) {
    TimedLaunch tl(ctx);
    ctx->launches++; ctx->last_kernel = "softmax_kernel";
    e = mnv1::launch_softmax(d_logits, n, classes, d_prob, d_top1, d_p1, ctx->stream);
  }
  if (e == cudaSuccess && prob) e = cudaMemcpyAsync(prob, d_prob, cnt * 4, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess && top1) e = cudaMemcpyAsync(top1, d_top1, n * 4, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess && top1_prob) e = cudaMemcpyAsync(top1_prob, d_p1, n * 4, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) return fail_cuda(ctx, e, "softmax");
  return MNV1_OK;
}

// ---------------------------------------------------------------- whole network
int mnv1_layer_table(mnv1_layer_info* out) {
  if (!out) return MNV1_EINVAL;
  const LayerDef* L = layer_defs();
  for (int i = 0; i < MNV1_NUM_LAYERS; ++i)
    out[i] = {i + 1, L[i].kind, L[i].cin, L[i].cout, L[i].hin, L[i].hout, L[i].stride, L[i].w_off, L[i].w_cnt,
              L[i].c_off};
  return MNV1_OK;
}

int mnv1_set_weights(mnv1_ctx* ctx, const float* weights, const float* scale, const float* shift, mnv1_act act) {
  GUARD(ctx);
  if (!ctx || !weights) return fail(ctx, MNV1_EINVAL, "set_weights: null argument");
  const LayerDef* L = layer_defs();
  for (auto& kv : ctx->graphs) cudaGraphExecDestroy(kv.second);
  ctx->graphs.clear();
  for (int i = 0; i < MNV1_NUM_LAYERS; ++i) {
    if (ctx->net[i]) { mnv1_filter_destroy(ctx, ctx->net[i]); ctx->net[i] = nullptr; }
    if (L[i].kind == MNV1_POOL) continue;
    // integer contexts: the reference's FC is its `pointwise` kernel, ReLU and requantisation included (MobileNet.c:2689-2754)
    const bool fc = L[i].kind == MNV1_FC && ctx->dtype != MNV1_U8;
    int rc = mnv1_filter_create(ctx, (mnv1_kind)L[i].kind, weights + L[i].w_off, L[i].cin, L[i].cout,
                                (scale && !fc) ? scale + L[i].c_off : nullptr, shift ? shift + L[i].c_off : nullptr,
                                fc ? MNV1_ACT_NONE : act, &ctx->net[i]);
    if (rc) return rc;
  }
  ctx->have_weights = true;
  return MNV1_OK;
}

int mnv1_load_weights(mnv1_ctx* ctx, const char* path, mnv1_act act) {
  GUARD(ctx);
  if (!ctx || !path) return fail(ctx, MNV1_EINVAL, "load_weights: null argument");
  std::vector<float> w, sc, sh;
  std::string err;
  int rc = mnv1::load_weight_file(path, &w, &sc, &sh, &err);
  if (rc) return fail(ctx, rc, err);
  return mnv1_set_weights(ctx, w.data(), sc.data(), sh.data(), act);
}
int mnv1_parse_weights(const char* path, float* weights, float* scale, float* shift) {
  if (!path || !weights) return fail(nullptr, MNV1_EINVAL, "parse_weights: null argument");
  std::vector<float> w, sc, sh;
  std::string err;
  int rc = mnv1::load_weight_file(path, &w, &sc, &sh, &err);
  if (rc) return fail(nullptr, rc, err);
  memcpy(weights, w.data(), w.size() * 4);
  if (scale) memcpy(scale, sc.data(), sc.size() * 4);
  if (shift) memcpy(shift, sh.data(), sh.size() * 4);
  return MNV1_OK;
}
int mnv1_save_weights_bin(const char* path, const float* weights, const float* scale, const float* shift) {
  if (!path || !weights) return fail(nullptr, MNV1_EINVAL, "save_weights_bin: null argument");
  std::string err;
  int rc = mnv1::save_weight_file_bin(path, weights, scale, shift, &err);
  if (rc) return fail(nullptr, rc, err);
  return MNV1_OK;
}
int mnv1_read_ppm(const char* path, uint8_t* out, int height, int width) {
  if (!path || !out) return fail(nullptr, MNV1_EINVAL, "read_ppm: null argument");
  std::string err;
  int rc = mnv1::read_ppm(path, out, height, width, &err);
  if (rc) return fail(nullptr, rc, err);
  return MNV1_OK;
}

static const size_t kImgBytes = 224 * 224 * 3;
static const size_t kMaxActElems = 802816;  // layer 3 output per image (64 x 112 x 112)

static int retire_all(mnv1_ctx* ctx);

int mnv1_plan(mnv1_ctx* ctx, int max_batch) {
  GUARD(ctx);
  if (!ctx || max_batch <= 0) return fail(ctx, MNV1_EINVAL, "plan: bad batch");
  if (max_batch <= ctx->plan_batch) return MNV1_OK;
  // Re-planning frees the slot buffers: batches still in flight are completed first and their results
  // delivered to the callers' arrays, so a later mnv1_forward_wait(ticket) finds them retired, not lost.
  int rc = retire_all(ctx);
  if (rc) return rc;
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  CK(ctx, cudaStreamSynchronize(ctx->copy_stream));
  CK(ctx, cudaStreamSynchronize(ctx->d2h_stream));
  free_plan(ctx);
  const size_t act_bytes = (size_t)max_batch * kMaxActElems * elem_size(ctx->dtype);
  cudaError_t e = cudaSuccess;
  for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaMalloc(&ctx->act[i], act_bytes);
  // fp32 pooled means [n][1024]; integer contexts keep their u8 pooled vector and u8 logits in the 2 KB per image behind it
  if (e == cudaSuccess) e = cudaMalloc(&ctx->d_pooled, (size_t)max_batch * (1024 * 4 + 2048));
  for (auto& sl : ctx->slots) {
    if (e == cudaSuccess) e = cudaMalloc(&sl.d_images, (size_t)max_batch * kImgBytes);
    if (e == cudaSuccess) e = cudaMalloc(&sl.d_logits, (size_t)max_batch * MNV1_NUM_CLASSES * 4);
    if (e == cudaSuccess) e = cudaMalloc(&sl.d_top1, (size_t)max_batch * 4);
    if (e == cudaSuccess) e = cudaMalloc(&sl.d_prob, (size_t)max_batch * 4);
    if (e == cudaSuccess) e = cudaMallocHost(&sl.h_images, (size_t)max_batch * kImgBytes);
    if (e == cudaSuccess) e = cudaMallocHost(&sl.h_logits, (size_t)max_batch * MNV1_NUM_CLASSES * 4);
    if (e == cudaSuccess) e = cudaMallocHost(&sl.h_top1, (size_t)max_batch * 4);
    if (e == cudaSuccess) e = cudaMallocHost(&sl.h_prob, (size_t)max_batch * 4);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&sl.ev_h2d, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&sl.ev_compute, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&sl.ev_done, cudaEventDisableTiming);
  }
  ctx->d_images = ctx->slots[0].d_images; ctx->d_logits = ctx->slots[0].d_logits;
  ctx->d_top1 = ctx->slots[0].d_top1; ctx->d_prob = ctx->slots[0].d_prob;
  ctx->h_images = ctx->slots[0].h_images; ctx->h_logits = ctx->slots[0].h_logits;
  ctx->h_top1 = ctx->slots[0].h_top1; ctx->h_prob = ctx->slots[0].h_prob;
  if (e != cudaSuccess) {
    free_plan(ctx);
    return fail(ctx, MNV1_ENOMEM, std::string("plan: allocation failed: ") + cudaGetErrorString(e));
  }
  ctx->plan_batch = max_batch;
  return MNV1_OK;
}

// Enqueue layers 1..last on the context stream.  Activations ping-pong between the two arena
// halves; returns the device pointer holding layer `last`'s output in *result.
static cudaError_t enqueue_layers(mnv1_ctx* ctx, const uint8_t* d_img, int n, int last, float* d_logits,
                                  int* d_top1, float* d_prob, const void** result, cudaEvent_t* evs) {
  const LayerDef* L = layer_defs();
  const void* cur = nullptr;
  int side = 0;
  cudaError_t e = cudaSuccess;
  for (int i = 0; i < last && e == cudaSuccess; ++i) {
    if (evs) cudaEventRecord(evs[i], ctx->stream);
    void* dst = ctx->act[side];
    const mnv1_filter* f = ctx->net[i];
    switch (L[i].kind) {
      case MNV1_CONVOLUTE:
        e = run_stem(ctx, dst, d_img, d_img + 1, d_img + 2, 3, (long)kImgBytes, f, n, L[i].hin, L[i].hin, L[i].stride);
        cur = dst; side ^= 1;
        break;
      case MNV1_DEPTHWISE:
        // depthwise + the pointwise that follows as one kernel where a fused variant exists: the
        // depthwise map then never leaves the SM
        if (ctx->use_fused && ctx->dtype == MNV1_BF16 && i + 1 < last && L[i + 1].kind == MNV1_POINTWISE) {
          cudaError_t fe = run_fused_block(ctx, dst, cur, f, ctx->net[i + 1], n, L[i].hin, L[i].hin, L[i].stride, mnv1::switches().fused_pair);
          if (fe != cudaErrorNotSupported) {
            e = fe;
            cur = dst; side ^= 1;
            ++i;  // the pointwise layer is done too
            if (evs) cudaEventRecord(evs[i], ctx->stream);
            break;
          }
        }
        e = run_depthwise(ctx, dst, cur, f, n, L[i].hin, L[i].hin, L[i].stride);
        cur = dst; side ^= 1;
        break;
      case MNV1_POINTWISE:
        e = run_pointwise(ctx, dst, cur, f, (long)n * L[i].hin * L[i].hin, false);
        cur = dst; side ^= 1;
        break;
      case MNV1_POOL:
        if (ctx->dtype == MNV1_U8) {   // kernel.cl:116-131: integer mean, u8
          uint8_t* pooled8 = (uint8_t*)ctx->d_pooled + (size_t)ctx->plan_batch * 4096;
          ctx->launches++; ctx->last_kernel = "pool_u8_kernel";
          e = mnv1::launch_pool_u8(pooled8, (const uint8_t*)cur, n, L[i].hin * L[i].hin, L[i].cout, ctx->u8_wrap, ctx->stream);
          cur = pooled8;
          if (last == i + 1 && e == cudaSuccess) {           // per-layer dump: as fp32 [n][1024]
            e = mnv1::launch_u8_to_f32(ctx->d_pooled, pooled8, (long)n * L[i].cout, ctx->stream);
            ctx->launches++;
            cur = ctx->d_pooled;
          }
          break;
        }
        if (last == i + 1) {  // pool alone (per-layer dump): fp32 means
          ctx->launches++; ctx->last_kernel = "pool_kernel";
          e = mnv1::launch_pool(ctx->dtype, ctx->d_pooled, cur, n, L[i].hin * L[i].hin, L[i].cout, true, ctx->stream);
          cur = ctx->d_pooled;
        }
        break;  // otherwise fused into the head below
      case MNV1_FC: {
        if (ctx->dtype == MNV1_U8) {   // the FC launch of `pointwise` (MobileNet.c:2681-2763), then the softmax over the u8 logits
          uint8_t* logits8 = (uint8_t*)ctx->d_pooled + (size_t)ctx->plan_batch * 4096 + (size_t)ctx->plan_batch * 1024;
          ctx->err.clear();
          ctx->launches++; ctx->last_kernel = "pointwise_i8_kernel";
          e = mnv1::launch_pointwise_i8(logits8, (const uint8_t*)cur, f, n, 1024, MNV1_NUM_CLASSES, ctx->u8_wrap, ctx->num_sms, ctx->stream, &ctx->err);
          if (e == cudaSuccess) { e = mnv1::launch_u8_to_f32(d_logits, logits8, (long)n * MNV1_NUM_CLASSES, ctx->stream); ctx->launches++; }
          if (e == cudaSuccess && (d_top1 || d_prob)) {
            e = mnv1::launch_softmax(d_logits, n, MNV1_NUM_CLASSES, nullptr, d_top1, d_prob, ctx->stream);
            ctx->launches++;
          }
          cur = d_logits;
          break;
        }
        int nl = 0;
        e = mnv1::launch_head(ctx->dtype, cur, n, 49, 1024, f, ctx->d_pooled, d_logits, d_top1, d_prob,
                              MNV1_NUM_CLASSES, ctx->gather, ctx->stream, &nl);
        ctx->launches += nl; ctx->last_kernel = "head";
        cur = d_logits;
        break;
      }
    }
  }
  if (evs) cudaEventRecord(evs[last], ctx->stream);
  if (result) *result = cur;
  return e;
}

static int check_ready(mnv1_ctx* ctx, int n) {
  if (!ctx) return fail(nullptr, MNV1_EINVAL, "null ctx");
  if (!ctx->have_weights) return fail(ctx, MNV1_ESTATE, "weights not loaded (mnv1_set_weights / mnv1_load_weights)");
  if (n <= 0) return fail(ctx, MNV1_EINVAL, "batch must be positive");
  if (n > ctx->plan_batch) {
    int rc = mnv1_plan(ctx, n);
    if (rc) return rc;
  }
  if (ctx->dtype == MNV1_U8 && (ctx->in_scale != 1.f || ctx->in_bias != 0.f))
    return fail(ctx, MNV1_EUNSUPPORTED, "integer contexts read the raw u8 pixels (input transform must be 1, 0)");
  return prepare_stem(ctx, ctx->net[0]);
}

int mnv1_forward_device(mnv1_ctx* ctx, const void* d_images, int n, void* d_logits, void* d_top1, void* d_prob) {
  GUARD(ctx);
  int rc = check_ready(ctx, n);
  if (rc) return rc;
  if (!d_images || !d_logits) return fail(ctx, MNV1_EINVAL, "forward_device: images and logits are required");
  if (ctx->gather.n_dst && n > ctx->g_max) return fail(ctx, MNV1_EINVAL, "forward: batch exceeds this rank's rows of the gather block");
  if (!ctx->use_graph) {
    CK(ctx, enqueue_layers(ctx, (const uint8_t*)d_images, n, MNV1_NUM_LAYERS, (float*)d_logits, (int*)d_top1,
                           (float*)d_prob, nullptr, nullptr));
    return MNV1_OK;
  }
  GraphKey key{d_images, n, d_logits, d_top1, d_prob};
  auto it = ctx->graphs.find(key);
  if (it == ctx->graphs.end()) {
    if (ctx->graphs.size() >= 16) {  // bounded cache
      for (auto& kv : ctx->graphs) cudaGraphExecDestroy(kv.second);
      ctx->graphs.clear();
    }
    const long before = ctx->launches;
    cudaGraph_t graph = nullptr;
    CK(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    cudaError_t e = enqueue_layers(ctx, (const uint8_t*)d_images, n, MNV1_NUM_LAYERS, (float*)d_logits, (int*)d_top1,
                                   (float*)d_prob, nullptr, nullptr);
    cudaError_t e2 = cudaStreamEndCapture(ctx->stream, &graph);
    ctx->graph_kernels = ctx->launches - before;  // kernels per replay
    ctx->launches = before;                        // capture enqueues nothing
    if (e != cudaSuccess) { if (graph) cudaGraphDestroy(graph); return fail_cuda(ctx, e, "graph capture (layer launch)"); }
    if (e2 != cudaSuccess) return fail_cuda(ctx, e2, "cudaStreamEndCapture");
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return fail_cuda(ctx, e, "cudaGraphInstantiate");
    it = ctx->graphs.emplace(key, exec).first;
  }
  CK(ctx, cudaGraphLaunch(it->second, ctx->stream));
  ctx->launches += ctx->graph_kernels;
  return MNV1_OK;
}

int mnv1_ctx_use_fused_blocks(mnv1_ctx* ctx, int on) {
  GUARD(ctx);
  if (!ctx) return MNV1_EINVAL;
  ctx->use_fused = on != 0;
  for (auto& kv : ctx->graphs) cudaGraphExecDestroy(kv.second);
  ctx->graphs.clear();
  return MNV1_OK;
}

// one depthwise(3x3) + pointwise(1x1) block through the fused kernel (MNV1_EUNSUPPORTED when the
// shape has no fused variant): out [n][cout][rows/stride][cols/stride]
int mnv1_dw_pw_block(mnv1_ctx* ctx, mnv1_buf* out, const mnv1_buf* in, const mnv1_filter* dw, const mnv1_filter* pw,
                     int rows, int cols, int stride) {
  GUARD(ctx);
  if (!ctx || !out || !in || !dw || !pw) return fail(ctx, MNV1_EINVAL, "dw_pw_block: null argument");
  if (dw->kind != MNV1_DEPTHWISE || pw->kind != MNV1_POINTWISE || pw->cin != dw->cout)
    return fail(ctx, MNV1_EINVAL, "dw_pw_block: filter mismatch");
  if (ctx->dtype != MNV1_BF16) return fail(ctx, MNV1_EUNSUPPORTED, "dw_pw_block: bf16 contexts only");
  int rc = check_fmap(ctx, in, dw->cout, rows, cols, "dw_pw_block(in)");
  if (rc) return rc;
  rc = check_fmap(ctx, out, pw->cout, rows / stride, cols / stride, "dw_pw_block(out)");
  if (rc) return rc;
  if (in->n != out->n) return fail(ctx, MNV1_EINVAL, "dw_pw_block: batch mismatch");
  TimedLaunch tl(ctx);
  cudaError_t e = run_fused_block(ctx, out->d, in->d, dw, pw, in->n, rows, cols, stride, true);
  if (e == cudaErrorNotSupported) return fail(ctx, MNV1_EUNSUPPORTED, "dw_pw_block: no fused variant for this shape");
  if (e != cudaSuccess) return fail_cuda(ctx, e, "dw_pw_block");
  return MNV1_OK;
}

// fused[i] = 1 when mnv1_forward* runs layer i+1 (a depthwise) and the pointwise after it as one kernel
int mnv1_fused_layers(mnv1_ctx* ctx, int* fused) {
  GUARD(ctx);
  if (!ctx || !fused) return fail(ctx, MNV1_EINVAL, "fused_layers: null argument");
  if (!ctx->have_weights) return fail(ctx, MNV1_ESTATE, "weights not loaded (mnv1_set_weights / mnv1_load_weights)");
  const LayerDef* L = layer_defs();
  for (int i = 0; i < MNV1_NUM_LAYERS; ++i) {
    fused[i] = 0;
    if (ctx->use_fused && ctx->dtype == MNV1_BF16 && L[i].kind == MNV1_DEPTHWISE && i + 1 < MNV1_NUM_LAYERS &&
        L[i + 1].kind == MNV1_POINTWISE)
      fused[i] = ((mnv1::switches().fused_pair && mnv1::fused_pair_supported(ctx->net[i], ctx->net[i + 1], L[i].hin, L[i].hin, L[i].stride)) ||
                  mnv1::fused_dw_pw_supported(ctx->net[i], ctx->net[i + 1], L[i].hin, L[i].hin, L[i].stride)) ? 1 : 0;
  }
  return MNV1_OK;
}

int mnv1_ctx_use_graph(mnv1_ctx* ctx, int on) {
  GUARD(ctx);
  if (!ctx) return MNV1_EINVAL;
  ctx->use_graph = on != 0;
  return MNV1_OK;
}

static bool is_pinned(const void* p) {
  cudaPointerAttributes pa;
  if (cudaPointerGetAttributes(&pa, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return pa.type == cudaMemoryTypeHost;
}

static int finish_slot(mnv1_ctx* ctx, mnv1_ctx::Slot& sl) {
  if (!sl.busy) return MNV1_OK;
  CK(ctx, cudaEventSynchronize(sl.ev_done));
  if (sl.u_logits) memcpy(sl.u_logits, sl.h_logits, (size_t)sl.n * MNV1_NUM_CLASSES * 4);
  if (sl.u_top1) memcpy(sl.u_top1, sl.h_top1, (size_t)sl.n * 4);
  if (sl.u_prob) memcpy(sl.u_prob, sl.h_prob, (size_t)sl.n * 4);
  sl.busy = false;
  return MNV1_OK;
}

static int retire_all(mnv1_ctx* ctx) {
  for (auto& sl : ctx->slots) {
    int rc = finish_slot(ctx, sl);
    if (rc) return rc;
  }
  return MNV1_OK;
}

int mnv1_forward_submit(mnv1_ctx* ctx, const uint8_t* images, int n, float* logits, int* top1, float* top1_prob,
                        long* ticket) {
  GUARD(ctx);
  int rc = check_ready(ctx, n);
  if (rc) return rc;
  if (!images || !ticket) return fail(ctx, MNV1_EINVAL, "forward_submit: images / ticket is null");
  mnv1_ctx::Slot& sl = ctx->slots[ctx->next_ticket % mnv1_ctx::kSlots];
  if ((rc = finish_slot(ctx, sl)) != MNV1_OK) return rc;  // the batch submitted kSlots calls ago
  sl.n = n; sl.ticket = ctx->next_ticket;
  // images: straight from the caller's buffer when it is page-locked, else through the slot's staging copy
  const uint8_t* src = images;
  if (!is_pinned(images)) { memcpy(sl.h_images, images, (size_t)n * kImgBytes); src = sl.h_images; }
  CK(ctx, cudaMemcpyAsync(sl.d_images, src, (size_t)n * kImgBytes, cudaMemcpyHostToDevice, ctx->copy_stream));
  CK(ctx, cudaEventRecord(sl.ev_h2d, ctx->copy_stream));
  CK(ctx, cudaStreamWaitEvent(ctx->stream, sl.ev_h2d, 0));
  rc = mnv1_forward_device(ctx, sl.d_images, n, sl.d_logits, sl.d_top1, sl.d_prob);
  if (rc) return rc;
  CK(ctx, cudaEventRecord(sl.ev_compute, ctx->stream));
  CK(ctx, cudaStreamWaitEvent(ctx->d2h_stream, sl.ev_compute, 0));
  sl.u_logits = nullptr; sl.u_top1 = nullptr; sl.u_prob = nullptr;
  if (logits) {
    float* dst = logits;
    if (!is_pinned(logits)) { dst = sl.h_logits; sl.u_logits = logits; }
    CK(ctx, cudaMemcpyAsync(dst, sl.d_logits, (size_t)n * MNV1_NUM_CLASSES * 4, cudaMemcpyDeviceToHost, ctx->d2h_stream));
  }
  if (top1) {
    int* dst = top1;
    if (!is_pinned(top1)) { dst = sl.h_top1; sl.u_top1 = top1; }
    CK(ctx, cudaMemcpyAsync(dst, sl.d_top1, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->d2h_stream));
  }
  if (top1_prob) {
    float* dst = top1_prob;
    if (!is_pinned(top1_prob)) { dst = sl.h_prob; sl.u_prob = top1_prob; }
    CK(ctx, cudaMemcpyAsync(dst, sl.d_prob, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->d2h_stream));
  }
  CK(ctx, cudaEventRecord(sl.ev_done, ctx->d2h_stream));
  sl.busy = true;
  *ticket = ctx->next_ticket++;
  return MNV1_OK;
}

int mnv1_forward_wait(mnv1_ctx* ctx, long ticket) {
  GUARD(ctx);
  if (!ctx) return fail(nullptr, MNV1_EINVAL, "null ctx");
  if (ticket < 0) return fail(ctx, MNV1_EINVAL, "forward_wait: unknown ticket");
  mnv1_ctx::Slot& sl = ctx->slots[ticket % mnv1_ctx::kSlots];
  if (ticket < 0 || ticket >= ctx->next_ticket) return fail(ctx, MNV1_EINVAL, "forward_wait: unknown ticket");
  if (sl.ticket != ticket) return MNV1_OK;  // already retired (results delivered) by a later submit or a re-plan
  return finish_slot(ctx, sl);
}

int mnv1_forward(mnv1_ctx* ctx, const uint8_t* images, int n, float* logits, int* top1, float* top1_prob) {
  GUARD(ctx);
  long ticket = -1;
  int rc = mnv1_forward_submit(ctx, images, n, logits, top1, top1_prob, &ticket);
  if (rc) return rc;
  return mnv1_forward_wait(ctx, ticket);
}

int mnv1_forward_upto(mnv1_ctx* ctx, const uint8_t* images, int n, int last_layer, float* host_nchw) {
  GUARD(ctx);
  int rc = check_ready(ctx, n);
  if (rc) return rc;
  if (!images || !host_nchw || last_layer < 1 || last_layer > MNV1_NUM_LAYERS)
    return fail(ctx, MNV1_EINVAL, "forward_upto: bad arguments");
  if ((rc = retire_all(ctx)) != MNV1_OK) return rc;
  CK(ctx, cudaMemcpyAsync(ctx->d_images, images, (size_t)n * kImgBytes, cudaMemcpyHostToDevice, ctx->stream));
  const void* res = nullptr;
  CK(ctx, enqueue_layers(ctx, ctx->d_images, n, last_layer, ctx->d_logits, ctx->d_top1, ctx->d_prob, &res, nullptr));
  const LayerDef& L = layer_defs()[last_layer - 1];
  if (L.kind == MNV1_POOL || L.kind == MNV1_FC) {  // fp32 [n][C] already planar
    CK(ctx, cudaMemcpyAsync(host_nchw, res, (size_t)n * L.cout * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return MNV1_OK;
  }
  mnv1_buf view;
  view.d = const_cast<void*>(res); view.n = n; view.c = L.cout; view.h = L.hout; view.w = L.hout;
  view.bytes = (size_t)n * L.cout * L.hout * L.hout * elem_size(ctx->dtype); view.owned = false;
  return mnv1_download_planar(ctx, &view, host_nchw);
}

// layers 1..last_layer on device-resident images, eagerly, no copies (kernel bring-up, ncu, bench.py's kernel names)
int mnv1_forward_prefix_device(mnv1_ctx* ctx, const void* d_images, int n, int last_layer) {
  GUARD(ctx);
  int rc = check_ready(ctx, n);
  if (rc) return rc;
  if (!d_images || last_layer < 1 || last_layer > MNV1_NUM_LAYERS) return fail(ctx, MNV1_EINVAL, "forward_prefix_device: bad arguments");
  CK(ctx, enqueue_layers(ctx, (const uint8_t*)d_images, n, last_layer, ctx->d_logits, ctx->d_top1, ctx->d_prob, nullptr, nullptr));
  return MNV1_OK;
}

int mnv1_profile_layers(mnv1_ctx* ctx, const void* d_images, int n, int iters, float* times_ms) {
  GUARD(ctx);
  int rc = check_ready(ctx, n);
  if (rc) return rc;
  if (!d_images || !times_ms || iters <= 0) return fail(ctx, MNV1_EINVAL, "profile_layers: bad arguments");
  cudaEvent_t evs[MNV1_NUM_LAYERS + 1];
  for (auto& e : evs) CK(ctx, cudaEventCreate(&e));
  for (int i = 0; i < MNV1_NUM_LAYERS; ++i) times_ms[i] = 0.f;
  cudaError_t e = cudaSuccess;
  for (int it = 0; it < iters + 1 && e == cudaSuccess; ++it) {  // first pass is a warm-up
    e = enqueue_layers(ctx, (const uint8_t*)d_images, n, MNV1_NUM_LAYERS, ctx->d_logits, ctx->d_top1, ctx->d_prob,
                       nullptr, evs);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess && it > 0)
      for (int i = 0; i < MNV1_NUM_LAYERS; ++i) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, evs[i], evs[i + 1]);
        times_ms[i] += ms / iters;
      }
  }
  for (auto& ev : evs) cudaEventDestroy(ev);
  if (e != cudaSuccess) return fail_cuda(ctx, e, "profile_layers");
  return MNV1_OK;
}

// ---------------------------------------------------------------- data-parallel logits gather
static void drop_graphs(mnv1_ctx* ctx) {
  for (auto& kv : ctx->graphs) cudaGraphExecDestroy(kv.second);
  ctx->graphs.clear();
}
static size_t gather_bytes(int total_rows) { return (size_t)total_rows * (MNV1_NUM_CLASSES * 4 + 8); }
static void gather_rebuild(mnv1_ctx* ctx) {
  mnv1::HeadGather g{};
  const long total = (long)ctx->g_world * ctx->g_rows;
  for (int d = 0; d < ctx->g_world; ++d) {
    uint8_t* base = (uint8_t*)ctx->g_peer[d];
    if (!base) continue;
    g.logits[g.n_dst] = (float*)base;
    g.top1[g.n_dst] = (int*)(base + (size_t)total * MNV1_NUM_CLASSES * 4);
    g.prob[g.n_dst] = (float*)(base + (size_t)total * (MNV1_NUM_CLASSES * 4 + 4));
    ++g.n_dst;
  }
  g.row0 = ctx->g_first;
  ctx->gather = g;
  drop_graphs(ctx);                                        // captured graphs carry the old pointers
}
static void gather_release(mnv1_ctx* ctx) {
  for (int d = 0; d < 8; ++d) {
    if (ctx->g_peer[d] && ctx->g_peer_ipc[d]) cudaIpcCloseMemHandle(ctx->g_peer[d]);
    ctx->g_peer[d] = nullptr; ctx->g_peer_ipc[d] = false;
  }
  cudaFree(ctx->g_block); ctx->g_block = nullptr;
  ctx->g_world = ctx->g_rank = ctx->g_rows = 0;
  ctx->g_first = 0; ctx->g_max = 0;
  ctx->gather = mnv1::HeadGather{};
}

int mnv1_gather_create(mnv1_ctx* ctx, int world, int rank, int rows_per_rank) {
  GUARD(ctx);
  if (!ctx || world < 1 || world > 8 || rank < 0 || rank >= world || rows_per_rank <= 0)
    return fail(ctx, MNV1_EINVAL, "gather_create: need 1 <= world <= 8, 0 <= rank < world, rows_per_rank > 0");
  if (ctx->dtype == MNV1_U8) return fail(ctx, MNV1_EUNSUPPORTED, "gather: fp32 / bf16 contexts (the softmax kernel does the stores)");
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  gather_release(ctx);
  cudaError_t e = cudaMalloc(&ctx->g_block, gather_bytes(world * rows_per_rank));
  if (e != cudaSuccess) return fail(ctx, MNV1_ENOMEM, std::string("gather_create: ") + cudaGetErrorString(e));
  CK(ctx, cudaMemset(ctx->g_block, 0, gather_bytes(world * rows_per_rank)));
  ctx->g_world = world; ctx->g_rank = rank; ctx->g_rows = rows_per_rank;
  ctx->g_first = (long)rank * rows_per_rank; ctx->g_max = rows_per_rank;
  ctx->g_peer[rank] = ctx->g_block;
  gather_rebuild(ctx);
  return MNV1_OK;
}
int mnv1_gather_set_rows(mnv1_ctx* ctx, long first_row, int max_rows) {
  GUARD(ctx);
  if (!ctx || !ctx->g_block) return fail(ctx, MNV1_ESTATE, "gather_set_rows: no gather block");
  if (first_row < 0 || max_rows <= 0 || first_row + max_rows > (long)ctx->g_world * ctx->g_rows)
    return fail(ctx, MNV1_EINVAL, "gather_set_rows: the window leaves the block of world * rows_per_rank rows");
  ctx->g_first = first_row; ctx->g_max = max_rows;
  gather_rebuild(ctx);
  return MNV1_OK;
}

int mnv1_gather_export(mnv1_ctx* ctx, void* handle64) {
  GUARD(ctx);
  if (!ctx || !handle64 || !ctx->g_block) return fail(ctx, MNV1_ESTATE, "gather_export: call mnv1_gather_create first");
  static_assert(sizeof(cudaIpcMemHandle_t) == MNV1_IPC_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  CK(ctx, cudaIpcGetMemHandle(&h, ctx->g_block));
  memcpy(handle64, &h, sizeof h);
  return MNV1_OK;
}
int mnv1_gather_import(mnv1_ctx* ctx, int peer_rank, const void* handle64) {
  GUARD(ctx);
  if (!ctx || !handle64 || !ctx->g_block || peer_rank < 0 || peer_rank >= ctx->g_world || peer_rank == ctx->g_rank)
    return fail(ctx, MNV1_EINVAL, "gather_import: bad peer rank or no gather block");
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof h);
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return fail_cuda(ctx, e, "cudaIpcOpenMemHandle (is the peer GPU NVLink / P2P reachable?)");
  if (ctx->g_peer[peer_rank] && ctx->g_peer_ipc[peer_rank]) cudaIpcCloseMemHandle(ctx->g_peer[peer_rank]);
  ctx->g_peer[peer_rank] = p; ctx->g_peer_ipc[peer_rank] = true;
  gather_rebuild(ctx);
  return MNV1_OK;
}
int mnv1_gather_attach(mnv1_ctx* ctx, mnv1_ctx* peer) {
  GUARD(ctx);
  if (!ctx || !peer || !ctx->g_block || !peer->g_block || peer->g_world != ctx->g_world || peer->g_rows != ctx->g_rows ||
      peer->g_rank == ctx->g_rank)
    return fail(ctx, MNV1_EINVAL, "gather_attach: both contexts need gather blocks of the same geometry and different ranks");
  if (peer->device != ctx->device) {
    int can = 0;
    CK(ctx, cudaDeviceCanAccessPeer(&can, ctx->device, peer->device));
    if (!can) return fail(ctx, MNV1_EUNSUPPORTED, "gather_attach: no peer access between the two devices");
    cudaError_t e = cudaDeviceEnablePeerAccess(peer->device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
    else if (e != cudaSuccess) return fail_cuda(ctx, e, "cudaDeviceEnablePeerAccess");
  }
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->g_peer[peer->g_rank] = peer->g_block; ctx->g_peer_ipc[peer->g_rank] = false;
  gather_rebuild(ctx);
  return MNV1_OK;
}
int mnv1_gather_destroy(mnv1_ctx* ctx) {
  GUARD(ctx);
  if (!ctx) return fail(nullptr, MNV1_EINVAL, "null ctx");
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  gather_release(ctx);
  drop_graphs(ctx);
  return MNV1_OK;
}
int mnv1_gather_ptrs(mnv1_ctx* ctx, void** logits, void** top1, void** top1_prob) {
  if (!ctx || !ctx->g_block) return fail(ctx, MNV1_ESTATE, "gather_ptrs: no gather block");
  const size_t total = (size_t)ctx->g_world * ctx->g_rows;
  if (logits) *logits = ctx->g_block;
  if (top1) *top1 = ctx->g_block + total * MNV1_NUM_CLASSES * 4;
  if (top1_prob) *top1_prob = ctx->g_block + total * (MNV1_NUM_CLASSES * 4 + 4);
  return MNV1_OK;
}
int mnv1_ctx_device(const mnv1_ctx* ctx) { return ctx ? ctx->device : -1; }

// Per-launch device time INSIDE the replayed graph.  For every launch boundary k of the schedule the graph of
// layers 1..k is captured and replayed `iters` times; cum_ms[k-1] is the median replay time (CUDA events on
// the context stream), -1 for a layer that ends inside a fused launch (a depthwise fused with its pointwise,
// the pool inside the head).  The difference of two consecutive boundaries is that launch's time with the
// programmatic-launch overlap of the real step, and the differences add up to the whole step by construction.
int mnv1_profile_prefixes(mnv1_ctx* ctx, const void* d_images, int n, int iters, float* cum_ms) {
  GUARD(ctx);
  int rc = check_ready(ctx, n);
  if (rc) return rc;
  if (!d_images || !cum_ms || iters <= 0) return fail(ctx, MNV1_EINVAL, "profile_prefixes: bad arguments");
  int fused[MNV1_NUM_LAYERS];
  if ((rc = mnv1_fused_layers(ctx, fused)) != MNV1_OK) return rc;
  const LayerDef* L = layer_defs();
  const long launches_before = ctx->launches;
  std::vector<int> cuts;
  std::vector<cudaGraphExec_t> execs;
  cudaError_t e = cudaSuccess;
  for (int k = 1; k <= MNV1_NUM_LAYERS && e == cudaSuccess; ++k) {
    cum_ms[k - 1] = -1.f;
    if (fused[k - 1]) continue;                                               // ends inside a fused dw+pw launch
    if (L[k - 1].kind == MNV1_POOL && ctx->dtype == MNV1_BF16) continue;       // one "head" row for pool + FC + softmax
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    e = cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) break;
    cudaError_t e1 = enqueue_layers(ctx, (const uint8_t*)d_images, n, k, ctx->d_logits, ctx->d_top1, ctx->d_prob, nullptr, nullptr);
    e = cudaStreamEndCapture(ctx->stream, &graph);
    if (e1 != cudaSuccess) e = e1;
    if (e == cudaSuccess) e = cudaGraphInstantiate(&exec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (e != cudaSuccess) break;
    cuts.push_back(k);
    execs.push_back(exec);
  }
  const int nc = (int)cuts.size();
  std::vector<cudaEvent_t> evs(2 * (size_t)nc);
  for (auto& ev : evs) if (e == cudaSuccess) e = cudaEventCreate(&ev);
  // Every repetition replays all the prefixes back to back, so that the two times a difference is made of are
  // taken milliseconds apart, under the same clocks; the per-launch time is the median over the repetitions
  // of that difference.
  std::vector<std::vector<float>> diff(nc, std::vector<float>(iters));
  for (int it = -2; it < iters && e == cudaSuccess; ++it) {                 // two warm-up rounds
    for (int j = 0; j < nc && e == cudaSuccess; ++j) {
      cudaEventRecord(evs[2 * j], ctx->stream);
      e = cudaGraphLaunch(execs[j], ctx->stream);
      cudaEventRecord(evs[2 * j + 1], ctx->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    // A repetition is a ~10 ms burst; pausing between bursts keeps every one of them at the clocks a short timed
    // loop sees (a B200 under a sustained load settles some 5 % lower under its power cap), so that the launch
    // times add up to what `K` back-to-back steps measure.
    std::this_thread::sleep_for(std::chrono::milliseconds(25));
    if (e != cudaSuccess || it < 0) continue;
    float prev = 0.f;
    for (int j = 0; j < nc; ++j) {
      float t = 0.f;
      cudaEventElapsedTime(&t, evs[2 * j], evs[2 * j + 1]);
      diff[j][it] = t - prev;
      prev = t;
    }
  }
  if (e == cudaSuccess) {
    float cum = 0.f;
    for (int j = 0; j < nc; ++j) {
      std::sort(diff[j].begin(), diff[j].end());
      cum += diff[j][iters / 2];
      cum_ms[cuts[j] - 1] = cum;
    }
  }
  for (auto ex : execs) cudaGraphExecDestroy(ex);
  for (auto& ev : evs) if (ev) cudaEventDestroy(ev);
  ctx->launches = launches_before;
  if (e != cudaSuccess) return fail_cuda(ctx, e, "profile_prefixes");
  return MNV1_OK;
}

// What a plain pinned cudaMemcpyAsync reaches on this box: `reps` back-to-back host-to-device copies of `bytes` on
// the context's copy stream, CUDA events.  The copies rotate over THREE pinned source buffers, like the three
// batches mnv1_forward_submit keeps in flight: a single buffer copied again and again is served from the CPU's
// last-level cache and overstates what the host's DRAM can feed the link.  The ceiling of mnv1_forward's upload.
struct mnv1_h2d_probe_t {
  mnv1_ctx* ctx = nullptr;
  size_t bytes = 0;
  void* h[3] = {nullptr, nullptr, nullptr};
  void* d = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
};

int mnv1_h2d_probe_close(mnv1_h2d_probe_t* p) {
  if (!p) return MNV1_OK;
  GUARD(p->ctx);
  if (p->e0) cudaEventDestroy(p->e0);
  if (p->e1) cudaEventDestroy(p->e1);
  cudaFree(p->d);
  for (int i = 0; i < 3; ++i) cudaFreeHost(p->h[i]);
  delete p;
  return MNV1_OK;
}

int mnv1_h2d_probe_open(mnv1_ctx* ctx, size_t bytes, mnv1_h2d_probe_t** out) {
  GUARD(ctx);
  if (!ctx || !bytes || !out) return fail(ctx, MNV1_EINVAL, "h2d_probe: bad arguments");
  mnv1_h2d_probe_t* p = new mnv1_h2d_probe_t;
  p->ctx = ctx; p->bytes = bytes;
  cudaError_t e = cudaMalloc(&p->d, bytes);
  for (int i = 0; i < 3 && e == cudaSuccess; ++i) {
    e = cudaHostAlloc(&p->h[i], bytes, mnv1::switches().h2d_wc ? cudaHostAllocWriteCombined : cudaHostAllocDefault);
    if (e == cudaSuccess) memset(p->h[i], i + 1, bytes);
  }
  if (e == cudaSuccess) e = cudaEventCreate(&p->e0);
  if (e == cudaSuccess) e = cudaEventCreate(&p->e1);
  if (e != cudaSuccess) { mnv1_h2d_probe_close(p); return fail_cuda(ctx, e, "h2d_probe_open"); }
  *out = p;
  return MNV1_OK;
}

int mnv1_h2d_probe_run(mnv1_h2d_probe_t* p, int reps, float* gbytes_per_s) {
  if (!p || reps <= 0 || !gbytes_per_s) return MNV1_EINVAL;
  mnv1_ctx* ctx = p->ctx;
  GUARD(ctx);
  cudaError_t e = cudaSuccess;
  for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaMemcpyAsync(p->d, p->h[i], p->bytes, cudaMemcpyHostToDevice, ctx->copy_stream);
  if (e == cudaSuccess) e = cudaEventRecord(p->e0, ctx->copy_stream);
  for (int i = 0; i < reps && e == cudaSuccess; ++i)
    e = cudaMemcpyAsync(p->d, p->h[(i + 2) % 3], p->bytes, cudaMemcpyHostToDevice, ctx->copy_stream);
  if (e == cudaSuccess) e = cudaEventRecord(p->e1, ctx->copy_stream);
  if (e == cudaSuccess) e = cudaEventSynchronize(p->e1);
  float ms = 0.f;
  if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, p->e0, p->e1);
  if (e != cudaSuccess) return fail_cuda(ctx, e, "h2d_probe_run");
  *gbytes_per_s = (float)((double)p->bytes * reps / (ms * 1e-3) / 1e9);
  return MNV1_OK;
}

int mnv1_h2d_probe(mnv1_ctx* ctx, size_t bytes, int reps, float* gbytes_per_s) {
  if (!ctx || !bytes || reps <= 0 || !gbytes_per_s) return fail(ctx, MNV1_EINVAL, "h2d_probe: bad arguments");
  mnv1_h2d_probe_t* p = nullptr;
  int rc = mnv1_h2d_probe_open(ctx, bytes, &p);
  if (rc) return rc;
  rc = mnv1_h2d_probe_run(p, reps, gbytes_per_s);
  mnv1_h2d_probe_close(p);
  return rc;
}

int mnv1_synth_images_device(mnv1_ctx* ctx, void* d_images, int n, long first, uint64_t seed) {
  GUARD(ctx);
  if (!ctx || !d_images || n < 0 || first < 0) return fail(ctx, MNV1_EINVAL, "synth_images: bad arguments");
  ctx->launches++; ctx->last_kernel = "synth_images_kernel";
  CK(ctx, mnv1::launch_synth_images((uint8_t*)d_images, first * (long)kImgBytes, (long)n * (long)kImgBytes, seed,
                                    ctx->stream));
  return MNV1_OK;
}

// pinned host memory for callers that want mnv1_forward to copy straight from their buffer
// (CL_MEM_ALLOC_HOST_PTR analogue)
int mnv1_host_alloc(size_t bytes, void** out) {
  if (!out) return MNV1_EINVAL;
  cudaError_t e = cudaMallocHost(out, bytes);
  if (e != cudaSuccess) return fail(nullptr, MNV1_ENOMEM, std::string("cudaMallocHost: ") + cudaGetErrorString(e));
  return MNV1_OK;
}
int mnv1_host_free(void* p) { cudaFreeHost(p); return MNV1_OK; }

}  // extern "C"

/*
 * mnv1.h — C-ABI of the B200-native MobileNet-V1 1.0-224 inference path.
 *
 * Drop-in boundary for the reference's host programs (MobileNet.c, MobileNet_13Layers.c,
 * MobileNet_L5.c).  The reference has no library API; its seam is the OpenCL host call
 * sequence repeated for every layer (MobileNet.c:322-408 is one instance) plus the four
 * kernel argument lists of kernel.cl.  Each entry point below names the reference lines it
 * replaces.  Plain C: opaque handles, plain pointers and sizes, int error codes, no
 * exceptions cross the boundary.  INTEGRATION.md shows the edits a maintainer makes in
 * MobileNet.c to bind to it.
 *
 * Conventions
 *  - every function returns 0 (MNV1_OK) or a negative MNV1_E* code; mnv1_last_error()
 *    gives the text (the reference's `printf("Error: ...") ; exit(1)` stays at the call site)
 *  - host tensors use the reference's layout: planar [N][C][H][W]
 *    (kernel.cl:14,56,73,90,103,107), filters OIHW / [C][3][3] / [Cout][Cin] in `findex`
 *    order (kernel.cl:18,30,42,77,106).  Device tensors are opaque (NHWC inside).
 *  - launches are asynchronous on the context's stream; download / softmax / forward /
 *    *_time_ms synchronise.  One context per GPU; a context is not thread-safe.
 */
#ifndef MNV1_H
#define MNV1_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MNV1_OK 0
#define MNV1_EINVAL (-1)       /* bad argument / shape mismatch                         */
#define MNV1_ECUDA (-2)        /* CUDA runtime or driver error (see mnv1_last_error)     */
#define MNV1_ENOMEM (-3)       /* host or device allocation failed                      */
#define MNV1_EIO (-4)          /* weight / image file could not be read                 */
#define MNV1_ESTATE (-5)       /* call order: weights not loaded, plan missing, ...      */
#define MNV1_EUNSUPPORTED (-6) /* shape outside what the sm_100a kernels implement       */

/* MNV1_U8: the reference's own integer arithmetic (kernel.cl:2-3,62,94: unsigned char maps x int filters summed
 * in an int, `if (sum <= 0) sum = 0`, stored to unsigned char): u8 activations, s8 filters, s32 accumulation on
 * tcgen05.mma.kind::i8 / DP4A; see mnv1_ctx_set_u8_store and the note at mnv1_filter_create. */
typedef enum { MNV1_F32 = 0, MNV1_BF16 = 1, MNV1_U8 = 2 } mnv1_dtype;
typedef enum { MNV1_ACT_NONE = 0, MNV1_ACT_RELU = 1, MNV1_ACT_RELU6 = 2 } mnv1_act;
/* padding of stride-2 layers: REF = 1 px top/left (kernel.cl:20,79 `< 0` test),
 * TFSAME = Keras/TF "SAME" on even sizes (0 top/left, 1 bottom/right). Stride 1: 1 px all round. */
typedef enum { MNV1_PAD_REF = 0, MNV1_PAD_TFSAME = 1 } mnv1_pad;
typedef enum {
  MNV1_CONVOLUTE = 0, /* kernel.cl:2   */
  MNV1_DEPTHWISE = 1, /* kernel.cl:62  */
  MNV1_POINTWISE = 2, /* kernel.cl:94  */
  MNV1_POOL = 3,      /* kernel.cl:116 */
  MNV1_FC = 4         /* pointwise reused at 1x1, MobileNet.c:2689 */
} mnv1_kind;

typedef struct mnv1_ctx mnv1_ctx;       /* replaces cl_context + cl_command_queue + cl_program */
typedef struct mnv1_buf mnv1_buf;       /* replaces cl_mem for images / feature maps           */
typedef struct mnv1_filter mnv1_filter; /* replaces cl_mem d_filter (+ folded BN, activation)  */

/* ---- context: replaces platform/device/context/queue/program bring-up (MobileNet.c:147-205)
 *      and the release calls (MobileNet.c:2809-2831) --------------------------------------- */
int mnv1_ctx_create(int device, mnv1_dtype dtype, mnv1_ctx** ctx);
int mnv1_ctx_destroy(mnv1_ctx* ctx);
const char* mnv1_last_error(const mnv1_ctx* ctx); /* ctx may be NULL (creation failures) */
int mnv1_ctx_set_stream(mnv1_ctx* ctx, void* cuda_stream);  /* cudaStream_t owned by the caller */
int mnv1_ctx_set_pad_mode(mnv1_ctx* ctx, mnv1_pad pad);     /* default MNV1_PAD_REF */
/* stem maps every u8 pixel x -> x*scale + bias before the conv; default 1, 0 (raw integers,
 * as kernel.cl reads them); Keras MobileNet preprocessing is 1/127.5, -1 */
int mnv1_ctx_set_input_transform(mnv1_ctx* ctx, float scale, float bias);
int mnv1_sync(mnv1_ctx* ctx);                               /* clFinish, MobileNet.c:302 */
/* device time of the most recent kernel launched through this context, in ms: replaces
 * clGetEventProfilingInfo(COMMAND_START/END), MobileNet.c:303-305 (which prints seconds) */
int mnv1_last_kernel_ms(mnv1_ctx* ctx, float* ms);
int mnv1_ctx_enable_timing(mnv1_ctx* ctx, int on); /* default on; off removes the event pair */

/* ---- buffers: replace clCreateBuffer / clEnqueueWriteBuffer / clEnqueueReadBuffer
 *      (MobileNet.c:340-342,350-351,395) ---------------------------------------------------- */
int mnv1_malloc(mnv1_ctx* ctx, int n, int c, int h, int w, mnv1_buf** buf); /* feature map */
int mnv1_malloc_u8(mnv1_ctx* ctx, size_t bytes, mnv1_buf** buf);            /* image planes */
int mnv1_free(mnv1_ctx* ctx, mnv1_buf* buf);
int mnv1_upload_u8(mnv1_ctx* ctx, mnv1_buf* buf, const uint8_t* host, size_t bytes);
int mnv1_upload_planar(mnv1_ctx* ctx, mnv1_buf* buf, const float* host_nchw);
int mnv1_download_planar(mnv1_ctx* ctx, mnv1_buf* buf, float* host_nchw);
void* mnv1_buf_device_ptr(mnv1_buf* buf);
/* integer contexts: planar u8 host arrays, the reference's own `unsigned char*` layout (MobileNet.c:116-143,
 * :350 clEnqueueWriteBuffer, :395 clEnqueueReadBuffer), no fp32 detour */
int mnv1_upload_planar_u8(mnv1_ctx* ctx, mnv1_buf* buf, const uint8_t* host_nchw);
int mnv1_download_planar_u8(mnv1_ctx* ctx, mnv1_buf* buf, uint8_t* host_nchw);
/* integer contexts: how an s32 result becomes the u8 that is stored.  wrap = 1: the C conversion to unsigned
 * char, i.e. modulo 256 — literally what kernel.cl:56,90,112 do; wrap = 0 (default): saturate to [0, 255]. */
int mnv1_ctx_set_u8_store(mnv1_ctx* ctx, int wrap);

/* ---- filters: replace readSquezeNetKernel + clCreateBuffer(d_filter) (MobileNet.c:31-47,
 *      241,248).  `w` is in the reference's flat order for that kernel kind; scale/shift
 *      (per output channel, may be NULL) are a folded BatchNorm or, for MNV1_FC, the bias. -- */
/* Integer contexts (MNV1_U8): `w` must hold integers in [-128, 127] (the int8_t filter buffers of
 * MobileNet.c:116,248); out = store_u8(act(acc + shift[c]) >> s) with shift an integer bias and scale = 2^-s the
 * same power of two for every channel (NULL: s = 0).  scale = shift = NULL, MNV1_ACT_RELU and the wrapping store
 * reproduce kernel.cl bit for bit. */
int mnv1_filter_create(mnv1_ctx* ctx, mnv1_kind kind, const float* w, int cin, int cout,
                       const float* scale, const float* shift, mnv1_act act,
                       mnv1_filter** filter);
int mnv1_filter_destroy(mnv1_ctx* ctx, mnv1_filter* filter);

/* ---- the four kernels; after `ctx` the argument order is kernel.cl's --------------------- */
/* kernel.cl:2-3 / clSetKernelArg at MobileNet.c:272-281.  in_r/g/b: u8 planes [n][rows][cols];
 * out: [n][op_size][rows/stride][cols/stride]. */
int mnv1_convolute(mnv1_ctx* ctx, mnv1_buf* out, const mnv1_buf* in_r, const mnv1_buf* in_g,
                   const mnv1_buf* in_b, const mnv1_filter* filter, int rows, int cols,
                   int filtersize, int stride, int op_size);
/* same, on the interleaved RGB payload `image[]` (MobileNet.c:29) without the host-side split
 * of MobileNet.c:218-238: in_rgb is u8 [n][rows][cols][3] */
int mnv1_convolute_rgb(mnv1_ctx* ctx, mnv1_buf* out, const mnv1_buf* in_rgb,
                       const mnv1_filter* filter, int rows, int cols, int filtersize, int stride,
                       int op_size);
/* kernel.cl:62 / MobileNet.c:363-370 */
int mnv1_depthwise(mnv1_ctx* ctx, mnv1_buf* out, const mnv1_buf* in, const mnv1_filter* filter,
                   int rows, int cols, int filtersize, int stride, int op_size);
/* kernel.cl:94 / MobileNet.c:453-459.  filtersize = number of input planes contracted (Cin);
 * the reference passes K_P = 1 — pass Cin. Also the FC layer (rows=cols=1). */
int mnv1_pointwise(mnv1_ctx* ctx, mnv1_buf* out, const mnv1_buf* in, const mnv1_filter* filter,
                   int rows, int cols, int filtersize, int op_size);
/* kernel.cl:116 / MobileNet.c:2640-2645.  filtersize x filtersize global average. */
int mnv1_pool(mnv1_ctx* ctx, mnv1_buf* out, const mnv1_buf* in, int rows, int cols,
              int filtersize, int op_size);
/* host softmax + argmax of MobileNet.c:2769-2792, on the device.  logits: [n][classes][1][1].
 * prob [n][classes] / top1 [n] (0-based; the reference prints location = index + 1) /
 * top1_prob [n] are host arrays, any may be NULL. */
int mnv1_softmax(mnv1_ctx* ctx, const mnv1_buf* logits, int classes, float* prob, int* top1,
                 float* top1_prob);

/* ---- whole network: the 29-block schedule of MobileNet.c:207-2763 in one call ------------- */
#define MNV1_NUM_LAYERS 29
#define MNV1_NUM_CLASSES 1000
#define MNV1_TOTAL_WEIGHTS 4209088L /* sum of the readSquezeNetKernel counts, SURVEY App. A */
#define MNV1_BN_CHANNELS 10944L
typedef struct {
  int index, kind, cin, cout, hin, hout, stride;
  long w_off, w_cnt, c_off;
} mnv1_layer_info;
int mnv1_layer_table(mnv1_layer_info* out29);
/* weights: MNV1_TOTAL_WEIGHTS floats in file order; scale/shift: MNV1_BN_CHANNELS + 1000
 * (last 1000 = FC bias in shift), may be NULL (=1 / =0).  Replaces the 29 readSquezeNetKernel
 * calls.  act applies to layers 1-27. */
int mnv1_set_weights(mnv1_ctx* ctx, const float* weights, const float* scale, const float* shift,
                     mnv1_act act);
/* weight file: either the reference's text format (whitespace separated decimals,
 * MobileNet.c:37-44, read sequentially instead of re-opened per layer) or "MNV1WTS1" binary
 * (see csrc/weights_io.cpp).  A file with only MNV1_TOTAL_WEIGHTS values gets scale=1, shift=0. */
int mnv1_load_weights(mnv1_ctx* ctx, const char* path, mnv1_act act);
int mnv1_save_weights_bin(const char* path, const float* weights, const float* scale,
                          const float* shift);
/* host-only parse of a weight file (text or binary) into caller arrays: weights
 * [MNV1_TOTAL_WEIGHTS], scale / shift [MNV1_BN_CHANNELS + 1000].  Needs no GPU. */
int mnv1_parse_weights(const char* path, float* weights, float* scale, float* shift);
/* P6 PPM reader that skips the header (fixes SURVEY App. C D-15); out = 224*224*3 bytes */
int mnv1_read_ppm(const char* path, uint8_t* out_rgb, int height, int width);
/* allocate the activation arena for batches up to max_batch and capture the CUDA graph */
int mnv1_plan(mnv1_ctx* ctx, int max_batch);
/* host in, host out.  images: u8 [n][224][224][3] (PPM payload order); logits [n][1000] fp32,
 * top1 [n], top1_prob [n] — any output may be NULL.  Includes H2D, the 28 launches, D2H. */
int mnv1_forward(mnv1_ctx* ctx, const uint8_t* images, int n, float* logits, int* top1,
                 float* top1_prob);
/* the same, split so that consecutive batches overlap: submit enqueues H2D (copy stream), the
 * graph (context stream) and D2H (third stream) and returns; wait blocks until that batch's
 * outputs are in the caller's arrays.  Three batches may be in flight; submitting a fourth first
 * retires the oldest.  Page-locked caller buffers (mnv1_host_alloc) are copied from / to directly. */
int mnv1_forward_submit(mnv1_ctx* ctx, const uint8_t* images, int n, float* logits, int* top1,
                        float* top1_prob, long* ticket);
int mnv1_forward_wait(mnv1_ctx* ctx, long ticket);
/* device in, device out, asynchronous on the context stream (no copies, no sync) */
int mnv1_forward_device(mnv1_ctx* ctx, const void* d_images_u8, int n, void* d_logits_f32,
                        void* d_top1_i32, void* d_top1_prob_f32);
/* run layers 1..last_layer eagerly and copy layer `last_layer`'s output to host planar fp32
 * (the per-layer parity dump; configs "first 5 layers" / "first 13 layers") */
int mnv1_forward_upto(mnv1_ctx* ctx, const uint8_t* images, int n, int last_layer,
                      float* host_nchw);
/* layers 1..last_layer on device-resident images, eagerly on the context stream, nothing copied */
int mnv1_forward_prefix_device(mnv1_ctx* ctx, const void* d_images_u8, int n, int last_layer);
/* per-layer device time (ms) of the last mnv1_profile_layers run; times[29] */
int mnv1_profile_layers(mnv1_ctx* ctx, const void* d_images_u8, int n, int iters, float* times_ms);
/* per-launch device time inside the replayed CUDA graph: cum_ms[k-1] = median time of the graph of layers
 * 1..k over `iters` replays, for every k at which a launch of the schedule ends (-1 elsewhere: a depthwise
 * fused with its pointwise, the pool inside the head kernel).  Consecutive differences are per-launch times
 * that add up to the whole step; what MobileNet.c:303-305,315 printed per layer, without the host round trip. */
int mnv1_profile_prefixes(mnv1_ctx* ctx, const void* d_images_u8, int n, int iters, float* cum_ms29);
/* GB/s of `reps` back-to-back pinned cudaMemcpyAsync host-to-device copies of `bytes` (the ceiling of the
 * upload inside mnv1_forward; replaces nothing in the reference, whose clEnqueueWriteBuffer is blocking) */
int mnv1_h2d_probe(mnv1_ctx* ctx, size_t bytes, int reps, float* gbytes_per_s);
/* the same in three steps, so that several ranks can allocate first (open: device buffer + three pinned source
 * buffers), then start copying TOGETHER (run, after the caller's barrier; may be repeated), then free (close) */
typedef struct mnv1_h2d_probe_t mnv1_h2d_probe_t;
int mnv1_h2d_probe_open(mnv1_ctx* ctx, size_t bytes, mnv1_h2d_probe_t** probe);
int mnv1_h2d_probe_run(mnv1_h2d_probe_t* probe, int reps, float* gbytes_per_s);
int mnv1_h2d_probe_close(mnv1_h2d_probe_t* probe);
/* fill d_images (u8 [n][224][224][3]) with the synthetic stream of SURVEY §8(d): images
 * first..first+n-1 of seed `seed`, generated on the device */
int mnv1_synth_images_device(mnv1_ctx* ctx, void* d_images_u8, int n, long first, uint64_t seed);
/* number of kernels launched through this context so far (bench.py's gpu_launches) */
long mnv1_launch_count(const mnv1_ctx* ctx);
/* name of the kernel the most recent entry point launched ("pointwise_tc_kernel", ...) */
const char* mnv1_last_kernel_name(const mnv1_ctx* ctx);
/* the CUDA-core GEMM that fp32 contexts always use, callable on any context: lets a test
 * cross-check the tcgen05 kernel on the device.  Same arguments as mnv1_pointwise. */
int mnv1_pointwise_simt(mnv1_ctx* ctx, mnv1_buf* out, const mnv1_buf* in, const mnv1_filter* filter,
                        int rows, int cols, int filtersize, int op_size);
/* depthwise(3x3) -> pointwise(1x1) as one kernel: the depthwise map stays in shared memory as the
 * GEMM's A operand (bf16 contexts; MNV1_EUNSUPPORTED when the shape has no fused variant).
 * mnv1_forward* uses it automatically; mnv1_ctx_use_fused_blocks(ctx, 0) turns that off. */
int mnv1_dw_pw_block(mnv1_ctx* ctx, mnv1_buf* out, const mnv1_buf* in, const mnv1_filter* dw_filter,
                     const mnv1_filter* pw_filter, int rows, int cols, int stride);
int mnv1_ctx_use_fused_blocks(mnv1_ctx* ctx, int on);
/* fused[i] (i = 0..MNV1_NUM_LAYERS-1) = 1 when mnv1_forward* runs depthwise layer i+1 and the pointwise
 * layer after it as one kernel (weights must be loaded); per-layer reports then merge the two rows */
int mnv1_fused_layers(mnv1_ctx* ctx, int* fused);
/* 1 (default): mnv1_forward* replays a captured CUDA graph; 0: launches the kernels eagerly */
int mnv1_ctx_use_graph(mnv1_ctx* ctx, int on);
/* page-locked host memory (CL_MEM_ALLOC_HOST_PTR analogue): mnv1_forward copies straight
 * from / to such buffers instead of staging */
int mnv1_host_alloc(size_t bytes, void** out);
int mnv1_host_free(void* p);
const char* mnv1_version(void);

/* ---- data parallelism over the batch (BASELINE config 5).  The reference is batch 1 on one
 *      queue (MobileNet.c:29,171); images are independent, so the batch is cut into contiguous shards,
 *      one context per GPU, weights replicated, no collective on the data path.  The only exchange is the
 *      logits gather, and it is done by the head kernel itself: every rank stores its rows straight into
 *      the gather blocks of all ranks through peer-mapped pointers over NVLink (no collective kernel). --- */
#define MNV1_IPC_HANDLE_BYTES 64
/* allocate this rank's gather block: rows [world * rows_per_rank] of logits[1000] f32, top1 i32, top1_prob f32;
 * from now on mnv1_forward* on this context (n <= rows_per_rank) also writes rows rank*rows_per_rank + i of
 * every attached / imported block.  bf16 contexts. */
int mnv1_gather_create(mnv1_ctx* ctx, int world, int rank, int rows_per_rank);
/* uneven shards: this rank's rows land at [first_row, first_row + max_rows) of the block instead (forward needs
 * n <= max_rows); the block keeps world * rows_per_rank rows, e.g. the shards of mnv1_dp_shard_weighted */
int mnv1_gather_set_rows(mnv1_ctx* ctx, long first_row, int max_rows);
/* peers in the SAME process (cudaDeviceEnablePeerAccess): make ctx store into peer's block */
int mnv1_gather_attach(mnv1_ctx* ctx, mnv1_ctx* peer);
/* peers in OTHER processes (one process per GPU): exchange the 64-byte CUDA IPC handle by any means */
int mnv1_gather_export(mnv1_ctx* ctx, void* handle64);
int mnv1_gather_import(mnv1_ctx* ctx, int peer_rank, const void* handle64);
/* device pointers of this rank's gathered arrays; complete once every rank's forward has finished */
int mnv1_gather_ptrs(mnv1_ctx* ctx, void** d_logits, void** d_top1, void** d_top1_prob);
int mnv1_gather_destroy(mnv1_ctx* ctx);
int mnv1_ctx_device(const mnv1_ctx* ctx);

/* one process driving several GPUs: a context and a worker thread per device, the gather wired all-to-all */
typedef struct mnv1_dp mnv1_dp;
int mnv1_dp_create(const int* devices, int n_devices, mnv1_dtype dtype, int max_batch_per_gpu, mnv1_dp** dp);
int mnv1_dp_destroy(mnv1_dp* dp);
const char* mnv1_dp_last_error(const mnv1_dp* dp);
int mnv1_dp_size(const mnv1_dp* dp);
mnv1_ctx* mnv1_dp_ctx(mnv1_dp* dp, int rank); /* for per-context settings; do not destroy */
int mnv1_dp_set_pad_mode(mnv1_dp* dp, mnv1_pad pad);
int mnv1_dp_set_input_transform(mnv1_dp* dp, float scale, float bias);
int mnv1_dp_set_weights(mnv1_dp* dp, const float* weights, const float* scale, const float* shift, mnv1_act act);
int mnv1_dp_load_weights(mnv1_dp* dp, const char* path, mnv1_act act);
/* shard of rank r of g for a batch of n: images [first, first + count) — contiguous, the remainder spread
 * over the first ranks */
int mnv1_dp_shard(int n, int rank, int world, int* first, int* count);
/* the same with shards in proportion to weight[0..world) (> 0; NULL = equal): floor of the exact share, the
 * leftover images to the largest fractions.  For boxes that feed their GPUs unevenly (see mnv1_dp_calibrate). */
int mnv1_dp_shard_weighted(int n, int rank, int world, const float* weight, int* first, int* count);
/* shard weights of mnv1_dp_forward* (NULL = equal shards again); every shard must still fit max_batch_per_gpu */
int mnv1_dp_set_shard_weights(mnv1_dp* dp, const float* weight);
/* measure each GPU's pinned host->device rate with ALL of them copying at once (mnv1_h2d_probe_*; a second pass
 * gives every GPU a copy count in proportion to its first rate, so that all finish together and the rates are those of
 * the steady state) and use the rates as shard weights; gbytes_per_s (may be NULL) receives the n_devices rates */
int mnv1_dp_calibrate(mnv1_dp* dp, float* gbytes_per_s);
/* host in / host out like mnv1_forward, the batch cut into contiguous shards; every GPU copies its shard of
 * the outputs straight into the caller's arrays.  submit/wait as for mnv1_forward_submit/_wait. */
int mnv1_dp_forward(mnv1_dp* dp, const uint8_t* images, int n, float* logits, int* top1, float* top1_prob);
int mnv1_dp_forward_submit(mnv1_dp* dp, const uint8_t* images, int n, float* logits, int* top1, float* top1_prob,
                           long* ticket);
int mnv1_dp_forward_wait(mnv1_dp* dp, long ticket);
/* device in / device out: rank r runs its n_per_gpu images (device pointer d_images[r], on device r);
 * when this returns every rank's gather block (mnv1_gather_ptrs on mnv1_dp_ctx(dp, r)) holds all
 * world * n_per_gpu rows. */
int mnv1_dp_forward_device(mnv1_dp* dp, const void* const* d_images, int n_per_gpu);

#ifdef __cplusplus
}
#endif
#endif /* MNV1_H */

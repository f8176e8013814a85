/*
 * mnv1_oracle.h — CPU oracle for the MobileNet-V1 1.0-224 hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked, imported or
 * executed by the product (libmnv1.so, the host drivers, the Python binding).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use it, and only as the checker / CPU baseline.
 *
 * What it restates (reference = anerisheth19/CNN-MobileNet-V1-...-OpenCL):
 *   kernel.cl:2-60    convolute   (stem 3x3, three planar R/G/B inputs)
 *   kernel.cl:62-92   depthwise   (per-channel 3x3)
 *   kernel.cl:94-114  pointwise   (1x1 conv / FC)
 *   kernel.cl:116-131 pool        (global average, /49 generalised to /(f*f))
 *   MobileNet.c:2769-2792  host softmax + argmax
 *   MobileNet.c:207-2763   the 29-layer schedule
 *
 * Semantics: the reference's INTENDED network (SURVEY.md §0, App. C lists the
 * defects that are deliberately not reproduced): accumulator reset per output
 * channel, depthwise reads its own input channel, pointwise contracts over all
 * Cin, stride is a stride (not a dilation), zero padding on every border.
 * Layout is the reference's: planar [N][C][H][W], weights OIHW / [C][3][3] /
 * [Cout][Cin], tap order i (row) outer, j (col) inner, R then G then B.
 *
 * PARITY PINNING: the reference ships no golden vectors, inputs or tests
 * ("parity unpinned" by the reference itself).  This oracle is pinned instead
 * against the reference's own kernel.cl compiled unchanged as C
 * (oracle/literal_shim.c -> oracle/_ref/libmnv1_literal.so) on the domain where
 * that code is well defined (tests/test_oracle_vs_literal.py), and against an
 * independent torch.nn.functional implementation everywhere else.
 */
#ifndef MNV1_ORACLE_H
#define MNV1_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { MNV1O_ACT_NONE = 0, MNV1O_ACT_RELU = 1, MNV1O_ACT_RELU6 = 2 };
/* padding convention for stride-2 layers (stride 1 is always 1 px each side)
 *   REF    : 1 px top/left (what kernel.cl's `< 0` test implies), window centred on s*y
 *   TFSAME : Keras/TF "SAME" on even sizes = 0 top/left, 1 bottom/right           */
enum { MNV1O_PAD_REF = 0, MNV1O_PAD_TFSAME = 1 };

/* values of mnv1o_epilogue.round_bf16 (the "how is the result stored" field) */
enum { MNV1O_STORE_F32 = 0, MNV1O_STORE_BF16 = 1, MNV1O_STORE_U8_SAT = 2, MNV1O_STORE_U8_WRAP = 3 };

typedef struct {
  const float* scale;   /* per output channel, NULL = 1  (folded BatchNorm)       */
  const float* shift;   /* per output channel, NULL = 0  (folded BatchNorm / bias) */
  int act;              /* MNV1O_ACT_*                                            */
  int round_bf16;       /* MNV1O_STORE_*: 1 = round the stored result to bfloat16 (RNE);
                           2 / 3 = the reference's integer arithmetic with a saturating / wrapping
                           u8 store: out = store(act(sum + shift[c]) >> s), scale[c] = 2^-s        */
} mnv1o_epilogue;

/* kernel.cl:2-60.  r,g,b: [n][rows][cols] u8 planes with element stride `pix_stride`
 * (1 = separate planes as MobileNet.c:218-238 builds them, 3 = interleaved PPM payload).
 * filter: [op_size][3(R,G,B)][3][3].  out: [n][op_size][rows/stride][cols/stride].
 * Input is mapped x -> x*in_scale + in_bias before the conv (1,0 = raw integers). */
void mnv1o_convolute(float* out, const uint8_t* r, const uint8_t* g, const uint8_t* b,
                     int pix_stride, long img_stride, const float* filter, int n, int rows,
                     int cols, int filtersize, int stride, int op_size, int pad_mode,
                     float in_scale, float in_bias, const mnv1o_epilogue* ep);

/* kernel.cl:62-92.  in: [n][op_size][rows][cols]; filter [op_size][3][3];
 * out [n][op_size][rows/stride][cols/stride]. */
void mnv1o_depthwise(float* out, const float* in, const float* filter, int n, int rows,
                     int cols, int filtersize, int stride, int op_size, int pad_mode,
                     const mnv1o_epilogue* ep);

/* kernel.cl:94-114.  in: [n][filtersize][rows][cols] (filtersize = Cin, the contraction
 * length the kernel loops over); filter [op_size][filtersize]; out [n][op_size][rows][cols]. */
void mnv1o_pointwise(float* out, const float* in, const float* filter, int n, int rows,
                     int cols, int filtersize, int op_size, const mnv1o_epilogue* ep);

/* kernel.cl:116-131.  in [n][op_size][filtersize][filtersize] -> out [n][op_size];
 * truncate=1 reproduces the integer `sum / 49` of kernel.cl:129. */
void mnv1o_pool(float* out, const float* in, int n, int rows, int cols, int filtersize,
                int op_size, int truncate, int round_bf16);

/* MobileNet.c:2769-2792 (with max-subtraction; 0-based argmax, the caller prints +1).
 * logits [n][classes] -> prob [n][classes] (may be NULL), top1 [n], top1_prob [n]. */
void mnv1o_softmax_argmax(const float* logits, int n, int classes, double* prob,
                          int* top1, double* top1_prob);

/* ---- whole network (MobileNet.c:207-2792 schedule) ------------- */
typedef struct {
  int kind;     /* 0 stem, 1 depthwise, 2 pointwise, 3 pool, 4 fc */
  int cin, cout, hin, hout, stride;
  long w_off;   /* offset into the flat weight file (reference order) */
  long w_cnt;
  long c_off;   /* offset of this layer's first channel in the scale/shift arrays */
} mnv1o_layer;
int  mnv1o_num_layers(void);                /* 29 */
const mnv1o_layer* mnv1o_layers(void);
long mnv1o_total_weights(void);             /* 4 209 088 */
long mnv1o_total_channels(void);            /* 10 944 + 1000 (fc "shift" = bias) */

/* images: [n][224][224][3] interleaved u8 (PPM payload order, MobileNet.c:29,215).
 * weights: flat, reference order.  scale/shift: per channel, per layer (c_off), may be NULL.
 * taps (may be NULL) is an array of 29 caller-allocated buffers (entries may be NULL);
 * taps[k] receives the planar [n][C][H][W] output of layer k+1.  Layers 1..last_layer
 * (1-based, <= 29) are run; final_out receives the last one's output (logits [n][1000]
 * when last_layer == 29).  round_bf16: bit0 = round every conv layer's stored output to
 * bf16, bit1 = also round the pooled vector (the per-layer mnv1_pool stores bf16; the
 * fused head keeps it in fp32). */
void mnv1o_forward(const uint8_t* images, int n, const float* weights, const float* scale,
                   const float* shift, int pad_mode, int act, int round_bf16,
                   float in_scale, float in_bias, int last_layer,
                   float* final_out, float** taps);

float mnv1o_round_bf16(float x);
#ifdef __cplusplus
}
#endif
#endif

/*
 * literal_shim.c — runs the reference's kernel.cl UNCHANGED as host C.
 *
 * TEST INFRASTRUCTURE ONLY (see mnv1_oracle.h).  kernel.cl is never copied into this
 * repo: it is #include'd at build time from where it lies (-DMNV1_KERNEL_CL="<path>",
 * default /root/reference/kernel.cl) and the result goes to oracle/_ref/ (git-ignored).
 *
 * kernel.cl is plain C99 apart from the OpenCL qualifiers and the two work-item
 * builtins; the macros below supply them.  An NDRange becomes two nested loops.
 *
 * Two uses:
 *  1. KATs that pin oracle/mnv1_oracle.c: the literal kernels are launched ONCE PER
 *     OUTPUT CHANNEL (op_size = 1, pre-offset pointers, filtersize = Cin for pointwise)
 *     which side-steps the defects listed in SURVEY.md App. C (D-01/03/04); on that
 *     domain literal == intended exactly (integers).
 *  2. the "reference CPU path" timing of bench.py --impl reference / cpu_baseline:
 *     the whole 29-layer schedule through these same literal kernels.
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define __kernel
#define __global
static __thread size_t lit_gid[2];
static __thread size_t lit_gsz[2];
static inline size_t get_global_id(int d) { return lit_gid[d]; }
static inline size_t get_global_size(int d) { return lit_gsz[d]; }

#ifndef MNV1_KERNEL_CL
#define MNV1_KERNEL_CL "/root/reference/kernel.cl"
#endif
#include MNV1_KERNEL_CL

#define NDRANGE(gx, gy, call)                        \
  do {                                               \
    lit_gsz[0] = (size_t)(gx);                       \
    lit_gsz[1] = (size_t)(gy);                       \
    for (size_t y_ = 0; y_ < (size_t)(gy); ++y_)     \
      for (size_t x_ = 0; x_ < (size_t)(gx); ++x_) { \
        lit_gid[0] = x_;                             \
        lit_gid[1] = y_;                             \
        call;                                        \
      }                                              \
  } while (0)

/* one launch each, exactly the kernel.cl argument lists (kernel.cl:2-3,62,94,116) */
void lit_convolute(unsigned char* out, unsigned char* r, unsigned char* g, unsigned char* b,
                   int* filter, int rows, int cols, int filtersize, int stride, int op_size,
                   int gx, int gy) {
  NDRANGE(gx, gy, convolute(out, r, g, b, filter, rows, cols, filtersize, stride, op_size));
}
void lit_depthwise(unsigned char* out, unsigned char* in, int* filter, int rows, int cols,
                   int filtersize, int stride, int op_size, int gx, int gy) {
  NDRANGE(gx, gy, depthwise(out, in, filter, rows, cols, filtersize, stride, op_size));
}
void lit_pointwise(unsigned char* out, unsigned char* in, int* filter, int rows, int cols,
                   int filtersize, int op_size, int gx, int gy) {
  NDRANGE(gx, gy, pointwise(out, in, filter, rows, cols, filtersize, op_size));
}
void lit_pool(unsigned char* out, unsigned char* in, int rows, int cols, int filtersize,
              int op_size, int gx, int gy) {
  NDRANGE(gx, gy, pool(out, in, rows, cols, filtersize, op_size));
}

/* ---- whole-network CPU timing path -------------------------------------------------
 * The MobileNet.c schedule through the literal kernels, per-output-channel
 * launches, OpenMP over (image, channel).  Buffers are planar u8 with one guard row after
 * every plane because kernel.cl has no upper-bound check (App. C D-08).  Stride-2 layers
 * run the literal (dilated) arithmetic at output resolution: the same number of MACs as
 * the intended layer, which is what the timing needs; the values are the reference's. */
typedef struct { int kind, cin, cout, hin, hout, stride; long w_off; } lit_layer;

static int lit_build_layers(lit_layer* L) {
  static const int dw_c[13] = {32, 64, 128, 128, 256, 256, 512, 512, 512, 512, 512, 512, 1024};
  static const int dw_s[13] = {1, 2, 1, 2, 1, 2, 1, 1, 1, 1, 1, 2, 1};
  static const int pw_o[13] = {64, 128, 128, 256, 256, 512, 512, 512, 512, 512, 512, 1024, 1024};
  long w = 0; int h = 112, k = 0;
  L[k++] = (lit_layer){0, 3, 32, 224, 112, 2, 0}; w = 864;
  for (int b = 0; b < 13; ++b) {
    int ho = h / dw_s[b];
    L[k++] = (lit_layer){1, dw_c[b], dw_c[b], h, ho, dw_s[b], w}; w += dw_c[b] * 9L; h = ho;
    L[k++] = (lit_layer){2, dw_c[b], pw_o[b], h, h, 1, w}; w += (long)dw_c[b] * pw_o[b];
  }
  L[k++] = (lit_layer){3, 1024, 1024, 7, 1, 1, w};
  L[k++] = (lit_layer){4, 1024, 1000, 1, 1, 1, w};
  return k;
}

/* images [n][224][224][3] interleaved u8; weights: flat int, reference order (4 209 088);
 * logits_u8 [n][1000].  Returns 0. */
int lit_forward(const unsigned char* images, int n, const int* weights, unsigned char* logits_u8) {
  lit_layer L[32];
  const int nl = lit_build_layers(L);
#pragma omp parallel for schedule(dynamic, 1)
  for (int im = 0; im < n; ++im) {
    /* MobileNet.c:218-238: de-interleave into three planes (+ guard rows) */
    const size_t plane = 224 * 224, guard = 2 * 224 + 8;
    unsigned char* rgb = (unsigned char*)calloc(3 * (plane + guard), 1);
    for (size_t p = 0; p < plane; ++p)
      for (int c = 0; c < 3; ++c) rgb[c * (plane + guard) + p] = images[(size_t)im * plane * 3 + p * 3 + c];
    unsigned char* cur = rgb; size_t cur_plane = plane + guard;
    for (int k = 0; k < nl; ++k) {
      const lit_layer* l = &L[k];
      size_t oplane = (size_t)l->hout * l->hout, og = 2 * (size_t)l->hout + 8;
      unsigned char* nxt = (unsigned char*)calloc((size_t)l->cout * (oplane + og) + 1024, 1);
      int* w = (int*)(weights + l->w_off);
      for (int fc = 0; fc < l->cout; ++fc) {
        unsigned char* o = nxt + (size_t)fc * (oplane + og);
        switch (l->kind) {
          case 0: /* rows=cols=2*hout so kernel.cl:14's (rows/2)*(cols/2) plane pitch is irrelevant at op_size=1 */
            lit_convolute(o, cur, cur + cur_plane, cur + 2 * cur_plane, w + fc * 27, 224, 224, 3, 2, 1,
                          l->hout, l->hout);
            break;
          case 1:
            lit_depthwise(o, cur + (size_t)fc * cur_plane, w + fc * 9, l->hout, l->hout, 3, l->stride, 1,
                          l->hout, l->hout);
            break;
          case 2: case 4: {
            /* pointwise strides input planes by rows*cols (kernel.cl:107): give it a dense view */
            static __thread unsigned char* dense = NULL; static __thread size_t dense_sz = 0;
            size_t ip = (size_t)l->hin * l->hin, need = ip * l->cin;
            if (fc == 0) {
              if (dense_sz < need) { free(dense); dense = (unsigned char*)malloc(need); dense_sz = need; }
              for (int ci = 0; ci < l->cin; ++ci) memcpy(dense + ci * ip, cur + (size_t)ci * cur_plane, ip);
            }
            lit_pointwise(o, dense, w + (size_t)fc * l->cin, l->hin, l->hin, l->cin, 1, l->hin, l->hin);
            break;
          }
          case 3:
            lit_pool(o, cur + (size_t)fc * cur_plane, 7, 7, 7, 1, 1, 1);
            break;
        }
      }
      free(cur);
      cur = nxt; cur_plane = oplane + og;
    }
    for (int c = 0; c < 1000; ++c) logits_u8[(size_t)im * 1000 + c] = cur[(size_t)c * cur_plane];
    free(cur);
  }
  return 0;
}

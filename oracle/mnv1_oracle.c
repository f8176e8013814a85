/*
 * mnv1_oracle.c — CPU oracle (TEST INFRASTRUCTURE ONLY; see mnv1_oracle.h).
 *
 * Plain C99 loops in the reference's planar-CHW layout and tap order.  Accumulation
 * is in double so that the oracle is the most accurate statement of the layer; on the
 * small-integer KAT domain every path (this file, kernel.cl compiled as C, the CUDA
 * kernels) is exact and the comparison is bit-for-bit.
 */
#include "mnv1_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

float mnv1o_round_bf16(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return x; /* NaN */
  uint32_t lsb = (u >> 16) & 1u;
  u += 0x7fffu + lsb; /* round to nearest even */
  u &= 0xffff0000u;
  memcpy(&x, &u, 4);
  return x;
}

/* Integer modes (round_bf16 = MNV1O_STORE_U8_SAT / _WRAP): the reference's own arithmetic, exactly —
 * `int sum` (kernel.cl:10,70,100), `if (sum <= 0) sum = 0` (:52-54,87-89,109-111), stored to `unsigned char`
 * (:56,90,112: the C conversion, i.e. modulo 256).  Generalised by an integer bias and a right shift:
 *   v = sum + shift[c];  ReLU;  v >>= s  (scale[c] = 2^-s);  store = v mod 256 (WRAP) or clamp(v, 0, 255) (SAT).
 * Computed in 64-bit integers from the (exact) double accumulator. */
static inline float epilogue_int(double acc, int c, const mnv1o_epilogue* ep) {
  long long v = llround(acc);
  if (ep->shift) v += llround((double)ep->shift[c]);
  if (ep->act != MNV1O_ACT_NONE && v < 0) v = 0;
  int s = 0;
  if (ep->scale) while (s < 31 && ep->scale[c] != ldexpf(1.0f, -s)) ++s;
  v = (v >= 0) ? (v >> s) : -((-v + (1LL << s) - 1) >> s); /* floor division by 2^s */
  if (ep->round_bf16 == MNV1O_STORE_U8_WRAP) v = ((v % 256) + 256) % 256;
  else v = v < 0 ? 0 : v > 255 ? 255 : v;
  return (float)v;
}

static inline float epilogue(double acc, int c, const mnv1o_epilogue* ep) {
  if (ep && ep->round_bf16 >= MNV1O_STORE_U8_SAT) return epilogue_int(acc, c, ep);
  float y = (float)acc;
  if (ep) {
    float s = ep->scale ? ep->scale[c] : 1.0f;
    float t = ep->shift ? ep->shift[c] : 0.0f;
    if (ep->scale || ep->shift) y = (float)((double)y * (double)s + (double)t);
    /* kernel.cl:52-54,87-89,109-111: `if (sum <= 0) sum = 0;` ; ReLU6 adds the cap */
    if (ep->act != MNV1O_ACT_NONE && y < 0.0f) y = 0.0f;
    if (ep->act == MNV1O_ACT_RELU6 && y > 6.0f) y = 6.0f;
    if (ep->round_bf16) y = mnv1o_round_bf16(y);
  }
  return y;
}

static inline int pad_lo(int stride, int pad_mode, int half) {
  return (stride == 2 && pad_mode == MNV1O_PAD_TFSAME) ? 0 : half;
}

/* kernel.cl:13-58 — for each filter: R taps, then G taps, then B taps, findex running
 * contiguously (27 per filter); i = row offset (outer), j = col offset (inner). */
void mnv1o_convolute(float* out, const uint8_t* r, const uint8_t* g, const uint8_t* b,
                     int pix_stride, long img_stride, const float* filter, int n, int rows,
                     int cols, int filtersize, int stride, int op_size, int pad_mode,
                     float in_scale, float in_bias, const mnv1o_epilogue* ep) {
  const int half = filtersize / 2, pl = pad_lo(stride, pad_mode, half);
  const int orows = rows / stride, ocols = cols / stride, ff = filtersize * filtersize;
  const uint8_t* planes[3] = {r, g, b};
#pragma omp parallel for collapse(2) schedule(static)
  for (int im = 0; im < n; ++im)
    for (int fc = 0; fc < op_size; ++fc) {
      float* o = out + ((long)im * op_size + fc) * orows * ocols;
      for (int ty = 0; ty < orows; ++ty)
        for (int tx = 0; tx < ocols; ++tx) {
          double sum = 0.0;
          for (int p = 0; p < 3; ++p) {
            const uint8_t* src = planes[p] + (long)im * img_stride;
            const float* w = filter + ((long)fc * 3 + p) * ff;
            for (int i = 0; i < filtersize; ++i) {
              int y = ty * stride + i - pl;
              if (y < 0 || y >= rows) continue;
              for (int j = 0; j < filtersize; ++j) {
                int x = tx * stride + j - pl;
                if (x < 0 || x >= cols) continue;
                float v = (float)src[((long)y * cols + x) * pix_stride] * in_scale + in_bias;
                sum += (double)v * (double)w[i * filtersize + j];
              }
            }
          }
          o[ty * ocols + tx] = epilogue(sum, fc, ep);
        }
    }
}

/* kernel.cl:72-91 with the channel offset the reference forgot (App. C D-03). */
void mnv1o_depthwise(float* out, const float* in, const float* filter, int n, int rows,
                     int cols, int filtersize, int stride, int op_size, int pad_mode,
                     const mnv1o_epilogue* ep) {
  const int half = filtersize / 2, pl = pad_lo(stride, pad_mode, half);
  const int orows = rows / stride, ocols = cols / stride;
#pragma omp parallel for collapse(2) schedule(static)
  for (int im = 0; im < n; ++im)
    for (int fc = 0; fc < op_size; ++fc) {
      const float* src = in + ((long)im * op_size + fc) * rows * cols;
      const float* w = filter + (long)fc * filtersize * filtersize;
      float* o = out + ((long)im * op_size + fc) * orows * ocols;
      for (int ty = 0; ty < orows; ++ty)
        for (int tx = 0; tx < ocols; ++tx) {
          double sum = 0.0;
          for (int i = 0; i < filtersize; ++i) {
            int y = ty * stride + i - pl;
            if (y < 0 || y >= rows) continue;
            for (int j = 0; j < filtersize; ++j) {
              int x = tx * stride + j - pl;
              if (x < 0 || x >= cols) continue;
              sum += (double)src[(long)y * cols + x] * (double)w[i * filtersize + j];
            }
          }
          o[ty * ocols + tx] = epilogue(sum, fc, ep);
        }
    }
}

/* kernel.cl:102-113 — per pixel, per output channel, contraction over `filtersize`
 * input planes; filter index runs [fc][i] (kernel.cl:106). */
void mnv1o_pointwise(float* out, const float* in, const float* filter, int n, int rows,
                     int cols, int filtersize, int op_size, const mnv1o_epilogue* ep) {
  const long hw = (long)rows * cols;
#pragma omp parallel for collapse(2) schedule(static)
  for (int im = 0; im < n; ++im)
    for (int fc = 0; fc < op_size; ++fc) {
      const float* src = in + (long)im * filtersize * hw;
      const float* w = filter + (long)fc * filtersize;
      float* o = out + ((long)im * op_size + fc) * hw;
      double* acc = (double*)calloc((size_t)hw, sizeof(double));
      for (int i = 0; i < filtersize; ++i) {
        const double wi = (double)w[i];
        const float* s = src + (long)i * hw;
        for (long p = 0; p < hw; ++p) acc[p] += (double)s[p] * wi;
      }
      for (long p = 0; p < hw; ++p) o[p] = epilogue(acc[p], fc, ep);
      free(acc);
    }
}

/* kernel.cl:124-130 — sum of filtersize^2 contiguous values per channel, / 49. */
void mnv1o_pool(float* out, const float* in, int n, int rows, int cols, int filtersize,
                int op_size, int truncate, int round_bf16) {
  (void)rows; (void)cols;
  const int cnt = filtersize * filtersize;
#pragma omp parallel for schedule(static)
  for (int im = 0; im < n; ++im)
    for (int fc = 0; fc < op_size; ++fc) {
      const float* s = in + ((long)im * op_size + fc) * cnt;
      double sum = 0.0;
      for (int i = 0; i < cnt; ++i) sum += (double)s[i];
      float y;
      if (truncate) y = (float)((long)sum / cnt);
      else y = (float)(sum / (double)cnt);
      if (round_bf16) y = mnv1o_round_bf16(y);
      out[(long)im * op_size + fc] = y;
    }
}

/* MobileNet.c:2769-2792.  The reference exponentiates raw logits; subtracting the row
 * maximum first is mathematically identical and keeps exp() in range. */
void mnv1o_softmax_argmax(const float* logits, int n, int classes, double* prob,
                          int* top1, double* top1_prob) {
  for (int im = 0; im < n; ++im) {
    const float* z = logits + (long)im * classes;
    double mx = z[0];
    int loc = 0;
    for (int k = 1; k < classes; ++k)
      if ((double)z[k] > mx) { mx = z[k]; loc = k; } /* strict '>' : first max wins (:2786) */
    double sum = 0.0;
    for (int k = 0; k < classes; ++k) sum += exp((double)z[k] - mx);
    if (prob)
      for (int k = 0; k < classes; ++k) prob[(long)im * classes + k] = exp((double)z[k] - mx) / sum;
    if (top1) top1[im] = loc;
    if (top1_prob) top1_prob[im] = 1.0 / sum;
  }
}

/* ---- layer schedule: MobileNet.c as written, SURVEY.md App. A (stride of L26 is 1,
 * App. B note) --------------------------------------------------------------------- */
#define NL 29
static mnv1o_layer g_layers[NL];
static int g_init = 0;
static void init_layers(void) {
  if (g_init) return;
  static const int dw_c[13] = {32, 64, 128, 128, 256, 256, 512, 512, 512, 512, 512, 512, 1024};
  static const int dw_s[13] = {1, 2, 1, 2, 1, 2, 1, 1, 1, 1, 1, 2, 1};
  static const int pw_o[13] = {64, 128, 128, 256, 256, 512, 512, 512, 512, 512, 512, 1024, 1024};
  long w = 0, c = 0;
  int h = 224, k = 0;
  g_layers[k] = (mnv1o_layer){0, 3, 32, 224, 112, 2, w, 864, c};
  w += 864; c += 32; h = 112; ++k;
  for (int b = 0; b < 13; ++b) {
    int ho = h / dw_s[b];
    g_layers[k] = (mnv1o_layer){1, dw_c[b], dw_c[b], h, ho, dw_s[b], w, (long)dw_c[b] * 9, c};
    w += (long)dw_c[b] * 9; c += dw_c[b]; h = ho; ++k;
    g_layers[k] = (mnv1o_layer){2, dw_c[b], pw_o[b], h, h, 1, w, (long)dw_c[b] * pw_o[b], c};
    w += (long)dw_c[b] * pw_o[b]; c += pw_o[b]; ++k;
  }
  g_layers[k] = (mnv1o_layer){3, 1024, 1024, 7, 1, 1, w, 0, c}; ++k;
  g_layers[k] = (mnv1o_layer){4, 1024, 1000, 1, 1, 1, w, 1024000L, c};
  g_init = 1;
}
int mnv1o_num_layers(void) { return NL; }
const mnv1o_layer* mnv1o_layers(void) { init_layers(); return g_layers; }
long mnv1o_total_weights(void) { init_layers(); return g_layers[NL - 1].w_off + g_layers[NL - 1].w_cnt; }
long mnv1o_total_channels(void) { init_layers(); return g_layers[NL - 1].c_off + 1000; }

void mnv1o_forward(const uint8_t* images, int n, const float* weights, const float* scale,
                   const float* shift, int pad_mode, int act, int round_bf16,
                   float in_scale, float in_bias, int last_layer,
                   float* final_out, float** taps) {
  init_layers();
  float* cur = NULL;
  long cur_elems = 0;
  for (int k = 0; k < last_layer && k < NL; ++k) {
    const mnv1o_layer* L = &g_layers[k];
    const int int_mode = round_bf16 >= MNV1O_STORE_U8_SAT;   /* the whole chain in the reference's integers */
    mnv1o_epilogue ep = {scale ? scale + L->c_off : NULL, shift ? shift + L->c_off : NULL, act,
                         int_mode ? round_bf16 : (round_bf16 & 1)};
    long out_elems = (long)n * L->cout * L->hout * L->hout;
    float* nxt = (float*)malloc((size_t)out_elems * sizeof(float));
    const float* w = weights + L->w_off;
    switch (L->kind) {
      case 0:
        mnv1o_convolute(nxt, images, images + 1, images + 2, 3, 3L * L->hin * L->hin, w, n, L->hin,
                        L->hin, 3, L->stride, L->cout, pad_mode, in_scale, in_bias, &ep);
        break;
      case 1:
        mnv1o_depthwise(nxt, cur, w, n, L->hin, L->hin, 3, L->stride, L->cout, pad_mode, &ep);
        break;
      case 2:
        mnv1o_pointwise(nxt, cur, w, n, L->hin, L->hin, L->cin, L->cout, &ep);
        break;
      case 3:
        if (int_mode) mnv1o_pool(nxt, cur, n, L->hin, L->hin, L->hin, L->cout, 1, 0);   /* kernel.cl:129 integer division */
        else mnv1o_pool(nxt, cur, n, L->hin, L->hin, L->hin, L->cout, 0, (round_bf16 >> 1) & 1);
        break;
      case 4: { /* FC = pointwise at rows=cols=1 (MobileNet.c:2689), bias, no activation,
                   logits kept in fp32 */
        mnv1o_epilogue fe = {NULL, shift ? shift + L->c_off : NULL, MNV1O_ACT_NONE, 0};
        if (int_mode) fe = ep;   /* the reference's FC is its `pointwise` kernel, ReLU and u8 store included (MobileNet.c:2689-2754) */
        mnv1o_pointwise(nxt, cur, w, n, 1, 1, L->cin, L->cout, &fe);
        break;
      }
    }
    if (taps && taps[k]) memcpy(taps[k], nxt, (size_t)out_elems * sizeof(float));
    free(cur);
    cur = nxt;
    cur_elems = out_elems;
  }
  if (final_out && cur) memcpy(final_out, cur, (size_t)cur_elems * sizeof(float));
  free(cur);
}

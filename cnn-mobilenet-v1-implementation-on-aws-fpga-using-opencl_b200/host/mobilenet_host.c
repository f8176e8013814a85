/* mobilenet_host.c — the reference's host structure on the mnv1 C-ABI.
 *
 * Keeps, from MobileNet.c: the image read + R/G/B split (:215-238), one block per layer with
 * "read filter -> create buffers -> set args -> launch -> wait -> event time -> printf"
 * (:322-408), the printed lines ("Kernel Execution time for Layer k: %f" in seconds, :315;
 * "Highest Probability of the element is present at location %d and it's value is %f.", :2792)
 * and the print-and-exit error handling (:295-299).
 * Replaces: the OpenCL bring-up (:147-205) by mnv1_ctx_create, clCreateBuffer/Write/ReadBuffer
 * by mnv1_malloc/upload/download, clEnqueueNDRangeKernel by the four mnv1_* kernels, the host
 * softmax (:2769-2792) by mnv1_softmax, readSquezeNetKernel's 29 re-opens of the text file
 * (:31-47) by one sequential parse.  Feature maps stay on the device between layers (the
 * reference copies them to the host and back, :306,340,350); pass -dump to get its per-layer
 * "Layer k op:" prints (:317-320) from a download.
 *
 * usage: prog [-w weights_file] [-i image.ppm] [-bf16 | -u8 [-wrap]] [-raw] [-ref-pad] [-dump] [-n batch]
 *   -w   weight file, default weights_c.txt (MobileNet.c:37); text or MNV1WTS1 binary
 *   -i   P6 PPM 224x224, default Cat_Image0.ppm (MobileNet.c:215)
 *   -u8  the reference's own integer arithmetic (u8 maps x int8 filters -> int -> ReLU -> u8, kernel.cl:2-3,62,94;
 *        implies -raw; the weight file must hold integers in [-128, 127]); -wrap stores like kernel.cl does
 *        (the C conversion to unsigned char, modulo 256) instead of saturating
 *   -raw keep u8 pixels as integers (what kernel.cl reads) instead of x/127.5-1
 *   -ref-pad  pad stride-2 layers top/left (kernel.cl's `< 0` test) instead of TF "SAME"
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/mnv1.h"
#include "mobilenet_host.h"

#define HEIGHT 224
#define WIDTH 224
#define K 3

#define CHECK(call)                                                           \
  do {                                                                        \
    int rc_ = (call);                                                         \
    if (rc_ != MNV1_OK) {                                                     \
      printf("Error: %s failed! %d (%s)\n", #call, rc_, mnv1_last_error(ctx)); \
      exit(1);                                                                \
    }                                                                         \
  } while (0)

int mobilenet_run(int num_layers, int argc, char** argv) {
  const char* weights_path = "weights_c.txt";
  const char* image_path = "Cat_Image0.ppm";
  mnv1_dtype dtype = MNV1_F32;
  int raw = 0, ref_pad = 0, dump = 0, batch = 1, wrap = 0;
  for (int a = 1; a < argc; ++a) {
    if (!strcmp(argv[a], "-w") && a + 1 < argc) weights_path = argv[++a];
    else if (!strcmp(argv[a], "-i") && a + 1 < argc) image_path = argv[++a];
    else if (!strcmp(argv[a], "-bf16")) dtype = MNV1_BF16;
    else if (!strcmp(argv[a], "-u8")) { dtype = MNV1_U8; raw = 1; }
    else if (!strcmp(argv[a], "-wrap")) wrap = 1;
    else if (!strcmp(argv[a], "-raw")) raw = 1;
    else if (!strcmp(argv[a], "-ref-pad")) ref_pad = 1;
    else if (!strcmp(argv[a], "-dump")) dump = 1;
    else if (!strcmp(argv[a], "-n") && a + 1 < argc) batch = atoi(argv[++a]);
    else { printf("usage: %s [-w weights] [-i image.ppm] [-bf16 | -u8 [-wrap]] [-raw] [-ref-pad] [-dump] [-n batch]\n", argv[0]); return 2; }
  }
  if (batch < 1) batch = 1;

  mnv1_ctx* ctx = NULL;
  int rc = mnv1_ctx_create(0, dtype, &ctx); /* MobileNet.c:147-205 */
  if (rc != MNV1_OK) {
    printf("Error: Failed to create a compute context! %d (%s)\n", rc, mnv1_last_error(NULL));
    return EXIT_FAILURE;
  }
  CHECK(mnv1_ctx_set_pad_mode(ctx, ref_pad ? MNV1_PAD_REF : MNV1_PAD_TFSAME));
  if (!raw) CHECK(mnv1_ctx_set_input_transform(ctx, 1.0f / 127.5f, -1.0f));
  if (dtype == MNV1_U8) CHECK(mnv1_ctx_set_u8_store(ctx, wrap));

  /* filter values for the whole network, parsed once (readSquezeNetKernel, MobileNet.c:31-47) */
  float* weights = (float*)malloc(sizeof(float) * MNV1_TOTAL_WEIGHTS);
  float* scale = (float*)malloc(sizeof(float) * (MNV1_BN_CHANNELS + MNV1_NUM_CLASSES));
  float* shift = (float*)malloc(sizeof(float) * (MNV1_BN_CHANNELS + MNV1_NUM_CLASSES));
  if (mnv1_parse_weights(weights_path, weights, scale, shift) != MNV1_OK) {
    printf("Error: Failed to read weights! (%s)\n", mnv1_last_error(NULL));
    return EXIT_FAILURE;
  }

  /* image: decode + separate R, G and B pixels (MobileNet.c:215-238); replicated over the batch */
  unsigned char* image = (unsigned char*)malloc((size_t)HEIGHT * WIDTH * K);
  if (mnv1_read_ppm(image_path, image, HEIGHT, WIDTH) != MNV1_OK) {
    printf("Error: Failed to read image! (%s)\n", mnv1_last_error(NULL));
    return EXIT_FAILURE;
  }
  const size_t plane = (size_t)HEIGHT * WIDTH;
  unsigned char* image_r = (unsigned char*)malloc(plane * batch);
  unsigned char* image_g = (unsigned char*)malloc(plane * batch);
  unsigned char* image_b = (unsigned char*)malloc(plane * batch);
  for (int n = 0; n < batch; ++n)
    for (size_t i = 0; i < plane; ++i) {
      image_r[n * plane + i] = image[3 * i];
      image_g[n * plane + i] = image[3 * i + 1];
      image_b[n * plane + i] = image[3 * i + 2];
    }
  mnv1_buf *d_image_r, *d_image_g, *d_image_b;
  CHECK(mnv1_malloc_u8(ctx, plane * batch, &d_image_r));
  CHECK(mnv1_malloc_u8(ctx, plane * batch, &d_image_g));
  CHECK(mnv1_malloc_u8(ctx, plane * batch, &d_image_b));
  CHECK(mnv1_upload_u8(ctx, d_image_r, image_r, plane * batch));
  CHECK(mnv1_upload_u8(ctx, d_image_g, image_g, plane * batch));
  CHECK(mnv1_upload_u8(ctx, d_image_b, image_b, plane * batch));

  mnv1_layer_info L[MNV1_NUM_LAYERS];
  mnv1_layer_table(L);
  if (num_layers > MNV1_NUM_LAYERS) num_layers = MNV1_NUM_LAYERS;

  mnv1_buf* d_input = NULL; /* output of the previous layer (d_image_l<k> in the reference) */
  double total_s = 0.0;
  for (int k = 0; k < num_layers; ++k) {
    const mnv1_layer_info* l = &L[k];
    const int op_size = l->cout, stride = l->stride;
    int rows = l->hin, cols = l->hin, filtersize = K;
    mnv1_filter* d_filter = NULL;
    mnv1_buf* d_output = NULL;
    /* integer contexts: every layer, the FC included, is the reference's `if (sum <= 0) sum = 0` (kernel.cl:52,87,109) */
    const int fc = l->kind == MNV1_FC && dtype != MNV1_U8;
    const mnv1_act act = dtype == MNV1_U8 ? MNV1_ACT_RELU : MNV1_ACT_RELU6;
    if (l->kind != MNV1_POOL)
      CHECK(mnv1_filter_create(ctx, (mnv1_kind)l->kind, weights + l->w_off, l->cin, l->cout, fc ? NULL : scale + l->c_off,
                               shift + l->c_off, fc ? MNV1_ACT_NONE : act, &d_filter));
    CHECK(mnv1_malloc(ctx, batch, op_size, l->hout, l->hout, &d_output));
    switch (l->kind) {
      case MNV1_CONVOLUTE: /* MobileNet.c:208-315 */
        CHECK(mnv1_convolute(ctx, d_output, d_image_r, d_image_g, d_image_b, d_filter, rows, cols, filtersize, stride, op_size));
        break;
      case MNV1_DEPTHWISE: /* e.g. MobileNet.c:322-408 */
        CHECK(mnv1_depthwise(ctx, d_output, d_input, d_filter, rows, cols, filtersize, stride, op_size));
        break;
      case MNV1_POINTWISE: /* e.g. MobileNet.c:410-498; filtersize is the contraction length Cin */
      case MNV1_FC:        /* MobileNet.c:2681-2763 */
        filtersize = l->cin;
        CHECK(mnv1_pointwise(ctx, d_output, d_input, d_filter, rows, cols, filtersize, op_size));
        break;
      case MNV1_POOL: /* MobileNet.c:2601-2679 */
        filtersize = l->hin;
        CHECK(mnv1_pool(ctx, d_output, d_input, rows, cols, filtersize, op_size));
        break;
    }
    CHECK(mnv1_sync(ctx)); /* clWaitForEvents + clFinish */
    float ms = 0.f;
    CHECK(mnv1_last_kernel_ms(ctx, &ms));
    total_s += ms / 1e3;
    printf("Kernel Execution time for Layer %d: %f\n", l->index, ms / 1e3);
    if (dump) {
      size_t cnt = (size_t)batch * op_size * l->hout * l->hout;
      float* host = (float*)malloc(sizeof(float) * cnt);
      CHECK(mnv1_download_planar(ctx, d_output, host));
      for (int i = 0; i < 20 && (size_t)i < cnt; ++i) printf("Layer %d op: %g\t", l->index, host[i]);
      printf("\n");
      free(host);
    }
    if (d_filter) mnv1_filter_destroy(ctx, d_filter);
    if (d_input) mnv1_free(ctx, d_input);
    d_input = d_output;
  }
  printf("Total kernel time for %d layer(s), batch %d: %f\n", num_layers, batch, total_s);

  if (num_layers == MNV1_NUM_LAYERS) { /* Layer 30 - Softmax (MobileNet.c:2769-2792) */
    int* location = (int*)malloc(sizeof(int) * batch);
    float* maximum = (float*)malloc(sizeof(float) * batch);
    CHECK(mnv1_softmax(ctx, d_input, MNV1_NUM_CLASSES, NULL, location, maximum));
    printf("Highest Probability of the element is present at location %d and it's value is %f.\n", location[0] + 1, maximum[0]);
    free(location); free(maximum);
  }

  /* Shutdown and cleanup (MobileNet.c:2794-2832) */
  mnv1_free(ctx, d_input);
  mnv1_free(ctx, d_image_r); mnv1_free(ctx, d_image_g); mnv1_free(ctx, d_image_b);
  free(image_r); free(image_g); free(image_b); free(image);
  free(weights); free(scale); free(shift);
  mnv1_ctx_destroy(ctx);
  return 0;
}

/* MobileNet_13Layers.c — layers 1-13 (reference MobileNet_13Layers.c, which stops after the 6th depthwise/pointwise pair).
 * Same command line as the other two host programs; see mobilenet_host.c. */
#include "mobilenet_host.h"

int main(int argc, char** argv) { return mobilenet_run(13, argc, argv); }

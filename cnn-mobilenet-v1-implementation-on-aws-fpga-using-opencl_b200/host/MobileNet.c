/* MobileNet.c — full network: 29 layers + softmax (reference MobileNet.c).
 * Same command line as the other two host programs; see mobilenet_host.c. */
#include "mobilenet_host.h"

int main(int argc, char** argv) { return mobilenet_run(29, argc, argv); }

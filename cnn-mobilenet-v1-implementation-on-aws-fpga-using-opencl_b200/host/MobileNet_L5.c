/* MobileNet_L5.c — layers 1-5 (reference MobileNet_L5.c).
 * Same command line as the other two host programs; see mobilenet_host.c. */
#include "mobilenet_host.h"

int main(int argc, char** argv) { return mobilenet_run(5, argc, argv); }

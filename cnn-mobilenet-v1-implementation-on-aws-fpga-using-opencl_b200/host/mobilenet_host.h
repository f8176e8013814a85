/* mobilenet_host.h — shared body of the three host programs (MobileNet.c, MobileNet_13Layers.c,
 * MobileNet_L5.c).  The reference repeats one ~85-line OpenCL block per layer
 * (MobileNet.c:322-408 is the layer-2 instance); here that block is written once, driven by the
 * layer table, and calls the mnv1_* C-ABI instead of cl*.  The three programs differ only in how
 * many layers they run, exactly like the reference's three files. */
#ifndef MOBILENET_HOST_H
#define MOBILENET_HOST_H
/* run layers 1..num_layers of the schedule; returns the process exit code */
int mobilenet_run(int num_layers, int argc, char** argv);
#endif

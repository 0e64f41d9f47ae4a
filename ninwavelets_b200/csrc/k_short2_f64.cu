#define NW_REAL double
#define NW_S2_MAXREG 128
#include "nw_kern_short2.cuh"

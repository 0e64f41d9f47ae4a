#define NW_REAL double
#define NW_S2_MAXREG 168
#include "nw_kern_short2.cuh"

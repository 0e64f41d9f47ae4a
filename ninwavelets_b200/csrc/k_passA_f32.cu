#define NW_REAL float
#include "nw_kern_passA.cuh"

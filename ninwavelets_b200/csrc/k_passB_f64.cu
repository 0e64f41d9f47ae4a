#define NW_REAL double
#include "nw_kern_passB.cuh"

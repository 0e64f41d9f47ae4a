#define NW_REAL double
#define NW_CFG 0
#define NW_CFG0_MAXREG 128
#include "nw_kern_long2.cuh"

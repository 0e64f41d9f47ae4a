#define NW_REAL float
#define NW_CFG 1
#define NW_CFG0_MAXREG 80
#include "nw_kern_long2.cuh"

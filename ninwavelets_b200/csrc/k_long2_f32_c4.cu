#define NW_REAL float
#define NW_CFG 4
#define NW_CFG0_MAXREG 80
#define NW_BIG_RADIX 1
#define NW_SP_A(X) X(11)
#define NW_SP_B(X) X(8) X(9) X(10)
#include "nw_kern_long2.cuh"

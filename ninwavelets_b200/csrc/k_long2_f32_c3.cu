#define NW_REAL float
#define NW_CFG 3
#define NW_CFG0_MAXREG 80
#define NW_SP_A(X) X(4) X(20) X(21) X(22)
#define NW_SP_B(X) X(4)
#include "nw_kern_long2.cuh"

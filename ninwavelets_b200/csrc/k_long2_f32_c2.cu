#define NW_REAL float
#define NW_CFG 2
#define NW_CFG0_MAXREG 80
#define NW_SP_A(X) X(2)
#define NW_SP_B(X) X(23) X(24) X(25) X(26)
#include "nw_kern_long2.cuh"

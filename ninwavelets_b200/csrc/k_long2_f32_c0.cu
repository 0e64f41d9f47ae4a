#define NW_REAL float
#define NW_CFG 0
#define NW_CFG0_MAXREG 80
#define NW_SP_A(X) X(3) X(4) X(5) X(14) X(15)
#define NW_SP_B(X) X(1) X(3) X(4) X(5) X(14) X(16)
#include "nw_kern_long2.cuh"

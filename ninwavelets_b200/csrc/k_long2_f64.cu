#define NW_REAL double
#define NW_CFG 0
#define NW_CFG0_MAXREG 128
#define NW_SP_A(X) X(2) X(4) X(5) X(13)
#define NW_SP_B(X) X(4) X(5) X(12) X(13)
#include "nw_kern_long2.cuh"

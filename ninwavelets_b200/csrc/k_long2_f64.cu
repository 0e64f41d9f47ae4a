#define NW_REAL double
#define NW_CFG 0
#ifndef NW_F64_MAXREG
#define NW_F64_MAXREG 168
#endif
#define NW_CFG0_MAXREG NW_F64_MAXREG
#define NW_SP_A(X) X(2) X(4) X(5) X(13)
#define NW_SP_B(X) X(4) X(5) X(12) X(13)
#include "nw_kern_long2.cuh"

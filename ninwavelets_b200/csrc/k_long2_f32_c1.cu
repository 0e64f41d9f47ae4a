#define NW_REAL float
#define NW_CFG 1
#define NW_CFG0_MAXREG 80
#define NW_SP_A(X) 
#define NW_SP_B(X) 
#include "nw_kern_long2.cuh"

#define NW_REAL float
#define NW_S2_MAXREG 80
#define NW_SP_S(X) X(6) X(7)
#include "nw_kern_short2.cuh"

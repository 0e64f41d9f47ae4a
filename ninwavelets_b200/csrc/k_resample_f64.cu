#define NW_REAL double
#define NW_RS_TAPS(X) X(8) X(10) X(12) X(14) X(16) X(18) X(20) X(22) X(24)
#include "nw_kern_resample.cuh"

#define NW_REAL float
#define NW_RS_TAPS(X) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16)
#include "nw_kern_resample.cuh"

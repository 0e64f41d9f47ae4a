#define NW_REAL float
#define NW_RSV_PQ 2
#define NW_RSV_MODE 2
#define NW_RSV_TAPS(X) X(4) X(6) X(8) X(10) X(12)
#include "nw_kern_resample_vec.cuh"

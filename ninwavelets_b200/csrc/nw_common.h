// Common definitions shared by the device code and the host planner.
//
// Everything that is marked NW_HD compiles both as sm_100a device code (the
// product) and as plain host C++ (tests/emul only: the same kernel bodies are
// run block-by-block on the CPU to check index algebra without a GPU).  The
// shipped library never executes the host instantiation.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define NW_HD __host__ __device__ __forceinline__
#define NW_D __device__ __forceinline__
#else
#define NW_HD inline
#define NW_D inline
#endif

#if defined(__CUDA_ARCH__)
#define NW_SYNC() __syncthreads()
#define NW_RESTRICT __restrict__
#else
// host instantiation (tests/emul): a barrier is a hook the emulator sets when it steps the
// threads of a block as fibers; with one stepping thread it is a no-op
namespace nw {
typedef void (*host_sync_fn)();
inline host_sync_fn& host_sync_hook() {
    static host_sync_fn f = nullptr;
    return f;
}
}  // namespace nw
#define NW_SYNC() do { if (nw::host_sync_hook()) nw::host_sync_hook()(); } while (0)
#define NW_RESTRICT
#endif

namespace nw {

// ---- complex value ---------------------------------------------------------
template <typename T>
struct alignas(2 * sizeof(T)) cx {
    T x, y;
};
template <typename T> NW_HD cx<T> mk(T a, T b) { cx<T> r; r.x = a; r.y = b; return r; }
template <typename T> NW_HD cx<T> operator+(cx<T> a, cx<T> b) { return mk<T>(a.x + b.x, a.y + b.y); }
template <typename T> NW_HD cx<T> operator-(cx<T> a, cx<T> b) { return mk<T>(a.x - b.x, a.y - b.y); }
template <typename T> NW_HD cx<T> cmul(cx<T> a, cx<T> b) {
    return mk<T>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
template <typename T> NW_HD cx<T> cmulc(cx<T> a, cx<T> b) {  // a * conj(b)
    return mk<T>(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
template <typename T> NW_HD cx<T> scale(cx<T> a, T s) { return mk<T>(a.x * s, a.y * s); }
// multiply by DIR*i  (DIR=+1: i*v, DIR=-1: -i*v)
template <int DIR, typename T> NW_HD cx<T> rot(cx<T> v) {
    return DIR > 0 ? mk<T>(-v.y, v.x) : mk<T>(v.y, -v.x);
}

// ---- streaming (evict-first) stores for results that are written once and never re-read by the kernels:
// they must not push the L2-resident intermediate out of the cache
#if defined(__CUDA_ARCH__)
NW_D void st_stream(float* p, float v) { __stcs(p, v); }
NW_D void st_stream(double* p, double v) { __stcs(p, v); }
NW_D void st_stream(cx<float>* p, cx<float> v) { __stcs((float2*)p, make_float2(v.x, v.y)); }
NW_D void st_stream(cx<double>* p, cx<double> v) { __stcs((double2*)p, make_double2(v.x, v.y)); }
#else
template <typename U> inline void st_stream(U* p, U v) { *p = v; }
#endif

// ---- exact division by a small runtime constant ------------------------------
// q = x / d for every x with x*d < 2^32 (checked by the planner): one
// multiply-high.  m == 0 encodes d == 1.
struct fastdiv {
    uint32_t d, m;
};
inline fastdiv make_fastdiv(uint32_t d) {
    fastdiv f;
    f.d = d;
    f.m = (d <= 1) ? 0u : (uint32_t)((((uint64_t)1 << 32) + d - 1) / d);
    return f;
}
NW_HD uint32_t fd_div(uint32_t x, fastdiv f) {
#if defined(__CUDA_ARCH__)
    return f.m ? __umulhi(x, f.m) : x;
#else
    return f.m ? (uint32_t)(((uint64_t)x * f.m) >> 32) : x;
#endif
}

// ---- enums shared with include/nwcwt.h (values must match) -------------------
enum { FAM_MORSE = 0, FAM_MORLET = 1, FAM_SHANNON = 2, FAM_TABLE = 3 };
enum { OUT_CWT = 0, OUT_ABS = 1, OUT_POWER = 2 };
enum { BL_NONE = 0, BL_MEAN = 1, BL_RATIO = 2, BL_PERCENT = 3, BL_LOG = 4, BL_ZSCORE = 5, BL_ZLOG = 6 };

static const int MAX_STAGES = 16;
static const int MAX_GENERIC_RADIX = 64;

// Radix plan of one P-point transform (Stockham, autosort).  ns[i] is the
// product of the radices before stage i; mod[i] divides by ns[i].
struct FftStages {
    int P;
    int nst;
    int radix[MAX_STAGES];
    int ns[MAX_STAGES];
    fastdiv div_ns[MAX_STAGES];   // b / ns
    fastdiv div_pr[MAX_STAGES];   // lin / (P / radix)   (b-fastest mapping)
};

// Radix plan of the packed in-place engine (nw_fft2.cuh).  ns[s] = product of the radices before
// stage s; stage s works in blocks of L_s = P / ns[s] with Q_s = L_s / radix[s] butterflies each.
static const int MAX_STAGES2 = 8;
struct Fft2Plan {
    int P;
    int nst;
    int radix[MAX_STAGES2];
    int ns[MAX_STAGES2];
    fastdiv div_q[MAX_STAGES2];   // bi / Q_s
    fastdiv div_r[MAX_STAGES2];   // x / radix[s]
};

}  // namespace nw

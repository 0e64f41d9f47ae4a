// libnwcwt.so - C ABI (include/nwcwt.h) over the sm_100a kernels.
//
// Replaces, on the device, WaveletBase.make_fft_wavelets / cwt / abs / power
// (reference base.py:258-279, 378-443), Baseline (base.py:46-68) applied per
// (signal, frequency) row, and the cupy branch (base.py:236-247, 398-404).
// There is no CPU path here: without a CUDA device every compute entry point
// fails with NWCWT_ERR_CUDA.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#include <string>
#include <vector>

#include "../../include/nwcwt.h"
#include "nw_common.h"
#include "nw_fft.cuh"
#include "nw_family.cuh"
#include "nw_kernels.cuh"
#include "nw_plan.h"
#include "nw_launch.h"

using namespace nw;

// ---------------------------------------------------------------------------------
// small kernels that live in this translation unit (the transform kernels are in k_*.cu)
// ---------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(512) nwcwt_baseline_rows_kernel(T* rows, long long N, int mode, int lo, int hi) {
    __shared__ double sh[2];
    baseline_rows_body<T>(rows, N, mode, lo, hi, sh, blockIdx.x, threadIdx.x, blockDim.x);
}

// Spectrum bank for inspection: one thread per (frequency, bin).  TO: the plan's precision (a chirp-z plan keeps its
// tables in fp64 whatever its I/O precision).
template <typename T, typename TO = T>
__global__ void __launch_bounds__(256) nwcwt_bank_kernel(SpecParams<T> sp, cx<TO>* bank, int F, long long N) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int f = blockIdx.y;
    if (i >= N || f >= F) return;
    const FreqRec r = sp.rec[f];
    cx<T> v = mk<T>((T)0, (T)0);
    if (i >= r.lo && i < r.hi) v = spec_times<T>(sp, r, f, (int)i, mk<T>((T)1, (T)0));
    bank[(size_t)f * (size_t)N + (size_t)i] = mk<TO>((TO)v.x, (TO)v.y);
}

// Epoch reductions (mneutils.py:53-55 mean of power, :68-71 inter-trial coherence).
template <typename T>
__global__ void __launch_bounds__(256) nwcwt_reduce_kernel(const void* in, T* out, long long E, long long count, int kind) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    if (kind == 0) {
        const T* p = (const T*)in;
        T acc = 0;
        for (long long e = 0; e < E; ++e) acc += p[(size_t)e * (size_t)count + (size_t)i];
        out[i] = acc / (T)E;
    } else {
        const cx<T>* p = (const cx<T>*)in;
        T ax = 0, ay = 0;
        for (long long e = 0; e < E; ++e) {
            const cx<T> z = p[(size_t)e * (size_t)count + (size_t)i];
            const T a = nw_hypot(z.x, z.y);
            ax += z.x / a;
            ay += z.y / a;
        }
        out[i] = nw_hypot(ax / (T)E, ay / (T)E);
    }
}

// ---- Bluestein (chirp-z) kernels: length-N transforms for ANY N through circular convolutions of a smooth length M >= 2N-1
// The chirp-z arithmetic runs in fp64 whatever the plan's precision (TI / TO are the plan's I/O types): its rounding error
// grows with the convolution length, and this is the compatibility path, not the fast one.
//   X[k] = conj(c[k]) sum_n (x[n] conj(c[n])) c[k - n],   z[n] = (1/N) c[n] sum_k (Y[k] c[k]) conj(c[n - k]),   c[n] = e^{i pi n^2 / N}
// Every convolution is  T(s; H) = ifft_M(H . fft_M(s))  of a REAL sequence s - exactly what the library's transform with a
// one-row TABLE plan of length M computes; complex sequences go through as their real and imaginary parts.
template <typename TI, typename T>
__global__ void __launch_bounds__(256) nwcwt_blu_pre_fwd_kernel(const TI* x, const cx<T>* chirp, T* in, long long N, long long M) {
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const size_t s = blockIdx.y;
    T re = 0, im = 0;
    if (m < N) {
        const T v = (T)x[s * (size_t)N + (size_t)m];
        const cx<T> c = chirp[m];
        re = v * c.x;
        im = -v * c.y;
    }
    in[(2 * s) * (size_t)M + (size_t)m] = re;
    in[(2 * s + 1) * (size_t)M + (size_t)m] = im;
}
// X[s][k] = conj(c[k]) (Tr[k] + i Ti[k])
template <typename T, typename TO>
__global__ void __launch_bounds__(256) nwcwt_blu_post_fwd_kernel(const cx<T>* conv, const cx<T>* chirp, cx<TO>* X, long long N, long long M) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N) return;
    const size_t s = blockIdx.y;
    const cx<T> a = conv[(2 * s) * (size_t)M + (size_t)k], b = conv[(2 * s + 1) * (size_t)M + (size_t)k];
    const cx<T> u = cmulc(mk<T>(a.x - b.y, a.y + b.x), chirp[k]);
    X[s * (size_t)N + (size_t)k] = mk<TO>((TO)u.x, (TO)u.y);
}
// b[k] = W_f(k) X[k] c[k] on the band, zero elsewhere and on the padding; rows are (signal, frequency) of this chunk
template <typename T>
__global__ void __launch_bounds__(256) nwcwt_blu_pre_inv_kernel(SpecParams<T> sp, const cx<T>* X, const cx<T>* chirp, T* in,
                                                                 long long N, long long M, int F, long long row0) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= M) return;
    const size_t r = blockIdx.y;
    const long long gr = row0 + (long long)r;
    const long long si = gr / F;
    const int fi = (int)(gr - si * F);
    const FreqRec rec = sp.rec[fi];
    cx<T> b = mk<T>((T)0, (T)0);
    if (k < N && k >= rec.lo && k < rec.hi) b = cmul(spec_times<T>(sp, rec, fi, (int)k, X[(size_t)si * (size_t)N + (size_t)k]), chirp[k]);
    in[(2 * r) * (size_t)M + (size_t)k] = b.x;
    in[(2 * r + 1) * (size_t)M + (size_t)k] = b.y;
}
// z[n] = c[n] (Tr[n] + i Ti[n])  ->  cwt / abs / power
template <typename T, typename TO>
__global__ void __launch_bounds__(256) nwcwt_blu_post_inv_kernel(const cx<T>* conv, const cx<T>* chirp, void* out, long long N,
                                                                  long long M, long long row0, int mode) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const size_t r = blockIdx.y;
    const cx<T> a = conv[(2 * r) * (size_t)M + (size_t)n], b = conv[(2 * r + 1) * (size_t)M + (size_t)n];
    const cx<T> z = cmul(mk<T>(a.x - b.y, a.y + b.x), chirp[n]);
    const size_t o = (size_t)(row0 + (long long)r) * (size_t)N + (size_t)n;
    if (mode == OUT_CWT) ((cx<TO>*)out)[o] = mk<TO>((TO)z.x, (TO)z.y);
    else ((TO*)out)[o] = (TO)real_out<T>(mode, z);
}
// table row of the one-row TABLE plan: H = S0 + i S1 (spectra of the filter's real and imaginary parts), or its conjugate
template <typename T>
__global__ void __launch_bounds__(256) nwcwt_blu_filter_kernel(const cx<T>* spec, cx<T>* hf, cx<T>* hi, long long M) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= M) return;
    const cx<T> a = spec[k], b = spec[(size_t)M + (size_t)k];
    const cx<T> h = mk<T>(a.x - b.y, a.y + b.x);
    hf[k] = h;
    hi[k] = mk<T>(h.x, -h.y);
}

// Pipe-rate microbenchmark for bench.py's FLOP roofline: 16 independent FFMA2 (fp32) or DFMA (fp64) chains per thread.
template <typename T>
__global__ void __launch_bounds__(512) nwcwt_fma_peak_kernel(T* out, int iters, T a, T b) {
    T v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = (T)threadIdx.x * (T)1e-6 + (T)i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = v[i] * a + b;   // fp32: the compiler does not pack these; see the float2 variant
    }
    T s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += v[i];
    if (s == (T)123.456) out[0] = s;
}
__global__ void __launch_bounds__(512) nwcwt_ffma2_peak_kernel(float* out, int iters, float a, float b) {
    float2 v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = make_float2(threadIdx.x * 1e-6f + i, threadIdx.x * 2e-6f + i);
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __ffma2_rn(v[i], a2, b2);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += v[i].x + v[i].y;
    if (s == 123.456f) out[0] = s;
}

// ---------------------------------------------------------------------------------
// plan object
// ---------------------------------------------------------------------------------
static thread_local std::string g_err;
static bool g_no_static = getenv("NWCWT_NO_STATIC") != nullptr;   // tuning: run-time plans only
static bool g_force_generic = false;   // nwcwt_debug_force_generic: run the generic kernels even where a fast path exists
static bool g_exact = false;           // nwcwt_debug_force_exact: exact length-N transforms for every row of a resampling plan
static int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            return fail(NWCWT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));    \
    } while (0)

struct nwcwt_plan {
    HostPlan hp;
    bool on_device = false;
    // device tables
    void *d_tw = nullptr, *d_twA = nullptr, *d_twB = nullptr, *d_twH = nullptr, *d_twL = nullptr;
    void *d_twA2 = nullptr, *d_twB2 = nullptr;   // fast long path
    int* d_ditpos = nullptr;                     // fast long path: decimation-in-time slot of every pass-A row
    // resampled rows (HostPlan::groups): one sub-plan per group, plus the group's interpolation tables
    struct Group {
        nwcwt_plan* sub = nullptr;
        void *d_coef = nullptr, *d_eq = nullptr, *d_coefq = nullptr;   // coefq: weights regrouped for the vector kernel
        int *d_t0 = nullptr, *d_fmap = nullptr;
        int t0min = 0;
    };
    std::vector<Group> groups;
    bool is_sub = false;
    // Bluestein (chirp-z) path for lengths the radix engine cannot factor: hp.path == 4, transforms of length blu_m
    nwcwt_plan *blu_f = nullptr, *blu_i = nullptr;   // one-row TABLE sub-plans of length blu_m: forward / inverse chirp filter
    long long blu_m = 0;
    void *d_chirp = nullptr;        // cx<T>[N]: exp(i pi n^2 / N)
    int blu_rows = 1;               // rows (pairs of real transforms) per chunk
    // fast long path: launch pairs of consecutive row groups alternate between auxiliary streams so
    // that one group's pass A fills the SMs the previous group's pass-B tail leaves idle
    static const int MAX_AUX = 4;
    cudaStream_t aux[MAX_AUX] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[MAX_AUX] = {nullptr, nullptr, nullptr, nullptr};
    int n_aux = 0;
    // L2 residency of the intermediate: persisting carve-out + access-policy window on the Tm ring
    size_t l2_persist_max = 0, l2_window_max = 0;
    const void* l2_window_base = nullptr;
    size_t l2_window_bytes = 0;
    void* d_rec = nullptr;
    void* d_table = nullptr;
    void* d_wtab = nullptr;                         // weight table of the bands (nw_plan.h: build_weight_table)
    // resampled short rows (nw_kernels4.cuh): the device array of Short3Group and everything it points to
    void* d_s3groups = nullptr;
    int s3_units = 0;
    std::vector<void*> s3_allocs;
    const std::vector<double>* eq_host = nullptr;   // sub-plan of a resampled group: the group's equaliser, folded into d_wtab
    // host-call resources
    void* h_in_dev[2] = {nullptr, nullptr};
    void* h_out_dev[2] = {nullptr, nullptr};
    void* h_ws[2] = {nullptr, nullptr};
    size_t h_in_bytes = 0, h_out_bytes = 0, h_ws_bytes = 0;
    long long h_chunk = 0;
    cudaStream_t h_stream[2] = {nullptr, nullptr};
    // CUDA-graph replay of the fast long path (nwcwt_transform): one step's launches - forward pairs, then per launch group
    // pass A / pass B / interpolation on the forked auxiliary streams - are captured once per argument set and replayed
    struct GraphEntry {
        const void* signals; void* out; long long S; int output, bl; long long lo, hi; void* ws; size_t ws_bytes; int flags;
        cudaGraphExec_t exec; long long launches; unsigned long long used;
    };
    std::vector<GraphEntry> graphs;
    cudaStream_t cap_stream = nullptr;
    unsigned long long graph_clock = 0;
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Every entry point works on the plan's device and leaves the calling thread's current device as it found it.
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != device) cudaSetDevice(device); else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// ---- launch accounting / optional per-class event timing (bench.py) -----------------------------
#include <atomic>
#include <mutex>
static std::atomic<long long> g_launches{0};
static bool g_profile = false;
static std::mutex g_prof_mu;
struct ProfEvent { cudaEvent_t a, b; int cls; };
static std::vector<ProfEvent> g_prof_events;
static const int N_PROF = 8;
static double g_prof_ms[N_PROF] = {0};
static long long g_prof_n[N_PROF] = {0};

struct LaunchScope {   // counts one launch; with profiling on, brackets it with events on `stream`
    cudaStream_t stream;
    ProfEvent ev;
    bool timed;
    LaunchScope(int cls, cudaStream_t s) : stream(s), timed(g_profile) {
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if (timed) {
            ev.cls = cls;
            cudaEventCreate(&ev.a);
            cudaEventCreate(&ev.b);
            cudaEventRecord(ev.a, stream);
        }
    }
    ~LaunchScope() {
        if (timed) {
            cudaEventRecord(ev.b, stream);
            std::lock_guard<std::mutex> lk(g_prof_mu);
            g_prof_events.push_back(ev);
        }
    }
};

template <typename T>
static void fill_twiddles(std::vector<cx<T>>& v, long long count, long long P, long long step) {
    // v[j] = exp(+2 pi i (j*step) / P), evaluated in long double
    v.resize((size_t)count);
    const long double tp = 6.283185307179586476925286766559005768L;
    for (long long j = 0; j < count; ++j) {
        const long long m = (j * step) % P;
        const long double a = tp * (long double)m / (long double)P;
        v[(size_t)j].x = (T)cosl(a);
        v[(size_t)j].y = (T)sinl(a);
    }
}

// fast long path: compiled launch shapes and compile-time plans per precision (k_long2_*.cu)
template <typename T> struct Long2Dispatch;
#define NW_BY_CFG(c, F, ...) \
    switch (c) { case 1: return F<float, 1>(__VA_ARGS__); case 2: return F<float, 2>(__VA_ARGS__); \
                 case 3: return F<float, 3>(__VA_ARGS__); default: return F<float, 0>(__VA_ARGS__); }
template <> struct Long2Dispatch<float> {
    static cudaError_t prepare(int c) { NW_BY_CFG(c, prepare_long2) }
    static bool has(int c, int pass, int sp) { NW_BY_CFG(c, has_static_plan, pass, sp) }
    static cudaError_t A(int c, int sp, const Long2Params<float>& P, dim3 g, int nt, size_t sm, cudaStream_t s) {
        NW_BY_CFG(c, launch_passA2, sp, P, g, nt, sm, s)
    }
    static cudaError_t B(int c, int sp, const Long2Params<float>& P, dim3 g, int nt, size_t sm, cudaStream_t s) {
        NW_BY_CFG(c, launch_passB2, sp, P, g, nt, sm, s)
    }
};
template <> struct Long2Dispatch<double> {
    static cudaError_t prepare(int) { return prepare_long2<double, 0>(); }
    static bool has(int, int pass, int sp) { return has_static_plan<double, 0>(pass, sp); }
    static cudaError_t A(int, int sp, const Long2Params<double>& P, dim3 g, int nt, size_t sm, cudaStream_t s) {
        return launch_passA2<double, 0>(sp, P, g, nt, sm, s);
    }
    static cudaError_t B(int, int sp, const Long2Params<double>& P, dim3 g, int nt, size_t sm, cudaStream_t s) {
        return launch_passB2<double, 0>(sp, P, g, nt, sm, s);
    }
};

// vector interpolation kernels (nw_resample.cuh: resample_vec_body) exist in fp32 only
template <typename T> struct RsVecDispatch {
    static cudaError_t prepare() { return cudaSuccess; }
    static cudaError_t run(int, int, int, int, const ResampleParams<T>&, dim3, size_t, cudaStream_t) { return cudaErrorInvalidValue; }
    static cudaError_t run_dir(int, int, int, const ResampleParams<T>&, dim3, cudaStream_t) { return cudaErrorInvalidValue; }
};
template <> struct RsVecDispatch<float> {
    static cudaError_t prepare() {
        cudaError_t e;
        if ((e = prepare_resample_vec<float, 4, OUT_POWER>()) != cudaSuccess) return e;
        if ((e = prepare_resample_vec<float, 4, OUT_ABS>()) != cudaSuccess) return e;
        if ((e = prepare_resample_vec<float, 2, OUT_POWER>()) != cudaSuccess) return e;
        return prepare_resample_vec<float, 2, OUT_ABS>();
    }
    static cudaError_t run(int PQ, int mode, int K, int R, const ResampleParams<float>& P, dim3 g, size_t sm, cudaStream_t s) {
        if (PQ == 4) return mode == OUT_POWER ? launch_resample_vec<float, 4, OUT_POWER>(K, R, P, g, sm, s)
                                              : launch_resample_vec<float, 4, OUT_ABS>(K, R, P, g, sm, s);
        return mode == OUT_POWER ? launch_resample_vec<float, 2, OUT_POWER>(K, R, P, g, sm, s)
                                 : launch_resample_vec<float, 2, OUT_ABS>(K, R, P, g, sm, s);
    }
    static cudaError_t run_dir(int PQ, int mode, int K, const ResampleParams<float>& P, dim3 g, cudaStream_t s) {
        if (PQ == 4) return mode == OUT_POWER ? launch_resample_dir<float, 4, OUT_POWER>(K, P, g, s)
                                              : launch_resample_dir<float, 4, OUT_ABS>(K, P, g, s);
        return mode == OUT_POWER ? launch_resample_dir<float, 2, OUT_POWER>(K, P, g, s)
                                 : launch_resample_dir<float, 2, OUT_ABS>(K, P, g, s);
    }
};
static bool g_rs_dir = !(getenv("NWCWT_RS_DIR") && atoi(getenv("NWCWT_RS_DIR")) == 0);   // A/B switch of the interpolation kernels

template <typename T>
static int upload_tw(void** dptr, long long count, long long P, long long step) {
    std::vector<cx<T>> v;
    fill_twiddles<T>(v, count, P, step);
    CUDA_TRY(cudaMalloc(dptr, v.size() * sizeof(cx<T>)));
    CUDA_TRY(cudaMemcpy(*dptr, v.data(), v.size() * sizeof(cx<T>), cudaMemcpyHostToDevice));
    return 0;
}

template <typename T> static int ensure_device_t(nwcwt_plan* pl);
template <typename T> static int ensure_device_short3(nwcwt_plan* pl);
template <typename T>
static int run_transform(nwcwt_plan* pl, const void* signals, void* out, void* spectra, long long S, int output,
                         int bl, long long blo, long long bhi, void* ws, size_t ws_bytes, cudaStream_t stream,
                         bool forward_only);

// Bluestein plan on the device: the chirp, the frequency records / tables of length N for the spectrum generator, and the
// chirp filter's spectrum (computed once with the sub-plan's own forward transform), installed as the sub-plan's table.
template <typename T>
static int ensure_device_bluestein(nwcwt_plan* pl) {
    HostPlan& hp = pl->hp;
    const long long N = hp.N, M = pl->blu_m;
    int rc;
    if ((rc = ensure_device_t<T>(pl->blu_f))) return rc;
    if ((rc = ensure_device_t<T>(pl->blu_i))) return rc;
    std::vector<cx<T>> c((size_t)N);
    std::vector<T> h(2 * (size_t)M, (T)0);   // filter c[n], n in (-N, N), wrapped onto M points: real parts, then imaginary parts
    const long double pi = 3.141592653589793238462643383279502884L;
    for (long long n = 0; n < N; ++n) {
        const long long q = (long long)(((unsigned long long)n * (unsigned long long)n) % (unsigned long long)(2 * N));
        const long double a = pi * (long double)q / (long double)N;
        const long double cr = cosl(a), ci = sinl(a);
        c[(size_t)n].x = (T)cr;
        c[(size_t)n].y = (T)ci;
        h[(size_t)n] = (T)cr;
        h[(size_t)M + (size_t)n] = (T)ci;
        if (n > 0) { h[(size_t)(M - n)] = (T)cr; h[(size_t)M + (size_t)(M - n)] = (T)ci; }
    }
    CUDA_TRY(cudaMalloc(&pl->d_chirp, sizeof(cx<T>) * (size_t)N));
    CUDA_TRY(cudaMemcpy(pl->d_chirp, c.data(), sizeof(cx<T>) * (size_t)N, cudaMemcpyHostToDevice));
    if (hp.F > 0) {
        CUDA_TRY(cudaMalloc(&pl->d_rec, sizeof(FreqRec) * hp.F));
        CUDA_TRY(cudaMemcpy(pl->d_rec, hp.rec.data(), sizeof(FreqRec) * hp.F, cudaMemcpyHostToDevice));
    }
    if (hp.family == FAM_TABLE) {
        const size_t cnt = (size_t)hp.F * (size_t)hp.table_len;
        std::vector<cx<T>> t(cnt);
        for (size_t i = 0; i < cnt; ++i) { t[i].x = (T)hp.table[2 * i]; t[i].y = (T)hp.table[2 * i + 1]; }
        CUDA_TRY(cudaMalloc(&pl->d_table, cnt * sizeof(cx<T>)));
        CUDA_TRY(cudaMemcpy(pl->d_table, t.data(), cnt * sizeof(cx<T>), cudaMemcpyHostToDevice));
    }
    // filter spectra: forward transforms of the filter's real and imaginary parts (the sub-plan's own forward kernels),
    // combined into the table rows of the two one-row TABLE sub-plans
    T* d_h = nullptr;
    cx<T>* d_spec = nullptr;
    void* ws = nullptr;
    size_t wsb = 0;
    nwcwt_workspace_bytes(pl->blu_f, 2, &wsb);
    CUDA_TRY(cudaMalloc((void**)&d_h, sizeof(T) * 2 * (size_t)M));
    CUDA_TRY(cudaMalloc((void**)&d_spec, sizeof(cx<T>) * 2 * (size_t)M));
    if (wsb) CUDA_TRY(cudaMalloc(&ws, wsb));
    CUDA_TRY(cudaMemcpy(d_h, h.data(), sizeof(T) * 2 * (size_t)M, cudaMemcpyHostToDevice));
    rc = run_transform<T>(pl->blu_f, d_h, nullptr, d_spec, 2, 0, 0, 0, 0, ws, wsb, 0, true);
    if (!rc) {
        nwcwt_blu_filter_kernel<T><<<(unsigned)((M + 255) / 256), 256>>>(d_spec, (cx<T>*)pl->blu_f->d_table, (cx<T>*)pl->blu_i->d_table, M);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) rc = fail(NWCWT_ERR_CUDA, std::string("bluestein filter: ") + cudaGetErrorString(e));
    }
    cudaFree(d_h);
    cudaFree(d_spec);
    if (ws) cudaFree(ws);
    if (rc) return rc;
    pl->on_device = true;
    return 0;
}

// Resampled short rows: per group the M-point twiddles, the centred bands with their weight table (1 / N and the
// group's equaliser folded in), the interpolation weights and the frequency map; then the array of Short3Group itself.
template <typename T>
static int ensure_device_short3(nwcwt_plan* pl) {
    HostPlan& hp = pl->hp;
    std::vector<Short3Group<T>> gs;
    int unit = 0;
    auto up = [&](const void* src, size_t bytes, void** dst) -> int {
        CUDA_TRY(cudaMalloc(dst, bytes ? bytes : 16));
        pl->s3_allocs.push_back(*dst);
        if (bytes) CUDA_TRY(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
        return 0;
    };
    int rc;
    for (const MrGroup& mg : hp.groups) {
        HostPlan sub = *mg.sub;   // build_weight_table sets FreqRec::woff
        std::vector<T> wtab;
        if (!build_weight_table<T>(sub, mg.D > 1 ? mg.eq.data() : nullptr, (size_t)256 << 20, wtab))
            return fail(NWCWT_ERR_UNSUPPORTED, "short rows: weight table too large");
        Short3Group<T> g;
        if (!short3_fill_group<T>(g, hp.N, sub.N, mg.D, mg.K, sub.F))
            return fail(NWCWT_ERR_UNSUPPORTED, "short rows: decimation without a kernel shape");
        g.unit0 = unit;
        unit += g.nunits;
        g.st = sub.stS;
        std::vector<cx<T>> tw;
        fill_twiddles<T>(tw, sub.N, sub.N, 1);
        void* d = nullptr;
        if ((rc = up(tw.data(), tw.size() * sizeof(cx<T>), &d))) return rc;
        g.tw = (const cx<T>*)d;
        if ((rc = up(sub.rec.data(), sub.rec.size() * sizeof(FreqRec), &d))) return rc;
        g.rec = (const FreqRec*)d;
        if ((rc = up(wtab.data(), wtab.size() * sizeof(T), &d))) return rc;
        g.wtab = (const T*)d;
        std::vector<T> coefq(mg.coef.size() + 4);
        if (mg.D > 1) resample_coefq<T>(mg.coef.data(), mg.D, mg.K, g.PQ, coefq.data());
        if ((rc = up(coefq.data(), coefq.size() * sizeof(T), &d))) return rc;
        g.coefq = (const T*)d;
        if ((rc = up(mg.fidx.data(), mg.fidx.size() * sizeof(int), &d))) return rc;
        g.fmap = (const int*)d;
        std::vector<int> pos((size_t)sub.N);
        for (int k = 0; k < (int)sub.N; ++k) pos[(size_t)k] = fft2_dit_pos(sub.stS, k);
        if ((rc = up(pos.data(), pos.size() * sizeof(int), &d))) return rc;
        g.ditpos = (const int*)d;
        gs.push_back(g);
    }
    CUDA_TRY(cudaMalloc(&pl->d_s3groups, gs.size() * sizeof(Short3Group<T>)));
    CUDA_TRY(cudaMemcpy(pl->d_s3groups, gs.data(), gs.size() * sizeof(Short3Group<T>), cudaMemcpyHostToDevice));
    pl->s3_units = unit;
    return 0;
}

template <typename T>
static int ensure_device_t(nwcwt_plan* pl) {
    HostPlan& hp = pl->hp;
    CUDA_TRY(cudaSetDevice(hp.device));
    if (pl->on_device) return 0;
    int rc;
    if (hp.path == 4) return ensure_device_bluestein<double>(pl);
    if (hp.path == 0) {
        if ((rc = upload_tw<T>(&pl->d_tw, hp.N, hp.N, 1))) return rc;
    } else {
        if (hp.generic_ok) {
            if ((rc = upload_tw<T>(&pl->d_twA, hp.N1, hp.N1, 1))) return rc;
            if ((rc = upload_tw<T>(&pl->d_twB, hp.N2, hp.N2, 1))) return rc;
        }
        const long long nL = 1LL << hp.lb, nH = (hp.N + nL - 1) / nL;
        if ((rc = upload_tw<T>(&pl->d_twL, nL, hp.N, 1))) return rc;
        if ((rc = upload_tw<T>(&pl->d_twH, nH, hp.N, nL))) return rc;
        if (hp.fast) {
            if ((rc = upload_tw<T>(&pl->d_twA2, hp.N1f, hp.N1f, 1))) return rc;
            if ((rc = upload_tw<T>(&pl->d_twB2, hp.N2f, hp.N2f, 1))) return rc;
            std::vector<int> pos((size_t)hp.N1f);
            for (int k1 = 0; k1 < hp.N1f; ++k1) pos[(size_t)k1] = fft2_dit_pos(hp.stA2, k1);
            CUDA_TRY(cudaMalloc((void**)&pl->d_ditpos, pos.size() * sizeof(int)));
            CUDA_TRY(cudaMemcpy(pl->d_ditpos, pos.data(), pos.size() * sizeof(int), cudaMemcpyHostToDevice));
        }
    }
    if (hp.path == 1 && hp.fast && !getenv("NWCWT_NO_WTAB")) {
        std::vector<T> tab;
        if (build_weight_table<T>(hp, pl->eq_host ? pl->eq_host->data() : nullptr, (size_t)512 << 20, tab)) {
            CUDA_TRY(cudaMalloc(&pl->d_wtab, tab.size() * sizeof(T)));
            CUDA_TRY(cudaMemcpy(pl->d_wtab, tab.data(), tab.size() * sizeof(T), cudaMemcpyHostToDevice));
        }
    }
    if (hp.F > 0) {
        CUDA_TRY(cudaMalloc(&pl->d_rec, sizeof(FreqRec) * hp.F));
        CUDA_TRY(cudaMemcpy(pl->d_rec, hp.rec.data(), sizeof(FreqRec) * hp.F, cudaMemcpyHostToDevice));
    }
    if (hp.family == FAM_TABLE) {
        const size_t cnt = (size_t)hp.F * (size_t)hp.table_len;
        std::vector<cx<T>> t(cnt);
        for (size_t i = 0; i < cnt; ++i) {
            t[i].x = (T)hp.table[2 * i];
            t[i].y = (T)hp.table[2 * i + 1];
        }
        CUDA_TRY(cudaMalloc(&pl->d_table, cnt * sizeof(cx<T>)));
        CUDA_TRY(cudaMemcpy(pl->d_table, t.data(), cnt * sizeof(cx<T>), cudaMemcpyHostToDevice));
    }
    // opt in to large dynamic shared memory
    if (hp.path == 0) {
        CUDA_TRY(prepare_short<T>());
        if (hp.short2) CUDA_TRY(prepare_short2<T>());
        if (hp.short3) {
            if ((rc = ensure_device_short3<T>(pl))) return rc;
            CUDA_TRY(prepare_short3<T>());
        }
    } else {
        CUDA_TRY(prepare_passA<T>());
        CUDA_TRY(prepare_passB<T>());
        if (hp.fast) {
            CUDA_TRY(Long2Dispatch<T>::prepare(hp.cfgA));
            CUDA_TRY(Long2Dispatch<T>::prepare(hp.cfgB));
            // measured (profiles/r01/shape_sweep.md): an L2 persistence window on the Tm ring makes the step
            // slower (24.2-26.3 vs 23.3 ms), so it is opt-in for experiments only
            if (getenv("NWCWT_L2_PERSIST")) {
                int v = 0;
                cudaDeviceGetAttribute(&v, cudaDevAttrMaxPersistingL2CacheSize, hp.device);
                pl->l2_persist_max = (size_t)(v > 0 ? v : 0);
                cudaDeviceGetAttribute(&v, cudaDevAttrMaxAccessPolicyWindowSize, hp.device);
                pl->l2_window_max = (size_t)(v > 0 ? v : 0);
                if (pl->l2_persist_max) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, pl->l2_persist_max);
            }
            int ns = pl->is_sub ? 1 : (hp.groups.empty() ? 2 : 3);   // resampled rows: three kernels per launch group (measured 9.0 vs 10.1 ms on cfg2)
            if (const char* e = getenv("NWCWT_STREAMS")) ns = pl->is_sub ? 1 : atoi(e);
            ns = ns < 1 ? 1 : ns > nwcwt_plan::MAX_AUX ? nwcwt_plan::MAX_AUX : ns;
            if (ns > 1) {
                for (int i = 0; i < ns; ++i) {
                    CUDA_TRY(cudaStreamCreateWithFlags(&pl->aux[i], cudaStreamNonBlocking));
                    CUDA_TRY(cudaEventCreateWithFlags(&pl->ev_join[i], cudaEventDisableTiming));
                }
                CUDA_TRY(cudaEventCreateWithFlags(&pl->ev_fork, cudaEventDisableTiming));
                pl->n_aux = ns;
            }
        }
    }
    // resampled groups: their sub-plans' tables and the interpolation tables
    for (size_t gi = 0; gi < pl->groups.size(); ++gi) {
        nwcwt_plan::Group& g = pl->groups[gi];
        const MrGroup& mg = hp.groups[gi];
        if (mg.D > 1) g.sub->eq_host = &mg.eq;
        if ((rc = ensure_device_t<T>(g.sub))) return rc;
        CUDA_TRY(cudaMalloc((void**)&g.d_fmap, sizeof(int) * mg.fidx.size()));
        CUDA_TRY(cudaMemcpy(g.d_fmap, mg.fidx.data(), sizeof(int) * mg.fidx.size(), cudaMemcpyHostToDevice));
        if (mg.D > 1) {
            ResampleVecShape vs;
            const bool vec = resample_vec_shape<T>(mg.D, mg.K, g.sub->hp.N, vs);
            if (!vec && !has_resample<T>(mg.K)) return fail(NWCWT_ERR_UNSUPPORTED, "no interpolation kernel for this tap count");
            if (vec) {
                std::vector<T> q(mg.coef.size());
                resample_coefq<T>(mg.coef.data(), mg.D, mg.K, vs.PQ, q.data());
                CUDA_TRY(cudaMalloc(&g.d_coefq, sizeof(T) * q.size()));
                CUDA_TRY(cudaMemcpy(g.d_coefq, q.data(), sizeof(T) * q.size(), cudaMemcpyHostToDevice));
            }
            std::vector<T> c(mg.coef.begin(), mg.coef.end()), e(mg.eq.begin(), mg.eq.end());
            CUDA_TRY(cudaMalloc(&g.d_coef, sizeof(T) * c.size()));
            CUDA_TRY(cudaMemcpy(g.d_coef, c.data(), sizeof(T) * c.size(), cudaMemcpyHostToDevice));
            CUDA_TRY(cudaMalloc(&g.d_eq, sizeof(T) * e.size()));
            CUDA_TRY(cudaMemcpy(g.d_eq, e.data(), sizeof(T) * e.size(), cudaMemcpyHostToDevice));
            CUDA_TRY(cudaMalloc((void**)&g.d_t0, sizeof(int) * mg.t0.size()));
            CUDA_TRY(cudaMemcpy(g.d_t0, mg.t0.data(), sizeof(int) * mg.t0.size(), cudaMemcpyHostToDevice));
            g.t0min = *std::min_element(mg.t0.begin(), mg.t0.end());
        }
    }
    if (!pl->groups.empty()) {
        CUDA_TRY(prepare_resample<T>());
        CUDA_TRY(RsVecDispatch<T>::prepare());
    }
    pl->on_device = true;
    return 0;
}

static int ensure_device(nwcwt_plan* pl) {
    return pl->hp.dtype == NWCWT_F32 ? ensure_device_t<float>(pl) : ensure_device_t<double>(pl);
}

template <typename T>
static SpecParams<T> make_spec(const nwcwt_plan* pl) {
    const HostPlan& hp = pl->hp;
    SpecParams<T> sp;
    sp.family = hp.family;
    sp.grid_off = hp.grid_off;
    sp.df = hp.df;
    sp.p0 = hp.p0;
    sp.p1 = hp.p1;
    sp.p2 = hp.family == FAM_MORSE ? hp.p0 / hp.p1 : hp.p2;   // (self.b / self.r), wavelets.py:72
    sp.norm = (T)(1.0 / (double)hp.data_len());                 // 1/N of ifft, base.py:406
    sp.rec = (const FreqRec*)pl->d_rec;
    sp.table = (const cx<T>*)pl->d_table;
    sp.table_len = hp.table_len;
    sp.wtab = (const T*)pl->d_wtab;
    return sp;
}

// ---------------------------------------------------------------------------------
// launches
// ---------------------------------------------------------------------------------
static int device_sms(int device) {
    static int cached[64] = {0};
    if (device < 64 && cached[device]) return cached[device];
    int n = 148;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device);
    if (device < 64) cached[device] = n;
    return n;
}

// fused short-row kernel on the packed engine: one CTA per signal pair (and frequency share)
template <typename T>
static int launch_short2(nwcwt_plan* pl, const void* signals, void* out, long long S, int output, int bl, long long blo,
                         long long bhi, cudaStream_t stream) {
    const HostPlan& hp = pl->hp;
    Short2Params<T> P;
    memset(&P, 0, sizeof(P));
    P.signals = (const T*)signals;
    P.out = out;
    P.N = (int)hp.N;
    P.F = hp.F;
    P.S = (int)S;
    P.tpsh = hp.tpshS;
    P.out_mode = output;
    P.bl_mode = bl;
    P.bl_lo = (int)blo;
    P.bl_hi = (int)bhi;
    P.st = hp.stS;
    P.tw = (const cx<T>*)pl->d_tw;
    P.sp = make_spec<T>(pl);
    const long long npairs = (S + 1) / 2;
    const int ngroups = (hp.F + (1 << hp.tpshS) - 1) >> hp.tpshS;
    const long long target = 6LL * device_sms(hp.device);
    long long fs = (target + npairs - 1) / npairs;
    if (fs > ngroups) fs = ngroups;
    if (fs < 1) fs = 1;
    P.fsplit = (int)fs;
    const long long grid = npairs * fs;
    if (grid > 2147483647LL) return fail(NWCWT_ERR_INVALID, "too many signals for one launch");
    int sp = g_no_static ? 0 : static_plan_id(hp.stS, hp.tpshS);
    if (!has_static_short2<T>(sp)) sp = 0;
    { LaunchScope ls(0, stream); CUDA_TRY(launch_short2<T>(sp, P, (unsigned)grid, hp.nthrS2, hp.smem_S2, stream)); }
    return 0;
}

// fused epoch reductions on the short-row kernel: one CTA per (channel, share of the frequency groups)
template <typename T>
static int launch_short2_epochs_t(nwcwt_plan* pl, const void* signals, void* out, long long C, long long E, int kind, void* acc,
                                  cudaStream_t stream) {
    const HostPlan& hp = pl->hp;
    Short2Params<T> P;
    memset(&P, 0, sizeof(P));
    P.signals = (const T*)signals;
    P.out = out;
    P.N = (int)hp.N;
    P.F = hp.F;
    P.S = (int)(C * E);
    P.tpsh = hp.tpshS;
    P.st = hp.stS;
    P.tw = (const cx<T>*)pl->d_tw;
    P.sp = make_spec<T>(pl);
    P.n_epochs = (int)E;
    P.acc = (cx<T>*)acc;
    const int ngroups = (hp.F + (1 << hp.tpshS) - 1) >> hp.tpshS;
    const long long target = 6LL * device_sms(hp.device);
    long long fs = (target + C - 1) / C;
    if (fs > ngroups) fs = ngroups;
    if (fs < 1) fs = 1;
    P.fsplit = (int)fs;
    const long long grid = C * fs;
    if (grid > 2147483647LL) return fail(NWCWT_ERR_INVALID, "too many channels for one launch");
    int sp = g_no_static ? 0 : static_plan_id(hp.stS, hp.tpshS);
    if (!has_static_short2<T>(sp)) sp = 0;
    { LaunchScope ls(0, stream); CUDA_TRY(launch_short2_epochs<T>(sp, kind, P, (unsigned)grid, hp.nthrS2, hp.smem_S2, stream)); }
    return 0;
}

// resampled short rows: one CTA per signal pair (and share of the work units)
template <typename T>
static int launch_short3_t(nwcwt_plan* pl, const void* signals, void* out, long long S, int output, int bl, long long blo,
                           long long bhi, cudaStream_t stream) {
    const HostPlan& hp = pl->hp;
    Short3Params<T> P;
    memset(&P, 0, sizeof(P));
    P.signals = (const T*)signals;
    P.out = out;
    P.N = (int)hp.N;
    P.F_out = hp.F;
    P.S = (int)S;
    P.out_mode = output;
    P.bl_mode = bl;
    P.bl_lo = (int)blo;
    P.bl_hi = (int)bhi;
    P.st = hp.stS;
    P.tw = (const cx<T>*)pl->d_tw;
    P.groups = (const Short3Group<T>*)pl->d_s3groups;
    P.ngroups = (int)hp.groups.size();
    P.nunits = pl->s3_units;
    const long long npairs = (S + 1) / 2;
    const long long target = 6LL * device_sms(hp.device);
    long long fs = (target + npairs - 1) / npairs;
    if (fs > P.nunits) fs = P.nunits;
    if (fs < 1) fs = 1;
    P.fsplit = (int)fs;
    const long long grid = npairs * fs;
    if (grid > 2147483647LL) return fail(NWCWT_ERR_INVALID, "too many signals for one launch");
    { LaunchScope ls(0, stream); CUDA_TRY(launch_short3<T>(P, (unsigned)grid, hp.nthrS3, hp.smem_S3, stream)); }
    return 0;
}

template <typename T>
static int launch_short(nwcwt_plan* pl, const void* signals, void* out, void* spectra, long long S, int output,
                        int bl, long long blo, long long bhi, cudaStream_t stream, bool forward_only) {
    const HostPlan& hp = pl->hp;
    if (hp.short3 && !forward_only && !g_force_generic && !g_exact && output != NWCWT_OUT_CWT)
        return launch_short3_t<T>(pl, signals, out, S, output, bl, blo, bhi, stream);
    if (hp.short2 && !forward_only && !g_force_generic)
        return launch_short2<T>(pl, signals, out, S, output, bl, blo, bhi, stream);
    ShortParams<T> P;
    memset(&P, 0, sizeof(P));
    P.signals = (const T*)signals;
    P.out = out;
    P.spectra = (cx<T>*)spectra;
    P.N = (int)hp.N;
    P.F = forward_only ? 0 : hp.F;
    P.S = (int)S;
    P.tsh = hp.tsh;
    P.pitch = hp.pitch;
    P.out_mode = output;
    P.bl_mode = bl;
    P.bl_lo = (int)blo;
    P.bl_hi = (int)bhi;
    P.st = hp.st;
    P.tw = (const cx<T>*)pl->d_tw;
    P.sp = make_spec<T>(pl);
    const int TT = 1 << hp.tsh;
    const int ngroups = forward_only ? 1 : (hp.F + TT - 1) / TT;
    // enough CTAs to fill the machine a few times over when there are few signals
    const long long target = 4LL * device_sms(hp.device);
    long long fs = (target + S - 1) / S;
    if (fs > ngroups) fs = ngroups;
    if (fs < 1) fs = 1;
    P.fsplit = (int)fs;
    const long long grid = S * fs;
    if (grid > 2147483647LL) return fail(NWCWT_ERR_INVALID, "too many signals for one launch");
    { LaunchScope ls(0, stream); CUDA_TRY(launch_short<T>(P, (unsigned)grid, hp.nthr_short, hp.smem_short, stream)); }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

template <typename T>
static LongParams<T> make_long(nwcwt_plan* pl) {
    const HostPlan& hp = pl->hp;
    LongParams<T> P;
    memset(&P, 0, sizeof(P));
    P.N = hp.N;
    P.N1 = hp.N1;
    P.N2 = hp.N2;
    P.tshA = hp.tshA;
    P.pitchA = hp.pitchA;
    P.tshB = hp.tshB;
    P.stA = hp.stA;
    P.stB = hp.stB;
    P.twA = (const cx<T>*)pl->d_twA;
    P.twB = (const cx<T>*)pl->d_twB;
    P.twH = (const cx<T>*)pl->d_twH;
    P.twL = (const cx<T>*)pl->d_twL;
    P.lb = hp.lb;
    P.tm_stride = hp.tm_stride;
    P.sp = make_spec<T>(pl);
    return P;
}

// intermediate ring of the long path: the forward transforms use the generic two-pass kernels,
// the inverse ones the fast kernels when the plan has them; one region serves both (and the groups' sub-plans)
// One stream slot of the rings holds the rows of one launch group of ANY of the plan's transforms (the plan's own exact
// rows or a resampled group's decimated rows): the slot size is the maximum over them, so that launch groups of
// different decimations running on different streams never overlap.
static size_t tm_slot_elems(const nwcwt_plan* pl) {
    size_t e = pl->hp.fast ? (size_t)pl->hp.ring2 * (size_t)pl->hp.tm_stride2 : 0;
    for (const nwcwt_plan::Group& g : pl->groups) e = std::max(e, tm_slot_elems(g.sub));
    return e;
}
static size_t y_slot_elems(const nwcwt_plan* pl) {
    size_t e = 0;
    for (size_t gi = 0; gi < pl->groups.size(); ++gi) {
        const HostPlan& sh = pl->groups[gi].sub->hp;
        if (pl->hp.groups[gi].D > 1) e = std::max(e, (size_t)sh.ring2 * (size_t)sh.N);
    }
    return e;
}
static size_t tm_ring_bytes(const nwcwt_plan* pl) {
    const HostPlan& hp = pl->hp;
    const size_t cs = cx_size(hp.dtype);
    size_t b = hp.generic_ok ? (size_t)hp.ring * hp.tm_stride * cs : 0;
    b = std::max<size_t>(b, (size_t)nwcwt_plan::MAX_AUX * tm_slot_elems(pl) * cs);
    return align_up(b, 256);
}
// decimated rows of the resampled groups: per stream slot, the rows of one launch group
static size_t y_ring_bytes(const nwcwt_plan* pl) {
    return align_up(y_slot_elems(pl) * cx_size(pl->hp.dtype) * nwcwt_plan::MAX_AUX, 256);
}

template <typename T>
static Long2Params<T> make_long2(const nwcwt_plan* pl) {
    const HostPlan& hp = pl->hp;
    Long2Params<T> P;
    memset(&P, 0, sizeof(P));
    P.N = hp.N;
    P.xstride = hp.data_len();
    P.N1 = hp.N1f;
    P.N2 = hp.N2f;
    P.F = hp.F;
    P.tpshA = hp.tpshA;
    P.tpshB = hp.tpshB;
    P.stA = hp.stA2;
    P.stB = hp.stB2;
    P.twA = (const cx<T>*)pl->d_twA2;
    P.twB = (const cx<T>*)pl->d_twB2;
    P.twH = (const cx<T>*)pl->d_twH;
    P.twL = (const cx<T>*)pl->d_twL;
    P.ditpos = pl->d_ditpos;
    P.lb = hp.lb;
    P.tm_stride = hp.tm_stride2;
    P.sp = make_spec<T>(pl);
    if (hp.stA2.nst >= 1) P.dstepA = make_fastdiv((uint32_t)std::max(1, hp.N1f / hp.stA2.radix[hp.stA2.nst - 1]));
    return P;
}

// Inverse transforms of gs signals x the frequencies of plan `ep` on the packed kernels: ep is the plan itself
// (every row, exact) or the sub-plan of one of its groups.  Rows go out in launch pairs of ep->ring2 rows that
// alternate between the plan's auxiliary streams (forked from / joined into `stream` by the caller).
//   grp == nullptr or D == 1: pass B writes the final output rows (through the group's frequency map)
//   D > 1: pass B writes the decimated rows into the stream slot's Y buffer and the interpolation kernel
//          (nw_resample.cuh) turns them into the output rows
template <typename T>
static int inverse_rows(nwcwt_plan* pl, nwcwt_plan* ep, int gidx, const cx<T>* X, cx<T>* Tm, cx<T>* Y, void* out_s0, int gs,
                        int output, cudaStream_t stream, int& launch_idx) {
    const HostPlan& hp = pl->hp;
    const HostPlan& eh = ep->hp;
    const MrGroup* mg = gidx >= 0 ? &hp.groups[(size_t)gidx] : nullptr;
    const nwcwt_plan::Group* dg = gidx >= 0 ? &pl->groups[(size_t)gidx] : nullptr;
    const int D = mg ? mg->D : 1;
    const size_t esz = (output == NWCWT_OUT_CWT) ? sizeof(cx<T>) : sizeof(T);
    Long2Params<T> Q = make_long2<T>(ep);
    Q.X = X;
    Q.xstride = hp.N;
    const unsigned tA = (unsigned)((eh.N2f + (2 << eh.tpshA) - 1) / (2 << eh.tpshA));
    const unsigned tB = (unsigned)((eh.N1f + (2 << eh.tpshB) - 1) / (2 << eh.tpshB));
    const long long rows = (long long)gs * eh.F;
    const int ns = pl->n_aux;
    const size_t tm_slot = tm_slot_elems(pl), y_slot = y_slot_elems(pl);   // slot sizes common to all groups of the plan
    // kernels specialised for this plan at compile time, where the library has them
    int spA = g_no_static ? 0 : static_plan_id(eh.stA2, eh.tpshA), spB = g_no_static ? 0 : static_plan_id(eh.stB2, eh.tpshB);
    if (!Long2Dispatch<T>::has(eh.cfgA, 0, spA)) spA = 0;
    if (!Long2Dispatch<T>::has(eh.cfgB, 1, spB)) spB = 0;
    ResampleParams<T> R;
    ResampleShape shp{1, 1, 0, 0, 0};
    ResampleVecShape vshp;
    ResampleDirShape dshp;
    bool vec = false, dir = false;
    if (D > 1) {
        vec = resample_vec_shape<T>(D, mg->K, eh.N, vshp) && dg->d_coefq;
        dir = vec && g_rs_dir && resample_dir_shape<T>(D, mg->K, eh.N, rows, eh.F, dshp) && dshp.PQ == vshp.PQ;
        if (dir) { static const int rs_ctas = env_int("NWCWT_RS_CTAS", 0); if (rs_ctas >= 1 && rs_ctas <= 8) dshp.ctas_per_sm = rs_ctas; }
        shp = resample_shape<T>(D, mg->K);
        memset(&R, 0, sizeof(R));
        R.ystride = eh.N;
        R.out = out_s0;
        R.N = hp.N;
        R.M = (int)eh.N;
        R.D = D;
        R.coef = (const T*)dg->d_coef;
        R.t0 = dg->d_t0;
        R.t0min = dg->t0min;
        R.fmap = dg->d_fmap;
        R.F = eh.F;
        R.F_out = hp.F;
        R.WR = shp.WR;
        R.WP = shp.WP;
        R.RS = shp.RS;
        R.dRD = make_fastdiv((uint32_t)(ResampleRun<T>::R * D));
        if (vec) { R.WR = vshp.WR; R.WP = vshp.WP; R.RS = (int)vshp.gbytes; R.dRD = vshp.dRD; R.dGT = make_fastdiv(vshp.items); R.dGT.d = vshp.items; }
        if (dir) { R.G = dshp.G; R.MW = dshp.MW; R.NSUB = dshp.NSUB; R.dG = make_fastdiv((uint32_t)dshp.G); R.dF = make_fastdiv((uint32_t)eh.F); R.dGT = make_fastdiv(dshp.items); R.dGT.d = dshp.items; }
        R.coefq = (const T*)dg->d_coefq;
        Q.eq = (const T*)dg->d_eq;
        Q.out_mode = NWCWT_OUT_CWT;
    } else {
        Q.out = out_s0;
        Q.out_mode = output;
        if (dg) { Q.fmap = dg->d_fmap; Q.F_out = hp.F; }
    }
    for (long long r0 = 0; r0 < rows; r0 += eh.ring2, ++launch_idx) {
        const int g = (int)std::min<long long>(eh.ring2, rows - r0);
        const int slot = ns > 1 ? launch_idx % ns : 0;
        cudaStream_t st = ns > 1 ? pl->aux[slot] : stream;
        Q.row0 = (int)r0;
        Q.Tm = Tm + (size_t)slot * tm_slot;
        Q.narrow = group_narrow(eh, r0, g) ? 1 : 0;
        cx<T>* yslot = nullptr;
        if (D > 1) {
            yslot = Y + (size_t)slot * y_slot;
            Q.out = (char*)yslot - (size_t)r0 * (size_t)eh.N * sizeof(cx<T>);   // pass B indexes rows from row0
        }
        { LaunchScope ls(3, st); CUDA_TRY(Long2Dispatch<T>::A(eh.cfgA, spA, Q, dim3(tA, g), eh.nthrA2, eh.smem_A2, st)); }
        { LaunchScope ls(4, st); CUDA_TRY(Long2Dispatch<T>::B(eh.cfgB, spB, Q, dim3(tB, g), eh.nthrB2, eh.smem_B2, st)); }
        if (D > 1) {
            R.y = yslot;
            R.row0 = (int)r0;
            const unsigned tiles = (unsigned)((eh.N + shp.C - 1) / shp.C);
            LaunchScope ls(6, st);
            R.nrows = g;
            if (dir) CUDA_TRY(RsVecDispatch<T>::run_dir(dshp.PQ, output, mg->K, R, dim3(resample_dir_grid(dshp, g, device_sms(hp.device))), st));
            else if (vec) CUDA_TRY(RsVecDispatch<T>::run(vshp.PQ, output, mg->K, vshp.R, R, dim3(resample_vec_grid(vshp, g, device_sms(hp.device))), vshp.smem, st));
            else CUDA_TRY(launch_resample<T>(mg->K, output, R, dim3(tiles, g), 32 * shp.WR * shp.WP, shp.smem, st));
        }
    }
    (void)esz;
    return 0;
}

// workspace layout of the long path: [ring] spectra of N, the Tm ring, the Y ring of the resampled groups
template <typename T>
static int launch_long(nwcwt_plan* pl, const void* signals, void* out, void* spectra_out, long long S, int output,
                       int bl, long long blo, long long bhi, void* ws, size_t ws_bytes, cudaStream_t stream,
                       bool forward_only) {
    const HostPlan& hp = pl->hp;
    const size_t xbytes = align_up((size_t)hp.ring * hp.N * sizeof(cx<T>), 256);
    const size_t tbytes = tm_ring_bytes(pl), ybytes = y_ring_bytes(pl);
    if (ws_bytes < xbytes + tbytes + ybytes || !ws) return fail(NWCWT_ERR_WORKSPACE, "workspace too small");
    cx<T>* X = (cx<T>*)ws;
    cx<T>* Tm = (cx<T>*)((char*)ws + xbytes);
    cx<T>* Y = (cx<T>*)((char*)ws + xbytes + tbytes);
    LongParams<T> P = make_long<T>(pl);
    P.Tm = Tm;
    const int TA = 1 << hp.tshA, TB = 1 << hp.tshB;
    const unsigned tilesA = hp.generic_ok ? (unsigned)((hp.N2 + TA - 1) / TA) : 0u, tilesB = hp.generic_ok ? (unsigned)((hp.N1 + TB - 1) / TB) : 0u;
    const size_t esz = (output == NWCWT_OUT_CWT) ? sizeof(cx<T>) : sizeof(T);
    for (long long s0 = 0; s0 < S; s0 += hp.ring) {
        const int gs = (int)std::min<long long>(hp.ring, S - s0);
        // forward transforms of gs signals (scipy.fftpack.fft, base.py:399)
        cx<T>* Xdst = forward_only ? (cx<T>*)spectra_out + (size_t)s0 * hp.N : X;
        if (hp.fast && !g_force_generic) {
            // packed kernels, conjugate transform; chunks of ring2 signals share the first Tm slot
            Long2Params<T> Qf = make_long2<T>(pl);
            Qf.Tm = Tm;
            Qf.out_mode = NWCWT_OUT_CWT;
            const unsigned tA = (unsigned)((hp.N2f + (2 << hp.tpshA) - 1) / (2 << hp.tpshA));
            const unsigned tB = (unsigned)((hp.N1f + (2 << hp.tpshB) - 1) / (2 << hp.tpshB));
            for (int c0 = 0; c0 < gs; c0 += hp.ring2) {
                const int gc = std::min(hp.ring2, gs - c0);
                Qf.signal = (const T*)signals + (size_t)(s0 + c0) * hp.N;
                Qf.out = Xdst + (size_t)c0 * hp.N;
                Qf.row0 = 0;
                { LaunchScope ls(1, stream); CUDA_TRY(Long2Dispatch<T>::A(hp.cfgA, -2, Qf, dim3(tA, gc), hp.nthrA2, hp.smem_A2, stream)); }
                { LaunchScope ls(2, stream); CUDA_TRY(Long2Dispatch<T>::B(hp.cfgB, -2, Qf, dim3(tB, gc), hp.nthrB2, hp.smem_B2, stream)); }
            }
        } else {
            if (!hp.generic_ok) return fail(NWCWT_ERR_UNSUPPORTED, "this length has no generic two-pass plan (packed kernels only)");
            P.signal = (const T*)signals + (size_t)s0 * hp.N;
            P.Xout = Xdst;
            { LaunchScope ls(1, stream); CUDA_TRY(launch_passA<T>(-1, P, dim3(tilesA, gs), hp.nthr_long, hp.smem_A, stream)); }
            { LaunchScope ls(2, stream); CUDA_TRY(launch_passB<T>(-1, P, dim3(tilesB, gs), hp.nthr_long, hp.smem_B, stream)); }
        }
        if (forward_only) continue;
        if (hp.fast && !g_force_generic) {
            // inverse transforms of all gs * F rows of the signal group
            void* out_s0 = (char*)out + (size_t)s0 * hp.F * (size_t)hp.N * esz;
            const int ns = pl->n_aux;
            if (ns > 1) {
                CUDA_TRY(cudaEventRecord(pl->ev_fork, stream));
                for (int i = 0; i < ns; ++i) CUDA_TRY(cudaStreamWaitEvent(pl->aux[i], pl->ev_fork, 0));
            }
            int launch_idx = 0, rc = 0;
            const bool resampled = !pl->groups.empty() && output != NWCWT_OUT_CWT && !g_exact;
            if (resampled) {
                for (size_t gi = 0; gi < pl->groups.size() && !rc; ++gi)
                    rc = inverse_rows<T>(pl, pl->groups[gi].sub, (int)gi, X, Tm, Y, out_s0, gs, output, stream, launch_idx);
            } else {
                rc = inverse_rows<T>(pl, pl, -1, X, Tm, Y, out_s0, gs, output, stream, launch_idx);
            }
            if (rc) return rc;
            if (ns > 1)
                for (int i = 0; i < ns; ++i) {
                    CUDA_TRY(cudaEventRecord(pl->ev_join[i], pl->aux[i]));
                    CUDA_TRY(cudaStreamWaitEvent(stream, pl->ev_join[i], 0));
                }
            if (bl != NWCWT_BL_NONE) {
                LaunchScope ls(5, stream);
                nwcwt_baseline_rows_kernel<T><<<(unsigned)((long long)gs * hp.F), 512, 0, stream>>>((T*)out_s0, hp.N, bl, (int)blo, (int)bhi);
            }
            continue;
        }
        if (!hp.generic_ok) return fail(NWCWT_ERR_UNSUPPORTED, "this length has no generic two-pass plan (packed kernels only)");
        for (int si = 0; si < gs; ++si) {
            P.X = X + (size_t)si * hp.N;
            char* out_s = (char*)out + (size_t)(s0 + si) * hp.F * (size_t)hp.N * esz;
            for (int f0 = 0; f0 < hp.F; f0 += hp.ring) {
                const int g = std::min(hp.ring, hp.F - f0);
                P.f0 = f0;
                P.out = out_s + (size_t)f0 * (size_t)hp.N * esz;
                P.out_mode = output;
                { LaunchScope ls(3, stream); CUDA_TRY(launch_passA<T>(1, P, dim3(tilesA, g), hp.nthr_long, hp.smem_A, stream)); }
                { LaunchScope ls(4, stream); CUDA_TRY(launch_passB<T>(1, P, dim3(tilesB, g), hp.nthr_long, hp.smem_B, stream)); }
            }
            if (bl != NWCWT_BL_NONE) {
                LaunchScope ls(5, stream);
                nwcwt_baseline_rows_kernel<T><<<hp.F, 512, 0, stream>>>((T*)out_s, hp.N, bl, (int)blo, (int)bhi);
            }
        }
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// workspace of a Bluestein plan: spectra of blu_rows signals, the real inputs and complex results of one chunk of
// convolutions (2 blu_rows sequences of length blu_m each), and the sub-plans' own workspace
static size_t bluestein_ws_bytes(const nwcwt_plan* pl) {
    const size_t rs = 8, R2 = 2 * (size_t)pl->blu_rows, M = (size_t)pl->blu_m;   // fp64 internals
    size_t sub = 0;
    nwcwt_workspace_bytes(pl->blu_f, (int64_t)R2, &sub);
    return align_up((size_t)pl->blu_rows * (size_t)pl->hp.N * 2 * rs, 256) + align_up(R2 * M * rs, 256) +
           align_up(R2 * M * 2 * rs, 256) + align_up(sub, 256);
}

template <typename TIO>
static int launch_bluestein(nwcwt_plan* pl, const void* signals, void* out, void* spectra_out, long long S, int output,
                            int bl, long long blo, long long bhi, void* ws, size_t ws_bytes, cudaStream_t stream,
                            bool forward_only) {
    typedef double T;   // internal precision of the chirp-z path
    const HostPlan& hp = pl->hp;
    const long long N = hp.N, M = pl->blu_m;
    const int RB = pl->blu_rows;
    if (ws_bytes < bluestein_ws_bytes(pl) || !ws) return fail(NWCWT_ERR_WORKSPACE, "workspace too small");
    const size_t R2 = 2 * (size_t)RB;
    char* p = (char*)ws;
    cx<T>* X = (cx<T>*)p;                 p += align_up((size_t)RB * (size_t)N * sizeof(cx<T>), 256);
    T* in = (T*)p;                        p += align_up(R2 * (size_t)M * sizeof(T), 256);
    cx<T>* conv = (cx<T>*)p;              p += align_up(R2 * (size_t)M * sizeof(cx<T>), 256);
    void* sub_ws = p;
    size_t sub_bytes = 0;
    nwcwt_workspace_bytes(pl->blu_f, (int64_t)R2, &sub_bytes);
    const cx<T>* chirp = (const cx<T>*)pl->d_chirp;
    const size_t esz = (output == NWCWT_OUT_CWT) ? sizeof(cx<TIO>) : sizeof(TIO);
    const unsigned gM = (unsigned)((M + 255) / 256), gN = (unsigned)((N + 255) / 256);
    SpecParams<T> sp = make_spec<T>(pl);
    int rc;
    for (long long s0 = 0; s0 < S; s0 += RB) {
        const int gs = (int)std::min<long long>(RB, S - s0);
        cx<TIO>* Xout = (cx<TIO>*)spectra_out + (size_t)s0 * (size_t)N;
        // forward (scipy.fftpack.fft, base.py:399): X = conj(c) . T(x conj(c); Hf)
        { LaunchScope ls(1, stream); nwcwt_blu_pre_fwd_kernel<TIO, T><<<dim3(gM, gs), 256, 0, stream>>>((const TIO*)signals + (size_t)s0 * (size_t)N, chirp, in, N, M); }
        if ((rc = run_transform<T>(pl->blu_f, in, conv, nullptr, 2 * gs, NWCWT_OUT_CWT, 0, 0, 0, sub_ws, sub_bytes, stream, false))) return rc;
        {
            LaunchScope ls(2, stream);
            if (forward_only) nwcwt_blu_post_fwd_kernel<T, TIO><<<dim3(gN, gs), 256, 0, stream>>>(conv, chirp, Xout, N, M);
            else nwcwt_blu_post_fwd_kernel<T, T><<<dim3(gN, gs), 256, 0, stream>>>(conv, chirp, X, N, M);
        }
        if (forward_only) continue;
        // inverse (ifft, base.py:404/406), rows (signal, frequency) in chunks: z = c . T(W X c; conj Hf) / N
        void* out_s0 = (char*)out + (size_t)s0 * (size_t)hp.F * (size_t)N * esz;
        const long long rows = (long long)gs * hp.F;
        for (long long r0 = 0; r0 < rows; r0 += RB) {
            const int g = (int)std::min<long long>(RB, rows - r0);
            { LaunchScope ls(3, stream); nwcwt_blu_pre_inv_kernel<T><<<dim3(gM, g), 256, 0, stream>>>(sp, X, chirp, in, N, M, hp.F, r0); }
            if ((rc = run_transform<T>(pl->blu_i, in, conv, nullptr, 2 * g, NWCWT_OUT_CWT, 0, 0, 0, sub_ws, sub_bytes, stream, false))) return rc;
            { LaunchScope ls(4, stream); nwcwt_blu_post_inv_kernel<T, TIO><<<dim3(gN, g), 256, 0, stream>>>(conv, chirp, out_s0, N, M, r0, output); }
        }
        if (bl != NWCWT_BL_NONE) {
            LaunchScope ls(5, stream);
            nwcwt_baseline_rows_kernel<TIO><<<(unsigned)rows, 512, 0, stream>>>((TIO*)out_s0, N, bl, (int)blo, (int)bhi);
        }
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

template <typename T>
static int run_transform(nwcwt_plan* pl, const void* signals, void* out, void* spectra, long long S, int output,
                         int bl, long long blo, long long bhi, void* ws, size_t ws_bytes, cudaStream_t stream,
                         bool forward_only) {
    if (pl->hp.path == 4)
        return launch_bluestein<T>(pl, signals, out, spectra, S, output, bl, blo, bhi, ws, ws_bytes, stream, forward_only);
    if (pl->hp.path == 0)
        return launch_short<T>(pl, signals, out, spectra, S, output, bl, blo, bhi, stream, forward_only);
    return launch_long<T>(pl, signals, out, spectra, S, output, bl, blo, bhi, ws, ws_bytes, stream, forward_only);
}

static int check_args(const nwcwt_plan* pl, int output, int bl, long long& blo, long long& bhi) {
    if (!pl) return fail(NWCWT_ERR_INVALID, "null plan");
    if (output < NWCWT_OUT_CWT || output > NWCWT_OUT_POWER) return fail(NWCWT_ERR_INVALID, "bad output mode");
    if (bl < NWCWT_BL_NONE || bl > NWCWT_BL_ZLOG) return fail(NWCWT_ERR_INVALID, "bad baseline mode");
    if (bl != NWCWT_BL_NONE) {
        if (output == NWCWT_OUT_CWT) return fail(NWCWT_ERR_INVALID, "baseline needs a real output (abs/power)");
        // python slice semantics of wave[lo:hi] for non-negative indices (base.py:49)
        if (blo < 0 || bhi < 0) return fail(NWCWT_ERR_INVALID, "negative baseline index");
        if (bhi > pl->hp.N) bhi = pl->hp.N;
        if (blo > bhi) blo = bhi;
    }
    if (pl->hp.F <= 0) return fail(NWCWT_ERR_INVALID, "plan has no frequencies");
    return 0;
}

// Graph replay of nwcwt_transform on the fast long path.  The launch sequence of a call depends only on the plan and on the
// call's arguments, so it is recorded once (stream capture on a stream of the plan; the auxiliary streams join the capture
// through the fork event) and replayed with one cudaGraphLaunch into the caller's stream: kernel-to-kernel dependencies are
// resolved on the device instead of through ~300 host launches and ~200 event calls per step.  A handful of argument sets is
// kept per plan (least recently used goes first); an argument set is recorded the second time it is seen.  Returns 1 when
// the call is not eligible or capture is unavailable (the caller then launches directly), 0 on success, a negative
// NWCWT_ERR_* code on failure.  NWCWT_GRAPH=0 switches it off.
static const int GRAPH_SLOTS = 4;
template <typename T>
static int transform_graph(nwcwt_plan* pl, const void* signals, void* out, long long S, int output, int bl, long long lo,
                           long long hi, void* ws, size_t ws_bytes, cudaStream_t stream) {
    static const int use_graph = env_int("NWCWT_GRAPH", 1);
    if (!use_graph || g_profile || pl->hp.path != 1 || !pl->hp.fast || g_force_generic || pl->n_aux < 2 || pl->l2_persist_max)
        return 1;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
        cudaGetLastError();
        return 1;                                   // the caller is capturing its own graph: launch into it directly
    }
    const int flags = (g_exact ? 1 : 0) | (g_no_static ? 2 : 0) | (g_rs_dir ? 4 : 0);
    nwcwt_plan::GraphEntry* hit = nullptr;
    for (nwcwt_plan::GraphEntry& g : pl->graphs)
        if (g.signals == signals && g.out == out && g.S == S && g.output == output && g.bl == bl && g.lo == lo && g.hi == hi &&
            g.ws == ws && g.ws_bytes == ws_bytes && g.flags == flags) { hit = &g; break; }
    if (!hit) {
        // first call with these arguments: remember them and launch directly - one-off calls never pay for a capture; the
        // second identical call (a processing loop that reuses its buffers) records the graph, later ones replay it
        if ((int)pl->graphs.size() >= GRAPH_SLOTS) {
            size_t victim = 0;
            for (size_t i = 1; i < pl->graphs.size(); ++i) if (pl->graphs[i].used < pl->graphs[victim].used) victim = i;
            if (pl->graphs[victim].exec) cudaGraphExecDestroy(pl->graphs[victim].exec);
            pl->graphs.erase(pl->graphs.begin() + (long)victim);
        }
        pl->graphs.push_back(nwcwt_plan::GraphEntry{signals, out, S, output, bl, lo, hi, ws, ws_bytes, flags, nullptr, 0, ++pl->graph_clock});
        return 1;
    }
    if (hit->launches < 0) return 1;                 // capture failed before for this argument set
    if (!hit->exec) {
        hit->launches = -1;
        if (!pl->cap_stream && cudaStreamCreateWithFlags(&pl->cap_stream, cudaStreamNonBlocking) != cudaSuccess) {
            cudaGetLastError();
            return 1;
        }
        const long long l0 = g_launches.load();
        if (cudaStreamBeginCapture(pl->cap_stream, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
            cudaGetLastError();
            return 1;
        }
        const int rc = launch_long<T>(pl, signals, out, nullptr, S, output, bl, lo, hi, ws, ws_bytes, pl->cap_stream, false);
        cudaGraph_t graph = nullptr;
        const cudaError_t ce = cudaStreamEndCapture(pl->cap_stream, &graph);
        const long long launches = g_launches.load() - l0;
        g_launches.fetch_sub(launches);              // nothing ran yet: replays are counted below
        if (rc || ce != cudaSuccess || !graph) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            return rc ? rc : 1;
        }
        cudaGraphExec_t exec = nullptr;
        const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess || !exec) {
            cudaGetLastError();
            return 1;
        }
        hit->exec = exec;
        hit->launches = launches;
    }
    hit->used = ++pl->graph_clock;
    CUDA_TRY(cudaGraphLaunch(hit->exec, stream));
    g_launches.fetch_add(hit->launches, std::memory_order_relaxed);
    return 0;
}

// ---------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------
extern "C" {

int nwcwt_plan_create(nwcwt_plan** out, const nwcwt_plan_desc* d);

int nwcwt_version(void) { return NWCWT_VERSION; }
int64_t nwcwt_launch_count(void) { return g_launches.load(); }

int nwcwt_fma_peak(int32_t device, int32_t dtype, double* lane_ops_per_s) {
    if (!lane_ops_per_s) return fail(NWCWT_ERR_INVALID, "null argument");
    DeviceGuard guard(device);
    void* buf = nullptr;
    CUDA_TRY(cudaMalloc(&buf, 64));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    const int sms = device_sms(device), grid = sms * 4, nthr = 512, iters = 4096;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0, 0);
        if (dtype == NWCWT_F32) nwcwt_ffma2_peak_kernel<<<grid, nthr>>>((float*)buf, iters, 0.999f, 1e-3f);
        else nwcwt_fma_peak_kernel<double><<<grid, nthr>>>((double*)buf, iters, 0.999, 1e-3);
        cudaEventRecord(e1, 0);
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    const double lanes = (dtype == NWCWT_F32 ? 2.0 : 1.0) * 16.0 * (double)iters * (double)nthr * (double)grid;
    *lane_ops_per_s = lanes / ((double)best * 1e-3);
    return 0;
}

int nwcwt_debug_force_generic(int32_t on) {
    g_force_generic = on != 0;
    return 0;
}

int nwcwt_debug_force_exact(int32_t on) {
    g_exact = on != 0;
    return 0;
}

int nwcwt_profile_enable(int32_t on) {
    g_profile = on != 0;
    return 0;
}

int nwcwt_profile_read(double ms[8], int64_t launches[8]) {
    if (!ms || !launches) return fail(NWCWT_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (ProfEvent& e : g_prof_events) {
        CUDA_TRY(cudaEventSynchronize(e.b));
        float t = 0;
        CUDA_TRY(cudaEventElapsedTime(&t, e.a, e.b));
        g_prof_ms[e.cls] += t;
        g_prof_n[e.cls] += 1;
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
    }
    g_prof_events.clear();
    for (int i = 0; i < N_PROF; ++i) {
        ms[i] = g_prof_ms[i];
        launches[i] = g_prof_n[i];
        g_prof_ms[i] = 0;
        g_prof_n[i] = 0;
    }
    return 0;
}
const char* nwcwt_last_error(void) { return g_err.c_str(); }

// Bluestein plan: smallest 2-3-5-smooth M >= 2 n - 1 that the engine plans, two one-row TABLE sub-plans of that length
static int plan_bluestein(nwcwt_plan* pl, std::string& err) {
    HostPlan& hp = pl->hp;
    const long long need = 2 * hp.N - 1;
    if (need >= (1LL << 30)) { err = "signal length too large for the chirp-z path"; return NWCWT_ERR_UNSUPPORTED; }
    long long M = 0;
    for (long long p5 = 1; p5 <= 4 * need; p5 *= 5)
        for (long long p3 = p5; p3 <= 4 * need; p3 *= 3) {
            long long v = p3;
            while (v < need) v *= 2;
            if (M == 0 || v < M) M = v;
        }
    std::vector<double> zero_table(2 * (size_t)M, 0.0);
    double f1 = 1.0;
    for (int which = 0; which < 2; ++which) {
        nwcwt_plan_desc d;
        memset(&d, 0, sizeof(d));
        d.device = hp.device; d.dtype = NWCWT_F64; d.family = NWCWT_TABLE; d.interpolate = 0;
        d.n = M; d.n_freqs = 1; d.resample = -1; d.sfreq = hp.sfreq; d.freqs = &f1;
        d.table = zero_table.data(); d.table_len = M; d.prune_eps = 0.0;
        nwcwt_plan* sub = nullptr;
        int rc = nwcwt_plan_create(&sub, &d);
        if (rc) { err = std::string("chirp-z sub-plan: ") + nwcwt_last_error(); return rc; }
        sub->is_sub = true;
        (which ? pl->blu_i : pl->blu_f) = sub;
    }
    hp.path = 4;
    pl->blu_m = M;
    const size_t per_row = (size_t)M * 48;
    pl->blu_rows = (int)std::max<size_t>(1, std::min<size_t>(64, ((size_t)256 << 20) / per_row));
    return 0;
}

int nwcwt_plan_create(nwcwt_plan** out, const nwcwt_plan_desc* d) {
    if (!out || !d) return fail(NWCWT_ERR_INVALID, "null argument");
    *out = nullptr;
    if (d->dtype != NWCWT_F32 && d->dtype != NWCWT_F64) return fail(NWCWT_ERR_INVALID, "bad dtype");
    if (d->family < NWCWT_MORSE || d->family > NWCWT_TABLE) return fail(NWCWT_ERR_INVALID, "bad family");
    if (d->n < 2) return fail(NWCWT_ERR_INVALID, "signal length must be >= 2");
    if (d->n >= (1LL << 31)) return fail(NWCWT_ERR_UNSUPPORTED, "signal length >= 2^31");
    if (d->n_freqs < 0 || (d->n_freqs > 0 && !d->freqs)) return fail(NWCWT_ERR_INVALID, "bad freqs");
    if (!(d->sfreq > 0)) return fail(NWCWT_ERR_INVALID, "sfreq must be positive");
    for (int i = 0; i < d->n_freqs; ++i)
        if (d->freqs[i] == 0) return fail(NWCWT_ERR_ZERO_FREQ, "freq == 0 (ZeroDivisionError, base.py:234-235)");
    if (d->family == NWCWT_MORLET && d->n_freqs > 0 && !d->aux)
        return fail(NWCWT_ERR_INVALID, "Morlet needs aux = peak_freq(freq)");
    if (d->family == NWCWT_TABLE && (d->table_len < 1 || (d->n_freqs > 0 && !d->table)))
        return fail(NWCWT_ERR_INVALID, "TABLE family needs table and table_len");
    nwcwt_plan* pl = new nwcwt_plan();
    HostPlan& hp = pl->hp;
    hp.device = d->device;
    hp.dtype = d->dtype;
    hp.family = d->family;
    hp.interpolate = d->interpolate ? 1 : 0;
    hp.N = d->n;
    hp.F = d->n_freqs;
    hp.sfreq = d->sfreq;
    hp.p0 = d->p0;
    hp.p1 = d->p1;
    hp.p2 = d->p2;
    hp.prune_eps = d->prune_eps < 0 ? (d->dtype == NWCWT_F32 ? 1e-12 : 1e-24) : d->prune_eps;
    hp.freqs.assign(d->freqs, d->freqs + d->n_freqs);
    if (d->family == NWCWT_MORLET) hp.aux.assign(d->aux, d->aux + d->n_freqs);
    if (d->family == NWCWT_TABLE) {
        hp.table_len = d->table_len;
        hp.table.assign(d->table, d->table + 2 * (size_t)d->n_freqs * (size_t)d->table_len);
        if (d->table_lens) hp.table_lens.assign(d->table_lens, d->table_lens + d->n_freqs);
    }
    hp.resample = d->resample >= 0 ? 1 : 0;
    hp.resample_tol = d->resample_tol;
    plan_geometry(hp);
    plan_bands(hp);
    std::string err;
    if (!plan_shape(hp, err)) {
        // no radix plan for this length (a prime factor > 64): Bluestein's algorithm on a smooth length >= 2n - 1
        int rc = plan_bluestein(pl, err);
        if (rc) {
            delete pl;
            return fail(rc, err);
        }
    }
    for (const MrGroup& mg : hp.groups) {
        if (hp.short3) break;   // short rows: the groups live inside one kernel (ensure_device_short3)
        nwcwt_plan::Group g;
        g.sub = new nwcwt_plan();
        g.sub->hp = *mg.sub;
        g.sub->is_sub = true;
        pl->groups.push_back(g);
    }
    *out = pl;
    return 0;
}

int nwcwt_plan_destroy(nwcwt_plan* pl) {
    if (!pl) return 0;
    for (nwcwt_plan::Group& g : pl->groups) {
        if (g.sub && g.sub->on_device) {
            cudaSetDevice(pl->hp.device);
            void* ptrs[] = {g.d_coef, g.d_eq, g.d_t0, g.d_fmap, g.d_coefq};
            for (void* q : ptrs) if (q) cudaFree(q);
        }
        nwcwt_plan_destroy(g.sub);
    }
    pl->groups.clear();
    if (!pl->graphs.empty() || pl->cap_stream) {
        cudaSetDevice(pl->hp.device);
        for (nwcwt_plan::GraphEntry& g : pl->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);   // exec is null for a set seen once
        pl->graphs.clear();
        if (pl->cap_stream) cudaStreamDestroy(pl->cap_stream);
        pl->cap_stream = nullptr;
    }
    if (pl->blu_f) nwcwt_plan_destroy(pl->blu_f);
    if (pl->blu_i) nwcwt_plan_destroy(pl->blu_i);
    pl->blu_f = pl->blu_i = nullptr;
    if (pl->on_device || pl->h_stream[0]) {
        cudaSetDevice(pl->hp.device);
        for (void* q : pl->s3_allocs) if (q) cudaFree(q);
        pl->s3_allocs.clear();
        if (pl->d_s3groups) cudaFree(pl->d_s3groups);
        void* ptrs[] = {pl->d_tw, pl->d_twA, pl->d_twB, pl->d_twH, pl->d_twL, pl->d_rec, pl->d_table, pl->d_wtab, pl->d_ditpos, pl->d_chirp, pl->d_twA2, pl->d_twB2,
                        pl->h_in_dev[0], pl->h_in_dev[1], pl->h_out_dev[0], pl->h_out_dev[1], pl->h_ws[0], pl->h_ws[1]};
        for (void* p : ptrs)
            if (p) cudaFree(p);
        for (int i = 0; i < 2; ++i)
            if (pl->h_stream[i]) cudaStreamDestroy(pl->h_stream[i]);
        for (int i = 0; i < nwcwt_plan::MAX_AUX; ++i) {
            if (pl->aux[i]) cudaStreamDestroy(pl->aux[i]);
            if (pl->ev_join[i]) cudaEventDestroy(pl->ev_join[i]);
        }
        if (pl->ev_fork) cudaEventDestroy(pl->ev_fork);
    }
    delete pl;
    return 0;
}

int nwcwt_plan_get_info(const nwcwt_plan* pl, nwcwt_plan_info* info) {
    if (!pl || !info) return fail(NWCWT_ERR_INVALID, "null argument");
    const HostPlan& hp = pl->hp;
    memset(info, 0, sizeof(*info));
    info->n = hp.N;
    info->n_freqs = hp.F;
    info->path = hp.path;
    info->band_bins = hp.band_bins;
    info->n_groups = (int32_t)std::min<size_t>(hp.groups.size(), 32);
    for (int g = 0; g < info->n_groups; ++g) {
        const MrGroup& mg = hp.groups[(size_t)g];
        info->group_D[g] = mg.D;
        info->group_K[g] = mg.K;
        info->group_rows[g] = (int32_t)mg.fidx.size();
        info->group_n1[g] = hp.short3 ? (int32_t)mg.sub->N : mg.sub->N1f;
        info->group_n2[g] = mg.sub->N2f;
        info->group_err[g] = mg.err;
    }
    if (hp.path == 4) {
        info->n1 = (int32_t)pl->blu_m;   // chirp-z: transforms of this length
        info->rows_per_launch = pl->blu_rows;
        return 0;
    }
    if (hp.path == 0) {
        info->batch = 1 << hp.tsh;
        info->n_stages[0] = hp.st.nst;
        for (int i = 0; i < hp.st.nst; ++i) info->radices[0][i] = hp.st.radix[i];
        info->smem_bytes = (int64_t)hp.smem_short;
        info->threads[0] = hp.nthr_short;
        if (hp.short2) {
            info->path = 3;
            info->batch = 1 << hp.tpshS;
            info->n_stages[0] = hp.stS.nst;
            for (int i = 0; i < hp.stS.nst; ++i) info->radices[0][i] = hp.stS.radix[i];
            info->smem_bytes = (int64_t)hp.smem_S2;
            info->threads[0] = hp.nthrS2;
        }
    } else if (hp.fast) {
        info->n1 = hp.N1f;
        info->n2 = hp.N2f;
        info->batch = 2 << hp.tpshB;
        info->n_stages[0] = hp.stA2.nst;
        info->n_stages[1] = hp.stB2.nst;
        for (int i = 0; i < hp.stA2.nst; ++i) info->radices[0][i] = hp.stA2.radix[i];
        for (int i = 0; i < hp.stB2.nst; ++i) info->radices[1][i] = hp.stB2.radix[i];
        info->smem_bytes = (int64_t)std::max(hp.smem_A2, hp.smem_B2);
        info->path = 2;
        info->threads[0] = hp.nthrA2;
        info->threads[1] = hp.nthrB2;
        info->rows_per_launch = hp.ring2;
    } else {
        info->n1 = hp.N1;
        info->n2 = hp.N2;
        info->batch = 1 << hp.tshA;
        info->n_stages[0] = hp.stA.nst;
        info->n_stages[1] = hp.stB.nst;
        for (int i = 0; i < hp.stA.nst; ++i) info->radices[0][i] = hp.stA.radix[i];
        for (int i = 0; i < hp.stB.nst; ++i) info->radices[1][i] = hp.stB.radix[i];
        info->smem_bytes = (int64_t)std::max(hp.smem_A, hp.smem_B);
        info->threads[0] = info->threads[1] = hp.nthr_long;
        info->rows_per_launch = hp.ring;
    }
    return 0;
}

int nwcwt_plan_get_bands(const nwcwt_plan* pl, int32_t* lo, int32_t* hi) {
    if (!pl || !lo || !hi) return fail(NWCWT_ERR_INVALID, "null argument");
    for (int i = 0; i < pl->hp.F; ++i) {
        lo[i] = pl->hp.rec[i].lo;
        hi[i] = pl->hp.rec[i].hi;
    }
    return 0;
}

int nwcwt_workspace_bytes(const nwcwt_plan* pl, int64_t n_signals, size_t* bytes) {
    if (!pl || !bytes) return fail(NWCWT_ERR_INVALID, "null argument");
    (void)n_signals;
    const HostPlan& hp = pl->hp;
    if (hp.path == 4) {
        *bytes = bluestein_ws_bytes(pl);
        return 0;
    }
    if (hp.path == 0) {
        *bytes = 0;
        return 0;
    }
    const size_t cs = cx_size(hp.dtype);
    *bytes = align_up((size_t)hp.ring * hp.N * cs, 256) + tm_ring_bytes(pl) + y_ring_bytes(pl);
    return 0;
}

int nwcwt_spectrum_bank(nwcwt_plan* pl, void* bank, void* stream) {
    if (!pl || !bank) return fail(NWCWT_ERR_INVALID, "null argument");
    DeviceGuard guard(pl->hp.device);
    int rc = ensure_device(pl);
    if (rc) return rc;
    const HostPlan& hp = pl->hp;
    if (hp.F <= 0) return 0;
    dim3 grid((unsigned)((hp.N + 255) / 256), (unsigned)hp.F);
    if (hp.path == 4 && hp.dtype == NWCWT_F32) {   // chirp-z plan: fp64 tables
        SpecParams<double> sp = make_spec<double>(pl);
        sp.norm = 1.0;
        nwcwt_bank_kernel<double, float><<<grid, 256, 0, (cudaStream_t)stream>>>(sp, (cx<float>*)bank, hp.F, hp.N);
    } else if (hp.dtype == NWCWT_F32) {
        SpecParams<float> sp = make_spec<float>(pl);
        sp.norm = 1.0f;
        nwcwt_bank_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(sp, (cx<float>*)bank, hp.F, hp.N);
    } else {
        SpecParams<double> sp = make_spec<double>(pl);
        sp.norm = 1.0;
        nwcwt_bank_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>(sp, (cx<double>*)bank, hp.F, hp.N);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int nwcwt_reduce_epochs(nwcwt_plan* pl, const void* in, void* out, int64_t E, int64_t count, int32_t kind,
                        void* stream) {
    if (!pl || !in || !out) return fail(NWCWT_ERR_INVALID, "null argument");
    if (E <= 0 || count <= 0 || kind < 0 || kind > 1) return fail(NWCWT_ERR_INVALID, "bad reduction arguments");
    DeviceGuard guard(pl->hp.device);
    const unsigned grid = (unsigned)((count + 255) / 256);
    LaunchScope ls(5, (cudaStream_t)stream);
    if (pl->hp.dtype == NWCWT_F32)
        nwcwt_reduce_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(in, (float*)out, E, count, kind);
    else
        nwcwt_reduce_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>(in, (double*)out, E, count, kind);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int nwcwt_transform_epochs(nwcwt_plan* pl, const void* signals, void* out, int64_t n_channels, int64_t n_epochs,
                           int32_t kind, void* ws, size_t ws_bytes, void* stream) {
    if (!pl || !signals || !out) return fail(NWCWT_ERR_INVALID, "null argument");
    if (n_channels <= 0 || n_epochs <= 0 || kind < 0 || kind > 1) return fail(NWCWT_ERR_INVALID, "bad epoch arguments");
    if (pl->hp.F <= 0) return fail(NWCWT_ERR_INVALID, "plan has no frequencies");
    if (pl->hp.path != 0 || !pl->hp.short2)
        return fail(NWCWT_ERR_UNSUPPORTED, "fused epoch reductions need rows that fit one CTA (short packed path); "
                                           "use nwcwt_transform + nwcwt_reduce_epochs for this length");
    const size_t need = kind == 1 ? (size_t)n_channels * (size_t)pl->hp.F * (size_t)pl->hp.N * cx_size(pl->hp.dtype) : 0;
    if (ws_bytes < need || (need && !ws)) return fail(NWCWT_ERR_WORKSPACE, "workspace too small");
    DeviceGuard guard(pl->hp.device);
    int rc = ensure_device(pl);
    if (rc) return rc;
    if (pl->hp.dtype == NWCWT_F32)
        return launch_short2_epochs_t<float>(pl, signals, out, n_channels, n_epochs, kind, ws, (cudaStream_t)stream);
    return launch_short2_epochs_t<double>(pl, signals, out, n_channels, n_epochs, kind, ws, (cudaStream_t)stream);
}

int nwcwt_baseline_rows(int32_t device, int32_t dtype, void* rows, int64_t n_rows, int64_t n, int32_t bl,
                        int64_t blo, int64_t bhi, void* stream) {
    if (!rows) return fail(NWCWT_ERR_INVALID, "null argument");
    if (bl <= NWCWT_BL_NONE || bl > NWCWT_BL_ZLOG) return fail(NWCWT_ERR_INVALID, "bad baseline mode");
    if (n_rows <= 0 || n <= 0) return 0;
    if (n_rows > 2147483647LL) return fail(NWCWT_ERR_INVALID, "too many rows");
    if (blo < 0 || bhi < 0) return fail(NWCWT_ERR_INVALID, "negative baseline index");
    if (bhi > n) bhi = n;
    if (blo > bhi) blo = bhi;
    DeviceGuard guard(device);
    LaunchScope ls(5, (cudaStream_t)stream);
    if (dtype == NWCWT_F32)
        nwcwt_baseline_rows_kernel<float><<<(unsigned)n_rows, 512, 0, (cudaStream_t)stream>>>((float*)rows, n, bl, (int)blo, (int)bhi);
    else
        nwcwt_baseline_rows_kernel<double><<<(unsigned)n_rows, 512, 0, (cudaStream_t)stream>>>((double*)rows, n, bl, (int)blo, (int)bhi);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int nwcwt_forward(nwcwt_plan* pl, const void* signals, void* spectra, int64_t S, void* ws, size_t ws_bytes,
                  void* stream) {
    if (!pl || !signals || !spectra) return fail(NWCWT_ERR_INVALID, "null argument");
    if (S <= 0) return 0;
    DeviceGuard guard(pl->hp.device);
    int rc = ensure_device(pl);
    if (rc) return rc;
    if (pl->hp.dtype == NWCWT_F32)
        return run_transform<float>(pl, signals, nullptr, spectra, S, 0, 0, 0, 0, ws, ws_bytes, (cudaStream_t)stream, true);
    return run_transform<double>(pl, signals, nullptr, spectra, S, 0, 0, 0, 0, ws, ws_bytes, (cudaStream_t)stream, true);
}

int nwcwt_transform(nwcwt_plan* pl, const void* signals, void* out, int64_t S, int32_t output, int32_t bl,
                    int64_t blo, int64_t bhi, void* ws, size_t ws_bytes, void* stream) {
    long long lo = blo, hi = bhi;
    int rc = check_args(pl, output, bl, lo, hi);
    if (rc) return rc;
    if (!signals || !out) return fail(NWCWT_ERR_INVALID, "null buffer");
    if (S <= 0) return 0;
    DeviceGuard guard(pl->hp.device);
    if ((rc = ensure_device(pl))) return rc;
    if (pl->hp.dtype == NWCWT_F32) {
        if ((rc = transform_graph<float>(pl, signals, out, S, output, bl, lo, hi, ws, ws_bytes, (cudaStream_t)stream)) <= 0) return rc;
        return run_transform<float>(pl, signals, out, nullptr, S, output, bl, lo, hi, ws, ws_bytes, (cudaStream_t)stream, false);
    }
    if ((rc = transform_graph<double>(pl, signals, out, S, output, bl, lo, hi, ws, ws_bytes, (cudaStream_t)stream)) <= 0) return rc;
    return run_transform<double>(pl, signals, out, nullptr, S, output, bl, lo, hi, ws, ws_bytes, (cudaStream_t)stream, false);
}

int nwcwt_transform_host(nwcwt_plan* pl, const void* signals, void* out, int64_t S, int32_t output, int32_t bl,
                         int64_t blo, int64_t bhi) {
    long long lo = blo, hi = bhi;
    int rc = check_args(pl, output, bl, lo, hi);
    if (rc) return rc;
    if (!signals || !out) return fail(NWCWT_ERR_INVALID, "null buffer");
    if (S <= 0) return 0;
    DeviceGuard guard(pl->hp.device);
    if ((rc = ensure_device(pl))) return rc;
    const HostPlan& hp = pl->hp;
    const size_t rs = hp.dtype == NWCWT_F32 ? 4 : 8;
    const size_t esz = output == NWCWT_OUT_CWT ? 2 * rs : rs;
    const size_t in_row = (size_t)hp.N * rs, out_sig = (size_t)hp.F * (size_t)hp.N * esz;
    // chunk: about 512 MB of output per slot, at least one signal
    long long chunk = (long long)std::max<size_t>(1, ((size_t)512 << 20) / std::max<size_t>(out_sig, 1));
    chunk = std::min<long long>(chunk, (S + 1) / 2 > 0 ? (S + 1) / 2 : 1);
    if (chunk < 1) chunk = 1;
    size_t wsb = 0;
    nwcwt_workspace_bytes(pl, chunk, &wsb);
    const size_t need_in = in_row * chunk, need_out = out_sig * chunk;
    if (pl->h_in_bytes < need_in || pl->h_out_bytes < need_out || pl->h_ws_bytes < wsb || !pl->h_stream[0]) {
        for (int i = 0; i < 2; ++i) {
            if (pl->h_in_dev[i]) cudaFree(pl->h_in_dev[i]);
            if (pl->h_out_dev[i]) cudaFree(pl->h_out_dev[i]);
            if (pl->h_ws[i]) cudaFree(pl->h_ws[i]);
            pl->h_in_dev[i] = pl->h_out_dev[i] = pl->h_ws[i] = nullptr;
            CUDA_TRY(cudaMalloc(&pl->h_in_dev[i], need_in));
            CUDA_TRY(cudaMalloc(&pl->h_out_dev[i], need_out));
            if (wsb) CUDA_TRY(cudaMalloc(&pl->h_ws[i], wsb));
            if (!pl->h_stream[i]) CUDA_TRY(cudaStreamCreateWithFlags(&pl->h_stream[i], cudaStreamNonBlocking));
        }
        pl->h_in_bytes = need_in;
        pl->h_out_bytes = need_out;
        pl->h_ws_bytes = wsb;
    }
    int slot = 0;
    for (long long s0 = 0; s0 < S; s0 += chunk, slot ^= 1) {
        const long long cs = std::min<long long>(chunk, S - s0);
        cudaStream_t st = pl->h_stream[slot];
        CUDA_TRY(cudaMemcpyAsync(pl->h_in_dev[slot], (const char*)signals + (size_t)s0 * in_row, in_row * cs,
                                 cudaMemcpyHostToDevice, st));
        if (hp.dtype == NWCWT_F32)
            rc = run_transform<float>(pl, pl->h_in_dev[slot], pl->h_out_dev[slot], nullptr, cs, output, bl, lo, hi,
                                      pl->h_ws[slot], wsb, st, false);
        else
            rc = run_transform<double>(pl, pl->h_in_dev[slot], pl->h_out_dev[slot], nullptr, cs, output, bl, lo, hi,
                                       pl->h_ws[slot], wsb, st, false);
        if (rc) return rc;
        CUDA_TRY(cudaMemcpyAsync((char*)out + (size_t)s0 * out_sig, pl->h_out_dev[slot], out_sig * cs,
                                 cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(cudaStreamSynchronize(pl->h_stream[0]));
    CUDA_TRY(cudaStreamSynchronize(pl->h_stream[1]));
    return 0;
}

}  // extern "C"

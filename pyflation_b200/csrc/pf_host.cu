// Host-side half of a host-bound solve (see PF_LAYOUT_PERT in include/pyflation_b200.h):
// the PCIe link carries only the perturbation rows; the background rows, identical for every
// mode, and the NaN prefix are written straight into the caller's yresult by host threads
// with non-temporal stores while the device->host copies are in flight.
#include <emmintrin.h>
#include <immintrin.h>
#include <math.h>

#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "pf_common.cuh"

namespace pf {

// One time step at a time: the start-row test is done once per mode and feeds all 2nf+1
// background rows (2nf+1 independent non-temporal store streams).
template <int NB>
static void fill_rows_nb(double* yout, int nvar, int64_t nk, int64_t ta, int64_t tb, const double* bg,
                         int64_t bg_pitch, int64_t bg_off, const int64_t* tsix, const double* ystart,
                         bool pert_nan) {
    const __m128d nan2 = _mm_set1_pd(NAN);
    const int nan_rows = pert_nan ? nvar : NB;          // rows NaN-filled before a mode's start
    for (int64_t t = ta; t < tb; ++t) {
        __m128d val[NB];
        double* row[NB];
        for (int v = 0; v < NB; ++v) {
            val[v] = _mm_set_pd(0.0, bg[t * bg_pitch + bg_off + v]);              // (re, im = 0)
            row[v] = yout + 2 * ((t * nvar + v) * nk);
        }
        for (int64_t k = 0; k < nk; ++k) {
            const int64_t s = tsix[k];
            if (t > s) {
                for (int v = 0; v < NB; ++v) _mm_stream_pd(row[v] + 2 * k, val[v]);
            } else if (t < s) {
                for (int v = 0; v < nan_rows; ++v) _mm_stream_pd(row[0] + 2 * ((int64_t)v * nk + k), nan2);
            } else {                                                             // rk4.py:106-107
                for (int v = 0; v < NB; ++v)
                    _mm_stream_pd(row[v] + 2 * k, _mm_loadu_pd(ystart + 2 * ((int64_t)v * nk + k)));
            }
        }
    }
    _mm_sfence();
}

// ---- sorted start rows (the normal case: k ascending) ------------------------------------------
// Row t of variable v is then three runs of columns: [0, a) started before t -> (bg[t][v], 0);
// [a, b) starting at t -> ystart[v][k]; [b, nk) not started -> NaN+NaNj, with a = #(tsix < t),
// b = #(tsix <= t) advancing monotonically with t.  Runs are written with the widest
// non-temporal store the CPU has (64 / 32 / 16 bytes, chosen at run time) and tsix is not
// re-read per element (8 of every 56-88 bytes of memory traffic in the per-element loop above).
static void stream_const_sse(double* dst, int64_t n, double re, double im) {
    const __m128d val = _mm_set_pd(im, re);
    for (int64_t i = 0; i < n; ++i) _mm_stream_pd(dst + 2 * i, val);
}
__attribute__((target("avx2"))) static void stream_const_avx2(double* dst, int64_t n, double re, double im) {
    int64_t i = 0;
    const __m128d v1 = _mm_set_pd(im, re);
    if (n > 0 && ((uintptr_t)dst & 31)) { _mm_stream_pd(dst, v1); i = 1; }
    const __m256d v2 = _mm256_set_pd(im, re, im, re);
    for (; i + 2 <= n; i += 2) _mm256_stream_pd(dst + 2 * i, v2);
    if (i < n) _mm_stream_pd(dst + 2 * i, v1);
}
__attribute__((target("avx512f"))) static void stream_const_avx512(double* dst, int64_t n, double re, double im) {
    int64_t i = 0;
    const __m128d v1 = _mm_set_pd(im, re);
    while (i < n && ((uintptr_t)(dst + 2 * i) & 63)) { _mm_stream_pd(dst + 2 * i, v1); ++i; }
    const __m512d v4 = _mm512_set_pd(im, re, im, re, im, re, im, re);
    for (; i + 4 <= n; i += 4) _mm512_stream_pd(dst + 2 * i, v4);
    for (; i < n; ++i) _mm_stream_pd(dst + 2 * i, v1);
}
typedef void (*stream_const_fn)(double*, int64_t, double, double);
static stream_const_fn pick_stream_const() {
    if (const char* e = getenv("PF_HOST_FILL_ISA")) {
        if (!strcmp(e, "sse2")) return stream_const_sse;
        if (!strcmp(e, "avx2")) return stream_const_avx2;
    }
    __builtin_cpu_init();
    if (__builtin_cpu_supports("avx512f")) return stream_const_avx512;
    if (__builtin_cpu_supports("avx2")) return stream_const_avx2;
    return stream_const_sse;
}

static void fill_rows_sorted(double* yout, int nb, int nvar, int64_t nk, int64_t ta, int64_t tb,
                             const double* bg, int64_t bg_pitch, int64_t bg_off, const int64_t* tsix,
                             const double* ystart, bool pert_nan) {
    static const stream_const_fn stream_const = pick_stream_const();
    const int nan_rows = pert_nan ? nvar : nb;
    int64_t a = std::lower_bound(tsix, tsix + nk, ta) - tsix;      // #(tsix < ta)
    int64_t b = a;
    for (int64_t t = ta; t < tb; ++t) {
        while (a < nk && tsix[a] < t) ++a;
        if (b < a) b = a;
        while (b < nk && tsix[b] <= t) ++b;
        for (int v = 0; v < nan_rows; ++v) {
            double* row = yout + 2 * ((t * nvar + v) * nk);
            if (v < nb) {
                stream_const(row, a, bg[t * bg_pitch + bg_off + v], 0.0);
                for (int64_t k = a; k < b; ++k)                     // the start row, rk4.py:106-107
                    _mm_stream_pd(row + 2 * k, _mm_loadu_pd(ystart + 2 * ((int64_t)v * nk + k)));
            }
            stream_const(row + 2 * b, nk - b, NAN, NAN);
        }
    }
    _mm_sfence();
}

// any number of background rows (nfields > 8): one row at a time
static void fill_rows_any(double* yout, int nb, int nvar, int64_t nk, int64_t ta, int64_t tb, const double* bg,
                          int64_t bg_pitch, int64_t bg_off, const int64_t* tsix, const double* ystart,
                          bool pert_nan) {
    const __m128d nan2 = _mm_set1_pd(NAN);
    const int nan_rows = pert_nan ? nvar : nb;
    for (int64_t t = ta; t < tb; ++t) {
        for (int v = 0; v < nan_rows; ++v) {
            double* row = yout + 2 * ((t * nvar + v) * nk);
            const __m128d val = _mm_set_pd(0.0, v < nb ? bg[t * bg_pitch + bg_off + v] : 0.0);
            for (int64_t k = 0; k < nk; ++k) {
                const int64_t s = tsix[k];
                if (t < s) _mm_stream_pd(row + 2 * k, nan2);
                else if (v < nb) _mm_stream_pd(row + 2 * k, t > s ? val : _mm_loadu_pd(ystart + 2 * ((int64_t)v * nk + k)));
            }
        }
    }
    _mm_sfence();
}

static void fill_rows(double* yout, int nb, int nvar, int64_t nk, int64_t ta, int64_t tb,
                      const double* bg, int64_t bg_pitch, int64_t bg_off, const int64_t* tsix,
                      const double* ystart, bool pert_nan, bool sorted) {
    if (sorted) {
        fill_rows_sorted(yout, nb, nvar, nk, ta, tb, bg, bg_pitch, bg_off, tsix, ystart, pert_nan);
        return;
    }
    switch (nb) {
#define PF_CASE(N) case N: fill_rows_nb<N>(yout, nvar, nk, ta, tb, bg, bg_pitch, bg_off, tsix, ystart, pert_nan); break;
        PF_CASE(3) PF_CASE(5) PF_CASE(7) PF_CASE(9) PF_CASE(11) PF_CASE(13) PF_CASE(15) PF_CASE(17)
#undef PF_CASE
        default: fill_rows_any(yout, nb, nvar, nk, ta, tb, bg, bg_pitch, bg_off, tsix, ystart, pert_nan); break;
    }
}

}  // namespace pf

using namespace pf;

extern "C" int pf_copy_rows_d2h(void* dst_host, size_t dst_pitch, const void* src_dev, size_t src_pitch,
                                size_t width_bytes, size_t height, void* stream) {
    if (height == 0 || width_bytes == 0) return PF_OK;
    if (!dst_host || !src_dev || dst_pitch < width_bytes || src_pitch < width_bytes) {
        set_error("pf_copy_rows_d2h: bad arguments");
        return PF_E_ARG;
    }
    PF_CUDA_CHECK(cudaMemcpy2DAsync(dst_host, dst_pitch, src_dev, src_pitch, width_bytes, height,
                                    cudaMemcpyDeviceToHost, (cudaStream_t)stream),
                  "pf_copy_rows_d2h");
    return PF_OK;
}

extern "C" int pf_copy_window_d2h(void* yresult_host, int64_t nvar, int64_t nk, int64_t row0, int64_t t0,
                                  const void* window_dev, int64_t wrows, int64_t kcols, int64_t nt,
                                  void* stream) {
    if (nt == 0 || kcols == 0 || wrows == 0) return PF_OK;
    if (!yresult_host || !window_dev || nvar < 1 || nk < 1 || row0 < 0 || t0 < 0 || wrows < 0 ||
        row0 + wrows > nvar || kcols < 0 || kcols > nk || nt < 0) {
        set_error("pf_copy_window_d2h: bad arguments");
        return PF_E_ARG;
    }
    // one 3-D copy: x = started columns (bytes), y = the window's variable rows, z = time steps
    cudaMemcpy3DParms p = {};
    p.srcPtr = make_cudaPitchedPtr(const_cast<void*>(window_dev), (size_t)nk * 16, (size_t)nk * 16, (size_t)wrows);
    p.dstPtr = make_cudaPitchedPtr(yresult_host, (size_t)nk * 16, (size_t)nk * 16, (size_t)nvar);
    p.dstPos = make_cudaPos(0, (size_t)row0, (size_t)t0);
    p.extent = make_cudaExtent((size_t)kcols * 16, (size_t)wrows, (size_t)nt);
    p.kind = cudaMemcpyDeviceToHost;
    PF_CUDA_CHECK(cudaMemcpy3DAsync(&p, (cudaStream_t)stream), "pf_copy_window_d2h");
    return PF_OK;
}

extern "C" int pf_host_fill_background(double* yout_host, int nfields, int64_t nk, int64_t t0, int64_t t1,
                                       const double* bg_host, int64_t bg_pitch, int64_t bg_off,
                                       const int64_t* tsix_host, const double* ystart_host, int pert_nan,
                                       int nthreads) {
    if (nfields < 1 || nfields > PF_MAX_NFIELDS_RT) { set_error("pf_host_fill_background: nfields=%d", nfields); return PF_E_NFIELDS; }
    if (nk < 0 || t0 < 0 || t1 < t0) { set_error("pf_host_fill_background: bad sizes"); return PF_E_ARG; }
    if (nk == 0 || t0 == t1) return PF_OK;
    if (!yout_host || !bg_host || !tsix_host || !ystart_host) { set_error("pf_host_fill_background: null buffer"); return PF_E_ARG; }
    if (((uintptr_t)yout_host & 15) != 0) { set_error("pf_host_fill_background: yresult must be 16-byte aligned"); return PF_E_ARG; }
    const int nb = nbg_of(nfields), nvar = nvar_of(nfields);
    int nt = nthreads > 0 ? nthreads : (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    const int64_t rows = t1 - t0;
    if (nt > rows) nt = (int)rows;
    static const bool no_runs = getenv("PF_HOST_FILL_PER_ELEMENT") != nullptr;
    const bool sorted = !no_runs && std::is_sorted(tsix_host, tsix_host + nk);
    std::vector<std::thread> pool;
    for (int w = 0; w < nt; ++w) {
        const int64_t ta = t0 + rows * w / nt, tb = t0 + rows * (w + 1) / nt;
        pool.emplace_back(fill_rows, yout_host, nb, nvar, nk, ta, tb, bg_host, bg_pitch, bg_off, tsix_host,
                          ystart_host, pert_nan != 0, sorted);
    }
    for (auto& th : pool) th.join();
    return PF_OK;
}

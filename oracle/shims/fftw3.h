/* TEST INFRASTRUCTURE — build shim, not product code.
 *
 * Minimal stand-in for the subset of the FFTW3 API that the reference's
 * fft.cpp uses (fftw_malloc, fftw_plan_dft_1d, fftw_execute), so that the
 * unmodified reference sources under /root/reference/src can be compiled in a
 * container that has no FFTW (reference dependency: "fftw3 >= 3.0", unpinned,
 * cmake/Modules/FindFFTW3.cmake:5).
 *
 * Arithmetic: plain fp64 DFT, unnormalised, sign as requested — the published
 * definition FFTW implements.  N = 64 uses three radix-4 passes with
 * precomputed twiddles (so the CPU baseline is not handicapped by a slow
 * FFT); any other power of two falls back to radix-2.  Results agree with
 * FFTW to ~1e-15 relative, not bit-for-bit (FFTW's codelet operation order is
 * not reproducible); nothing in the reference pins FFT output bitwise.
 */
#ifndef B200RX_SHIM_FFTW3_H
#define B200RX_SHIM_FFTW3_H

#include <stdlib.h>
#include <string.h>
#include <math.h>

typedef double fftw_complex[2];

#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)

struct shim_fftw_plan_s {
    int n;
    int sign;
    fftw_complex *in;
    fftw_complex *out;
    double *tw;  /* n entries of (cos, sign*sin)(2*pi*k/n) */
    int *rev;    /* digit reversal (base 4 for n=64, base 2 otherwise) */
};
typedef struct shim_fftw_plan_s *fftw_plan;

static inline void *fftw_malloc(size_t n)
{
    void *p = NULL;
    if (posix_memalign(&p, 64, n ? n : 64)) return NULL;
    return p;
}
static inline void fftw_free(void *p) { free(p); }

static inline fftw_plan fftw_plan_dft_1d(int n, fftw_complex *in, fftw_complex *out, int sign, unsigned flags)
{
    (void)flags;
    fftw_plan p = (fftw_plan)malloc(sizeof(*p));
    p->n = n; p->sign = sign; p->in = in; p->out = out;
    p->tw = (double *)malloc(sizeof(double) * 2 * n);
    p->rev = (int *)malloc(sizeof(int) * n);
    for (int k = 0; k < n; k++) {
        double a = 2.0 * M_PI * (double)k / (double)n;
        p->tw[2 * k] = cos(a);
        p->tw[2 * k + 1] = (sign < 0 ? -1.0 : 1.0) * sin(a);
    }
    if (n == 64) {
        for (int k = 0; k < 64; k++)
            p->rev[k] = ((k & 3) << 4) | (k & 12) | ((k >> 4) & 3);
    } else {
        int bits = 0;
        while ((1 << bits) < n) bits++;
        for (int k = 0; k < n; k++) {
            int r = 0;
            for (int b = 0; b < bits; b++) if (k & (1 << b)) r |= 1 << (bits - 1 - b);
            p->rev[k] = r;
        }
    }
    return p;
}

static inline void fftw_destroy_plan(fftw_plan p)
{
    if (!p) return;
    free(p->tw); free(p->rev); free(p);
}

static inline void shim_fft64_radix4(const struct shim_fftw_plan_s *p)
{
    double xr[64], xi[64];
    for (int k = 0; k < 64; k++) { xr[k] = p->in[p->rev[k]][0]; xi[k] = p->in[p->rev[k]][1]; }
    const double sg = (p->sign < 0) ? -1.0 : 1.0; /* multiply by sign*i */
    for (int len = 4; len <= 64; len <<= 2) {
        const int q = len >> 2;
        const int tstep = 64 / len;
        for (int base = 0; base < 64; base += len) {
            for (int j = 0; j < q; j++) {
                const int i0 = base + j, i1 = i0 + q, i2 = i1 + q, i3 = i2 + q;
                const double w1r = p->tw[2 * (j * tstep)],     w1i = p->tw[2 * (j * tstep) + 1];
                const double w2r = p->tw[2 * (2 * j * tstep)], w2i = p->tw[2 * (2 * j * tstep) + 1];
                const double w3r = p->tw[2 * (3 * j * tstep)], w3i = p->tw[2 * (3 * j * tstep) + 1];
                const double ar = xr[i0], ai = xi[i0];
                const double br = xr[i1] * w1r - xi[i1] * w1i, bi = xr[i1] * w1i + xi[i1] * w1r;
                const double cr = xr[i2] * w2r - xi[i2] * w2i, ci = xr[i2] * w2i + xi[i2] * w2r;
                const double dr = xr[i3] * w3r - xi[i3] * w3i, di = xr[i3] * w3i + xi[i3] * w3r;
                const double t0r = ar + cr, t0i = ai + ci, t1r = ar - cr, t1i = ai - ci;
                const double t2r = br + dr, t2i = bi + di;
                /* (b - d) * (sign * i) */
                const double t3r = -sg * (bi - di), t3i = sg * (br - dr);
                xr[i0] = t0r + t2r; xi[i0] = t0i + t2i;
                xr[i1] = t1r + t3r; xi[i1] = t1i + t3i;
                xr[i2] = t0r - t2r; xi[i2] = t0i - t2i;
                xr[i3] = t1r - t3r; xi[i3] = t1i - t3i;
            }
        }
    }
    for (int k = 0; k < 64; k++) { p->out[k][0] = xr[k]; p->out[k][1] = xi[k]; }
}

static inline void shim_fft_radix2(const struct shim_fftw_plan_s *p)
{
    const int n = p->n;
    double *xr = (double *)malloc(sizeof(double) * 2 * n), *xi = xr + n;
    for (int k = 0; k < n; k++) { xr[k] = p->in[p->rev[k]][0]; xi[k] = p->in[p->rev[k]][1]; }
    for (int len = 2; len <= n; len <<= 1) {
        const int h = len >> 1, tstep = n / len;
        for (int base = 0; base < n; base += len)
            for (int j = 0; j < h; j++) {
                const double wr = p->tw[2 * j * tstep], wi = p->tw[2 * j * tstep + 1];
                const int a = base + j, b = a + h;
                const double tr = xr[b] * wr - xi[b] * wi, ti = xr[b] * wi + xi[b] * wr;
                xr[b] = xr[a] - tr; xi[b] = xi[a] - ti;
                xr[a] += tr; xi[a] += ti;
            }
    }
    for (int k = 0; k < n; k++) { p->out[k][0] = xr[k]; p->out[k][1] = xi[k]; }
    free(xr);
}

static inline void fftw_execute(const fftw_plan p)
{
    if (p->n == 64) shim_fft64_radix4(p);
    else shim_fft_radix2(p);
}

#endif /* B200RX_SHIM_FFTW3_H */

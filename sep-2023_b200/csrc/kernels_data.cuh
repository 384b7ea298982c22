// kernels_data.cuh -- data-side operators between the forward and the reverse-time loop of a shot (sm_100a):
// trace windows + weights, band-pass, normalised cross-correlation misfit, least-squares source-signature update.
//
// para_file.json carries their switches (`if_win`, `filter`, `if_cross_misfit`, `if_src_update`, Src/Parameter.cpp:139-176);
// the operators are live code in Src/utilities.cu (cuda_window :790-842, bp_filter1d :1115-1168, cuda_find_normfact :1011-1040,
// cuda_normal_misfit :1058-1087, cuda_normal_adjoint_source :1090-1113, source_update :1170-1276, cuda_spectrum_update :904-975,
// source_update_adj :1280-1326, amp_ratio_comp :1328-1356), their call sites in libCUFD.cu:353-457 are written out but
// commented.  Here they act on the DAS component, in the order of those call sites (sepfwi.cu: run_conditioning).
//
// The reference pads every trace to 2 nt samples and goes through cuFFT (R2C, per-bin gain, C2R).  nt is arbitrary
// (2 x 4001 has no small radix), the band-pass touches a few hundred bins and the source update runs once per shot and
// evaluation, so the transforms here are direct sums over a host-built twiddle table of length 2 nt, indexed by
// (k n) mod 2 nt kept incrementally -- exact phases, no cuFFT, one thread per output bin / output sample.
#pragma once
#include "common.cuh"

namespace sepfwi {

constexpr int DF_NT = 128;        // threads per block of the transforms
constexpr int DF_CHUNK = 1024;    // samples / bins staged in shared memory per pass

// cuda_window with per-trace windows and weights (utilities.cu:790-842) or, with win_start == nullptr, its window-less overload
// (utilities.cu:844-884: end tapers of ratio x record length, no weights); grid (ceil(nt / 256), nrec)
__global__ void __launch_bounds__(256) k_win_traces(float *data, const int nrec, const int nt, const float dt, const float *win_start,
                                                     const float *win_end, const float *weights, const float src_weight, const float ratio)
{
    const int it = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (it >= nt || r >= nrec) return;
    const float PI = 3.14159265358979323846f;
    const float t = it * dt, t_max = nt * dt;
    float t0 = win_start ? win_start[r] : 0.0f, t3 = win_start ? win_end[r] : t_max;
    t0 = fminf(fmaxf(t0, 0.0f), t_max); t3 = fminf(fmaxf(t3, 0.0f), t_max);
    const float offset = (t3 - t0) * ratio;
    if (offset <= 0.0f) return;                       // "Window error 1": the trace is left as it is (:815-819)
    if (!win_start && 2.0f * offset >= t3 - t0) return;      // "Window error 2" (:859-862)
    const float t1 = t0 + offset, t2 = t3 - offset;
    float amp;
    if (t >= t0 && t < t1) amp = sinf(PI / 2.0f * (t - t0) / (t1 - t0));
    else if (t >= t1 && t < t2) amp = 1.0f;
    else if (t >= t2 && t < t3) amp = cosf(PI / 2.0f * (t - t2) / (t3 - t2));
    else amp = 0.0f;
    data[(size_t)r * nt + it] *= win_start ? amp * amp * weights[r] * src_weight : amp * amp;
}

// Forward transform of zero-padded real traces at the contiguous bins k0 .. k0 + nb - 1 of the length-n2 DFT:
//   X[tr][b] = sum_{n < nt} w(n) x[tr][n] exp(-2 pi i (k0 + b) n / n2)
// w = the 0.01 end taper of cuda_window on the PADDED length (source_update :1197-1200) when taper != 0 -- only its rising
// half can touch the nt <= n2 / 2 data samples.  grid (ceil(nb / DF_NT), ntr); tw[j] = (cos, sin)(2 pi j / n2).
__global__ void __launch_bounds__(DF_NT) k_dft_fwd(const float *x, const int nt, const int n2, const int taper, const int k0, const int nb,
                                                    const float2 *tw, float2 *X)
{
    __shared__ float sx[DF_CHUNK];
    const int tr = blockIdx.y, b = blockIdx.x * DF_NT + threadIdx.x, k = k0 + b;
    const float *xr = x + (size_t)tr * nt;
    const float toff = 0.01f * (float)n2;             // taper length in samples: offset / dt = n2 * 0.01
    float re = 0.f, im = 0.f;
    int idx = 0;                                      // (k n) mod n2
    for (int n0 = 0; n0 < nt; n0 += DF_CHUNK) {
        const int m = min(DF_CHUNK, nt - n0);
        __syncthreads();
        for (int j = threadIdx.x; j < m; j += DF_NT) {
            float v = xr[n0 + j];
            if (taper) {
                const float n = (float)(n0 + j);
                if (n < toff) { const float a = sinf(1.57079632679489662f * n / toff); v *= a * a; }
            }
            sx[j] = v;
        }
        __syncthreads();
        if (b < nb) {
#pragma unroll 4
            for (int j = 0; j < m; j++) {
                const float2 w = __ldg(tw + idx);
                re = fmaf(sx[j], w.x, re); im = fmaf(-sx[j], w.y, im);
                idx += k; if (idx >= n2) idx -= n2;
            }
        }
    }
    if (b < nb) X[(size_t)tr * nb + b] = make_float2(re, im);
}

// Inverse (C2R semantics: bins 0 and n2 / 2 once, the others twice) restricted to the bins k0 .. k0 + nb - 1, cropped to nt:
//   y[tr][n] = scale * sum_b h_b Re( G_b X[tr][b] exp(+2 pi i (k0 + b) n / n2) ),  G_b = cgain[b] (complex) * rgain[b] (real)
// scale = scale0 * (*dscale) when dscale != NULL (amp_ratio of source_update_adj lives on the device).  grid (ceil(nt / DF_NT), ntr)
__global__ void __launch_bounds__(DF_NT) k_dft_inv(const float2 *X, const int nb, const int k0, const float2 *cgain, const float *rgain,
                                                    const int n2, const float2 *tw, const float scale0, const float *dscale, float *y, const int nt)
{
    __shared__ float2 sX[DF_CHUNK];
    const int tr = blockIdx.y, n = blockIdx.x * DF_NT + threadIdx.x;
    const float2 *Xr = X + (size_t)tr * nb;
    float acc = 0.f;
    for (int b0 = 0; b0 < nb; b0 += DF_CHUNK) {
        const int m = min(DF_CHUNK, nb - b0);
        __syncthreads();
        for (int j = threadIdx.x; j < m; j += DF_NT) {
            float2 v = Xr[b0 + j];
            if (cgain) { const float2 g = cgain[b0 + j]; v = make_float2(v.x * g.x - v.y * g.y, v.x * g.y + v.y * g.x); }
            float h = 2.0f;
            const int k = k0 + b0 + j;
            if (k == 0 || 2 * k == n2) h = 1.0f;
            if (rgain) h *= rgain[b0 + j];
            sX[j] = make_float2(v.x * h, v.y * h);
        }
        __syncthreads();
        if (n < nt) {
            int idx = (int)(((long long)(k0 + b0) * n) % n2);
#pragma unroll 4
            for (int j = 0; j < m; j++) {
                const float2 w = __ldg(tw + idx);
                acc = fmaf(sX[j].x, w.x, acc); acc = fmaf(-sX[j].y, w.y, acc);     // Re((a + i b)(c + i s)) = a c - b s
                idx += n; if (idx >= n2) idx -= n2;
            }
        }
    }
    if (n < nt) y[(size_t)tr * nt + n] = acc * scale0 * (dscale ? *dscale : 1.0f);
}

// cuda_spectrum_update (utilities.cu:904-975), one block per frequency: coef = sum_r conj(cal) obs / (sum_r |cal|^2 + 1e-6);
// the source spectrum is multiplied by it (the calculated data get it as the complex gain of their inverse transform)
__global__ void __launch_bounds__(256) k_spectrum_coef(const float2 *Fo, const float2 *Fc, const int ntr, const int nb, float2 *Fs, float2 *coef)
{
    __shared__ float sn[256][2], sd[256];
    const int b = blockIdx.x;
    float nr = 0.f, ni = 0.f, dd = 0.f;
    for (int r = threadIdx.x; r < ntr; r += 256) {
        const float2 o = Fo[(size_t)r * nb + b], c = Fc[(size_t)r * nb + b];
        nr += c.x * o.x + c.y * o.y; ni += c.x * o.y - c.y * o.x;      // conj(c) o
        dd += c.x * c.x + c.y * c.y;
    }
    sn[threadIdx.x][0] = nr; sn[threadIdx.x][1] = ni; sd[threadIdx.x] = dd;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { sn[threadIdx.x][0] += sn[threadIdx.x + o][0]; sn[threadIdx.x][1] += sn[threadIdx.x + o][1]; sd[threadIdx.x] += sd[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float den = sd[0] + 1e-6f;
        const float2 c = make_float2(sn[0][0] / den, sn[0][1] / den);
        coef[b] = c;
        const float2 s = Fs[b];
        Fs[b] = make_float2(s.x * c.x - s.y * c.y, s.x * c.y + s.y * c.x);
    }
}

// cuda_find_normfact x 3 (utilities.cu:1011-1040): per trace obs.obs, cal.cal, obs.cal, each + DIVCONST; one block per trace
__global__ void __launch_bounds__(256) k_normfacts(const float *obs, const float *cal, const int nt, float *nf /*[3][ntr]*/, const int ntr)
{
    __shared__ float s3[256][3];
    const int r = blockIdx.x;
    float a = 0.f, b = 0.f, c = 0.f;
    for (int i = threadIdx.x; i < nt; i += 256) {
        const float o = obs[(size_t)r * nt + i], q = cal[(size_t)r * nt + i];
        a += o * o; b += q * q; c += o * q;
    }
    s3[threadIdx.x][0] = a; s3[threadIdx.x][1] = b; s3[threadIdx.x][2] = c;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) for (int j = 0; j < 3; j++) s3[threadIdx.x][j] += s3[threadIdx.x + o][j];
        __syncthreads();
    }
    if (threadIdx.x < 3) nf[(size_t)threadIdx.x * ntr + r] = s3[0][threadIdx.x] + 1e-9f;
}

// cuda_normal_misfit (utilities.cu:1058-1087): -2 sum_r cross / (sqrt(obs) sqrt(cal)) w_r src_weight ; one block, fixed order
__global__ void __launch_bounds__(256) k_cross_misfit(const float *nf, const int ntr, const float *weights, const float src_weight, double *out)
{
    __shared__ double sm[256];
    double acc = 0.0;
    for (int r = threadIdx.x; r < ntr; r += 256)
        acc += (double)(nf[2 * (size_t)ntr + r] / (sqrtf(nf[r]) * sqrtf(nf[(size_t)ntr + r])) * (weights ? weights[r] : 1.0f) * src_weight);
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) *out = -2.0 * sm[0];
}

// cuda_normal_adjoint_source (utilities.cu:1090-1113); grid (ceil(nt / 256), ntr)
__global__ void __launch_bounds__(256) k_cross_adjoint(const float *obs, const float *cal, const float *nf, const int ntr, const int nt,
                                                        const float *weights, const float src_weight, float *res)
{
    const int it = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (it >= nt || r >= ntr) return;
    const float on = nf[r], cn = nf[(size_t)ntr + r], cr = nf[2 * (size_t)ntr + r];
    const size_t i = (size_t)r * nt + it;
    res[i] = (obs[i] - cr / cn * cal[i]) / (sqrtf(on) * sqrtf(cn)) * (weights ? weights[r] : 1.0f) * src_weight;
}

// amp_ratio_comp (utilities.cu:1328-1356): max |obs| / max |cal| (0 when the calculated data vanish); two launches:
// k_absmax2 accumulates the two maxima (bit patterns of non-negative floats order like integers), k_amp_ratio divides.
__global__ void __launch_bounds__(256) k_absmax2(const float *a, const float *b, const size_t n, unsigned *mx /*[2], zeroed*/)
{
    float ma = 0.f, mb = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { ma = fmaxf(ma, fabsf(a[i])); mb = fmaxf(mb, fabsf(b[i])); }
    for (int o = 16; o > 0; o >>= 1) { ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, o)); mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, o)); }
    if ((threadIdx.x & 31) == 0) { atomicMax(mx, __float_as_uint(ma)); atomicMax(mx + 1, __float_as_uint(mb)); }
}
__global__ void k_amp_ratio(const unsigned *mx, float *ratio)
{
    const float mo = __uint_as_float(mx[0]), mc = __uint_as_float(mx[1]);
    *ratio = mc != 0.0f ? mo / mc : 0.0f;
}

// gpuMinus + cuda_cal_objective (utilities.cu:154-205) of ONE slot's conditioned traces: res = obs - cal, sample 0 forced to 0,
// partial sums of res^2 per block (fixed summation order downstream)
__global__ void __launch_bounds__(256) k_residual_one(const float *obs, const float *cal, float *res, const int ntr, const int nt, double *partial)
{
    const size_t n = (size_t)ntr * nt;
    double acc = 0.0;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        const int it = (int)(k % nt);
        const float r = it > 0 ? obs[k] - cal[k] : 0.0f;
        res[k] = r;
        acc += (double)r * r;
    }
    __shared__ double sm[256];
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) partial[blockIdx.x] = sm[0];
}
__global__ void k_sum_partials_one(const double *partial, const int nblk, double *out)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) { double acc = 0.0; for (int k = 0; k < nblk; k++) acc += partial[k]; *out = acc; }
}

// the slot's scaled source amplitudes back to stf units (inverse of stage_batch's 1500^2 dt)
__global__ void k_scale_copy(const float *src, float *dst, const int n, const float scale)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i] * scale;
}

}  // namespace sepfwi

// sepfwi.cu -- host engine and C ABI of libsepfwi.so (see include/sepfwi.h).
//
// One handle per GPU owns persistent device arenas (state, boundary ring store, traces,
// gradients) sized at creation; nothing is allocated, parsed or read from disk inside the
// time loops -- the reference re-does all of that on every call (libCUFD.cu:47-152,215-223).
// Shots are propagated `max_batch` at a time: every kernel takes the slot index from
// blockIdx.z / blockIdx.y, so small grids still fill the 148 SMs.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <vector>

#include "../../include/sepfwi.h"
#include "common.cuh"
#include "kernels_base.cuh"
#include "kernels_stream.cuh"
#include "kernels_resident.cuh"
#include "kernels_data.cuh"

using namespace sepfwi;

static thread_local char g_err[1024] = "";

static int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return fail(e_ == cudaErrorMemoryAllocation ? SEPFWI_ENOMEM : SEPFWI_ECUDA,           \
                        "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));       \
    } while (0)

struct sepfwi_handle {
    sepfwi_params p;
    int device;
    Dims d;
    int B;                 // slots
    bool sponge;
    // device arenas
    float *state = nullptr, *model = nullptr, *cz = nullptr, *cx = nullptr, *cxs = nullptr, *cxv = nullptr, *cxa = nullptr, *damp = nullptr;
    float *ring = nullptr, *trace = nullptr, *grad = nullptr, *gstf = nullptr;
    float *dense[3] = {nullptr, nullptr, nullptr};   // [nz][nx] staging for model in / gradient out
    int *maxcp = nullptr;
    double *partial = nullptr, *misfit = nullptr;
    // slot tables (device) + pinned host mirrors
    int *t_int = nullptr, *h_int = nullptr;          // packed ints
    float *t_flt = nullptr, *h_flt = nullptr;        // packed floats
    size_t n_int = 0, n_flt = 0;
    SlotTab tab;
    // offsets (in elements) of the packed tables
    size_t o_zs, o_xs, o_nrec, o_zrec, o_xrec, o_injN, o_injCell, o_injField, o_injPtr, o_injRec, o_sInjPtr, o_sInj;
    int nStrips = 0;
    bool stream = false;   // register-streaming kernels (kernels = 0 / 3); otherwise the unfused baseline kernels
    bool stream_sponge = false;   // sponge flavour through k_stream_sponge (kernels = 0 / 3)
    TmaMaps tmaps;                // tensor maps of the state / model / gradient blocks (TMA operand path of the interior warps)
    bool tma = false;             // ... valid and enabled (SEPFWI_TMA=0 keeps the cp.async path)
    int nSM = 148;
    size_t smem_optin = 0; // largest opt-in dynamic shared memory per block on this device
    bool pdl = true;       // programmatic dependent launch inside the time loops (SEPFWI_PDL=0 disables)
    int warps_per_sm = SW_MINB * SW_WPB;   // resident warps per SM of the streaming kernels (register-bound)
    // tuning / debugging knobs, read from the environment ONCE at creation (never inside a launch path):
    // SEPFWI_LZ / SEPFWI_LZE chunk heights, SEPFWI_FORCE 1/2 force the interior / edge code path (timing only),
    // SEPFWI_PLAN_DEBUG prints the plans, SEPFWI_RES_DEBUG the resident kernel's timing switches
    int tune_lz = 0, tune_lze = 0, tune_force = 0, res_dbg = 0, res_forced_rpt = 0;
    int pair_lz[3] = {0, 0, 0}, pair_le[3] = {0, 0, 0};      // chunk heights chosen jointly for the merged reverse-time launch (0: none)
    int tune_merge = -1;   // SEPFWI_MERGE: reconstruction + adjoint sweep of a time step in one launch: -1 where it pays (stream_plan_pair), 0 never, 1 always
    int smem_pad = 0;      // SEPFWI_SMEM_PAD: extra dynamic shared memory per streaming CTA (experiment: shrinks the L1 carve-out)
    bool plan_debug = false, res_fake_refuse = false;
    bool res_attr_done[16] = {false};      // cudaFuncSetAttribute issued for k_resident_fwd<RPT> on this handle's device
    int4 *work[3] = {nullptr, nullptr, nullptr};   // work lists of the streaming kernels (device): forward, reconstruction, adjoint
    size_t work_cap[3] = {0, 0, 0};
    size_t o_amp, o_rxz, o_w, o_injCoef;
    bool use_w = false;
    // shared-memory-resident forward loop (kernels_resident.cuh): used when the tiles of a shot fit the SMs
    bool resident = false;
    int *res_dev = nullptr, *res_host = nullptr;    // [flags B*nSM | err | tilePtr B*(nSM+1) | tileRec B*maxRec | ringPtr nSM+1] device / pinned mirror
    int2 *res_ring = nullptr;                       // [ringLen] ring entries bucketed by tile
    int res_ring_rpt = 0;                           // tiling the ring entries were built for
    size_t ro_flags = 0, ro_err = 0, ro_tptr = 0, ro_trec = 0, ro_rptr = 0, res_n = 0;
    int *res_errh = nullptr;                        // pinned copy of the error flag
    int res_used = 0;                               // launches of the resident kernel since creation (introspection)
    // wavefield snapshots of the sponge flavour (sepfwi_forward_snapshots): destination, stride in time steps, pointer space
    float *snap_dst = nullptr;
    int snap_step = 0, snap_mem = SEPFWI_MEM_HOST;
    // host copies
    std::vector<float> hcz, hcx;
    float courant = 0.f;
    bool have_model = false;
    long long launches = 0;
    // data-side operators (kernels_data.cuh): options, host-built tables, device scratch (allocated by sepfwi_set_data_options)
    sepfwi_data_options dopt;
    bool dopt_any = false;
    int d_n2 = 0, d_nfft = 0, d_k0 = 0, d_nband = 0;   // padded length 2 nt, bins of the full spectrum, band-pass bins [k0, k0 + nband)
    float2 *d_tw = nullptr, *d_F[2] = {nullptr, nullptr}, *d_Fs = nullptr, *d_coef = nullptr;
    float *d_gain = nullptr, *d_nf = nullptr, *d_win = nullptr, *d_ratio = nullptr, *d_src = nullptr, *h_win = nullptr, *h_src = nullptr;
    unsigned *d_mx = nullptr;
    double *hj = nullptr; float *hg = nullptr;       // pinned staging: per-shot misfits, stf gradients
    size_t hj_cap = 0, hg_cap = 0;
    double last_misfit = 0.0;                        // misfit of the last sepfwi_gradient call in double precision
    uint64_t staged_sig = 0;                         // signature of the batch the device slot tables hold (0: none)
    uint64_t plan_sig[3] = {0, 0, 0};                // ... and of the work lists (forward, reconstruction, adjoint)
    StreamArgs plan_sa[3];
    int4 *work_host[3] = {nullptr, nullptr, nullptr};    // pinned mirrors of the work lists
    cudaEvent_t ev[4];
    float fwd_ms = 0.f, bwd_ms = 0.f;
    static const int NBLK_RES = 64;
    // per-kernel profile (sepfwi_set_profile): CUDA-event pairs around every launch of the first
    // prof_steps time steps of each loop, on the launching stream
    int prof_steps = 0;
    struct ProfRec { int kind; cudaEvent_t e0, e1; };
    std::vector<ProfRec> prof_recs;
    double prof_ms[SEPFWI_NKERNEL] = {0};
    long long prof_n[SEPFWI_NKERNEL] = {0};
};

static const char *k_names[SEPFWI_NKERNEL] = {"ring_save", "stress_fwd", "velocity_fwd", "record", "velocity_bwd",
                                              "stress_bwd", "velocity_adj", "inject", "stress_adj", "stream_fwd", "stream_recon", "stream_adj", "resident_fwd", "stream_bwd"};

// Launch with programmatic stream serialization (the kernels call griddepcontrol.launch_dependents / .wait themselves).
template <typename... KA, typename... A>
static cudaError_t launch_pdl(void (*kern)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, A... args)
{
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, args...);
}

// launch `stmt` and, in profile mode, bracket it with an event pair
#define LAUNCH(h, KND, prof_on, st, stmt)                                     \
    do {                                                                       \
        if (prof_on) {                                                         \
            sepfwi_handle::ProfRec r_; r_.kind = (KND);                        \
            cudaEventCreate(&r_.e0); cudaEventCreate(&r_.e1);                  \
            cudaEventRecord(r_.e0, st); stmt; cudaEventRecord(r_.e1, st);      \
            (h)->prof_recs.push_back(r_);                                      \
        } else { stmt; }                                                       \
        (h)->launches++;                                                       \
    } while (0)

static void prof_collect(sepfwi_handle *h)
{
    for (auto &r : h->prof_recs) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) { h->prof_ms[r.kind] += ms; h->prof_n[r.kind]++; }
        cudaEventDestroy(r.e0); cudaEventDestroy(r.e1);
    }
    h->prof_recs.clear();
}

extern "C" int sepfwi_set_profile(sepfwi_handle *h, int nsteps)
{
    if (!h) return fail(SEPFWI_EINVAL, "null handle");
    h->prof_steps = nsteps > 0 ? nsteps : 0;
    for (int k = 0; k < SEPFWI_NKERNEL; k++) { h->prof_ms[k] = 0; h->prof_n[k] = 0; }
    return 0;
}
extern "C" int sepfwi_get_profile(sepfwi_handle *h, double *ms, long long *count)
{
    if (!h || !ms || !count) return fail(SEPFWI_EINVAL, "null argument");
    for (int k = 0; k < SEPFWI_NKERNEL; k++) { ms[k] = h->prof_ms[k]; count[k] = h->prof_n[k]; }
    return 0;
}
extern "C" const char *sepfwi_kernel_name(int k) { return k >= 0 && k < SEPFWI_NKERNEL ? k_names[k] : ""; }


extern "C" const char *sepfwi_last_error(void) { return g_err; }
// used by the other translation units of the library to report through the same channel
extern "C" int sepfwi_set_error_(int code, const char *msg) { return fail(code, "%s", msg); }
extern "C" int sepfwi_version(void) { return 100; }

// ------------------------------------------------------------------------------------------
// CPML profiles on the host, same arithmetic as cpmlInit (utilities.cu:243-359): Rcoef 8e-4,
// NPOWER 8, K_MAX 2, alpha_max = pi f0, profile 0.25 d + 0.75 d^8, CpAve fixed to 3000.
// Output rows: 1/K, a, b, 1/K_half, a_half, b_half (K itself kept for sepfwi_get_cpml).
static void build_cpml(int N, int nPml, float dh, float f0, float dt, std::vector<float> &out, std::vector<float> &Kraw)
{
    out.assign((size_t)NCOEF * N, 0.f);
    Kraw.assign((size_t)2 * N, 1.f);
    const double PI = 3.141592653589793238462643383279502884197169;
    const float alpha_max = (float)(2.0 * PI * (f0 / 2.0));
    const float npower = 8.0f, kmax = 2.0f, w1 = 0.25f, w2 = 0.75f;
    const float thick = nPml * dh;
    const float d0 = (float)(-(npower + 1) * 3000.0f * log((double)0.0008f) / (2.0 * thick));
    auto prof = [&](float depth, float &damp, float &K, float &alpha) {
        if (depth >= 0.0f) {
            const float dn = depth / thick;
            damp = (float)(d0 * (w1 * dn + w2 * pow(dn, npower) + 0.0f * pow(dn, 2 * npower)));
            K = (float)(1.0 + (kmax - 1.0) * pow(dn, npower));
            alpha = (float)(alpha_max * (1.0 - dn));
        }
    };
    for (int i = 0; i < N; i++) {
        float damp = 0.f, K = 1.f, alpha = 0.f, damph = 0.f, Kh = 1.f, alphah = 0.f;
        prof((nPml - i) * dh, damp, K, alpha);
        prof((float)((nPml - i - 0.5) * dh), damph, Kh, alphah);
        prof((nPml - N + i) * dh, damp, K, alpha);
        {   // the reference evaluates K_half on the right edge with powf (utilities.cu:323)
            float depth = (float)((nPml - N + i + 0.5) * dh);
            if (depth >= 0.0f) {
                const float dn = depth / thick;
                damph = (float)(d0 * (w1 * dn + w2 * pow(dn, npower) + 0.0f * pow(dn, 2 * npower)));
                Kh = (float)(1.0 + (kmax - 1.0) * powf(dn, npower));
                alphah = (float)(alpha_max * (1.0 - dn));
            }
        }
        if (alpha < 0.f) alpha = 0.f;
        if (alphah < 0.f) alphah = 0.f;
        const float b = expf(-(damp / K + alpha) * dt), bh = expf(-(damph / Kh + alphah) * dt);
        float a = 0.f, ah = 0.f;
        if (fabs(damp) > 1.0e-6) a = (float)(damp * (b - 1.0) / (K * (damp + K * alpha)));
        if (fabs(damph) > 1.0e-6) ah = (float)(damph * (bh - 1.0) / (Kh * (damph + Kh * alphah)));
        out[(size_t)C_RK * N + i] = 1.0f / K;   out[(size_t)C_A * N + i] = a;   out[(size_t)C_B * N + i] = b;
        out[(size_t)C_RKH * N + i] = 1.0f / Kh; out[(size_t)C_AH * N + i] = ah; out[(size_t)C_BH * N + i] = bh;
        Kraw[i] = K; Kraw[N + i] = Kh;
    }
}

// Source-time-function end taper, cuda_window without weights (utilities.cu:844-884), ratio 0.001.
static void taper_stf(int nt, float dt, float ratio, float *s)
{
    const double PI = 3.141592653589793238462643383279502884197169;
    const float t0 = 0.f, t3 = nt * dt, off = nt * dt * ratio;
    if (2.0 * off >= t3 - t0) return;
    const float t1 = t0 + off, t2 = t3 - off;
    for (int i = 0; i < nt; i++) {
        const float t = i * dt;
        float w;
        if (t >= t0 && t < t1) w = (float)sin(PI / 2.0 * (t - t0) / (t1 - t0));
        else if (t >= t1 && t < t2) w = 1.0f;
        else if (t >= t2 && t < t3) w = (float)cos(PI / 2.0 * (t - t2) / (t3 - t2));
        else w = 0.0f;
        s[i] *= w * w;
    }
}

static int fill_dims(const sepfwi_params &p, Dims &d)
{
    if (p.nz <= 0 || p.nx <= 0 || p.nSteps < 2 || p.nPml < 4 || p.nPad < 0) return fail(SEPFWI_EINVAL, "bad grid parameters (need nPml >= 4, nSteps >= 2)");
    if (!(p.dz > 0.f && p.dx > 0.f && p.dt > 0.f)) return fail(SEPFWI_EINVAL, "dz, dx, dt must be positive");
    memset(&d, 0, sizeof(d));
    d.nz = p.nz; d.nx = p.nx; d.nPml = p.nPml; d.nPad = p.nPad; d.nSteps = p.nSteps;
    d.nzA = p.nz - p.nPad;
    if (d.nzA - 2 * p.nPml < 8 || d.nx - 2 * p.nPml < 8) return fail(SEPFWI_EINVAL, "interior smaller than 8 cells");
    d.ldx = (p.nx + 31) / 32 * 32;
    d.z1 = d.nzA - 1 - p.nPml; d.x1 = p.nx - 1 - p.nPml;
    d.nzB = d.nzA - 2 * p.nPml + 4; d.nxB = p.nx - 2 * p.nPml + 4;
    d.ringLen = 2 * 5 * (d.nzB + d.nxB);
    d.maxRec = p.max_nrec > 0 ? p.max_nrec : 1;
    d.dt = p.dt;
    d.c1z = (float)(9.0 / 8.0) / p.dz; d.c2z = (float)(1.0 / 24.0) / p.dz;
    d.c1x = (float)(9.0 / 8.0) / p.dx; d.c2x = (float)(1.0 / 24.0) / p.dx;
    d.rdz = 1.0f / p.dz; d.rdx = 1.0f / p.dx;
    d.fsz = (size_t)d.nzA * d.ldx;
    return 0;
}

extern "C" int sepfwi_ring_len(const sepfwi_params *p)
{
    if (!p) return fail(SEPFWI_EINVAL, "null params");
    return 2 * 5 * ((p->nz - 2 * p->nPml - p->nPad + 4) + (p->nx - 2 * p->nPml + 4));
}

static KArgs kargs(const sepfwi_handle *h)
{
    KArgs a;
    a.d = h->d; a.state = h->state; a.model = h->model; a.cz = h->cz; a.cx = h->cx; a.cxs = h->cxs; a.cxv = h->cxv; a.cxa = h->cxa; a.damp = h->damp;
    a.ring = h->ring; a.trace = h->trace; a.grad = h->grad; a.gstf = h->gstf; a.t = h->tab;
    if (!h->use_w) a.t.w = nullptr;
    return a;
}

static void free_data_scratch(sepfwi_handle *h)
{
    void *dp[] = {h->d_tw, h->d_F[0], h->d_F[1], h->d_Fs, h->d_coef, h->d_gain, h->d_nf, h->d_win, h->d_ratio, h->d_src, h->d_mx};
    for (void *q : dp) if (q) cudaFree(q);
    h->d_tw = nullptr; h->d_F[0] = h->d_F[1] = nullptr; h->d_Fs = h->d_coef = nullptr;
    h->d_gain = h->d_nf = h->d_win = h->d_ratio = h->d_src = nullptr; h->d_mx = nullptr;
    if (h->h_win) cudaFreeHost(h->h_win);
    if (h->h_src) cudaFreeHost(h->h_src);
    h->h_win = h->h_src = nullptr;
}

extern "C" int sepfwi_destroy(sepfwi_handle *h)
{
    if (!h) return 0;
    cudaSetDevice(h->device);
    free_data_scratch(h);
    float *fp[] = {h->state, h->model, h->cz, h->cx, h->cxs, h->cxv, h->cxa, h->damp, h->ring, h->trace, h->grad, h->gstf,
                   h->dense[0], h->dense[1], h->dense[2], h->t_flt};
    for (float *q : fp) if (q) cudaFree(q);
    if (h->maxcp) cudaFree(h->maxcp);
    for (int k = 0; k < 3; k++) { if (h->work[k]) cudaFree(h->work[k]); if (h->work_host[k]) cudaFreeHost(h->work_host[k]); }
    if (h->partial) cudaFree(h->partial);
    if (h->misfit) cudaFree(h->misfit);
    if (h->t_int) cudaFree(h->t_int);
    if (h->h_int) cudaFreeHost(h->h_int);
    if (h->h_flt) cudaFreeHost(h->h_flt);
    if (h->res_dev) cudaFree(h->res_dev);
    if (h->res_ring) cudaFree(h->res_ring);
    if (h->res_host) cudaFreeHost(h->res_host);
    if (h->res_errh) cudaFreeHost(h->res_errh);
    if (h->hj) cudaFreeHost(h->hj);
    if (h->hg) cudaFreeHost(h->hg);
    for (int i = 0; i < 4; i++) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    delete h;
    return 0;
}

// Tensor maps for the TMA operand path: rank-4 view [slot][array][z][x] of the state block, rank-3 views [array][z][x] of the model
// and (per slot) gradient blocks; boxes of one 128-float row.  The encoder is a driver entry point fetched at run time.
static bool build_tensor_maps(sepfwi_handle *h)
{
    typedef CUresult (*encode_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                 const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn || qr != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return false;
    }
    encode_t enc = (encode_t)fn;
    const Dims &d = h->d;
    memset(&h->tmaps, 0, sizeof(h->tmaps));
    const cuuint32_t one[4] = {1, 1, 1, 1};
    {
        const cuuint64_t dims[4] = {(cuuint64_t)d.ldx, (cuuint64_t)d.nzA, (cuuint64_t)(d.sstride / d.fsz), (cuuint64_t)h->B};
        const cuuint64_t str[3] = {(cuuint64_t)d.ldx * 4, (cuuint64_t)d.fsz * 4, (cuuint64_t)d.sstride * 4};
        const cuuint32_t box[4] = {128, 1, 1, 1};
        if (enc(&h->tmaps.state, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, h->state, dims, str, box, one, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return false;
    }
    {
        const cuuint64_t dims[3] = {(cuuint64_t)d.ldx, (cuuint64_t)d.nzA, (cuuint64_t)NMODEL};
        const cuuint64_t str[2] = {(cuuint64_t)d.ldx * 4, (cuuint64_t)d.fsz * 4};
        const cuuint32_t box[3] = {128, 1, 1};
        if (enc(&h->tmaps.model, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, h->model, dims, str, box, one, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return false;
    }
    if (h->grad) {
        const cuuint64_t dims[3] = {(cuuint64_t)d.ldx, (cuuint64_t)d.nzA, (cuuint64_t)3 * h->B};
        const cuuint64_t str[2] = {(cuuint64_t)d.ldx * 4, (cuuint64_t)d.fsz * 4};
        const cuuint32_t box[3] = {128, 1, 1};
        if (enc(&h->tmaps.grad, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, h->grad, dims, str, box, one, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return false;
    } else h->tmaps.grad = h->tmaps.model;
    return true;
}

extern "C" int sepfwi_create(const sepfwi_params *pp, int device, sepfwi_handle **out)
{
    if (!pp || !out) return fail(SEPFWI_EINVAL, "null argument");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(SEPFWI_ECUDA, "no CUDA device available (%s); libsepfwi has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(SEPFWI_EINVAL, "device %d out of range (have %d)", device, ndev);
    CU(cudaSetDevice(device));
    sepfwi_handle *h = new sepfwi_handle();
    for (int i = 0; i < 4; i++) h->ev[i] = nullptr;
    h->p = *pp; h->device = device;
    memset(&h->dopt, 0, sizeof(h->dopt));
    h->sponge = pp->flavour == SEPFWI_FLAVOUR_SPONGE;
    int rc = fill_dims(*pp, h->d);
    if (rc) { delete h; return rc; }
    h->B = std::max(1, pp->max_batch);
    h->d.nTrace = h->sponge ? NTRACE : 4;
    if (pp->kernels != 0 && pp->kernels != 1 && pp->kernels != 3) { delete h; return fail(SEPFWI_EINVAL, "kernels must be 0 (default), 1 (unfused baseline) or 3 (streaming only)"); }
    if (h->sponge && pp->with_adjoint) { delete h; return fail(SEPFWI_EINVAL, "the sponge flavour is forward-only (the reference has no adjoint for it)"); }
    const Dims &d = h->d;
    const int B = h->B;
#define ALLOC(ptr, n)                                                                                    \
    do {                                                                                                 \
        cudaError_t e2 = cudaMalloc((void **)&(ptr), (n));                                               \
        if (e2 != cudaSuccess) {                                                                         \
            int c = fail(SEPFWI_ENOMEM, "cudaMalloc of %zu bytes for %s failed: %s", (size_t)(n), #ptr,  \
                         cudaGetErrorString(e2));                                                        \
            sepfwi_destroy(h);                                                                           \
            return c;                                                                                    \
        }                                                                                                \
    } while (0)
    h->d.sstride = (size_t)(pp->with_adjoint ? NSTATE : NSTATE_FWD) * d.fsz;
    ALLOC(h->state, (size_t)B * d.sstride * sizeof(float));
    ALLOC(h->model, (size_t)NMODEL * d.fsz * sizeof(float));
    ALLOC(h->cz, (size_t)NCOEF * d.nzA * sizeof(float));
    ALLOC(h->cx, (size_t)NCOEF * d.nx * sizeof(float));
    ALLOC(h->cxs, (size_t)NCOEF * d.ldx * sizeof(float));
    ALLOC(h->cxv, (size_t)NCOEF * d.ldx * sizeof(float));
    ALLOC(h->cxa, (size_t)NCOEF * d.ldx * sizeof(float));
    ALLOC(h->trace, (size_t)B * d.nTrace * d.maxRec * d.nSteps * sizeof(float));
    ALLOC(h->gstf, (size_t)B * d.nSteps * sizeof(float));
    for (int k = 0; k < 3; k++) ALLOC(h->dense[k], (size_t)d.nz * d.nx * sizeof(float));
    ALLOC(h->maxcp, sizeof(int));
    ALLOC(h->partial, (size_t)B * sepfwi_handle::NBLK_RES * sizeof(double));
    ALLOC(h->misfit, (size_t)B * sizeof(double));
    if (pp->with_adjoint) {
        ALLOC(h->ring, (size_t)B * NFIELD * d.nSteps * d.ringLen * sizeof(float));
        ALLOC(h->grad, (size_t)B * 3 * d.fsz * sizeof(float));
    }
    if (h->sponge) ALLOC(h->damp, d.fsz * sizeof(float));
    CU(cudaMemset(h->model, 0, (size_t)NMODEL * d.fsz * sizeof(float)));

    // packed slot tables
    const int maxCon = 8 * d.maxRec, maxInj = maxCon;
    size_t oi = 0;
    auto takei = [&](size_t n) { size_t o = oi; oi += n; return o; };
    h->o_zs = takei(B); h->o_xs = takei(B); h->o_nrec = takei(B);
    h->o_zrec = takei((size_t)B * d.maxRec); h->o_xrec = takei((size_t)B * d.maxRec);
    h->o_injN = takei(B); h->o_injCell = takei((size_t)B * maxInj); h->o_injField = takei((size_t)B * maxInj);
    h->o_injPtr = takei((size_t)B * (maxInj + 1)); h->o_injRec = takei((size_t)B * maxCon);
    h->nStrips = (d.nx + SW_OWN - 1) / SW_OWN;
    h->o_sInjPtr = takei((size_t)B * h->nStrips * (d.nzA + 1)); h->o_sInj = takei((size_t)B * 2 * maxInj);
    h->n_int = oi;
    h->stream = !h->sponge && (pp->kernels == 0 || pp->kernels == 3);
    if (const char *e = getenv("SEPFWI_PDL")) h->pdl = atoi(e) != 0;
    if (const char *e = getenv("SEPFWI_LZ")) h->tune_lz = atoi(e);
    if (const char *e = getenv("SEPFWI_LZE")) h->tune_lze = atoi(e);
    if (const char *e = getenv("SEPFWI_FORCE")) h->tune_force = atoi(e);
    if (const char *e = getenv("SEPFWI_SMEM_PAD")) h->smem_pad = atoi(e);
    if (const char *e = getenv("SEPFWI_MERGE")) h->tune_merge = atoi(e);
    if (const char *e = getenv("SEPFWI_RES_DEBUG")) h->res_dbg = atoi(e);
    if (const char *e = getenv("SEPFWI_RESIDENT_RPT")) h->res_forced_rpt = atoi(e);
    h->plan_debug = getenv("SEPFWI_PLAN_DEBUG") != nullptr;
    h->res_fake_refuse = getenv("SEPFWI_RES_FAKE_REFUSE") != nullptr;
    {
        cudaDeviceProp prop;
        CU(cudaGetDeviceProperties(&prop, device));
        h->nSM = prop.multiProcessorCount;
        h->smem_optin = (size_t)prop.sharedMemPerBlockOptin;
    }
    h->stream_sponge = h->sponge && pp->kernels != 1;
    if (h->stream_sponge) CU(cudaFuncSetAttribute(k_stream_sponge, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SP_SMEM + h->smem_pad));
    if (h->stream) {
        CU(cudaFuncSetAttribute(k_stream_recon, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RC_SMEM + h->smem_pad));
        CU(cudaFuncSetAttribute(k_stream_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FR_SMEM + h->smem_pad));
        CU(cudaFuncSetAttribute(k_stream_adj, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AR_SMEM + h->smem_pad));
        CU(cudaFuncSetAttribute(k_stream_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RC_SMEM));
    }
    h->resident = h->stream && pp->kernels == 0 && d.nPml <= RS_PW;
    if (const char *e = getenv("SEPFWI_RESIDENT")) h->resident = h->resident && atoi(e) != 0;
    if (h->resident) {
        int coop = 0;
        CU(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
        if (!coop) h->resident = false;
    }
    if (h->resident) {
        size_t o = 0;
        h->ro_flags = o; o += (size_t)B * h->nSM * RS_FLAGW;
        h->ro_err = o; o += 1;
        h->ro_tptr = o; o += (size_t)B * (h->nSM + 1);
        h->ro_trec = o; o += (size_t)B * d.maxRec;
        h->ro_rptr = o; o += (size_t)h->nSM + 1;
        h->res_n = o;
        ALLOC(h->res_dev, o * sizeof(int));
        ALLOC(h->res_ring, (size_t)std::max(1, d.ringLen) * sizeof(int2));
        CU(cudaMallocHost((void **)&h->res_host, o * sizeof(int)));
        CU(cudaMallocHost((void **)&h->res_errh, sizeof(int)));
        memset(h->res_host, 0, o * sizeof(int));
        *h->res_errh = 0;
    }
    size_t of = 0;
    auto takef = [&](size_t n) { size_t o = of; of += n; return o; };
    h->o_amp = takef((size_t)B * d.nSteps); h->o_rxz = takef(B); h->o_w = takef((size_t)B * d.maxRec * 3);
    h->o_injCoef = takef((size_t)B * maxCon);
    h->n_flt = of;
    ALLOC(h->t_int, h->n_int * sizeof(int));
    ALLOC(h->t_flt, h->n_flt * sizeof(float));
    CU(cudaMallocHost((void **)&h->h_int, h->n_int * sizeof(int)));
    CU(cudaMallocHost((void **)&h->h_flt, h->n_flt * sizeof(float)));
    memset(h->h_int, 0, h->n_int * sizeof(int));
    memset(h->h_flt, 0, h->n_flt * sizeof(float));
    SlotTab &t = h->tab;
    t.zs = h->t_int + h->o_zs; t.xs = h->t_int + h->o_xs; t.nrec = h->t_int + h->o_nrec;
    t.zrec = h->t_int + h->o_zrec; t.xrec = h->t_int + h->o_xrec;
    t.injN = h->t_int + h->o_injN; t.injCell = h->t_int + h->o_injCell; t.injField = h->t_int + h->o_injField;
    t.injPtr = h->t_int + h->o_injPtr; t.injRec = h->t_int + h->o_injRec;
    t.amp = h->t_flt + h->o_amp; t.rxz = h->t_flt + h->o_rxz; t.w = h->t_flt + h->o_w; t.injCoef = h->t_flt + h->o_injCoef;
    t.maxInj = maxInj; t.maxCon = maxCon;
    t.sInjPtr = h->t_int + h->o_sInjPtr; t.sInj = h->t_int + h->o_sInj; t.nStrips = h->nStrips;

    // CPML profiles / sponge
    if (!h->sponge) {
        std::vector<float> kz, kx;
        build_cpml(d.nzA, d.nPml, pp->dz, pp->f0, pp->dt, h->hcz, kz);
        build_cpml(d.nx, d.nPml, pp->dx, pp->f0, pp->dt, h->hcx, kx);
        // keep raw K for introspection: store behind the six rows
        h->hcz.insert(h->hcz.end(), kz.begin(), kz.end());
        h->hcx.insert(h->hcx.end(), kx.begin(), kx.end());
        CU(cudaMemcpy(h->cz, h->hcz.data(), (size_t)NCOEF * d.nzA * sizeof(float), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(h->cx, h->hcx.data(), (size_t)NCOEF * d.nx * sizeof(float), cudaMemcpyHostToDevice));
        // quad-aligned copies for the streaming kernels, neutral wherever the reference kernel skips the CPML branch
        std::vector<float> ts((size_t)NCOEF * d.ldx), tv((size_t)NCOEF * d.ldx), ta((size_t)NCOEF * d.ldx);
        for (int x = 0; x < d.ldx; x++) {
            const bool ins = x < d.nx && ((x < d.nPml) || (x > d.nx - d.nPml - 1));     // el_stress.cu:63
            const bool inv = x < d.nx && ((x < d.nPml) || (x > d.nx - d.nPml));         // el_velocity.cu:56,71
            for (int k = 0; k < NCOEF; k++) {
                const float neutral = (k == C_RK || k == C_RKH) ? 1.0f : 0.0f;
                ts[(size_t)k * d.ldx + x] = ins ? h->hcx[(size_t)k * d.nx + x] : neutral;
                tv[(size_t)k * d.ldx + x] = inv ? h->hcx[(size_t)k * d.nx + x] : neutral;
                ta[(size_t)k * d.ldx + x] = x < d.nx ? h->hcx[(size_t)k * d.nx + x] : neutral;
            }
        }
        CU(cudaMemcpy(h->cxs, ts.data(), ts.size() * sizeof(float), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(h->cxv, tv.data(), tv.size() * sizeof(float), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(h->cxa, ta.data(), ta.size() * sizeof(float), cudaMemcpyHostToDevice));
    } else {
        // multiplicative sponge sin^2(pi/2 i/ndamp) from all four sides, elasticSolver.py:74-79
        std::vector<float> dm(d.fsz, 1.0f);
        const int nd = d.nPml;
        for (int z = 0; z < d.nzA; z++)
            for (int x = 0; x < d.nx; x++) {
                double v = 1.0;
                auto w = [&](int i) { double s_ = sin(M_PI / 2 * i / nd); return s_ * s_; };
                if (x < nd) v *= w(x);
                if (x >= d.nx - nd) v *= w(d.nx - 1 - x);
                if (z < nd) v *= w(z);
                if (z >= d.nzA - nd) v *= w(d.nzA - 1 - z);
                dm[(size_t)z * d.ldx + x] = (float)v;
            }
        CU(cudaMemcpy(h->damp, dm.data(), d.fsz * sizeof(float), cudaMemcpyHostToDevice));
    }
    for (int i = 0; i < 4; i++) CU(cudaEventCreate(&h->ev[i]));
    if (h->stream) {
        // TMA operand path of the forward kernel's interior warps: measured slower than the cp.async ring on every grid (round 2,
        // profiles/README.md: one row per tensor copy has a longer latency than 32 x 16-byte cp.async, and the march is latency-bound),
        // so it is opt-in: SEPFWI_TMA=1
        const char *e = getenv("SEPFWI_TMA");
        h->tma = e && atoi(e) != 0 && build_tensor_maps(h);
    }
    *out = h;
    return 0;
}

extern "C" int sepfwi_get_cpml(sepfwi_handle *h, int axis, float *out)
{
    if (!h || !out || h->sponge) return fail(SEPFWI_EINVAL, "bad argument");
    const int N = axis == 0 ? h->d.nzA : h->d.nx;
    const std::vector<float> &c = axis == 0 ? h->hcz : h->hcx;
    const float *K = c.data() + (size_t)NCOEF * N;
    for (int i = 0; i < N; i++) {
        out[0 * N + i] = K[i];           out[1 * N + i] = c[(size_t)C_A * N + i];  out[2 * N + i] = c[(size_t)C_B * N + i];
        out[3 * N + i] = K[N + i];       out[4 * N + i] = c[(size_t)C_AH * N + i]; out[5 * N + i] = c[(size_t)C_BH * N + i];
    }
    return 0;
}

extern "C" long long sepfwi_launch_count(sepfwi_handle *h) { return h ? h->launches : 0; }
extern "C" long long sepfwi_resident_launches(sepfwi_handle *h) { return h ? h->res_used : 0; }

extern "C" int sepfwi_last_misfit(sepfwi_handle *h, double *misfit)
{
    if (!h || !misfit) return fail(SEPFWI_EINVAL, "null argument");
    *misfit = h->last_misfit;
    return 0;
}

// Device bytes one concurrent shot ("slot") costs on a handle created with these parameters -- what sepfwi_create allocates
// per unit of max_batch (state block, boundary ring, traces, gradients, slot tables); host arithmetic only.
extern "C" long long sepfwi_bytes_per_slot(const sepfwi_params *pp)
{
    if (!pp) return fail(SEPFWI_EINVAL, "null params");
    Dims d;
    int rc = fill_dims(*pp, d);
    if (rc) return rc;
    const bool sponge = pp->flavour == SEPFWI_FLAVOUR_SPONGE;
    const size_t nTrace = sponge ? NTRACE : 4;
    size_t b = (size_t)(pp->with_adjoint ? NSTATE : NSTATE_FWD) * d.fsz * sizeof(float);
    b += nTrace * (size_t)d.maxRec * d.nSteps * sizeof(float) + (size_t)d.nSteps * sizeof(float);
    if (pp->with_adjoint) b += (size_t)NFIELD * d.nSteps * d.ringLen * sizeof(float) + 3 * d.fsz * sizeof(float);
    const size_t nStrips = (d.nx + SW_OWN - 1) / SW_OWN;
    b += (size_t)(50 * d.maxRec + 8 + nStrips * (d.nzA + 1)) * sizeof(int) + (size_t)(d.nSteps + 11 * d.maxRec) * sizeof(float);
    return (long long)b;
}

extern "C" int sepfwi_last_timing(sepfwi_handle *h, float *fwd_ms, float *bwd_ms)
{
    if (!h) return fail(SEPFWI_EINVAL, "null handle");
    if (fwd_ms) *fwd_ms = h->fwd_ms;
    if (bwd_ms) *bwd_ms = h->bwd_ms;
    return 0;
}

extern "C" int sepfwi_set_model(sepfwi_handle *h, const float *lam, const float *mu, const float *rho, int mem, void *stream)
{
    if (!h || !lam || !mu || !rho) return fail(SEPFWI_EINVAL, "null argument");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const Dims &d = h->d;
    const size_t nb = (size_t)d.nz * d.nx * sizeof(float);
    const float *src[3] = {lam, mu, rho};
    const float *dev[3];
    for (int k = 0; k < 3; k++) {
        if (mem == SEPFWI_MEM_HOST) {
            CU(cudaMemcpyAsync(h->dense[k], src[k], nb, cudaMemcpyHostToDevice, st));
            dev[k] = h->dense[k];
        } else dev[k] = src[k];
    }
    CU(cudaMemsetAsync(h->maxcp, 0, sizeof(int), st));
    dim3 blk(32, 8), grd((d.nx + 31) / 32, (d.nz + 7) / 8);
    if (h->sponge) k_model_prep<true><<<grd, blk, 0, st>>>(d, dev[0], dev[1], dev[2], h->model, h->maxcp);
    else k_model_prep<false><<<grd, blk, 0, st>>>(d, dev[0], dev[1], dev[2], h->model, h->maxcp);
    h->launches++;
    CU(cudaGetLastError());
    int bits = 0;
    CU(cudaMemcpyAsync(&bits, h->maxcp, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    float cpmax;
    memcpy(&cpmax, &bits, 4);
    const float dh = h->p.dz < h->p.dx ? h->p.dz : h->p.dx;
    h->courant = (float)(cpmax * h->p.dt * sqrtf(2.0f) * (1.0 / 24.0 + 9.0 / 8.0) / dh);   // utilities.cu:235
    h->have_model = true;
    if (h->courant > 1.0f) {
        h->have_model = false;
        return fail(SEPFWI_ECOURANT, "Courant_number = %g > 1 (max Cp %g, dt %g, dh %g)", h->courant, cpmax, h->p.dt, dh);
    }
    return 0;
}

extern "C" int sepfwi_courant(sepfwi_handle *h, float *c)
{
    if (!h || !c) return fail(SEPFWI_EINVAL, "null argument");
    *c = h->courant;
    return 0;
}

// ------------------------------------------------------------------------------------------
// Per-batch host staging of the slot tables.
struct InjEntry { int field, cell, rec; float coef; };

// 64-bit mix of a byte range (word-wise; only used to recognise "the same shots as last time")
static uint64_t hash_bytes(uint64_t hsh, const void *p, size_t n)
{
    const unsigned char *b = (const unsigned char *)p;
    size_t i = 0;
    for (; i + 8 <= n; i += 8) { uint64_t w; memcpy(&w, b + i, 8); hsh = (hsh ^ w) * 0x9E3779B97F4A7C15ull; hsh ^= hsh >> 29; }
    uint64_t w = 0;
    if (i < n) memcpy(&w, b + i, n - i);
    hsh = (hsh ^ w ^ (uint64_t)n) * 0xC2B2AE3D27D4EB4Full; hsh ^= hsh >> 32;
    return hsh;
}

static uint64_t batch_signature(const sepfwi_handle *h, int nb, const sepfwi_shot *shots, bool need_inject)
{
    uint64_t g = 0x5EF0F1ull + (uint64_t)nb * 131 + (need_inject ? 7 : 0);
    for (int s = 0; s < nb; s++) {
        const sepfwi_shot &sh = shots[s];
        const int hd[3] = {sh.zs, sh.xs, sh.nrec};
        g = hash_bytes(g, hd, sizeof(hd));
        g = hash_bytes(g, &sh.src_rxz, sizeof(float));
        if (sh.nrec > 0 && sh.zrec && sh.xrec) { g = hash_bytes(g, sh.zrec, (size_t)sh.nrec * 4); g = hash_bytes(g, sh.xrec, (size_t)sh.nrec * 4); }
        if (sh.stf) g = hash_bytes(g, sh.stf, (size_t)h->d.nSteps * 4);
        if (sh.weights) g = hash_bytes(g, sh.weights, (size_t)sh.nrec * 12);
        else g = hash_bytes(g, "nw", 2);
    }
    return g | 1ull;      // never 0 (0 = nothing staged)
}

// Fills the pinned mirrors and uploads them -- unless the device tables already hold exactly this batch (an inversion calls
// with the same shots on every evaluation: then nothing is recomputed, allocated or copied).
static int stage_batch(sepfwi_handle *h, int nb, const sepfwi_shot *shots, bool need_inject, cudaStream_t st)
{
    const Dims &d = h->d;
    const int maxCon = h->tab.maxCon, maxInj = h->tab.maxInj;
    for (int s = 0; s < nb; s++) {
        const sepfwi_shot &sh = shots[s];
        if (sh.nrec < 0 || sh.nrec > d.maxRec) return fail(SEPFWI_EINVAL, "shot %d: nrec %d exceeds max_nrec %d", s, sh.nrec, d.maxRec);
        if (sh.zs < 2 || sh.zs > d.nzA - 3 || sh.xs < 2 || sh.xs > d.nx - 3) return fail(SEPFWI_EINVAL, "shot %d: source (%d,%d) outside the active grid", s, sh.zs, sh.xs);
        if (!sh.stf) return fail(SEPFWI_EINVAL, "shot %d: null stf", s);
        if (sh.nrec > 0 && (!sh.zrec || !sh.xrec)) return fail(SEPFWI_EINVAL, "shot %d: null receiver arrays", s);
        for (int r = 0; r < sh.nrec; r++) {
            const int z = sh.zrec[r], x = sh.xrec[r];
            if (z < 1 || z > d.nzA - 2 || x < 1 || x > d.nx - 2) return fail(SEPFWI_EINVAL, "shot %d: receiver %d at (%d,%d) outside the grid", s, r, z, x);
        }
    }
    const uint64_t sig = batch_signature(h, nb, shots, need_inject);
    if (sig == h->staged_sig) return 0;
    // the previous batch's upload may still be reading the pinned mirrors
    if (h->staged_sig) CU(cudaStreamSynchronize(st));
    h->staged_sig = 0;
    h->use_w = false;
    for (int s = 0; s < nb; s++) if (shots[s].weights) h->use_w = true;
    for (int s = 0; s < nb; s++) {
        const sepfwi_shot &sh = shots[s];
        h->h_int[h->o_zs + s] = sh.zs; h->h_int[h->o_xs + s] = sh.xs; h->h_int[h->o_nrec + s] = sh.nrec;
        for (int r = 0; r < sh.nrec; r++) {
            h->h_int[h->o_zrec + (size_t)s * d.maxRec + r] = sh.zrec[r];
            h->h_int[h->o_xrec + (size_t)s * d.maxRec + r] = sh.xrec[r];
            float *w = h->h_flt + h->o_w + ((size_t)s * d.maxRec + r) * 3;
            if (sh.weights) { w[0] = sh.weights[3 * r]; w[1] = sh.weights[3 * r + 1]; w[2] = sh.weights[3 * r + 2]; }
            else { w[0] = h->p.fiber == SEPFWI_FIBER_EXX ? 1.f : 0.f; w[1] = h->p.fiber == SEPFWI_FIBER_EZZ ? 1.f : 0.f; w[2] = 0.f; }
        }
        // source amplitude per step
        float *amp = h->h_flt + h->o_amp + (size_t)s * d.nSteps;
        memcpy(amp, sh.stf, (size_t)d.nSteps * sizeof(float));
        if (!h->sponge) {
            taper_stf(d.nSteps, h->p.dt, 0.001f, amp);                       // Src_Rec.cu:137
            const float scale = (float)pow(1500.0, 2);                       // utilities.cu:531
            for (int it = 0; it < d.nSteps; it++) amp[it] = scale * amp[it] * h->p.dt;
        } else {
            for (int it = 0; it < d.nSteps; it++) amp[it] = (float)((double)amp[it] * (double)h->p.dt / 2.0);   // elasticSolver.py:259
        }
        h->h_flt[h->o_rxz + s] = sh.src_rxz;
        if (need_inject) {
            std::vector<InjEntry> v;
            v.reserve((size_t)sh.nrec * 2);
            for (int r = 0; r < sh.nrec; r++) {
                const int z = sh.zrec[r], x = sh.xrec[r];
                const float *w = h->h_flt + h->o_w + ((size_t)s * d.maxRec + r) * 3;
                const int c = z * d.ldx + x;
                // reference-race compatibility: first receiver of a 32-receiver block whose `-=` lands on the
                // cell the previous receiver (last thread of the previous block) adds to loses that update
                const bool seam = h->p.ref_race_compat && r > 0 && (r % 32) == 0;
                const bool drop_x = seam && sh.zrec[r - 1] == z && sh.xrec[r - 1] == x - 1;
                const bool drop_z = seam && sh.xrec[r - 1] == x && sh.zrec[r - 1] == z - 1;
                if (w[0] != 0.f) { v.push_back({F_VX, c, r, w[0]}); if (!drop_x) v.push_back({F_VX, c - 1, r, -w[0]}); }
                if (w[1] != 0.f) { v.push_back({F_VZ, c, r, w[1]}); if (!drop_z) v.push_back({F_VZ, c - d.ldx, r, -w[1]}); }
                if (w[2] != 0.f) {
                    v.push_back({F_VX, c + d.ldx, r, 0.5f * w[2]}); v.push_back({F_VX, c, r, -0.5f * w[2]});
                    v.push_back({F_VZ, c + 1, r, 0.5f * w[2]});     v.push_back({F_VZ, c, r, -0.5f * w[2]});
                }
            }
            std::stable_sort(v.begin(), v.end(), [](const InjEntry &a, const InjEntry &b) {
                return a.field != b.field ? a.field < b.field : a.cell < b.cell; });
            if ((int)v.size() > maxCon) return fail(SEPFWI_EINVAL, "too many injection terms");
            int *cell = h->h_int + h->o_injCell + (size_t)s * maxInj, *fld = h->h_int + h->o_injField + (size_t)s * maxInj;
            int *ptr = h->h_int + h->o_injPtr + (size_t)s * (maxInj + 1), *rec = h->h_int + h->o_injRec + (size_t)s * maxCon;
            float *coef = h->h_flt + h->o_injCoef + (size_t)s * maxCon;
            int nt = 0;
            for (size_t k = 0; k < v.size(); k++) {
                if (k == 0 || v[k].field != v[k - 1].field || v[k].cell != v[k - 1].cell) { cell[nt] = v[k].cell; fld[nt] = v[k].field; ptr[nt] = (int)k; nt++; }
                rec[k] = v[k].rec; coef[k] = v[k].coef;
            }
            ptr[nt] = (int)v.size();
            h->h_int[h->o_injN + s] = nt;
            // bucket the targets by streaming strip (columns [120 k - 4, 120 k + 124)) and row
            const int nS = h->nStrips, nR = d.nzA + 1;
            int *sp = h->h_int + h->o_sInjPtr + (size_t)s * nS * nR, *sl = h->h_int + h->o_sInj + (size_t)s * 2 * maxInj;
            std::fill(sp, sp + (size_t)nS * nR, 0);
            auto for_strips = [&](int m, auto &&fn) {
                const int z = cell[m] / d.ldx, x = cell[m] % d.ldx;
                for (int k = std::max(0, (x - 124) / SW_OWN); k < nS && k * SW_OWN - 4 <= x; k++)
                    if (x >= k * SW_OWN - 4 && x < k * SW_OWN + SW_OWN + 4) fn(k, z);
            };
            // counting sort over (strip, row); entry [k][z] first holds the count of row z-1 ... then the running offset
            std::vector<int> cnt((size_t)nS * d.nzA, 0);
            for (int m = 0; m < nt; m++) for_strips(m, [&](int k, int z) { cnt[(size_t)k * d.nzA + z]++; });
            int run = 0;
            for (int k = 0; k < nS; k++) {
                for (int z = 0; z < d.nzA; z++) { sp[(size_t)k * nR + z] = run; run += cnt[(size_t)k * d.nzA + z]; }
                sp[(size_t)k * nR + d.nzA] = run;
            }
            std::vector<int> cur2((size_t)nS * d.nzA);
            for (int k = 0; k < nS; k++) for (int z = 0; z < d.nzA; z++) cur2[(size_t)k * d.nzA + z] = sp[(size_t)k * nR + z];
            for (int m = 0; m < nt; m++) for_strips(m, [&](int k, int z) { sl[cur2[(size_t)k * d.nzA + z]++] = m; });
        } else h->h_int[h->o_injN + s] = 0;
    }
    for (int s = nb; s < h->B; s++) {
        h->h_int[h->o_nrec + s] = 0; h->h_int[h->o_injN + s] = 0;
        std::fill(h->h_int + h->o_sInjPtr + (size_t)s * h->nStrips * (d.nzA + 1), h->h_int + h->o_sInjPtr + (size_t)(s + 1) * h->nStrips * (d.nzA + 1), 0);
    }
    CU(cudaMemcpyAsync(h->t_int, h->h_int, h->n_int * sizeof(int), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(h->t_flt, h->h_flt, h->n_flt * sizeof(float), cudaMemcpyHostToDevice, st));
    h->staged_sig = sig;
    for (int k = 0; k < 3; k++) h->plan_sig[k] = 0;      // the adjoint plan looks at the staged injection tables
    return 0;
}

static int max_nrec(int nb, const sepfwi_shot *shots)
{
    int m = 0;
    for (int s = 0; s < nb; s++) m = std::max(m, shots[s].nrec);
    return m;
}

// Latency-regime cost of the CTAs of ONE shot for chunk heights (Lz interior, le edge), in interior-row units, launch order
// (edge items first): an edge row costs rho interior rows, an item pays its lead-in rows and a prologue (calibrated on the B200
// with forced-height sweeps, tools/sweep_plan.sh).  model: 0 forward, 1 reconstruction, 2 adjoint.
static void l2_cta_costs(const sepfwi_handle *h, int nb, int model, int Lz, int le, std::vector<float> &cta)
{
    const Dims &d = h->d;
    static const double narr[3] = {20.0, 26.0, 20.0};
    const double ws = (double)nb * d.nzA * d.nx * 4.0 * narr[model];
    const double bl = std::min(1.0, std::max(0.0, (ws - 60.0e6) / 140.0e6));
    static const double rho_res[3] = {2.0, 3.0, 4.0}, lead[3] = {2.2, 2.6, 2.2};
    const double ci = 1.0 + 2.0 * bl, rho = rho_res[model] * (1.0 + 0.25 * bl), P = 3.0, H = lead[model];
    const int nStrips = (d.nx + SW_OWN - 1) / SW_OWN, zi0 = d.nPml + 5, zi1 = d.nzA - d.nPml - 5;
    auto strip_inner = [&](int sx) { const int x0 = sx * SW_OWN; return x0 - 4 >= d.nPml + 3 && x0 + SW_OWN + 3 <= d.nx - d.nPml - 4; };
    std::vector<float> cost;
    auto piece = [&](int z0, int z1, int L, bool is_edge) {
        const int n = (z1 - z0 + L - 1) / L;
        for (int c = 0; c < n; c++) {
            const int a0 = z0 + (int)((long long)(z1 - z0) * c / n), a1 = z0 + (int)((long long)(z1 - z0) * (c + 1) / n);
            int rows = a1 - a0;
            if (model == 1) {      // reconstruction: rows outside interior + ring are skipped, items without any are not launched
                rows = std::max(0, std::min(a1, d.z1 + 3) - std::max(a0, d.nPml - 2));
                if (rows == 0) continue;
            }
            cost.push_back(is_edge ? (float)((rows + H + P) * rho) : (float)((((rows + 4 + 1) / 2) * 2 - 4 + H + P) * ci));
        }
    };
    for (int sx = 0; sx < nStrips; sx++) { piece(0, zi0, le, true); piece(zi1, d.nzA, le, true); }
    for (int sx = 0; sx < nStrips; sx++) if (!strip_inner(sx)) piece(zi0, zi1, le, true);
    for (int sx = 0; sx < nStrips; sx++) if (strip_inner(sx)) piece(zi0, zi1, Lz, false);
    cta.clear();
    for (size_t i = 0; i < cost.size(); i += SW_WPB) {
        float dmax = 0.f;
        for (size_t j = i; j < std::min(cost.size(), i + SW_WPB); j++) dmax = std::max(dmax, cost[j]);
        cta.push_back(dmax);
    }
}

// makespan of CTAs dispatched in order onto `nslots` slots (every CTA goes to the slot that frees first)
static double list_schedule(const std::vector<const std::vector<float> *> &per_shot, int nb, int nslots, std::vector<double> &slot)
{
    slot.assign(nslots, 0.0);
    auto cmp = [](double x, double y) { return x > y; };
    std::make_heap(slot.begin(), slot.end(), cmp);
    double end = 0.0;
    for (int b = 0; b < nb; b++)
        for (const std::vector<float> *v : per_shot)
            for (float c : *v) {
                std::pop_heap(slot.begin(), slot.end(), cmp);
                const double t = slot.back() + c;
                slot.back() = t;
                std::push_heap(slot.begin(), slot.end(), cmp);
                end = std::max(end, t);
            }
    return end;
}

// Work list of the streaming kernels: 120-column strips x row chunks, one warp each.
//   * rows [0, nPml+2) and [nzA-nPml-2, nzA) and the strips that touch the x CPML / rim are "edge" items (slower per row:
//     CPML memory variables, no look-ahead): they get shorter chunks and are listed first so they never form the tail;
//   * interior chunks are Lz rows with (Lz + 4) a multiple of the 6-row unroll, Lz chosen so that the item count fills whole
//     waves of nSM x 8 resident warps (2 CTAs x 4 warps at 255 registers).  SEPFWI_LZ / SEPFWI_LZE override the two heights.
static int stream_plan(sepfwi_handle *h, int nb, int which /*0 fwd, 1 recon, 2 adj*/, StreamArgs &sa, cudaStream_t st,
                       std::vector<int4> *items_out = nullptr)
{
    const Dims &d = h->d;
    // same batch, same kernel: the device work list is still valid
    const uint64_t want = h->staged_sig ? (h->staged_sig ^ ((uint64_t)nb << 48)) | 1ull : 0;
    if (!items_out && want && h->plan_sig[which] == want) { sa = h->plan_sa[which]; return 0; }
    memset(&sa, 0, sizeof(sa));
    const int model = which;
    const int nStrips = (d.nx + SW_OWN - 1) / SW_OWN;
    // interior rows [zi0, zi1) / strips: the adjoint sweep keeps CPML memory on strips nPml + 2 wide (el_stress_adj.cu:67-72),
    // a warp recomputes a 2-cell halo, and the reverse sweep restores a ring that reaches 3 cells into the interior:
    // the branch-free variants need a margin of nPml + 5
    const int zi0 = d.nPml + 5, zi1 = d.nzA - d.nPml - 5;
    auto strip_inner = [&](int sx) { const int x0 = sx * SW_OWN; return x0 - 4 >= d.nPml + 3 && x0 + SW_OWN + 3 <= d.nx - d.nPml - 4; };
    // Two regimes (round 2, calibrated with tools/sweep_plan.sh on the B200):
    //  * the batch's working set stays in the 126 MB L2 (one shot of the BASELINE grids, the reference's 19-shot experiment): the
    //    launch is one to two waves of latency-bound items -- a CPML row of the adjoint sweep takes ~2.5 us against 0.4 us for an
    //    interior row -- and what counts is when the last CTA finishes.  Cost = the makespan of a list-scheduling simulation of
    //    the launch (CTAs of four items in launch order onto nSM x 2 CTA slots, item cost = rows x cost per row).
    //  * beyond L2 (batches, the 8000 x 2000 grid): HBM throughput, several waves; the round-1 wave model below.
    static const double narr[3] = {20.0, 26.0, 20.0};                 // arrays a launch touches per cell
    const double ws = (double)nb * d.nzA * d.nx * 4.0 * narr[model];
    const double bl = std::min(1.0, std::max(0.0, (ws - 60.0e6) / 140.0e6));      // 0: L2-resident ... 1: HBM-bound
    int best = 8, Le = 8;
    double bestc = 1e300;
    if (bl < 0.5) {
        const int nslots = h->nSM * (h->warps_per_sm / SW_WPB);
        // (measured: below 8 interior / 2 edge rows per item the 4-row halo and the prologue only add work)
        static const int lzs[] = {8, 10, 12, 14, 16, 20, 24, 28, 32, 40, 48, 64, 96, 128};
        static const int les2[] = {2, 3, 4, 6, 8, 12, 16, 24, 32, 48, 64};
        std::vector<float> cta;
        std::vector<double> slot;
        for (int Lz : lzs)
            for (int le : les2) {
                l2_cta_costs(h, nb, model, Lz, le, cta);
                const double end = list_schedule({&cta}, nb, nslots, slot);
                if (end < bestc - 1e-9) { bestc = end; best = Lz; Le = le; }
            }
    } else {
    // HBM regime (calibrated on the 8000 x 2000 grid and on 8 - 64-shot batches of the C3 grid, tools/sweep_plan.sh): all resident
    // warps share the memory bandwidth, so the launch takes (total row-work) / (warp slots) -- plus what a partly filled last wave
    // loses: measured, a last wave filled to a fraction f of the CTA slots costs g(f) = min(1.5 f, 0.5 + 0.5 f) of a full one
    // (f = 0.11: 0.10, 0.28: 0.51, 0.45: 0.69, 0.8: 0.85 on the reconstruction kernel), because the memory system is not
    // saturated by few warps.  An edge row costs edge_thr interior rows in throughput and, as a lone
    // item under a saturated memory system, edge_lat in latency (adjoint sweep: 37-row edge items took 290 us against 205 for 12).
    static const double edge_thr[3] = {1.15, 1.1, 1.8}, edge_lat0[3] = {2.0, 1.6, 2.5}, edge_lat[3] = {2.5, 2.0, 5.0};
    const int nslots = h->nSM * (h->warps_per_sm / SW_WPB);
    static const int les[] = {2, 3, 4, 6, 8, 12, 16, 24, 32, 48, 64, 96, 128};
    auto pieces = [&](int z0, int z1, int L, bool xlive, double &n, double &rows_sum) {       // live items of one strip segment
        const int np = (z1 - z0 + L - 1) / L;
        for (int c = 0; c < np; c++) {
            const int a0 = z0 + (int)((long long)(z1 - z0) * c / np), a1 = z0 + (int)((long long)(z1 - z0) * (c + 1) / np);
            int rows = a1 - a0;
            if (model == 1) { rows = xlive ? std::max(0, std::min(a1, d.z1 + 3) - std::max(a0, d.nPml - 2)) : 0; if (rows == 0) continue; }
            n += 1.0; rows_sum += rows;
        }
    };
    for (int Lz = 8; Lz <= 64; Lz += 2)       // (taller chunks measured worse: 98 rows +5 - 10 % on the 8000 x 2000 grid)
        for (int le : les) {
            double nE = 0, rE = 0, nI = 0, rI = 0;
            for (int sx = 0; sx < nStrips; sx++) {
                const int x0 = sx * SW_OWN;
                const bool xlive = !(x0 + SW_OWN <= d.nPml - 2 || x0 > d.x1 + 2);
                pieces(0, zi0, le, xlive, nE, rE); pieces(zi1, d.nzA, le, xlive, nE, rE);
                if (strip_inner(sx)) pieces(zi0, zi1, Lz, xlive, nI, rI); else pieces(zi0, zi1, le, xlive, nE, rE);
            }
            if (nE + nI == 0) continue;
            const double tI = 2 * ((Lz + 4 + 1) / 2) + 3.0, tE = le + 4 + 3.0;       // rows an item marches (+3: prologue)
            const double work = (rE + nE * 7.0) * edge_thr[model] + (rI + nI * (tI - Lz));
            const double ctas = ceil((nE + nI) / SW_WPB) * nb, wv = ctas / nslots, f = wv - floor(wv);
            const double g = std::min(1.5 * f, 0.5 + 0.5 * f), tail_item = nI > 0 ? tI : tE * edge_thr[model];
            // + half of the longest item: the items of the last shot(s) start late and finish unevenly
            const double longest = std::max(nI > 0 ? tI : 0.0, tE * edge_lat0[model]);
            double c = work * nb / ((double)nslots * SW_WPB) + (g - f) * tail_item + 0.5 * longest;
            c = std::max(c, std::max(nI > 0 ? tI : 0.0, tE * edge_lat[model]));      // never shorter than its longest item
            if (c < bestc - 1e-9) { bestc = c; best = Lz; Le = le; }
        }
    }
    if (h->pair_lz[which] >= 1) { best = h->pair_lz[which]; Le = h->pair_le[which]; }      // chosen jointly (stream_plan_pair)
    if (h->tune_lz >= 1) best = h->tune_lz;
    if (h->tune_lze >= 1) Le = h->tune_lze;
    sa.force = h->tune_force;
    if (h->plan_debug) fprintf(stderr, "stream_plan[%d]: nb %d Lz %d Le %d cost %.1f\n", which, nb, best, Le, bestc);
    std::vector<int4> edge, inner;
    auto split = [&](int z0, int z1, int L, int sx, bool is_edge) {
        const int n = (z1 - z0 + L - 1) / L;
        for (int c = 0; c < n; c++) {       // equal pieces
            const int a0 = z0 + (int)((long long)(z1 - z0) * c / n), a1 = z0 + (int)((long long)(z1 - z0) * (c + 1) / n);
            (is_edge ? edge : inner).push_back(make_int4(sx * SW_OWN, a0, a1, is_edge ? 1 : 0));
        }
    };
    // adjoint sweep: strips that hold residual-injection targets (a vertical fiber puts targets on every row of one strip) are
    // slower per row (staging + dependent table loads): half-height chunks, listed first, so they do not form the tail
    std::vector<char> heavy(nStrips, 0);
    if (model == 2 && h->h_int)
        for (int s_ = 0; s_ < nb; s_++)
            for (int sx = 0; sx < nStrips; sx++) {
                const int *sp = h->h_int + h->o_sInjPtr + ((size_t)s_ * h->nStrips + sx) * (d.nzA + 1);
                int rows = 0;
                for (int z = zi0; z < zi1; z++) rows += sp[z + 1] > sp[z] ? 1 : 0;
                if (4 * rows >= zi1 - zi0) heavy[sx] = 1;      // targets on a quarter of the interior rows or more
            }
    for (int sx = 0; sx < nStrips; sx++) { split(0, zi0, Le, sx, true); split(zi1, d.nzA, Le, sx, true); }
    for (int sx = 0; sx < nStrips; sx++) {
        if (strip_inner(sx)) split(zi0, zi1, heavy[sx] ? std::max(8, best / 2) : best, sx, false);
        else split(zi0, zi1, Le, sx, true);
    }
    // interior items: chunk-major so that concurrently running warps read neighbouring strips of the same rows; heavy strips first
    std::stable_sort(inner.begin(), inner.end(), [&](const int4 &p, const int4 &q) {
        const int hp = heavy[p.x / SW_OWN] ? 0 : 1, hq = heavy[q.x / SW_OWN] ? 0 : 1;
        return hp != hq ? hp < hq : p.y < q.y; });
    if (which == 1) {
        // the reverse sweep only touches interior + ring = [nPml-2, z1+2] x [nPml-2, x1+2]: items outside it would return at once but
        // still hold a warp of their CTA until its other warps finish -- they are not launched at all
        auto dead = [&](const int4 &w) {
            return std::max(w.y, d.nPml - 2) >= std::min(w.z, d.z1 + 3) || w.x + SW_OWN <= d.nPml - 2 || w.x > d.x1 + 2;
        };
        edge.erase(std::remove_if(edge.begin(), edge.end(), dead), edge.end());
    }
    edge.insert(edge.end(), inner.begin(), inner.end());
    if (items_out) { *items_out = edge; return 0; }      // host-only planning (sepfwi_plan_stream): nothing is uploaded
    if (edge.size() > h->work_cap[which]) {
        // kernels of earlier calls may still read the old list
        CU(cudaStreamSynchronize(st));
        if (h->work[which]) cudaFree(h->work[which]);
        if (h->work_host[which]) cudaFreeHost(h->work_host[which]);
        h->work[which] = nullptr; h->work_host[which] = nullptr; h->work_cap[which] = 0;
        const size_t cap = edge.size() + edge.size() / 4 + 64;
        CU(cudaMalloc((void **)&h->work[which], cap * sizeof(int4)));
        CU(cudaMallocHost((void **)&h->work_host[which], cap * sizeof(int4)));
        h->work_cap[which] = cap;
    } else if (h->plan_sig[which]) CU(cudaStreamSynchronize(st));      // the pinned mirror may still be in flight / the list in use
    memcpy(h->work_host[which], edge.data(), edge.size() * sizeof(int4));
    CU(cudaMemcpyAsync(h->work[which], h->work_host[which], edge.size() * sizeof(int4), cudaMemcpyHostToDevice, st));
    sa.work = h->work[which]; sa.nWork = (int)edge.size();
    h->plan_sa[which] = sa; h->plan_sig[which] = want;
    return 0;
}


// Work lists of the reverse-time step (adjoint sweep `sa`, reconstruction `sr`) and the decision whether both go into ONE launch
// (k_stream_bwd: the two sweeps of a step only read the adjoint buffer and write disjoint arrays).  Measured on the B200: one launch
// saves the inter-launch gap and lets the CTAs of one sweep fill the slots the other leaves idle in its last wave -- C3 single shot
// 54 -> 41 us per step, 19-shot reference experiment 90 -> 80, 8 shots 230 -> 211, 8000 x 2000 503 -> 485 -- but with many waves per
// launch (64 shots: 1450 -> 1565 us) the adjoint CTAs only lose: the merged launch needs the reconstruction kernel's 108 KB of
// shared memory per CTA, which leaves them a quarter of their L1.  In the latency regime the chunk heights of both sweeps are
// chosen together: makespan of the merged grid (per shot: adjoint CTAs, then reconstruction CTAs) by list scheduling.
static int stream_plan_pair(sepfwi_handle *h, int nb, StreamArgs &sa, StreamArgs &sr, bool &merged, cudaStream_t st,
                            std::vector<int4> *items_a = nullptr, std::vector<int4> *items_r = nullptr, int heights[4] = nullptr)
{
    const Dims &d = h->d;
    const uint64_t want = h->staged_sig ? (h->staged_sig ^ ((uint64_t)nb << 48)) | 1ull : 0;
    const bool cached = want && h->plan_sig[1] == want && h->plan_sig[2] == want;
    const int nslots = h->nSM * (h->warps_per_sm / SW_WPB);
    h->pair_lz[1] = h->pair_lz[2] = 0;
    if (!cached) {
        const double cells = (double)nb * d.nzA * d.nx * 4.0;
        const bool l2 = (cells * 26.0 - 60.0e6) / 140.0e6 < 0.5;      // both sweeps in the latency regime (see stream_plan)
        if (l2 && h->tune_merge != 0 && h->tune_lz < 1 && h->tune_lze < 1) {
            static const int lzs[] = {8, 12, 16, 20, 24, 32, 48, 64}, les[] = {2, 3, 4, 6, 8, 12, 16};
            constexpr int NL = 8, NE = 7;
            std::vector<std::vector<float>> ca(NL * NE), cr(NL * NE);
            for (int i = 0; i < NL; i++)
                for (int j = 0; j < NE; j++) { l2_cta_costs(h, nb, 2, lzs[i], les[j], ca[i * NE + j]); l2_cta_costs(h, nb, 1, lzs[i], les[j], cr[i * NE + j]); }
            for (auto &v : ca) v.insert(v.begin(), 0.05f);      // the CTA that writes the stf gradient
            std::vector<double> slot;
            double bestc = 1e300;
            int ba = 0, br = 0;
            for (int x = 0; x < NL * NE; x++)
                for (int y = 0; y < NL * NE; y++) {
                    // candidates with the same interior height but no interior items are duplicates: skipped by their equal cost vectors
                    const double e = list_schedule({&ca[x], &cr[y]}, nb, nslots, slot);
                    if (e < bestc - 1e-9) { bestc = e; ba = x; br = y; }
                }
            h->pair_lz[2] = lzs[ba / NE]; h->pair_le[2] = les[ba % NE];
            h->pair_lz[1] = lzs[br / NE]; h->pair_le[1] = les[br % NE];
            if (h->plan_debug) fprintf(stderr, "stream_plan_pair: nb %d adjoint Lz %d Le %d, reconstruction Lz %d Le %d, makespan %.1f\n", nb, h->pair_lz[2], h->pair_le[2], h->pair_lz[1], h->pair_le[1], bestc);
        }
    }
    if (heights) { heights[0] = h->pair_lz[2]; heights[1] = h->pair_le[2]; heights[2] = h->pair_lz[1]; heights[3] = h->pair_le[1]; }
    int rc = stream_plan(h, nb, 2, sa, st, items_a);
    if (!rc) rc = stream_plan(h, nb, 1, sr, st, items_r);
    h->pair_lz[1] = h->pair_lz[2] = 0;
    if (rc) return rc;
    if (items_a) { sa.nWork = (int)items_a->size(); sr.nWork = (int)items_r->size(); }      // host-only planning: nothing was uploaded
    const double waves = (double)nb * (1 + (sa.nWork + SW_WPB - 1) / SW_WPB + (sr.nWork + SW_WPB - 1) / SW_WPB) / nslots;
    merged = h->tune_merge < 0 ? waves <= 6.5 : h->tune_merge != 0;
    return 0;
}

// ------------------------------------------------------------------------------------------
// Shared-memory-resident forward loop (kernels_resident.cuh): tiling plan, tables, cooperative launch.
struct ResPlan { int rpt = 0, ntx = 0, ntz = 0, orows = 0, per_launch = 0; };
static const int RES_UNAVAILABLE = 12345;   // run_resident: the cooperative launch was refused; nothing has run
static int resident_check(sepfwi_handle *h);

template <int RPT>
static cudaError_t launch_resident(sepfwi_handle *h, dim3 grid, const KArgs &a, const ResArgs &ra, cudaStream_t st)
{
    static_assert(RPT < 16, "res_attr_done is indexed by RPT");
    if (!h->res_attr_done[RPT]) {      // per handle (one handle = one device, one thread at a time): no shared mutable state
        cudaError_t e = cudaFuncSetAttribute(k_resident_fwd<RPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_smem_bytes(RPT));
        if (e != cudaSuccess) return e;
        h->res_attr_done[RPT] = true;
    }
    void *args[2] = {(void *)&a, (void *)&ra};
    return cudaLaunchCooperativeKernel((void *)k_resident_fwd<RPT>, grid, dim3(RS_NT), args, rs_smem_bytes(RPT), st);
}

static const int k_res_rpts[] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13};

// Pick rows-per-thread (tile height 8 RPT - 8) so that the tiles of as many shots as possible are co-resident.
// Every tile may hold the CPML memory of one z side and one x side only (RS_PW rows / columns each).
static ResPlan resident_plan(const sepfwi_handle *h, int nb)
{
    ResPlan best;
    if (!h->resident || nb < 1) return best;
    const Dims &d = h->d;
    const int ntx = (d.nx + RS_OW - 1) / RS_OW;
    for (int t = 0; t < ntx; t++) {
        const int lo = std::max(t * RS_OW - 4, 0), hi = std::min(t * RS_OW + RS_EW - 5, d.nx - 1);
        if (lo < d.nPml && hi > d.nx - d.nPml - 1) return best;
    }
    const int forced = h->res_forced_rpt;
    double bestc = 1e300;
    for (int rpt : k_res_rpts) {
        if (forced && rpt != forced) continue;
        const int ORmax = RS_NG * rpt - 8, ER = RS_NG * rpt, ntz = (d.nzA + ORmax - 1) / ORmax;
        const int OR = (d.nzA + ntz - 1) / ntz;      // equal tiles
        const int ctas = ntx * ntz;
        if (ctas > h->nSM || OR < 4 || rs_smem_bytes(rpt) > h->smem_optin) continue;
        bool ok = true;
        for (int t = 0; t < ntz && ok; t++) {
            const int lo = std::max(t * OR - 4, 0), hi = std::min(t * OR - 4 + ER - 1, d.nzA - 1);
            if (lo < d.nPml && hi > d.nzA - d.nPml - 1) ok = false;
        }
        if (!ok) continue;
        const int per = std::min(nb, h->nSM / ctas);
        const int launches = (nb + per - 1) / per;
        // measured us per time step of one launch (B200): exchange + barriers ~1.2, ~0.6 per row of a thread, more once the
        // coefficient registers spill (rpt > 11)
        const double c = launches * (1.2 + 0.6 * rpt + (rpt > 11 ? 2.0 * (rpt - 11) : 0.0));
        if (c < bestc) { bestc = c; best.rpt = rpt; best.ntx = ntx; best.ntz = ntz; best.orows = OR; best.per_launch = per; }
    }
    // the launch-per-step streaming kernels (measured): latency floor 10 us per step, up to 30 us when the grid is mostly CPML
    // (the edge warps' longer rows), ~7e10 interior and ~2.5e10 CPML cell-updates/s once a batch fills the SMs
    const double cells = (double)d.nzA * d.nx;
    const double inner = (double)std::max(0, d.nzA - 2 * (d.nPml + 5)) * std::max(0, d.nx - 2 * (d.nPml + 5));
    const double t_stream = std::max(10.0 + 20.0 * (cells - inner) / cells, 1e6 * nb * (inner / 7.0e10 + (cells - inner) / 2.5e10));
    if (best.rpt && !forced && bestc >= t_stream) best = ResPlan();
    if (h->plan_debug) fprintf(stderr, "resident_plan: nb %d -> rpt %d, %d x %d tiles of %d x %d, %d shots per launch\n", nb, best.rpt, best.ntx, best.ntz, best.orows, RS_OW, best.per_launch);
    return best;
}

// The tiling `sepfwi_forward` would use for `nshots` concurrent shots on a device with `nsm` SMs and `smem_optin` bytes of
// opt-in shared memory per block -- pure host arithmetic (no CUDA call), exposed so that the planner can be tested without a GPU.
// out = {rows per thread (0: streaming kernels), tiles in x, tiles in z, own rows per tile, shots per launch}
extern "C" int sepfwi_plan_resident(const sepfwi_params *pp, int nshots, int nsm, size_t smem_optin, int out[5])
{
    if (!pp || !out || nshots < 1 || nsm < 1) return fail(SEPFWI_EINVAL, "bad argument");
    sepfwi_handle h;
    h.p = *pp;
    int rc = fill_dims(*pp, h.d);
    if (rc) return rc;
    h.nSM = nsm; h.smem_optin = smem_optin;
    h.sponge = pp->flavour == SEPFWI_FLAVOUR_SPONGE;
    h.stream = !h.sponge && (pp->kernels == 0 || pp->kernels == 3);
    h.resident = h.stream && pp->kernels == 0 && h.d.nPml <= RS_PW;
    const ResPlan pl = resident_plan(&h, nshots);
    out[0] = pl.rpt; out[1] = pl.ntx; out[2] = pl.ntz; out[3] = pl.orows; out[4] = pl.per_launch;
    return 0;
}

// Work list of the streaming kernel `which` (0 forward, 1 reconstruction, 2 adjoint) for `nshots` concurrent shots on a device
// with `nsm` SMs -- host arithmetic only.  items: up to `cap` entries {first owned column, first row, end row, 1 if edge item};
// *n receives the number of items the plan has (may exceed cap).
extern "C" int sepfwi_plan_stream(const sepfwi_params *pp, int nshots, int nsm, int which, int *items, int cap, int *n)
{
    if (!pp || !n || nshots < 1 || nsm < 1 || which < 0 || which > 2 || (cap > 0 && !items)) return fail(SEPFWI_EINVAL, "bad argument");
    sepfwi_handle h;
    h.p = *pp;
    int rc = fill_dims(*pp, h.d);
    if (rc) return rc;
    h.nSM = nsm;
    StreamArgs sa;
    std::vector<int4> v;
    rc = stream_plan(&h, nshots, which, sa, nullptr, &v);
    if (rc) return rc;
    *n = (int)v.size();
    for (int i = 0; i < (int)v.size() && i < cap; i++) { items[4 * i] = v[i].x; items[4 * i + 1] = v[i].y; items[4 * i + 2] = v[i].z; items[4 * i + 3] = v[i].w; }
    return 0;
}

// The plan of the reverse-time step `sepfwi_gradient` would use for `nshots` concurrent shots on a device with `nsm` SMs -- pure
// host arithmetic, for tests.  out = {1 if reconstruction and adjoint sweep share one launch, jointly chosen chunk heights
// (adjoint interior, adjoint edge, reconstruction interior, reconstruction edge; 0 = each kernel planned on its own),
// adjoint work items per shot, reconstruction work items per shot}
extern "C" int sepfwi_plan_backward(const sepfwi_params *pp, int nshots, int nsm, int out[7])
{
    if (!pp || !out || nshots < 1 || nsm < 1) return fail(SEPFWI_EINVAL, "bad argument");
    sepfwi_handle h;
    h.p = *pp;
    int rc = fill_dims(*pp, h.d);
    if (rc) return rc;
    h.nSM = nsm;
    StreamArgs sa, sr;
    std::vector<int4> va, vr;
    bool merged = false;
    int hts[4] = {0, 0, 0, 0};
    rc = stream_plan_pair(&h, nshots, sa, sr, merged, nullptr, &va, &vr, hts);
    if (rc) return rc;
    out[0] = merged ? 1 : 0;
    for (int i = 0; i < 4; i++) out[1 + i] = hts[i];
    out[5] = (int)va.size(); out[6] = (int)vr.size();
    return 0;
}

static void ring_cell_host(const Dims &d, int idx, int &z, int &x)       // inverse of the ring map, as kernels_base.cuh ring_cell
{
    const int L = 5;
    if (idx < L * d.nzB) { int j = idx / d.nzB, i = idx - j * d.nzB; z = i + d.nPml - 2; x = j + d.nPml - 2; }
    else if (idx < 2 * L * d.nzB) { int q = idx - L * d.nzB; int j = q / d.nzB, i = q - j * d.nzB; z = i + d.nPml - 2; x = d.nx - d.nPml - j + 1; }
    else if (idx < L * (2 * d.nzB + d.nxB)) { int q = idx - 2 * L * d.nzB; int i = q / d.nxB, j = q - i * d.nxB; z = i + d.nPml - 2; x = j + d.nPml - 2; }
    else { int q = idx - L * (2 * d.nzB + d.nxB); int i = q / d.nxB, j = q - i * d.nxB; z = d.nzA - d.nPml - i + 1; x = j + d.nPml - 2; }
}

// Whole forward time loop of slots [s0, s0+n) in one cooperative launch.  Expects the state / trace arenas zeroed.
static int run_resident(sepfwi_handle *h, const ResPlan &pl, int s0, int n, int mask, bool save_ring, cudaStream_t st)
{
    const Dims &d = h->d;
    const int nT = pl.ntx * pl.ntz, OR = pl.orows;
    int *hp = h->res_host;
    if (s0 > 0) {       // the pinned tables below are still being read by the previous launch's copy
        CU(cudaStreamSynchronize(st));
        int rc = resident_check(h);
        if (rc) return rc;
    }
    // receivers by tile
    for (int q = 0; q < n; q++) {
        const int s = s0 + q, nrec = h->h_int[h->o_nrec + s];
        const int *zr = h->h_int + h->o_zrec + (size_t)s * d.maxRec, *xr = h->h_int + h->o_xrec + (size_t)s * d.maxRec;
        int *tp = hp + h->ro_tptr + (size_t)q * (nT + 1), *tr = hp + h->ro_trec + (size_t)q * d.maxRec;
        std::fill(tp, tp + nT + 1, 0);
        for (int r = 0; r < nrec; r++) tp[(zr[r] / OR) * pl.ntx + xr[r] / RS_OW + 1]++;
        for (int t = 0; t < nT; t++) tp[t + 1] += tp[t];
        std::vector<int> cur(tp, tp + nT);
        for (int r = 0; r < nrec; r++) tr[cur[(zr[r] / OR) * pl.ntx + xr[r] / RS_OW]++] = r;
    }
    if (save_ring && h->res_ring_rpt != pl.rpt * 1024 + pl.orows) {
        std::vector<int> cnt(nT + 1, 0);
        std::vector<int2> ent(d.ringLen);
        std::vector<int> own(d.ringLen);
        for (int idx = 0; idx < d.ringLen; idx++) {
            int z, x; ring_cell_host(d, idx, z, x);
            own[idx] = (z / OR) * pl.ntx + x / RS_OW;
            cnt[own[idx] + 1]++;
        }
        for (int t = 0; t < nT; t++) cnt[t + 1] += cnt[t];
        std::vector<int> cur(cnt.begin(), cnt.end() - 1);
        for (int idx = 0; idx < d.ringLen; idx++) {
            int z, x; ring_cell_host(d, idx, z, x);
            const int t = own[idx], tz0 = (t / pl.ntx) * OR, tx0 = (t % pl.ntx) * RS_OW;
            ent[cur[t]++] = make_int2(idx, (z - tz0 + 4) * RS_EW + (x - tx0 + 4));
        }
        memcpy(hp + h->ro_rptr, cnt.data(), (size_t)(nT + 1) * sizeof(int));
        CU(cudaMemcpyAsync(h->res_ring, ent.data(), (size_t)d.ringLen * sizeof(int2), cudaMemcpyHostToDevice, st));
        CU(cudaStreamSynchronize(st));      // `ent` is a pageable temporary
        h->res_ring_rpt = pl.rpt * 1024 + pl.orows;
    }
    // inboxes: zero, except the slots of neighbours that do not exist; then the error flag
    for (int q = 0; q < n; q++)
        for (int t = 0; t < nT; t++) {
            int *box = hp + h->ro_flags + ((size_t)q * nT + t) * RS_FLAGW;
            std::fill(box, box + RS_FLAGW, 0);
            for (int k = 0; k < 9; k++) {
                const int nz_ = t / pl.ntx + k / 3 - 1, nx_ = t % pl.ntx + k % 3 - 1;
                if (k != 4 && !(nz_ >= 0 && nz_ < pl.ntz && nx_ >= 0 && nx_ < pl.ntx)) box[k] = 0x7fffffff;
            }
        }
    hp[h->ro_err] = 0;
    CU(cudaMemcpyAsync(h->res_dev, hp, h->res_n * sizeof(int), cudaMemcpyHostToDevice, st));
    if (save_ring)      // sigma(0) = 0: the stress rings of time 0 are never written by the kernel
        for (int q = 0; q < n; q++)
            for (int f = F_SZZ; f <= F_SXX; f++)
                CU(cudaMemsetAsync(h->ring + (((size_t)(s0 + q) * NFIELD + f) * d.nSteps) * d.ringLen, 0, (size_t)d.ringLen * sizeof(float), st));
    KArgs a = kargs(h);
    ResArgs ra;
    memset(&ra, 0, sizeof(ra));
    ra.ntx = pl.ntx; ra.ntz = pl.ntz; ra.orows = pl.orows; ra.slot0 = s0; ra.mask = mask; ra.fiber = h->p.fiber; ra.save_ring = save_ring ? 1 : 0;
    ra.dbg = h->res_dbg;
    ra.flags = h->res_dev + h->ro_flags; ra.err = h->res_dev + h->ro_err;
    ra.tilePtr = h->res_dev + h->ro_tptr; ra.tileRec = h->res_dev + h->ro_trec;
    ra.ringPtr = h->res_dev + h->ro_rptr; ra.ringEnt = h->res_ring;
    const dim3 grid(nT, n);
    cudaError_t e = cudaErrorInvalidValue;
    if (h->res_fake_refuse) e = cudaErrorCooperativeLaunchTooLarge;     // test hook for the fallback below
    else
    switch (pl.rpt) {
    case 3: e = launch_resident<3>(h, grid, a, ra, st); break;
    case 4: e = launch_resident<4>(h, grid, a, ra, st); break;
    case 5: e = launch_resident<5>(h, grid, a, ra, st); break;
    case 6: e = launch_resident<6>(h, grid, a, ra, st); break;
    case 7: e = launch_resident<7>(h, grid, a, ra, st); break;
    case 8: e = launch_resident<8>(h, grid, a, ra, st); break;
    case 9: e = launch_resident<9>(h, grid, a, ra, st); break;
    case 10: e = launch_resident<10>(h, grid, a, ra, st); break;
    case 11: e = launch_resident<11>(h, grid, a, ra, st); break;
    case 12: e = launch_resident<12>(h, grid, a, ra, st); break;
    case 13: e = launch_resident<13>(h, grid, a, ra, st); break;
    }
    if (e != cudaSuccess) {
        // the device cannot co-schedule the tiles (fewer SMs than planned for, a partitioned GPU, ...): not an error of the
        // call -- the caller falls back to the launch-per-step streaming kernels and stops planning resident launches
        cudaGetLastError();
        fail(SEPFWI_ECUDA, "resident forward kernel (RPT %d, %d x %d tiles, %d shots): %s", pl.rpt, pl.ntx, pl.ntz, n, cudaGetErrorString(e));
        return RES_UNAVAILABLE;
    }
    CU(cudaMemcpyAsync(h->res_errh, h->res_dev + h->ro_err, sizeof(int), cudaMemcpyDeviceToHost, st));
    h->res_used++;
    return 0;
}

// after the stream has been synchronised: did a tile of the resident kernel give up waiting for a neighbour?
static int resident_check(sepfwi_handle *h)
{
    if (h->res_errh && *h->res_errh) { *h->res_errh = 0; return fail(SEPFWI_ECUDA, "resident forward kernel: a neighbour tile never arrived (co-residency lost)"); }
    return 0;
}

// Forward time loop of one batch.  CPML flavour: libCUFD.cu:268-332; sponge: elasticSolver.py:241-276.
static int run_forward(sepfwi_handle *h, int nb, int mrec, int mask, bool save_ring, cudaStream_t st)
{
    const Dims &d = h->d;
    KArgs a = kargs(h);
    const size_t slot_stride = d.sstride * sizeof(float);
    const size_t live = (size_t)NSTATE_FWD * d.fsz * sizeof(float);
    for (int s = 0; s < nb; s++)
        CU(cudaMemsetAsync((char *)h->state + s * slot_stride, 0, live, st));
    CU(cudaMemsetAsync(h->trace, 0, (size_t)nb * d.nTrace * d.maxRec * d.nSteps * sizeof(float), st));
    dim3 blk(BX, BY), grd((d.nx + BX - 1) / BX, (d.nzA + BY - 1) / BY, nb);
    dim3 rgrd((mrec + 127) / 128, nb), ringgrd((d.ringLen + 255) / 256, nb);
    CU(cudaEventRecord(h->ev[0], st));
    ResPlan rpl = resident_plan(h, nb);
    if (rpl.rpt) {
        // the whole time loop of `per_launch` shots per cooperative launch, tiles resident in shared memory
        for (int s0 = 0; s0 < nb; s0 += rpl.per_launch) {
            const bool pr = h->prof_steps > 0;
            int rc = 0;
            LAUNCH(h, SEPFWI_K_RESIDENT_FWD, pr, st, (rc = run_resident(h, rpl, s0, std::min(rpl.per_launch, nb - s0), mrec > 0 ? mask : 0, save_ring, st)));
            if (rc == RES_UNAVAILABLE && s0 == 0) { h->resident = false; rpl = ResPlan(); break; }
            if (rc) return rc == RES_UNAVAILABLE ? SEPFWI_ECUDA : rc;
        }
    }
    if (rpl.rpt) {
    } else if (h->stream) {
        StreamArgs sa;
        int rc = stream_plan(h, nb, 0, sa, st);
        if (rc) return rc;
        sa.fiber = h->p.fiber; sa.save_ring = save_ring ? 1 : 0; sa.nrecMax = mrec; sa.tma = h->tma ? 1 : 0;
        const int items = (mrec > 0 ? mrec : 0) + (save_ring ? d.ringLen : 0);
        sa.nAux = items > 0 ? std::min(h->nSM, (items + 2 * SW_NT - 1) / (2 * SW_NT)) : 0;      // two items per thread while that stays below one CTA per SM
        dim3 sgrd(sa.nAux + (sa.nWork + SW_WPB - 1) / SW_WPB, nb);
        for (int it = 0; it <= d.nSteps - 2; it++) {
            const bool pr = it < h->prof_steps;
            sa.it = it; sa.mask = (it >= 1 && mrec > 0) ? mask : 0;
            cudaError_t le = cudaSuccess;
            LAUNCH(h, SEPFWI_K_STREAM_FWD, pr, st, (le = launch_pdl(k_stream_fwd, sgrd, dim3(SW_NT), FR_SMEM + h->smem_pad, st, h->pdl, a, sa, h->tmaps)));
            CU(le);
        }
        const int par = (d.nSteps - 1) & 1;   // buffer that holds the final state
        if (mrec > 0) LAUNCH(h, SEPFWI_K_RECORD, false, st, (k_record<false><<<rgrd, 128, 0, st>>>(a, d.nSteps - 1, mask, h->p.fiber, par)));
    } else if (!h->sponge) {
        for (int it = 0; it <= d.nSteps - 2; it++) {
            const bool pr = it < h->prof_steps;
            if (save_ring) LAUNCH(h, SEPFWI_K_RING_SAVE, pr, st, (k_ring_save<<<ringgrd, 256, 0, st>>>(a, it)));
            LAUNCH(h, SEPFWI_K_STRESS_FWD, pr, st, (k_stress_fwd<false><<<grd, blk, 0, st>>>(a, it)));
            LAUNCH(h, SEPFWI_K_VELOCITY_FWD, pr, st, (k_velocity_fwd<false><<<grd, blk, 0, st>>>(a)));
            if (mrec > 0) LAUNCH(h, SEPFWI_K_RECORD, pr, st, (k_record<false><<<rgrd, 128, 0, st>>>(a, it + 1, mask, h->p.fiber, 0)));
        }
    } else if (h->stream_sponge) {
        // sponge flavour, one launch per time step (elasticSolver.py:241-276): step `it` reads buffer it & 1; the traces of sample
        // it - 1 are written by the launch of step it from its input state, the last sample after the loop
        StreamArgs sa;
        int rc = stream_plan(h, nb, 0, sa, st);
        if (rc) return rc;
        sa.fiber = h->p.fiber; sa.nrecMax = mrec;
        sa.nAux = mrec > 0 ? std::min(h->nSM, (mrec + 2 * SW_NT - 1) / (2 * SW_NT)) : 0;
        dim3 sgrd(sa.nAux + (sa.nWork + SW_WPB - 1) / SW_WPB, nb);
        const int nzI = d.nzA - 2 * d.nPml, nxI = d.nx - 2 * d.nPml;
        for (int it = 0; it < d.nSteps; it++) {
            const bool pr = it < h->prof_steps;
            sa.it = it; sa.mask = (it >= 1 && mrec > 0) ? mask : 0;
            cudaError_t le = cudaSuccess;
            LAUNCH(h, SEPFWI_K_STREAM_FWD, pr, st, (le = launch_pdl(k_stream_sponge, sgrd, dim3(SW_NT), SP_SMEM + h->smem_pad, st, h->pdl, a, sa)));
            CU(le);
            if (h->snap_dst && it % h->snap_step == 0) {
                static const int fld[4] = {F_SXX, F_SZZ, F_VX, F_VZ};
                const int buf = ((it + 1) & 1) ? S_FWD1 : S_FWD;          // where the state after step `it` lives
                float *dst = h->snap_dst + (size_t)(it / h->snap_step) * 4 * nzI * nxI;
                for (int f = 0; f < 4; f++)
                    CU(cudaMemcpy2DAsync(dst + (size_t)f * nzI * nxI, (size_t)nxI * sizeof(float),
                                         h->state + (size_t)(buf + fld[f]) * d.fsz + (size_t)d.nPml * d.ldx + d.nPml, (size_t)d.ldx * sizeof(float),
                                         (size_t)nxI * sizeof(float), nzI,
                                         h->snap_mem == SEPFWI_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, st));
            }
        }
        if (mrec > 0) LAUNCH(h, SEPFWI_K_RECORD, false, st, (k_record<true><<<rgrd, 128, 0, st>>>(a, d.nSteps - 1, mask, h->p.fiber, d.nSteps & 1)));
    } else {
        const int nzI = d.nzA - 2 * d.nPml, nxI = d.nx - 2 * d.nPml;
        for (int it = 0; it < d.nSteps; it++) {
            const bool pr = it < h->prof_steps;
            LAUNCH(h, SEPFWI_K_VELOCITY_FWD, pr, st, (k_velocity_fwd<true><<<grd, blk, 0, st>>>(a)));
            LAUNCH(h, SEPFWI_K_STRESS_FWD, pr, st, (k_stress_fwd<true><<<grd, blk, 0, st>>>(a, it)));
            if (mrec > 0) LAUNCH(h, SEPFWI_K_RECORD, pr, st, (k_record<true><<<rgrd, 128, 0, st>>>(a, it, mask, h->p.fiber, 0)));
            if (h->snap_dst && it % h->snap_step == 0) {
                // interior of sxx, szz, vx, vz of slot 0 after step `it` (elasticSolver.py:279-284)
                static const int fld[4] = {F_SXX, F_SZZ, F_VX, F_VZ};
                float *dst = h->snap_dst + (size_t)(it / h->snap_step) * 4 * nzI * nxI;
                for (int f = 0; f < 4; f++)
                    CU(cudaMemcpy2DAsync(dst + (size_t)f * nzI * nxI, (size_t)nxI * sizeof(float),
                                         h->state + (size_t)(S_FWD + fld[f]) * d.fsz + (size_t)d.nPml * d.ldx + d.nPml, (size_t)d.ldx * sizeof(float),
                                         (size_t)nxI * sizeof(float), nzI,
                                         h->snap_mem == SEPFWI_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, st));
            }
        }
    }
    CU(cudaEventRecord(h->ev[1], st));
    CU(cudaGetLastError());
    return 0;
}

extern "C" int sepfwi_forward(sepfwi_handle *h, int nshots, const sepfwi_shot *shots, int mem, void *stream)
{
    if (!h || !shots || nshots < 0) return fail(SEPFWI_EINVAL, "bad argument");
    if (!h->have_model) return fail(SEPFWI_EINVAL, "sepfwi_set_model has not succeeded on this handle");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const Dims &d = h->d;
    const cudaMemcpyKind kind = mem == SEPFWI_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    h->fwd_ms = 0.f; h->bwd_ms = 0.f;
    for (int s0 = 0; s0 < nshots; s0 += h->B) {
        const int nb = std::min(h->B, nshots - s0);
        int mask = 0;
        for (int s = 0; s < nb; s++)
            for (int c = 0; c < NTRACE; c++)
                if (shots[s0 + s].out[c]) {
                    if (c >= d.nTrace) return fail(SEPFWI_EINVAL, "component %d is only available in the sponge flavour", c);
                    mask |= 1 << c;
                }
        int rc = stage_batch(h, nb, shots + s0, false, st);
        if (rc) return rc;
        const int mrec = max_nrec(nb, shots + s0);
        rc = run_forward(h, nb, mrec, mask, false, st);
        if (rc) return rc;
        for (int s = 0; s < nb; s++)
            for (int c = 0; c < d.nTrace; c++)
                if (shots[s0 + s].out[c] && shots[s0 + s].nrec > 0)
                    CU(cudaMemcpyAsync(shots[s0 + s].out[c], h->trace + ((size_t)s * d.nTrace + c) * d.maxRec * d.nSteps,
                                       (size_t)shots[s0 + s].nrec * d.nSteps * sizeof(float), kind, st));
        CU(cudaStreamSynchronize(st));
        prof_collect(h);
        rc = resident_check(h);
        if (rc) return rc;
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
        h->fwd_ms += ms;
    }
    return 0;
}

extern "C" int sepfwi_forward_snapshots(sepfwi_handle *h, const sepfwi_shot *shot, int save_step, float *snap, int mem, void *stream)
{
    if (!h || !shot || !snap || save_step < 1) return fail(SEPFWI_EINVAL, "bad argument");
    if (!h->sponge) return fail(SEPFWI_EINVAL, "wavefield snapshots belong to the sponge flavour (elasticSolver.forward(save_wavefield=True))");
    h->snap_dst = snap; h->snap_step = save_step; h->snap_mem = mem;
    const int rc = sepfwi_forward(h, 1, shot, mem, stream);
    h->snap_dst = nullptr;
    return rc;
}

// ------------------------------------------------------------------------------------------
// Data-side operators (kernels_data.cuh): options + scratch, and the per-shot chain of libCUFD.cu:353-457.
extern "C" int sepfwi_set_data_options(sepfwi_handle *h, const sepfwi_data_options *o)
{
    if (!h) return fail(SEPFWI_EINVAL, "null handle");
    if (h->sponge) return fail(SEPFWI_EINVAL, "the sponge flavour is forward-only");
    CU(cudaSetDevice(h->device));
    CU(cudaDeviceSynchronize());
    free_data_scratch(h);
    memset(&h->dopt, 0, sizeof(h->dopt));
    h->dopt_any = false;
    if (!o) return 0;
    sepfwi_data_options q = *o;
    if (q.win_ratio <= 0.f) q.win_ratio = 0.005f;                        // libCUFD.cu:63
    if (q.win_ratio > 0.5f) return fail(SEPFWI_EINVAL, "win_ratio must not exceed 0.5 (utilities.cu:803-806)");
    if (q.if_filter && !(q.filter[0] >= 0.f && q.filter[0] < q.filter[1] && q.filter[1] <= q.filter[2] && q.filter[2] < q.filter[3]))
        return fail(SEPFWI_EINVAL, "filter needs corner frequencies 0 <= f0 < f1 <= f2 < f3");
    h->dopt = q;
    h->dopt_any = q.if_win || q.if_filter || q.if_cross_misfit || q.if_src_update;
    if (!h->dopt_any) return 0;
    const Dims &d = h->d;
    const int nt = d.nSteps, n2 = 2 * nt, nfft = n2 / 2 + 1;
    h->d_n2 = n2; h->d_nfft = nfft;
    // twiddles (cos, sin)(2 pi j / n2) in double precision
    std::vector<float2> tw(n2);
    for (int j = 0; j < n2; j++) { const double th = 2.0 * M_PI * (double)j / (double)n2; tw[j] = make_float2((float)cos(th), (float)sin(th)); }
    // band-pass gain per bin, cuda_bp_filter1d (utilities.cu:733-765) in its float arithmetic; the band = the bins it leaves non-zero
    std::vector<float> gain;
    h->d_k0 = 0; h->d_nband = 0;
    if (q.if_filter) {
        const float PI = 3.141592653589793238462643383279502884197169f;
        const float df = (float)(1.0 / h->p.dt / n2);
        int kmin = nfft, kmax = -1;
        std::vector<float> g(nfft, 0.f);
        for (int k = 0; k < nfft; k++) {
            const float freq = k * df;
            float a;
            if (freq >= q.filter[0] && freq < q.filter[1]) a = (float)sin(PI / 2.0 * (freq - q.filter[0]) / (q.filter[1] - q.filter[0]));
            else if (freq >= q.filter[1] && freq < q.filter[2]) a = 1.0f;
            else if (freq >= q.filter[2] && freq < q.filter[3]) a = (float)cos(PI / 2.0 * (freq - q.filter[2]) / (q.filter[3] - q.filter[2]));
            else a = 0.0f;
            g[k] = a * a;
            if (g[k] != 0.f) { kmin = std::min(kmin, k); kmax = std::max(kmax, k); }
        }
        if (kmax < kmin) return fail(SEPFWI_EINVAL, "the band-pass leaves no frequency bin (df = %g Hz)", df);
        h->d_k0 = kmin; h->d_nband = kmax - kmin + 1;
        gain.assign(g.begin() + kmin, g.begin() + kmax + 1);
    }
    const size_t nbins_buf = q.if_src_update ? (size_t)nfft : (size_t)std::max(h->d_nband, 1);
    auto dalloc = [&](void **ptr, size_t bytes) { return cudaMalloc(ptr, bytes) == cudaSuccess; };
    bool ok = dalloc((void **)&h->d_tw, (size_t)n2 * sizeof(float2)) && dalloc((void **)&h->d_F[0], (size_t)d.maxRec * nbins_buf * sizeof(float2)) &&
              dalloc((void **)&h->d_nf, (size_t)3 * d.maxRec * sizeof(float)) && dalloc((void **)&h->d_win, (size_t)3 * d.maxRec * sizeof(float)) &&
              dalloc((void **)&h->d_ratio, sizeof(float)) && dalloc((void **)&h->d_mx, 2 * sizeof(unsigned)) &&
              dalloc((void **)&h->d_src, (size_t)h->B * nt * sizeof(float)) && dalloc((void **)&h->d_gain, (size_t)std::max(h->d_nband, 1) * sizeof(float));
    if (ok && q.if_src_update)
        ok = dalloc((void **)&h->d_F[1], (size_t)d.maxRec * nfft * sizeof(float2)) && dalloc((void **)&h->d_Fs, (size_t)nfft * sizeof(float2)) &&
             dalloc((void **)&h->d_coef, (size_t)nfft * sizeof(float2));
    if (ok) ok = cudaMallocHost((void **)&h->h_win, (size_t)3 * d.maxRec * h->B * sizeof(float)) == cudaSuccess &&
                 cudaMallocHost((void **)&h->h_src, (size_t)h->B * nt * sizeof(float)) == cudaSuccess;
    if (!ok) { cudaGetLastError(); free_data_scratch(h); h->dopt_any = false; return fail(SEPFWI_ENOMEM, "device scratch of the data-side operators"); }
    CU(cudaMemcpy(h->d_tw, tw.data(), (size_t)n2 * sizeof(float2), cudaMemcpyHostToDevice));
    if (q.if_filter) CU(cudaMemcpy(h->d_gain, gain.data(), gain.size() * sizeof(float), cudaMemcpyHostToDevice));
    return 0;
}

// band-pass of ntr traces in place: forward transform at the band's bins, gain, inverse (bp_filter1d, utilities.cu:1115-1168)
static void band_pass(sepfwi_handle *h, float *data, int ntr, cudaStream_t st)
{
    const int nt = h->d.nSteps;
    k_dft_fwd<<<dim3((h->d_nband + DF_NT - 1) / DF_NT, ntr), DF_NT, 0, st>>>(data, nt, h->d_n2, 0, h->d_k0, h->d_nband, h->d_tw, h->d_F[0]);
    k_dft_inv<<<dim3((nt + DF_NT - 1) / DF_NT, ntr), DF_NT, 0, st>>>(h->d_F[0], h->d_nband, h->d_k0, nullptr, h->d_gain, h->d_n2, h->d_tw,
                                                                     1.0f / (float)h->d_n2, nullptr, data, nt);
    h->launches += 2;
}

// Conditioned residual + misfit of slot s (shot sh): libCUFD.cu:353-457 on the DAS traces.  obs / cal are modified in place.
static int run_conditioning(sepfwi_handle *h, int s, const sepfwi_shot &sh, cudaStream_t st)
{
    const Dims &d = h->d;
    const sepfwi_data_options &o = h->dopt;
    const int nt = d.nSteps, ntr = sh.nrec, n2 = h->d_n2, nfft = h->d_nfft;
    const size_t cs = (size_t)d.maxRec * d.nSteps;
    float *tb = h->trace + (size_t)s * d.nTrace * cs;
    float *obs = tb + T_OBS * cs, *cal = tb + T_ETT * cs, *res = tb + T_RES * cs;
    double *part = h->partial + (size_t)s * sepfwi_handle::NBLK_RES, *mis = h->misfit + s;
    if (ntr <= 0) { CU(cudaMemsetAsync(mis, 0, sizeof(double), st)); return 0; }
    const float srcw = sh.src_weight != 0.f ? sh.src_weight : 1.0f;
    const dim3 gt((nt + 255) / 256, ntr);
    float *wst = h->d_win, *wen = h->d_win + d.maxRec, *wwt = h->d_win + 2 * (size_t)d.maxRec;
    if (o.if_win || o.if_cross_misfit) {
        // windows and trace weights of this shot: pinned staging row of the slot, one asynchronous copy
        if (o.if_win && (!sh.win_start || !sh.win_end)) return fail(SEPFWI_EINVAL, "if_win needs win_start / win_end of every shot");
        float *hw = h->h_win + (size_t)s * 3 * d.maxRec;
        for (int r = 0; r < ntr; r++) {
            hw[r] = sh.win_start ? sh.win_start[r] : 0.f; hw[d.maxRec + r] = sh.win_end ? sh.win_end[r] : 0.f;
            hw[2 * (size_t)d.maxRec + r] = sh.trace_weights ? sh.trace_weights[r] : 1.0f;
        }
        CU(cudaMemcpyAsync(h->d_win, hw, (size_t)3 * d.maxRec * sizeof(float), cudaMemcpyHostToDevice, st));
    }
    if (o.if_win) {                                                   // :353-363
        k_win_traces<<<gt, 256, 0, st>>>(obs, ntr, nt, h->p.dt, wst, wen, wwt, srcw, o.win_ratio);
        k_win_traces<<<gt, 256, 0, st>>>(cal, ntr, nt, h->p.dt, wst, wen, wwt, srcw, o.win_ratio);
    } else {                                                          // :363-367: without windows the records still get their end tapers
        k_win_traces<<<gt, 256, 0, st>>>(obs, ntr, nt, h->p.dt, nullptr, nullptr, nullptr, 1.0f, o.win_ratio);
        k_win_traces<<<gt, 256, 0, st>>>(cal, ntr, nt, h->p.dt, nullptr, nullptr, nullptr, 1.0f, o.win_ratio);
    }
    h->launches += 2;
    if (o.if_filter) { band_pass(h, obs, ntr, st); band_pass(h, cal, ntr, st); }                    // :370-373
    if (o.if_cross_misfit) { k_normfacts<<<ntr, 256, 0, st>>>(obs, cal, nt, h->d_nf, ntr); h->launches++; }   // :376-384
    if (o.if_src_update) {                                            // :387-394, source_update utilities.cu:1170-1276
        const dim3 gf((nfft + DF_NT - 1) / DF_NT, ntr), gi((nt + DF_NT - 1) / DF_NT, ntr);
        const float *amp = h->t_flt + h->o_amp + (size_t)s * d.nSteps;      // tapered source, scaled by 1500^2 dt (linear: undone on output)
        k_dft_fwd<<<gf, DF_NT, 0, st>>>(obs, nt, n2, 1, 0, nfft, h->d_tw, h->d_F[0]);
        k_dft_fwd<<<gf, DF_NT, 0, st>>>(cal, nt, n2, 1, 0, nfft, h->d_tw, h->d_F[1]);
        k_dft_fwd<<<dim3(gf.x, 1), DF_NT, 0, st>>>(amp, nt, n2, 0, 0, nfft, h->d_tw, h->d_Fs);
        k_spectrum_coef<<<nfft, 256, 0, st>>>(h->d_F[0], h->d_F[1], ntr, nfft, h->d_Fs, h->d_coef);
        k_dft_inv<<<gi, DF_NT, 0, st>>>(h->d_F[1], nfft, 0, h->d_coef, nullptr, n2, h->d_tw, 1.0f / (float)n2, nullptr, cal, nt);
        float *srcn = h->d_src + (size_t)s * nt;
        k_dft_inv<<<dim3(gi.x, 1), DF_NT, 0, st>>>(h->d_Fs, nfft, 0, nullptr, nullptr, n2, h->d_tw, 1.0f / (float)n2, nullptr, srcn, nt);
        const float unscale = (float)(1.0 / (pow(1500.0, 2) * (double)h->p.dt));
        k_scale_copy<<<(nt + 255) / 256, 256, 0, st>>>(srcn, srcn, nt, unscale);
        CU(cudaMemsetAsync(h->d_mx, 0, 2 * sizeof(unsigned), st));
        k_absmax2<<<64, 256, 0, st>>>(obs, cal, (size_t)ntr * nt, h->d_mx);        // amp_ratio_comp :1328-1356 (trace rows are contiguous)
        k_amp_ratio<<<1, 1, 0, st>>>(h->d_mx, h->d_ratio);
        h->launches += 9;
    }
    if (!o.if_cross_misfit) {                                         // :397-400
        k_residual_one<<<sepfwi_handle::NBLK_RES, 256, 0, st>>>(obs, cal, res, ntr, nt, part);
        k_sum_partials_one<<<1, 32, 0, st>>>(part, sepfwi_handle::NBLK_RES, mis);
    } else {                                                          // :401-407
        CU(cudaMemsetAsync(res, 0, (size_t)ntr * nt * sizeof(float), st));
        k_cross_misfit<<<1, 256, 0, st>>>(h->d_nf, ntr, wwt, srcw, mis);
    }
    h->launches += 2;
    if (o.if_src_update) {                                            // :430-433, source_update_adj utilities.cu:1280-1326
        k_dft_fwd<<<dim3((nfft + DF_NT - 1) / DF_NT, ntr), DF_NT, 0, st>>>(res, nt, n2, 1, 0, nfft, h->d_tw, h->d_F[0]);
        k_dft_inv<<<dim3((nt + DF_NT - 1) / DF_NT, ntr), DF_NT, 0, st>>>(h->d_F[0], nfft, 0, h->d_coef, nullptr, n2, h->d_tw, 1.0f / (float)n2,
                                                                         h->d_ratio, res, nt);
        h->launches += 2;
    }
    if (o.if_cross_misfit) { k_cross_adjoint<<<gt, 256, 0, st>>>(obs, cal, h->d_nf, ntr, nt, wwt, srcw, res); h->launches++; }   // :436-443
    if (o.if_filter) band_pass(h, res, ntr, st);                      // :446-448
    if (o.if_win) k_win_traces<<<gt, 256, 0, st>>>(res, ntr, nt, h->p.dt, wst, wen, wwt, srcw, o.win_ratio);       // :450-457
    else k_win_traces<<<gt, 256, 0, st>>>(res, ntr, nt, h->p.dt, nullptr, nullptr, nullptr, 1.0f, o.win_ratio);
    h->launches++;
    CU(cudaGetLastError());
    return 0;
}

// The data-side chain on given traces of ONE shot (no wave propagation): obs and syn [nrec][nSteps] in space `mem`; outputs (each may be
// NULL) the adjoint source `res`, the conditioned synthetic `syn_out`, the shot's misfit contribution (already times 0.5) and -- through
// shot->src_updated -- the updated source.  Uses slot 0 of the handle.
extern "C" int sepfwi_condition(sepfwi_handle *h, const sepfwi_shot *shot, const float *obs, const float *syn, float *res, float *syn_out,
                                double *misfit, int mem, void *stream)
{
    if (!h || !shot || !obs || !syn) return fail(SEPFWI_EINVAL, "null argument");
    if (h->sponge) return fail(SEPFWI_EINVAL, "the sponge flavour is forward-only");
    if (!h->dopt_any) return fail(SEPFWI_EINVAL, "no data-side option is set (sepfwi_set_data_options)");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const Dims &d = h->d;
    int rc = stage_batch(h, 1, shot, false, st);
    if (rc) return rc;
    const size_t cs = (size_t)d.maxRec * d.nSteps, n = (size_t)shot->nrec * d.nSteps * sizeof(float);
    const cudaMemcpyKind kin = mem == SEPFWI_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    const cudaMemcpyKind kout = mem == SEPFWI_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    if (n) {
        CU(cudaMemcpyAsync(h->trace + T_OBS * cs, obs, n, kin, st));
        CU(cudaMemcpyAsync(h->trace + T_ETT * cs, syn, n, kin, st));
    }
    rc = run_conditioning(h, 0, *shot, st);
    if (rc) return rc;
    if (n && res) CU(cudaMemcpyAsync(res, h->trace + T_RES * cs, n, kout, st));
    if (n && syn_out) CU(cudaMemcpyAsync(syn_out, h->trace + T_ETT * cs, n, kout, st));
    double J = 0.0;
    CU(cudaMemcpyAsync(&J, h->misfit, sizeof(double), cudaMemcpyDeviceToHost, st));
    if (h->dopt.if_src_update && shot->src_updated)
        CU(cudaMemcpyAsync(shot->src_updated, h->d_src, (size_t)d.nSteps * sizeof(float), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (misfit) *misfit = 0.5 * J;
    return 0;
}

// Backward time loop of one batch, libCUFD.cu:500-653.
static int run_backward(sepfwi_handle *h, int nb, int minj, cudaStream_t st)
{
    const Dims &d = h->d;
    KArgs a = kargs(h);
    const size_t slot_stride = d.sstride * sizeof(float);
    for (int s = 0; s < nb; s++)   // adjoint fields + adjoint memory variables restart from zero (libCUFD.cu:503-517)
        CU(cudaMemsetAsync((char *)h->state + s * slot_stride + (size_t)S_ADJ * d.fsz * sizeof(float), 0,
                           (size_t)(NSTATE - NSTATE_FWD) * d.fsz * sizeof(float), st));
    CU(cudaMemsetAsync(h->gstf, 0, (size_t)nb * d.nSteps * sizeof(float), st));
    dim3 blk(BX, BY), grd((d.nx + BX - 1) / BX, (d.nzA + BY - 1) / BY, nb);
    const int rx = d.x1 + 2 - (d.nPml - 2) + 1, rz = d.z1 + 2 - (d.nPml - 2) + 1;
    dim3 rgrd((rx + BX - 1) / BX, (rz + BY - 1) / BY, nb);
    dim3 igrd((minj + 127) / 128, nb);
    StreamArgs sa, sr;
    bool merged = false;
    if (h->stream) {
        int rc = stream_plan_pair(h, nb, sa, sr, merged, st);
        if (rc) return rc;
    }
    CU(cudaEventRecord(h->ev[2], st));
    if (h->stream) {
        int q = (d.nSteps - 1) & 1, pa = 0;     // forward state nSteps-1 sits in buffer q; adjoint starts in buffer 0
        for (int it = d.nSteps - 2; it >= 0; it--) {
            const bool pr = (d.nSteps - 2 - it) < h->prof_steps;
            cudaError_t le = cudaSuccess;
            sr.it = it; sr.q = q; sr.pa = pa;
            if (merged) {
                sa.it = it; sa.q = q; sa.pa = pa;
                const int nAdj = 1 + (sa.nWork + SW_WPB - 1) / SW_WPB, nRec = (sr.nWork + SW_WPB - 1) / SW_WPB;
                LAUNCH(h, SEPFWI_K_STREAM_BWD, pr, st, (le = launch_pdl(k_stream_bwd, dim3(nAdj + nRec, nb), dim3(SW_NT), RC_SMEM, st, h->pdl, a, sr, sa, nAdj)));
                CU(le);
                q ^= 1; pa ^= 1;
                continue;
            }
            LAUNCH(h, SEPFWI_K_STREAM_RECON, pr, st, (le = launch_pdl(k_stream_recon, dim3((sr.nWork + SW_WPB - 1) / SW_WPB, nb), dim3(SW_NT), RC_SMEM + h->smem_pad, st, h->pdl, a, sr)));
            CU(le);
            sa.it = it; sa.q = q; sa.pa = pa;
            LAUNCH(h, SEPFWI_K_STREAM_ADJ, pr, st, (le = launch_pdl(k_stream_adj, dim3(1 + (sa.nWork + SW_WPB - 1) / SW_WPB, nb), dim3(SW_NT), AR_SMEM + h->smem_pad, st, h->pdl, a, sa)));
            CU(le);
            q ^= 1; pa ^= 1;
        }
    } else
    for (int it = d.nSteps - 2; it >= 0; it--) {
        const bool pr = (d.nSteps - 2 - it) < h->prof_steps;
        LAUNCH(h, SEPFWI_K_VELOCITY_BWD, pr, st, (k_velocity_bwd<<<rgrd, blk, 0, st>>>(a, it)));
        LAUNCH(h, SEPFWI_K_STRESS_BWD, pr, st, (k_stress_bwd<<<rgrd, blk, 0, st>>>(a, it)));
        LAUNCH(h, SEPFWI_K_VELOCITY_ADJ, pr, st, (k_velocity_adj<<<grd, blk, 0, st>>>(a)));
        if (minj > 0) LAUNCH(h, SEPFWI_K_INJECT, pr, st, (k_inject<<<igrd, 128, 0, st>>>(a, it)));
        LAUNCH(h, SEPFWI_K_STRESS_ADJ, pr, st, (k_stress_adj<<<grd, blk, 0, st>>>(a)));
    }
    CU(cudaEventRecord(h->ev[3], st));
    CU(cudaGetLastError());
    return 0;
}

// Pinned staging for the small per-call results (per-shot misfit, stf gradients): grown on demand, never per call.
static int ensure_result_staging(sepfwi_handle *h, size_t n_misfit, size_t n_gstf)
{
    if (n_misfit > h->hj_cap) {
        if (h->hj) cudaFreeHost(h->hj);
        h->hj = nullptr; h->hj_cap = 0;
        CU(cudaMallocHost((void **)&h->hj, n_misfit * sizeof(double)));
        h->hj_cap = n_misfit;
    }
    if (n_gstf > h->hg_cap) {
        if (h->hg) cudaFreeHost(h->hg);
        h->hg = nullptr; h->hg_cap = 0;
        CU(cudaMallocHost((void **)&h->hg, n_gstf * sizeof(float)));
        h->hg_cap = n_gstf;
    }
    return 0;
}

extern "C" int sepfwi_gradient(sepfwi_handle *h, int nshots, const sepfwi_shot *shots, int with_adj,
                               float *misfit, float *glam, float *gmu, float *grho, int mem, void *stream)
{
    if (!h || !shots || nshots < 0) return fail(SEPFWI_EINVAL, "bad argument");
    if (!h->have_model) return fail(SEPFWI_EINVAL, "sepfwi_set_model has not succeeded on this handle");
    if (h->sponge) return fail(SEPFWI_EINVAL, "the sponge flavour is forward-only");
    if (with_adj && !h->p.with_adjoint) return fail(SEPFWI_EINVAL, "handle was created without with_adjoint");
    if (with_adj && (!glam || !gmu || !grho)) return fail(SEPFWI_EINVAL, "null gradient output");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const Dims &d = h->d;
    const cudaMemcpyKind kin = mem == SEPFWI_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    const cudaMemcpyKind kout = mem == SEPFWI_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    int rc = ensure_result_staging(h, (size_t)nshots, with_adj ? (size_t)nshots * d.nSteps : 0);
    if (rc) return rc;
    if (with_adj) CU(cudaMemsetAsync(h->grad, 0, (size_t)h->B * 3 * d.fsz * sizeof(float), st));
    h->fwd_ms = 0.f; h->bwd_ms = 0.f;
    const size_t cs = (size_t)d.maxRec * d.nSteps;
    const bool timing = h->prof_steps > 0 || nshots > h->B;      // several batches: the event pairs are reused, read them per batch
    for (int s0 = 0; s0 < nshots; s0 += h->B) {
        const int nb = std::min(h->B, nshots - s0);
        rc = stage_batch(h, nb, shots + s0, with_adj != 0, st);
        if (rc) return rc;
        const int mrec = max_nrec(nb, shots + s0);
        rc = run_forward(h, nb, mrec, 1 << T_ETT, with_adj != 0, st);
        if (rc) return rc;
        for (int s = 0; s < nb; s++) {
            const sepfwi_shot &sh = shots[s0 + s];
            if (sh.nrec > 0) {
                if (!sh.obs_ett) return fail(SEPFWI_EINVAL, "shot %d: null obs_ett", s0 + s);
                CU(cudaMemcpyAsync(h->trace + ((size_t)s * d.nTrace + T_OBS) * cs, sh.obs_ett,
                                   (size_t)sh.nrec * d.nSteps * sizeof(float), kin, st));
            }
        }
        // the synthetic traces leave before the data-side operators modify them in place
        for (int s = 0; s < nb; s++)
            if (shots[s0 + s].out[T_ETT] && shots[s0 + s].nrec > 0)
                CU(cudaMemcpyAsync(shots[s0 + s].out[T_ETT], h->trace + ((size_t)s * d.nTrace + T_ETT) * cs,
                                   (size_t)shots[s0 + s].nrec * d.nSteps * sizeof(float), kout, st));
        KArgs a = kargs(h);
        if (!h->dopt_any) {
            k_residual<<<dim3(sepfwi_handle::NBLK_RES, nb), 256, 0, st>>>(a, h->partial, sepfwi_handle::NBLK_RES);
            k_sum_partials<<<1, 32 * ((nb + 31) / 32), 0, st>>>(h->partial, sepfwi_handle::NBLK_RES, h->misfit, nb);
            h->launches += 2;
        } else {
            // shot by shot (the spectra scratch and the window table are shared by the slots; stream order keeps them apart)
            for (int s = 0; s < nb; s++) {
                rc = run_conditioning(h, s, shots[s0 + s], st);
                if (rc) return rc;
            }
            if (h->dopt.if_src_update)
                CU(cudaMemcpyAsync(h->h_src, h->d_src, (size_t)nb * d.nSteps * sizeof(float), cudaMemcpyDeviceToHost, st));
        }
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(h->hj + s0, h->misfit, nb * sizeof(double), cudaMemcpyDeviceToHost, st));
        if (with_adj) {
            int minj = 0;
            for (int s = 0; s < nb; s++) minj = std::max(minj, h->h_int[h->o_injN + s]);
            rc = run_backward(h, nb, minj, st);
            if (rc) return rc;
            // one asynchronous copy of the batch's stf gradients into pinned staging (handed out after the final sync)
            CU(cudaMemcpyAsync(h->hg + (size_t)s0 * d.nSteps, h->gstf, (size_t)nb * d.nSteps * sizeof(float), cudaMemcpyDeviceToHost, st));
        }
        if (h->dopt_any && h->dopt.if_src_update) {
            // updated source time functions of this batch: handed out per batch (the staging rows are reused)
            CU(cudaStreamSynchronize(st));
            for (int s = 0; s < nb; s++)
                if (shots[s0 + s].src_updated) memcpy(shots[s0 + s].src_updated, h->h_src + (size_t)s * d.nSteps, (size_t)d.nSteps * sizeof(float));
        }
        if (timing || s0 + h->B >= nshots) {
            // the last batch is followed by the gradient reduction below: no sync here unless timing needs the events
            if (timing) {
                CU(cudaStreamSynchronize(st));
                prof_collect(h);
                rc = resident_check(h);
                if (rc) return rc;
                float ms = 0.f;
                CU(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
                h->fwd_ms += ms;
                if (with_adj) { CU(cudaEventElapsedTime(&ms, h->ev[2], h->ev[3])); h->bwd_ms += ms; }
            }
        }
    }
    if (with_adj) {
        float *outs[3] = {glam, gmu, grho};
        dim3 blk(32, 8), grd((d.nx + 31) / 32, (d.nz + 7) / 8);
        for (int k = 0; k < 3; k++) {
            float *dst = mem == SEPFWI_MEM_DEVICE ? outs[k] : h->dense[k];
            k_grad_reduce<<<grd, blk, 0, st>>>(d, h->grad, h->B, k, dst);
            h->launches++;
            if (mem == SEPFWI_MEM_HOST)
                CU(cudaMemcpyAsync(outs[k], dst, (size_t)d.nz * d.nx * sizeof(float), cudaMemcpyDeviceToHost, st));
        }
        CU(cudaGetLastError());
    }
    // ONE synchronisation per call: the misfit is a host scalar of the C ABI, so the call cannot return earlier
    CU(cudaStreamSynchronize(st));
    if (!timing) {
        prof_collect(h);
        rc = resident_check(h);
        if (rc) return rc;
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
        h->fwd_ms = ms;
        if (with_adj) { CU(cudaEventElapsedTime(&ms, h->ev[2], h->ev[3])); h->bwd_ms = ms; }
    }
    double J = 0.0;
    for (int s = 0; s < nshots; s++) J += h->hj[s];
    if (misfit) *misfit = (float)(0.5 * J);
    h->last_misfit = 0.5 * J;
    if (with_adj)
        for (int s = 0; s < nshots; s++)
            if (shots[s].gstf) memcpy(shots[s].gstf, h->hg + (size_t)s * d.nSteps, (size_t)d.nSteps * sizeof(float));
    return 0;
}

// ------------------------------------------------------------------------------------------
// Ring save / restore of one host field (test entry points for the index map).
static int ring_io(sepfwi_handle *h, float *field, float *bnd, int restore)
{
    if (!h || !field || !bnd) return fail(SEPFWI_EINVAL, "null argument");
    CU(cudaSetDevice(h->device));
    const Dims &d = h->d;
    float *df = nullptr, *db = nullptr;
    CU(cudaMalloc((void **)&df, d.fsz * sizeof(float)));
    CU(cudaMalloc((void **)&db, (size_t)d.ringLen * sizeof(float)));
    CU(cudaMemset(df, 0, d.fsz * sizeof(float)));
    CU(cudaMemcpy2D(df, (size_t)d.ldx * sizeof(float), field, (size_t)d.nx * sizeof(float), (size_t)d.nx * sizeof(float), d.nzA, cudaMemcpyHostToDevice));
    if (restore) CU(cudaMemcpy(db, bnd, (size_t)d.ringLen * sizeof(float), cudaMemcpyHostToDevice));
    k_ring_copy_field<<<(d.ringLen + 255) / 256, 256>>>(d, df, db, restore);
    h->launches++;
    CU(cudaGetLastError());
    if (restore) CU(cudaMemcpy2D(field, (size_t)d.nx * sizeof(float), df, (size_t)d.ldx * sizeof(float), (size_t)d.nx * sizeof(float), d.nzA, cudaMemcpyDeviceToHost));
    else CU(cudaMemcpy(bnd, db, (size_t)d.ringLen * sizeof(float), cudaMemcpyDeviceToHost));
    cudaFree(df); cudaFree(db);
    return 0;
}
extern "C" int sepfwi_ring_save(sepfwi_handle *h, const float *field, float *bnd) { return ring_io(h, (float *)field, bnd, 0); }
extern "C" int sepfwi_ring_restore(sepfwi_handle *h, float *field, const float *bnd) { return ring_io(h, field, (float *)bnd, 1); }

// TEST INFRASTRUCTURE ONLY -- C entry points around the reference's OWN data-side operators, compiled by oracle/Makefile into
// oracle/_ref/libcufd_ref.so together with the reference's sources (this file is ours; nothing of the reference is copied: the
// kernels and host functions it calls are declared by the reference's Src/utilities.h and defined in Src/utilities.cu).
//
// The operators (cuda_window, bp_filter1d, cuda_find_normfact, cuda_normal_misfit, cuda_normal_adjoint_source, source_update,
// source_update_adj) are live code in the reference, but their call sites in the shot driver are commented out
// (Src/libCUFD.cu:353-457).  ref_dataops_chain() executes them in exactly the order and with exactly the launch configurations
// of those commented lines, so that oracle/dataops.py and the product's sepfwi_condition can be pinned against the reference's own
// kernels run on the GPU (tests/golden/make_dataops_golden.py, tests/test_gpu_dataops.py).
//
// All pointers are HOST pointers; traces are [nrec][nt] row-major like the reference's d_data (ip = idr * nt + idt).
#include <cuda_runtime.h>
#include <cuComplex.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "utilities.h"

namespace {

struct DevBuf {
    float *p = nullptr;
    size_t n = 0;
    explicit DevBuf(size_t count, const float *host = nullptr) : n(count) {
        cudaMalloc((void **)&p, (count ? count : 1) * sizeof(float));
        if (host) cudaMemcpy(p, host, count * sizeof(float), cudaMemcpyHostToDevice);
        else cudaMemset(p, 0, (count ? count : 1) * sizeof(float));
    }
    ~DevBuf() { cudaFree(p); }
    void get(float *host) const { if (host) cudaMemcpy(host, p, n * sizeof(float), cudaMemcpyDeviceToHost); }
};

// The reference zero-fills only the first nSteps columns of its 2 nSteps padded scratch rows (intialArrayGPU is launched with the
// unpadded grid, utilities.cu:1133, 1191-1192, 1298) and relies on fresh cudaMalloc memory for the rest.  Make that assumption true:
// park zeroed blocks of the sizes it is about to allocate in the allocator's free lists.
void prezero(size_t floats, int copies)
{
    std::vector<void *> v;
    for (int i = 0; i < copies; i++) {
        void *q = nullptr;
        if (cudaMalloc(&q, floats * sizeof(float)) == cudaSuccess) { cudaMemset(q, 0, floats * sizeof(float)); v.push_back(q); }
    }
    cudaDeviceSynchronize();
    for (void *q : v) cudaFree(q);
}

}  // namespace

extern "C" {

// cuda_window with per-trace windows (utilities.cu:790-842), launch configuration of libCUFD.cu:351,355
int ref_window_traces(int nt, int nrec, float dt, const float *win_start, const float *win_end, const float *weights, float src_weight,
                      float ratio, float *data)
{
    DevBuf d(size_t(nt) * nrec, data), a(nrec, win_start), b(nrec, win_end), w(nrec, weights);
    dim3 threads(TX, TY), blocks((nt + TX - 1) / TX, (nrec + TY - 1) / TY);
    cuda_window<<<blocks, threads>>>(nt, nrec, dt, a.p, b.p, w.p, src_weight, ratio, d.p);
    int rc = (int)cudaDeviceSynchronize();
    d.get(data);
    return rc;
}

// cuda_window without windows (utilities.cu:844-884)
int ref_window_simple(int nt, int nrec, float dt, float ratio, float *data)
{
    DevBuf d(size_t(nt) * nrec, data);
    dim3 threads(TX, TY), blocks((nt + TX - 1) / TX, (nrec + TY - 1) / TY);
    cuda_window<<<blocks, threads>>>(nt, nrec, dt, ratio, d.p);
    int rc = (int)cudaDeviceSynchronize();
    d.get(data);
    return rc;
}

// bp_filter1d (utilities.cu:1115-1168)
int ref_bp_filter(int nt, float dt, int nrec, float *data, const float *filter4)
{
    DevBuf d(size_t(nt) * nrec, data);
    float f[4] = {filter4[0], filter4[1], filter4[2], filter4[3]};
    prezero(size_t(2) * nt * nrec, 1);
    bp_filter1d(nt, dt, nrec, d.p, f);
    int rc = (int)cudaDeviceSynchronize();
    d.get(data);
    return rc;
}

// cuda_find_normfact (utilities.cu:1011-1040), launch configuration of libCUFD.cu:378
int ref_normfact(int nt, int nrec, const float *a, const float *b, float *out)
{
    DevBuf da(size_t(nt) * nrec, a), db(size_t(nt) * nrec, b), o(nrec);
    cuda_find_normfact<<<nrec, 512>>>(nt, nrec, da.p, db.p, o.p);
    int rc = (int)cudaDeviceSynchronize();
    o.get(out);
    return rc;
}

// The whole data-side section of one shot in the order of Src/libCUFD.cu:351-457 with its comments removed.
//   obs, cal [nrec][nt] and src [nt] are conditioned in place; res [nrec][nt] receives the adjoint source;
//   misfit = what the driver adds to h_l2Obj for this shot (before the final 0.5, libCUFD.cu:776).
int ref_dataops_chain(int nt, int nrec, float dt, float *obs, float *cal, float *src, int if_win, const float *win_start,
                      const float *win_end, const float *weights, float src_weight, float win_ratio, int if_filter, const float *filter4,
                      int if_cross_misfit, int if_src_update, float *res, float *misfit, float *amp_ratio_out)
{
    const size_t n = size_t(nt) * nrec;
    DevBuf d_obs(n, obs), d_cal(n, cal), d_res(n), d_src(nt, src), ws(nrec, win_start), we(nrec, win_end), wt(nrec, weights);
    DevBuf d_obs_nf(nrec), d_cal_nf(nrec), d_cross_nf(nrec), d_obj(1);
    cuFloatComplex *d_coef = nullptr;
    cudaMalloc((void **)&d_coef, sizeof(cuFloatComplex) * (nt + 1));
    cudaMemset(d_coef, 0, sizeof(cuFloatComplex) * (nt + 1));
    float filt[4] = {0.f, 0.f, 0.f, 0.f};
    if (filter4) for (int i = 0; i < 4; i++) filt[i] = filter4[i];
    float amp_ratio = 1.0f;
    dim3 threads(TX, TY);                                              // libCUFD.cu:92
    dim3 blocksT((nt + TX - 1) / TX, (nrec + TY - 1) / TY);            // :351

    if (if_win) {                                                      // :354-363
        cuda_window<<<blocksT, threads>>>(nt, nrec, dt, ws.p, we.p, wt.p, src_weight, win_ratio, d_obs.p);
        cuda_window<<<blocksT, threads>>>(nt, nrec, dt, ws.p, we.p, wt.p, src_weight, win_ratio, d_cal.p);
    } else {                                                           // :363-367
        cuda_window<<<blocksT, threads>>>(nt, nrec, dt, win_ratio, d_obs.p);
        cuda_window<<<blocksT, threads>>>(nt, nrec, dt, win_ratio, d_cal.p);
    }
    if (if_filter) {                                                   // :371-374
        prezero(2 * n, 1); bp_filter1d(nt, dt, nrec, d_obs.p, filt);
        prezero(2 * n, 1); bp_filter1d(nt, dt, nrec, d_cal.p, filt);
    }
    if (if_cross_misfit) {                                             // :377-384
        cuda_find_normfact<<<nrec, 512>>>(nt, nrec, d_obs.p, d_obs.p, d_obs_nf.p);
        cuda_find_normfact<<<nrec, 512>>>(nt, nrec, d_cal.p, d_cal.p, d_cal_nf.p);
        cuda_find_normfact<<<nrec, 512>>>(nt, nrec, d_obs.p, d_cal.p, d_cross_nf.p);
    }
    if (if_src_update) {                                               // :388-393
        prezero(2 * n, 2);
        amp_ratio = source_update(nt, dt, nrec, d_obs.p, d_cal.p, d_src.p, d_coef);
    }
    if (!if_cross_misfit) {                                            // :398-401
        gpuMinus<<<blocksT, threads>>>(d_res.p, d_obs.p, d_cal.p, nt, nrec);
        cuda_cal_objective<<<1, 512>>>(d_obj.p, d_res.p, nt * nrec);
    } else {                                                           // :402-407
        cuda_normal_misfit<<<1, 512>>>(nrec, d_cross_nf.p, d_obs_nf.p, d_cal_nf.p, d_obj.p, wt.p, src_weight);
    }
    if (if_src_update) {                                               // :430-433
        prezero(2 * n, 1);
        source_update_adj(nt, dt, nrec, d_res.p, amp_ratio, d_coef);
    }
    if (if_cross_misfit)                                               // :437-442
        cuda_normal_adjoint_source<<<blocksT, threads>>>(nt, nrec, d_obs_nf.p, d_cal_nf.p, d_cross_nf.p, d_obs.p, d_cal.p, d_res.p, wt.p,
                                                         src_weight);
    if (if_filter) { prezero(2 * n, 1); bp_filter1d(nt, dt, nrec, d_res.p, filt); }     // :445-447
    if (if_win) cuda_window<<<blocksT, threads>>>(nt, nrec, dt, ws.p, we.p, wt.p, src_weight, win_ratio, d_res.p);      // :450-454
    else cuda_window<<<blocksT, threads>>>(nt, nrec, dt, win_ratio, d_res.p);                                          // :454-457
    int rc = (int)cudaDeviceSynchronize();
    d_obs.get(obs); d_cal.get(cal); d_src.get(src); d_res.get(res); d_obj.get(misfit);
    if (amp_ratio_out) *amp_ratio_out = amp_ratio;
    cudaFree(d_coef);
    return rc;
}

}  // extern "C"

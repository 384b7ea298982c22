/*
 * sepfwi.h -- C ABI of libsepfwi.so: the B200-native (sm_100a) replacement for the
 * data-parallel hot path of seisfwi/SEP-2023 (2-D isotropic elastic velocity-stress
 * FD propagator with CPML, its adjoint, and the FWI gradient built on it).
 *
 * Reference interfaces replaced (paths under DAS_Waveform_Inversion/Ops/FWI/Src/):
 *   sepfwi_cufd      <- extern "C" cufd(...)                    libCUFD.h:6-10, libCUFD.cu:32-820
 *                       (same argument list; `const std::string para_fname` becomes `const char*`)
 *   sepfwi_create    <- Parameter / Cpml / Bnd constructors     Parameter.cpp:17-178, Cpml.cu:7-118,
 *                                                               Boundary.cu:5-43 (done once, not per call)
 *   sepfwi_set_model <- Model constructor + host transposes     Model.cu:15-93, libCUFD.cu:67-77
 *   sepfwi_forward   <- cufd(calc_id = 2) shot loop             libCUFD.cu:170-347
 *   sepfwi_forward_snapshots <- elasticSolver.forward_it(isrc, True)   DAS_Waveform_Modeling/src/elasticSolver.py:185-305
 *   sepfwi_gradient  <- cufd(calc_id = 0 / 1) shot loop         libCUFD.cu:170-724, 775-780
 *   sepfwi_ring_*    <- Bnd::field_from_bnd / field_to_bnd      Boundary.cu:55-101, utilities.cu:362-425
 *   sepfwi_set_data_options <- if_win / filter / if_cross_misfit / if_src_update   Parameter.cpp:139-176, utilities.cu:733-1356
 *
 * Plain C: pointers, sizes and POD structs only -- no torch, no C++ types.  All
 * functions return 0 on success or a negative SEPFWI_E* code; sepfwi_last_error()
 * returns a message for the calling thread.  Nothing in this library calls exit().
 * There is no CPU fallback: without a CUDA device every compute entry fails.
 *
 * Threading: a handle is bound to one device and must be used by one thread at a
 * time; different handles (one per GPU) may be used concurrently, as the
 * reference does from OpenMP threads (Torch_Fwi.cpp:71).
 */
#ifndef SEPFWI_H
#define SEPFWI_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SEPFWI_OK          0
#define SEPFWI_EINVAL     -1   /* bad argument / inconsistent sizes            */
#define SEPFWI_ECUDA      -2   /* CUDA runtime error (message has the detail)  */
#define SEPFWI_ECOURANT   -3   /* CFL number > 1 (utilities.cu:225-241)        */
#define SEPFWI_EIO        -4   /* para/survey/Shot_*.bin file problem          */
#define SEPFWI_ENOMEM     -5   /* device allocation failed                     */

/* DAS component recorded as `ett` and injected by the adjoint. */
#define SEPFWI_FIBER_EXX   0   /* horizontal fiber: vx[z,x]-vx[z,x-1]   (recording_exx, utilities.cu:593-602) */
#define SEPFWI_FIBER_EZZ   1   /* vertical fiber:   vz[z,x]-vz[z-1,x]   (recording_ezz, utilities.cu:620-628) */

/* Scheme flavour. */
#define SEPFWI_FLAVOUR_CPML   0 /* TorchFWI: fp32, CPML, stress->source->velocity->record(it+1)               */
#define SEPFWI_FLAVOUR_SPONGE 1 /* DAS_Waveform_Modeling elasticSolver: sponge, velocity->stress->source->record(it) */

/* Output components of sepfwi_forward (bit mask). */
#define SEPFWI_OUT_PR   1
#define SEPFWI_OUT_VX   2
#define SEPFWI_OUT_VZ   4
#define SEPFWI_OUT_ETT  8
#define SEPFWI_OUT_EXX 16      /* sponge flavour only: exx, ezz, exz divided by dx/dz */
#define SEPFWI_OUT_EZZ 32
#define SEPFWI_OUT_EXZ 64

/* Pointer spaces. */
#define SEPFWI_MEM_HOST    0
#define SEPFWI_MEM_DEVICE  1

typedef struct sepfwi_params {
    int   nz, nx;          /* padded grid exactly as para_file.json: nz = nz_orig + 2 nPml + nPad, nx = nx_orig + 2 nPml */
    int   nPml, nPad;      /* nPoints_pml, nPad (sponge flavour: nPml = ndamp, nPad = 0)                                */
    int   nSteps;
    float dz, dx, dt, f0;
    int   fiber;           /* SEPFWI_FIBER_*                                                                           */
    int   flavour;         /* SEPFWI_FLAVOUR_*                                                                         */
    int   max_batch;       /* shots propagated concurrently on the device (>= 1)                                       */
    int   max_nrec;        /* upper bound of receivers per shot                                                        */
    int   with_adjoint;    /* 1: allocate boundary store + adjoint state (needed by sepfwi_gradient with_adj)          */
    int   kernels;         /* 0: default (shared-memory-resident forward loop where the tiles of a shot fit the SMs, register-streaming
                              kernels otherwise), 1: unfused baseline kernels (cross-check), 3: streaming kernels only */
    int   ref_race_compat; /* 0 (default): race-free adjoint source.  1: reproduce the reference's lost update in
                              res_injection_exx/_ezz (utilities.cu:613-614,639-640, launched 32 receivers per block):
                              when receiver 32k subtracts at the cell receiver 32k-1 adds to, the subtraction is dropped.
                              Only for bit-level comparisons against the reference's own (racy) gradients.                 */
    int   reserved[5];
} sepfwi_params;

typedef struct sepfwi_shot {
    int   zs, xs;          /* source position, padded-grid indices (survey index + nPml, Src_Rec.cu:87,92)             */
    int   nrec;
    const int   *zrec;     /* HOST pointers, padded-grid indices (Src_Rec.cu:108,115)                                  */
    const int   *xrec;
    const float *stf;      /* HOST pointer, nSteps raw samples; the 0.001 end taper of Src_Rec.cu:137 is applied inside
                              (CPML flavour).  Sponge flavour: used as given, amplitude dt/2 (elasticSolver.py:259-260)  */
    float src_rxz;         /* sxx/szz ratio used by the stf gradient only (utilities.cu:719-730); reference default 1   */
    const float *obs_ett;  /* [nrec][nSteps] observed DAS data (gradient only); space given by `mem`                   */
    float *out[7];         /* forward: pr, vx, vz, ett, exx, ezz, exz traces [nrec][nSteps] (NULL = skip); space `mem`  */
    float *gstf;           /* gradient: HOST pointer, nSteps floats (NULL = skip)                                      */
    const float *weights;  /* optional HOST [nrec][3]: ett = w0*exx + w1*ezz + w2*exz (NULL = pure fiber component)    */
    /* data-side options (sepfwi_set_data_options); all optional, HOST pointers */
    const float *win_start;     /* [nrec] window start / end in seconds, survey_file.json "win_start" / "win_end" (Src_Rec.cu:145-170) */
    const float *win_end;
    const float *trace_weights; /* [nrec] "weights" (Src_Rec.cu:176-193); NULL = 1                                    */
    float src_weight;           /* "src_weight" (Src_Rec.cu:195-201); 0 = unset = 1                                   */
    float *src_updated;         /* out, if_src_update: nSteps samples of the updated source time function (NULL = skip) */
} sepfwi_shot;

/* Data-side operators applied to observed and synthetic DAS traces between the forward and the reverse-time loop, in the order of
 * the (commented) call sites libCUFD.cu:353-457; the switches are those of para_file.json (Parameter.cpp:139-176).  All off (the
 * default, and the reference's live behaviour): residual = obs - syn, misfit = 0.5 sum residual^2.  As soon as one switch is on the
 * whole section runs as written there: with if_win off, observed data, synthetic data and the residual still get the window-less end
 * tapers of win_ratio x record length (cuda_window utilities.cu:844-884, call sites libCUFD.cu:363-367, 454-457). */
typedef struct sepfwi_data_options {
    int   if_win;           /* per-trace windows + trace weights + source weight, cuda_window utilities.cu:790-842 (needs win_start / win_end) */
    float win_ratio;        /* taper fraction of the window length; 0 = 0.005 (libCUFD.cu:63)                            */
    int   if_filter;        /* band-pass, bp_filter1d utilities.cu:1115-1168                                             */
    float filter[4];        /* corner frequencies f0 < f1 <= f2 < f3 in Hz ("filter" in para_file.json)                  */
    int   if_cross_misfit;  /* normalised zero-lag cross-correlation misfit + its adjoint source, utilities.cu:1011-1113  */
    int   if_src_update;    /* least-squares source-signature update of the synthetic data + residual, utilities.cu:1170-1356 */
    int   reserved[6];
} sepfwi_data_options;

typedef struct sepfwi_handle sepfwi_handle;

const char *sepfwi_last_error(void);
int sepfwi_version(void);

int sepfwi_create(const sepfwi_params *p, int device, sepfwi_handle **out);
int sepfwi_destroy(sepfwi_handle *h);

/* Select the data-side operators of later sepfwi_gradient calls (NULL = all off).  Allocates its scratch on the device. */
int sepfwi_set_data_options(sepfwi_handle *h, const sepfwi_data_options *o);

/* The data-side chain alone on given traces of one shot: obs, syn [nrec][nSteps] in space `mem` -> adjoint source `res`, conditioned
 * synthetic `syn_out` (each may be NULL), the shot's misfit contribution (0.5 included); shot->src_updated receives the updated source
 * when if_src_update is set.  Needs sepfwi_set_data_options with at least one switch on. */
int sepfwi_condition(sepfwi_handle *h, const sepfwi_shot *shot, const float *obs, const float *syn, float *res, float *syn_out,
                     double *misfit, int mem, void *stream);

/* lambda, mu in MPa, rho in kg/m^3, row-major [nz][nx] (the tensors FWIFunction receives,
 * FWI_ops.py:46-51).  Sponge flavour: lambda, mu in Pa.  Computes the averaged
 * coefficients and the CFL number; returns SEPFWI_ECOURANT if it exceeds 1. */
int sepfwi_set_model(sepfwi_handle *h, const float *lam, const float *mu, const float *rho, int mem, void *stream);
int sepfwi_courant(sepfwi_handle *h, float *courant);

/* Forward modelling of `nshots` shots (any number; processed max_batch at a time). */
int sepfwi_forward(sepfwi_handle *h, int nshots, const sepfwi_shot *shots, int mem, void *stream);

/* Sponge flavour only: forward modelling of ONE shot that also stores the interior (without the ndamp sponge cells) of
 * sxx, szz, vx, vz after every time step `it` with it % save_step == 0 -- elasticSolver.forward_it(isrc, save_wavefield=True),
 * DAS_Waveform_Modeling/src/elasticSolver.py:231-237,279-284,298-303.
 *   snap : [(nSteps-1)/save_step + 1][4][nz - 2 nPml][nx - 2 nPml] floats in space `mem`, field order sxx, szz, vx, vz, each
 *          image row-major [z][x] (the reference stores [x][z]; the Python layer transposes). */
int sepfwi_forward_snapshots(sepfwi_handle *h, const sepfwi_shot *shot, int save_step, float *snap, int mem, void *stream);

/* Misfit (+ gradient when with_adj) of `nshots` shots.
 *   misfit : HOST float, 0.5 * sum over shots and samples of (obs-syn)^2      (libCUFD.cu:427,776)
 *   glam, gmu, grho : [nz][nx] row-major in space `mem`, OVERWRITTEN with the sum over the given shots
 *                     (gradients w.r.t. the MPa inputs for lambda, mu: el_stress.cu:108-115)
 * `syn_ett`, when non-NULL in shots[i].out[3], receives the synthetic traces. */
int sepfwi_gradient(sepfwi_handle *h, int nshots, const sepfwi_shot *shots, int with_adj,
                    float *misfit, float *glam, float *gmu, float *grho, int mem, void *stream);

/* Drop-in for the reference's `cufd`: host pointers, reads para_file.json /
 * survey_file.json, reads/writes Shot_{pr,vx,vz,ett}{id}.bin exactly like
 * libCUFD.cu:215-223,755-769.  calc_id 0 = misfit, 1 = misfit + gradient, 2 = observed data.
 * grad_stf rows are written at the LOCAL shot index like libCUFD.cu:671-673. */
int sepfwi_cufd(float *misfit, float *grad_Lambda, float *grad_Mu, float *grad_Den, float *grad_stf,
                const float *Lambda, const float *Mu, const float *Den, const float *stf,
                int calc_id, int gpu_id, int group_size, const int *shot_ids, const char *para_fname);

/* Frees the handles (device memory) sepfwi_cufd caches between calls. */
int sepfwi_cufd_clear_cache(void);

/* Boundary-ring geometry and save/restore of one [nz][nx] host field -- exposed so the index
 * map of utilities.cu:362-425 can be checked bit-exactly.  `bnd` is a HOST buffer of
 * sepfwi_ring_len() floats. */
int sepfwi_ring_len(const sepfwi_params *p);
int sepfwi_ring_save(sepfwi_handle *h, const float *field, float *bnd);
int sepfwi_ring_restore(sepfwi_handle *h, float *field, const float *bnd);

/* Introspection for tests and the bench: CPML profiles as built on the host
 * (order K, a, b, K_half, a_half, b_half; z arrays have nz-nPad entries), kernel launch count
 * since creation, and device time of the last forward / backward time loops in ms. */
int sepfwi_get_cpml(sepfwi_handle *h, int axis /*0=z,1=x*/, float *out6xN);
long long sepfwi_launch_count(sepfwi_handle *h);
long long sepfwi_resident_launches(sepfwi_handle *h);   /* cooperative launches of the resident forward loop so far */
/* Host-only: the tiling the resident forward loop would use for `nshots` concurrent shots on a device with `nsm` SMs and
 * `smem_optin` bytes of opt-in shared memory per block; out = {rows per thread (0 = streaming kernels instead), tiles in x,
 * tiles in z, own rows per tile, shots per cooperative launch}.  Makes no CUDA call. */
int sepfwi_plan_resident(const sepfwi_params *p, int nshots, int nsm, size_t smem_optin, int out[5]);
/* Host-only: the work list of streaming kernel `which` (0 forward, 1 reconstruction + imaging, 2 adjoint) -- one entry per warp,
 * {first owned column, first row, end row, 1 if the warp takes the CPML path}; *n = number of entries (may exceed cap). */
int sepfwi_plan_stream(const sepfwi_params *p, int nshots, int nsm, int which, int *items4, int cap, int *n);
/* Host-only: the plan of the reverse-time step for `nshots` concurrent shots on `nsm` SMs.  out = {1 if the reconstruction and the
 * adjoint sweep of a step share one launch (k_stream_bwd), jointly chosen chunk heights (adjoint interior, adjoint edge,
 * reconstruction interior, reconstruction edge; 0 = planned separately), adjoint items per shot, reconstruction items per shot}. */
int sepfwi_plan_backward(const sepfwi_params *p, int nshots, int nsm, int out[7]);
int sepfwi_last_timing(sepfwi_handle *h, float *fwd_ms, float *bwd_ms);
/* The misfit of the last sepfwi_gradient call in double precision (the float* argument mirrors the reference's float). */
int sepfwi_last_misfit(sepfwi_handle *h, double *misfit);
/* Host-only: device bytes per unit of max_batch that sepfwi_create would allocate for these parameters (max_batch ignored). */
long long sepfwi_bytes_per_slot(const sepfwi_params *p);

/* Per-kernel device timing: with nsteps > 0 every launch of the first nsteps time steps of each
 * time loop is bracketed by a CUDA-event pair on the launching stream; sepfwi_get_profile returns
 * accumulated milliseconds and launch counts per kernel kind since the last sepfwi_set_profile. */
enum { SEPFWI_K_RING_SAVE = 0, SEPFWI_K_STRESS_FWD, SEPFWI_K_VELOCITY_FWD, SEPFWI_K_RECORD, SEPFWI_K_VELOCITY_BWD,
       SEPFWI_K_STRESS_BWD, SEPFWI_K_VELOCITY_ADJ, SEPFWI_K_INJECT, SEPFWI_K_STRESS_ADJ, SEPFWI_K_STREAM_FWD, SEPFWI_K_STREAM_RECON,
       SEPFWI_K_STREAM_ADJ, SEPFWI_K_RESIDENT_FWD, SEPFWI_K_STREAM_BWD /* reconstruction + adjoint sweep in one launch */, SEPFWI_NKERNEL };
int sepfwi_set_profile(sepfwi_handle *h, int nsteps);
int sepfwi_get_profile(sepfwi_handle *h, double *ms /*[SEPFWI_NKERNEL]*/, long long *count /*[SEPFWI_NKERNEL]*/);
const char *sepfwi_kernel_name(int kind);

#ifdef __cplusplus
}
#endif
#endif

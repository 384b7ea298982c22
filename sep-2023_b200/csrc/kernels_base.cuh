// kernels_base.cuh -- baseline (unfused) sm_100a kernels: one launch per half step.
//
// These are the first correct CUDA path and stay in the library as the on-GPU
// cross-check of the streaming / resident kernels (sepfwi_params.kernels = 1).  Written from the
// numerical specification in SURVEY.md App. A; reference lines are cited per kernel
// (paths under DAS_Waveform_Inversion/Ops/FWI/Src/).
//
// Differences from the reference kernels, all deliberate:
//   * x-fastest layout, 32x8 thread blocks, dead nPad rows not stored;
//   * source injection is folded into the stress kernels, stf gradient and ring
//     restore into the reverse-time kernels (12 launches/step -> 4 fwd, 4 bwd);
//   * gradient "sprays" (atomicAdd in el_stress.cu:112-123, el_velocity.cu:104-110)
//     are evaluated as deterministic gathers;
//   * the adjoint memory variables are only maintained where a CPML coefficient
//     can ever multiply them (nPml+2 wide strips) instead of over the whole grid
//     (el_stress_adj.cu:67-72,88-95).
#pragma once
#include "common.cuh"

namespace sepfwi {

#define BX 32
#define BY 8

// ----------------------------------------------------------------------------------------
// Forward stress update + explosive source.   el_stress.cu:50-87, utilities.cu:524-552
// SPONGE: elasticSolver.py:348-386 + :254-260 (multiplicative sponge, then source).
template <bool SPONGE>
__global__ void __launch_bounds__(BX *BY) k_stress_fwd(const KArgs a, const int it)
{
    const Dims &d = a.d;
    const int x = blockIdx.x * BX + threadIdx.x, z = blockIdx.y * BY + threadIdx.y, s = blockIdx.z;
    if (z < 2 || z > d.nzA - 3 || x < 2 || x > d.nx - 3) return;
    const int ld = d.ldx;
    const size_t i = (size_t)z * ld + x;
    float *st = slot_state(a, s);
    const float *vz = st + (S_FWD + F_VZ) * d.fsz, *vx = st + (S_FWD + F_VX) * d.fsz;
    float *szz = st + (S_FWD + F_SZZ) * d.fsz, *sxx = st + (S_FWD + F_SXX) * d.fsz, *sxz = st + (S_FWD + F_SXZ) * d.fsz;

    float dvz_dz = dzb(vz, i, ld, d.c1z, d.c2z);
    float dvx_dx = dxb(vx, i, d.c1x, d.c2x);
    float dvx_dz = dzf(vx, i, ld, d.c1z, d.c2z);
    float dvz_dx = dxf(vz, i, d.c1x, d.c2x);
    if (!SPONGE) {
        const bool zp = (z < d.nPml) || (z > d.nzA - d.nPml - 1);
        const bool xp = (x < d.nPml) || (x > d.nx - d.nPml - 1);
        if (zp) {
            const float *c = a.cz + z;
            float *p0 = st + (S_FPSI + P_VZ_Z) * d.fsz + i, *p1 = st + (S_FPSI + P_VX_Z) * d.fsz + i;
            float m0 = c[C_B * d.nzA] * *p0 + c[C_A * d.nzA] * dvz_dz;
            float m1 = c[C_BH * d.nzA] * *p1 + c[C_AH * d.nzA] * dvx_dz;
            *p0 = m0; *p1 = m1;
            dvz_dz = dvz_dz * c[C_RK * d.nzA] + m0;
            dvx_dz = dvx_dz * c[C_RKH * d.nzA] + m1;
        }
        if (xp) {
            const float *c = a.cx + x;
            float *p0 = st + (S_FPSI + P_VX_X) * d.fsz + i, *p1 = st + (S_FPSI + P_VZ_X) * d.fsz + i;
            float m0 = c[C_B * d.nx] * *p0 + c[C_A * d.nx] * dvx_dx;
            float m1 = c[C_BH * d.nx] * *p1 + c[C_AH * d.nx] * dvz_dx;
            *p0 = m0; *p1 = m1;
            dvx_dx = dvx_dx * c[C_RK * d.nx] + m0;
            dvz_dx = dvz_dx * c[C_RKH * d.nx] + m1;
        }
    }
    const float lam = a.model[M_LAM * d.fsz + i], mu = a.model[M_MU * d.fsz + i], mua = a.model[M_MUAVE * d.fsz + i];
    const float l2u = lam + 2.0f * mu;
    float nzz = szz[i] + (l2u * dvz_dz + lam * dvx_dx) * d.dt;
    float nxx = sxx[i] + (lam * dvz_dz + l2u * dvx_dx) * d.dt;
    float nxz = sxz[i] + mua * (dvx_dz + dvz_dx) * d.dt;
    if (SPONGE) { const float g = a.damp[i]; nzz *= g; nxx *= g; nxz *= g; }
    if (z == a.t.zs[s] && x == a.t.xs[s]) {
        const float amp = a.t.amp[(size_t)s * d.nSteps + it];
        nzz += amp; nxx += amp;
    }
    szz[i] = nzz; sxx[i] = nxx; sxz[i] = nxz;
}

// ----------------------------------------------------------------------------------------
// Forward velocity update.   el_velocity.cu:45-82 ; SPONGE: elasticSolver.py:310-345 + :247-248
template <bool SPONGE>
__global__ void __launch_bounds__(BX *BY) k_velocity_fwd(const KArgs a)
{
    const Dims &d = a.d;
    const int x = blockIdx.x * BX + threadIdx.x, z = blockIdx.y * BY + threadIdx.y, s = blockIdx.z;
    if (z < 2 || z > d.nzA - 3 || x < 2 || x > d.nx - 3) return;
    const int ld = d.ldx;
    const size_t i = (size_t)z * ld + x;
    float *st = slot_state(a, s);
    float *vz = st + (S_FWD + F_VZ) * d.fsz, *vx = st + (S_FWD + F_VX) * d.fsz;
    const float *szz = st + (S_FWD + F_SZZ) * d.fsz, *sxx = st + (S_FWD + F_SXX) * d.fsz, *sxz = st + (S_FWD + F_SXZ) * d.fsz;

    float dszz_dz = dzf(szz, i, ld, d.c1z, d.c2z);
    float dsxz_dx = dxb(sxz, i, d.c1x, d.c2x);
    float dsxz_dz = dzb(sxz, i, ld, d.c1z, d.c2z);
    float dsxx_dx = dxf(sxx, i, d.c1x, d.c2x);
    if (!SPONGE) {
        const bool zp = (z < d.nPml) || (z > d.nzA - d.nPml - 1);
        const bool xp = (x < d.nPml) || (x > d.nx - d.nPml);       // el_velocity.cu:56,71: one column narrower than the stress kernel
        if (zp) {
            const float *c = a.cz + z;
            float *p0 = st + (S_FPSI + P_SZZ_Z) * d.fsz + i, *p1 = st + (S_FPSI + P_SXZ_Z) * d.fsz + i;
            float m0 = c[C_BH * d.nzA] * *p0 + c[C_AH * d.nzA] * dszz_dz;
            float m1 = c[C_B * d.nzA] * *p1 + c[C_A * d.nzA] * dsxz_dz;
            *p0 = m0; *p1 = m1;
            dszz_dz = dszz_dz * c[C_RKH * d.nzA] + m0;
            dsxz_dz = dsxz_dz * c[C_RK * d.nzA] + m1;
        }
        if (xp) {
            const float *c = a.cx + x;
            float *p0 = st + (S_FPSI + P_SXZ_X) * d.fsz + i, *p1 = st + (S_FPSI + P_SXX_X) * d.fsz + i;
            float m0 = c[C_B * d.nx] * *p0 + c[C_A * d.nx] * dsxz_dx;
            float m1 = c[C_BH * d.nx] * *p1 + c[C_AH * d.nx] * dsxx_dx;
            *p0 = m0; *p1 = m1;
            dsxz_dx = dsxz_dx * c[C_RK * d.nx] + m0;
            dsxx_dx = dsxx_dx * c[C_RKH * d.nx] + m1;
        }
    }
    float nvz = vz[i] + (dszz_dz + dsxz_dx) * a.model[M_BYCA * d.fsz + i] * d.dt;
    float nvx = vx[i] + (dsxz_dz + dsxx_dx) * a.model[M_BYCB * d.fsz + i] * d.dt;
    if (SPONGE) { const float g = a.damp[i]; nvz *= g; nvx *= g; }
    vz[i] = nvz; vx[i] = nvx;
}

// ----------------------------------------------------------------------------------------
// Boundary ring save of the five forward fields at time index `it`, reference layout.
// Bnd::field_from_bnd, Boundary.cu:55-76; from_bnd, utilities.cu:362-392.
__device__ __forceinline__ void ring_cell(const Dims &d, int idx, int &z, int &x)
{
    const int L = 5;
    if (idx < L * d.nzB) { int j = idx / d.nzB, i = idx - j * d.nzB; z = i + d.nPml - 2; x = j + d.nPml - 2; }
    else if (idx < 2 * L * d.nzB) { int q = idx - L * d.nzB; int j = q / d.nzB, i = q - j * d.nzB; z = i + d.nPml - 2; x = d.nx - d.nPml - j + 1; }
    else if (idx < L * (2 * d.nzB + d.nxB)) { int q = idx - 2 * L * d.nzB; int i = q / d.nxB, j = q - i * d.nxB; z = i + d.nPml - 2; x = j + d.nPml - 2; }
    else { int q = idx - L * (2 * d.nzB + d.nxB); int i = q / d.nxB, j = q - i * d.nxB; z = d.nzA - d.nPml - i + 1; x = j + d.nPml - 2; }
}

__global__ void k_ring_save(const KArgs a, const int it)
{
    const Dims &d = a.d;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x, s = blockIdx.y;
    if (idx >= d.ringLen) return;
    int z, x; ring_cell(d, idx, z, x);
    const size_t i = (size_t)z * d.ldx + x;
    const float *st = slot_state(a, s);
#pragma unroll
    for (int f = 0; f < NFIELD; f++)
        a.ring[(((size_t)s * NFIELD + f) * d.nSteps + it) * d.ringLen + idx] = st[(S_FWD + f) * d.fsz + i];
}

// Generic single-field ring save / restore used by the sepfwi_ring_* test entry points.
__global__ void k_ring_copy_field(const Dims d, float *field, float *bnd, const int restore)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= d.ringLen) return;
    int z, x; ring_cell(d, idx, z, x);
    const size_t i = (size_t)z * d.ldx + x;
    if (restore) field[i] = bnd[idx]; else bnd[idx] = field[i];
}

// ----------------------------------------------------------------------------------------
// Trace recording at sample `samp`.  recording / _vx / _vz / _exx / _ezz, utilities.cu:593-703.
// SPONGE flavour: elasticSolver.py:263-276 (pr halved, strain rates divided by the spacing).
template <bool SPONGE>
__global__ void k_record(const KArgs a, const int samp, const int mask, const int fiber, const int par)
{
    const Dims &d = a.d;
    const int r = blockIdx.x * blockDim.x + threadIdx.x, s = blockIdx.y;
    if (r >= a.t.nrec[s]) return;
    const int z = a.t.zrec[(size_t)s * d.maxRec + r], x = a.t.xrec[(size_t)s * d.maxRec + r];
    const int ld = d.ldx;
    const size_t i = (size_t)z * ld + x;
    const float *st = slot_state(a, s) + (size_t)(par ? S_FWD1 : S_FWD) * d.fsz;   // par: which ping-pong buffer holds the state
    const float *vz = st + F_VZ * d.fsz, *vx = st + F_VX * d.fsz;
    const float *szz = st + F_SZZ * d.fsz, *sxx = st + F_SXX * d.fsz;
    float *tr = a.trace + (size_t)s * d.nTrace * d.maxRec * d.nSteps + (size_t)r * d.nSteps + samp;
    const size_t cs = (size_t)d.maxRec * d.nSteps;
    float exx = vx[i] - vx[i - 1];
    float ezz = vz[i] - vz[i - ld];
    if (SPONGE) { exx *= d.rdx; ezz *= d.rdz; }
    if (mask & (1 << T_PR)) tr[T_PR * cs] = SPONGE ? 0.5f * (sxx[i] + szz[i]) : szz[i] + sxx[i];
    if (mask & (1 << T_VX)) tr[T_VX * cs] = vx[i];
    if (mask & (1 << T_VZ)) tr[T_VZ * cs] = vz[i];
    float exz = 0.f;
    const float *w = a.t.w ? a.t.w + ((size_t)s * d.maxRec + r) * 3 : nullptr;
    if ((mask & (1 << T_EXZ)) || w) {
        exz = SPONGE ? 0.5f * ((vx[i + ld] - vx[i]) * d.rdz + (vz[i + 1] - vz[i]) * d.rdx)
                     : 0.5f * ((vx[i + ld] - vx[i]) + (vz[i + 1] - vz[i]));
    }
    if (mask & (1 << T_ETT)) tr[T_ETT * cs] = w ? w[0] * exx + w[1] * ezz + w[2] * exz : (fiber == 0 ? exx : ezz);
    if (d.nTrace > T_EXX) {
        if (mask & (1 << T_EXX)) tr[T_EXX * cs] = exx;
        if (mask & (1 << T_EZZ)) tr[T_EZZ * cs] = ezz;
        if (mask & (1 << T_EXZ)) tr[T_EXZ * cs] = exz;
    }
}

// ----------------------------------------------------------------------------------------
// Residual res = obs - syn with sample 0 forced to 0 (gpuMinus, utilities.cu:154-167) and the
// per-block partial sums of res^2 (cuda_cal_objective, utilities.cu:169-205, there one block).
__global__ void __launch_bounds__(256) k_residual(const KArgs a, double *partial, const int nblk)
{
    const Dims &d = a.d;
    const int s = blockIdx.y;
    const size_t n = (size_t)a.t.nrec[s] * d.nSteps;
    float *tb = a.trace + (size_t)s * d.nTrace * d.maxRec * d.nSteps;
    const size_t cs = (size_t)d.maxRec * d.nSteps;
    const float *syn = tb + T_ETT * cs, *obs = tb + T_OBS * cs;
    float *res = tb + T_RES * cs;
    double acc = 0.0;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        const int it = (int)(k % d.nSteps);
        const float r = it > 0 ? obs[k] - syn[k] : 0.0f;
        res[k] = r;
        acc += (double)r * r;
    }
    __shared__ double sm[256];
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[(size_t)s * nblk + blockIdx.x] = sm[0];
}

__global__ void k_sum_partials(const double *partial, const int nblk, double *out, const int nslot)
{
    // one thread per slot: fixed summation order => deterministic misfit
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nslot) return;
    double acc = 0.0;
    for (int k = 0; k < nblk; k++) acc += partial[(size_t)s * nblk + k];
    out[s] = acc;
}

// ----------------------------------------------------------------------------------------
// Reverse-time kernels.  Region = interior U ring = [nPml-2, z1+2] x [nPml-2, x1+2].
__device__ __forceinline__ bool in_interior(const Dims &d, int z, int x)
{ return z >= d.nPml && z <= d.z1 && x >= d.nPml && x <= d.x1; }

// Velocity reconstruction + density imaging + stf gradient + ring restore of vz, vx.
// el_velocity.cu:84-117 ; source_grad utilities.cu:719-730 ; to_bnd utilities.cu:395-425.
__global__ void __launch_bounds__(BX *BY) k_velocity_bwd(const KArgs a, const int it)
{
    const Dims &d = a.d;
    const int s = blockIdx.z;
    float *st = slot_state(a, s);
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && threadIdx.y == 0) {
        const size_t is = (size_t)a.t.zs[s] * d.ldx + a.t.xs[s];
        a.gstf[(size_t)s * d.nSteps + it] =
            -(st[(S_ADJ + F_SZZ) * d.fsz + is] + a.t.rxz[s] * st[(S_ADJ + F_SXX) * d.fsz + is]) * d.dt;
    }
    const int x = d.nPml - 2 + blockIdx.x * BX + threadIdx.x, z = d.nPml - 2 + blockIdx.y * BY + threadIdx.y;
    if (z > d.z1 + 2 || x > d.x1 + 2) return;
    const int ld = d.ldx;
    const size_t i = (size_t)z * ld + x;
    float *vz = st + (S_FWD + F_VZ) * d.fsz, *vx = st + (S_FWD + F_VX) * d.fsz;
    const float *szz = st + (S_FWD + F_SZZ) * d.fsz, *sxx = st + (S_FWD + F_SXX) * d.fsz, *sxz = st + (S_FWD + F_SXZ) * d.fsz;
    const float *vza = st + (S_ADJ + F_VZ) * d.fsz, *vxa = st + (S_ADJ + F_VX) * d.fsz;
    const float *bya = a.model + M_BYCA * d.fsz, *byb = a.model + M_BYCB * d.fsz;

    float nvz = vz[i], nvx = vx[i];
    float g = 0.f;
    bool touch = false;
    if (in_interior(d, z, x)) {
        const float A = dzf(szz, i, ld, d.c1z, d.c2z) + dxb(sxz, i, d.c1x, d.c2x);
        const float B = dzb(sxz, i, ld, d.c1z, d.c2z) + dxf(sxx, i, d.c1x, d.c2x);
        const float ba = bya[i], bb = byb[i];
        nvz -= A * ba * d.dt;
        nvx -= B * bb * d.dt;
        g = vza[i] * A * d.dt * (0.5f * ba * ba) + vxa[i] * B * d.dt * (0.5f * bb * bb);
        touch = true;
    }
    if (z <= d.z1 && in_interior(d, z - 1, x)) {          // spray to (z+1, x), el_velocity.cu:107-108
        const size_t k = i - ld;
        const float A = dzf(szz, k, ld, d.c1z, d.c2z) + dxb(sxz, k, d.c1x, d.c2x);
        const float ba = bya[k];
        g += vza[k] * A * d.dt * (0.5f * ba * ba);
        touch = true;
    }
    if (in_interior(d, z, x - 1)) {                        // spray to (z, x+1), unconditional (el_velocity.cu:109)
        const size_t k = i - 1;
        const float B = dzb(sxz, k, ld, d.c1z, d.c2z) + dxf(sxx, k, d.c1x, d.c2x);
        const float bb = byb[k];
        g += vxa[k] * B * d.dt * (0.5f * bb * bb);
        touch = true;
    }
    if (touch) a.grad[((size_t)s * 3 + 2) * d.fsz + i] += g;
    int ridx[4];
    if (ring_indices(d, z, x, ridx) > 0) {
        const float *rb = a.ring + (((size_t)s * NFIELD) * d.nSteps + it) * d.ringLen + ridx[0];
        const size_t fs = (size_t)d.nSteps * d.ringLen;
        nvz = rb[F_VZ * fs];
        nvx = rb[F_VX * fs];
    }
    vz[i] = nvz; vx[i] = nvx;
}

// Shear-modulus imaging term of one staggered cell (el_stress.cu:112-115), 0 outside the interior.
__device__ __forceinline__ float shear_scale(const KArgs &a, const float *vz, const float *vx, const float *sxza,
                                             int z, int x)
{
    const Dims &d = a.d;
    if (!in_interior(d, z, x)) return 0.f;
    const int ld = d.ldx;
    const size_t i = (size_t)z * ld + x;
    const float mua = a.model[M_MUAVE * d.fsz + i];
    if (mua == 0.0f) return 0.f;
    const float *mu = a.model + M_MU * d.fsz;
    const float D3 = dzf(vx, i, ld, d.c1z, d.c2z) + dxf(vz, i, d.c1x, d.c2x);
    const float rs = 1.0f / mu[i] + 1.0f / mu[i + ld] + 1.0f / mu[i + 1] + 1.0f / mu[i + ld + 1];
    return -sxza[i] * D3 * d.dt * mua / rs * 1e6f;
}

// Source removal + stress reconstruction + lambda/mu imaging + ring restore of szz, sxz, sxx.
// add_source(isFor=false) utilities.cu:541-550 ; el_stress.cu:89-128 ; to_bnd.
__global__ void __launch_bounds__(BX *BY) k_stress_bwd(const KArgs a, const int it)
{
    const Dims &d = a.d;
    const int s = blockIdx.z;
    const int x = d.nPml - 2 + blockIdx.x * BX + threadIdx.x, z = d.nPml - 2 + blockIdx.y * BY + threadIdx.y;
    if (z > d.z1 + 2 || x > d.x1 + 2) return;
    const int ld = d.ldx;
    const size_t i = (size_t)z * ld + x;
    float *st = slot_state(a, s);
    const float *vz = st + (S_FWD + F_VZ) * d.fsz, *vx = st + (S_FWD + F_VX) * d.fsz;
    float *szz = st + (S_FWD + F_SZZ) * d.fsz, *sxx = st + (S_FWD + F_SXX) * d.fsz, *sxz = st + (S_FWD + F_SXZ) * d.fsz;
    const float *szza = st + (S_ADJ + F_SZZ) * d.fsz, *sxxa = st + (S_ADJ + F_SXX) * d.fsz, *sxza = st + (S_ADJ + F_SXZ) * d.fsz;

    float nzz = szz[i], nxx = sxx[i], nxz = sxz[i];
    if (z == a.t.zs[s] && x == a.t.xs[s]) {
        const float amp = a.t.amp[(size_t)s * d.nSteps + it];
        nzz -= amp; nxx -= amp;
    }
    if (in_interior(d, z, x)) {
        const float D1 = dzb(vz, i, ld, d.c1z, d.c2z), D2 = dxb(vx, i, d.c1x, d.c2x);
        const float D3 = dzf(vx, i, ld, d.c1z, d.c2z) + dxf(vz, i, d.c1x, d.c2x);
        const float lam = a.model[M_LAM * d.fsz + i], mu = a.model[M_MU * d.fsz + i], mua = a.model[M_MUAVE * d.fsz + i];
        const float l2u = lam + 2.0f * mu;
        nzz -= (l2u * D1 + lam * D2) * d.dt;
        nxx -= (lam * D1 + l2u * D2) * d.dt;
        nxz -= mua * D3 * d.dt;
        const float za = szza[i], xa = sxxa[i];
        a.grad[((size_t)s * 3 + 0) * d.fsz + i] += -(za + xa) * (D1 + D2) * d.dt * 1e6f;
        float gm = (-2.0f * za * D1 * d.dt - 2.0f * xa * D2 * d.dt) * 1e6f;
        // gather form of the four-point spray
        float sh = shear_scale(a, vz, vx, sxza, z, x) + shear_scale(a, vz, vx, sxza, z - 1, x) +
                   shear_scale(a, vz, vx, sxza, z, x - 1) + shear_scale(a, vz, vx, sxza, z - 1, x - 1);
        gm += sh / (mu * mu);
        a.grad[((size_t)s * 3 + 1) * d.fsz + i] += gm;
    } else if (x == d.x1 + 1 && z >= d.nPml && z <= d.z1) {
        // the reference's unconditional (z, x+1) spray lands one column outside the interior (el_stress.cu:119)
        const float mu = a.model[M_MU * d.fsz + i];
        float sh = shear_scale(a, vz, vx, sxza, z, x - 1);
        a.grad[((size_t)s * 3 + 1) * d.fsz + i] += sh / (mu * mu);
    }
    int ridx[4];
    if (ring_indices(d, z, x, ridx) > 0) {
        const float *rb = a.ring + (((size_t)s * NFIELD) * d.nSteps + it) * d.ringLen + ridx[0];
        const size_t fs = (size_t)d.nSteps * d.ringLen;
        nzz = rb[F_SZZ * fs]; nxz = rb[F_SXZ * fs]; nxx = rb[F_SXX * fs];
    }
    szz[i] = nzz; sxx[i] = nxx; sxz[i] = nxz;
}

// ----------------------------------------------------------------------------------------
// Adjoint sweep.  Memory-variable strips: a coefficient `a` is non-zero only inside the
// PML and the stencils that multiply it reach two cells further, so the adjoint memory
// variables are kept on z < nPml+2 | z > nzA-nPml-3 (z arrays) and the same in x.
__device__ __forceinline__ bool zstrip(const Dims &d, int z) { return z < d.nPml + 2 || z > d.nzA - d.nPml - 3; }
__device__ __forceinline__ bool xstrip(const Dims &d, int x) { return x < d.nPml + 2 || x > d.nx - d.nPml - 3; }

// Adjoint velocity update.   el_velocity_adj.cu:55-103
__global__ void __launch_bounds__(BX *BY) k_velocity_adj(const KArgs a)
{
    const Dims &d = a.d;
    const int x = blockIdx.x * BX + threadIdx.x, z = blockIdx.y * BY + threadIdx.y, s = blockIdx.z;
    if (z < 2 || z > d.nzA - 3 || x < 2 || x > d.nx - 3) return;
    const int ld = d.ldx;
    const size_t i = (size_t)z * ld + x;
    float *st = slot_state(a, s);
    float *vz = st + (S_ADJ + F_VZ) * d.fsz, *vx = st + (S_ADJ + F_VX) * d.fsz;
    const float *szz = st + (S_ADJ + F_SZZ) * d.fsz, *sxx = st + (S_ADJ + F_SXX) * d.fsz, *sxz = st + (S_ADJ + F_SXZ) * d.fsz;
    float *ps = st + S_APSI * d.fsz;
    const float *cz = a.cz + z, *cx = a.cx + x;
    const float rKx = cx[C_RK * d.nx], ax = cx[C_A * d.nx], rKxh = cx[C_RKH * d.nx], axh = cx[C_AH * d.nx];
    const float rKz = cz[C_RK * d.nzA], az = cz[C_A * d.nzA], rKzh = cz[C_RKH * d.nzA], azh = cz[C_AH * d.nzA];
    const float lam = a.model[M_LAM * d.fsz + i], mu = a.model[M_MU * d.fsz + i], mua = a.model[M_MUAVE * d.fsz + i];
    const float l2u = lam + 2.0f * mu;
    const bool zp = (z < d.nPml) || (z > d.nzA - d.nPml - 1);
    const bool xp = (x < d.nPml) || (x > d.nx - d.nPml - 1);

    // vx
    float acc = (lam * -dxf(szz, i, d.c1x, d.c2x) + l2u * -dxf(sxx, i, d.c1x, d.c2x)) * rKx * d.dt +
                mua * rKzh * -dzb(sxz, i, ld, d.c1z, d.c2z) * d.dt;
    if (ax != 0.f) acc += ax * -dxf(ps + P_VX_X * d.fsz, i, d.c1x, d.c2x);
    if (azh != 0.f) acc += azh * -dzb(ps + P_VX_Z * d.fsz, i, ld, d.c1z, d.c2z);
    const float nvx = vx[i] + acc;
    vx[i] = nvx;
    const float bb = a.model[M_BYCB * d.fsz + i], ba = a.model[M_BYCA * d.fsz + i];
    if (xp) { float *p = ps + P_SXX_X * d.fsz + i; *p = cx[C_BH * d.nx] * *p + bb * nvx * d.dt; }
    if (zp) { float *p = ps + P_SXZ_Z * d.fsz + i; *p = cz[C_B * d.nzA] * *p + bb * nvx * d.dt; }
    // vz
    acc = (l2u * -dzf(szz, i, ld, d.c1z, d.c2z) + lam * -dzf(sxx, i, ld, d.c1z, d.c2z)) * rKz * d.dt +
          mua * rKxh * -dxb(sxz, i, d.c1x, d.c2x) * d.dt;
    if (az != 0.f) acc += az * -dzf(ps + P_VZ_Z * d.fsz, i, ld, d.c1z, d.c2z);
    if (axh != 0.f) acc += axh * -dxb(ps + P_VZ_X * d.fsz, i, d.c1x, d.c2x);
    const float nvz = vz[i] + acc;
    vz[i] = nvz;
    if (xp) { float *p = ps + P_SXZ_X * d.fsz + i; *p = cx[C_B * d.nx] * *p + ba * nvz * d.dt; }
    if (zp) { float *p = ps + P_SZZ_Z * d.fsz + i; *p = cz[C_BH * d.nzA] * *p + ba * nvz * d.dt; }
}

// Adjoint source: residual injection, deterministic gather over unique target cells.
// res_injection_exx / _ezz, utilities.cu:605-641 (racy scatter in the reference).
__global__ void k_inject(const KArgs a, const int it)
{
    const Dims &d = a.d;
    const int k = blockIdx.x * blockDim.x + threadIdx.x, s = blockIdx.y;
    if (k >= a.t.injN[s]) return;
    const size_t tb = (size_t)s * a.t.maxInj, cb = (size_t)s * a.t.maxCon;
    const int p0 = a.t.injPtr[(size_t)s * (a.t.maxInj + 1) + k], p1 = a.t.injPtr[(size_t)s * (a.t.maxInj + 1) + k + 1];
    const float *res = a.trace + ((size_t)s * d.nTrace + T_RES) * d.maxRec * d.nSteps + it;
    float *f = slot_state(a, s) + (size_t)(S_ADJ + a.t.injField[tb + k]) * d.fsz + a.t.injCell[tb + k];
    float v = *f;
    for (int p = p0; p < p1; p++) v += a.t.injCoef[cb + p] * res[(size_t)a.t.injRec[cb + p] * d.nSteps];
    *f = v;
}

// Adjoint stress update.   el_stress_adj.cu:50-97
__global__ void __launch_bounds__(BX *BY) k_stress_adj(const KArgs a)
{
    const Dims &d = a.d;
    const int x = blockIdx.x * BX + threadIdx.x, z = blockIdx.y * BY + threadIdx.y, s = blockIdx.z;
    if (z < 2 || z > d.nzA - 3 || x < 2 || x > d.nx - 3) return;
    const int ld = d.ldx;
    const size_t i = (size_t)z * ld + x;
    float *st = slot_state(a, s);
    const float *vz = st + (S_ADJ + F_VZ) * d.fsz, *vx = st + (S_ADJ + F_VX) * d.fsz;
    float *szz = st + (S_ADJ + F_SZZ) * d.fsz, *sxx = st + (S_ADJ + F_SXX) * d.fsz, *sxz = st + (S_ADJ + F_SXZ) * d.fsz;
    float *ps = st + S_APSI * d.fsz;
    const float *cz = a.cz + z, *cx = a.cx + x;
    const float rKx = cx[C_RK * d.nx], ax = cx[C_A * d.nx], rKxh = cx[C_RKH * d.nx], axh = cx[C_AH * d.nx];
    const float rKz = cz[C_RK * d.nzA], az = cz[C_A * d.nzA], rKzh = cz[C_RKH * d.nzA], azh = cz[C_AH * d.nzA];
    const float lam = a.model[M_LAM * d.fsz + i], mu = a.model[M_MU * d.fsz + i], mua = a.model[M_MUAVE * d.fsz + i];
    const float l2u = lam + 2.0f * mu;
    const float bb = a.model[M_BYCB * d.fsz + i], ba = a.model[M_BYCA * d.fsz + i];
    const bool zs = zstrip(d, z), xs = xstrip(d, x);

    float acc = -dxf(vz, i, d.c1x, d.c2x) * rKx * ba * d.dt + -dzf(vx, i, ld, d.c1z, d.c2z) * rKz * bb * d.dt;
    if (ax != 0.f) acc += ax * -dxf(ps + P_SXZ_X * d.fsz, i, d.c1x, d.c2x);
    if (az != 0.f) acc += az * -dzf(ps + P_SXZ_Z * d.fsz, i, ld, d.c1z, d.c2z);
    const float nxz = sxz[i] + acc;
    sxz[i] = nxz;
    if (xs) { float *p = ps + P_VZ_X * d.fsz + i; *p = cx[C_BH * d.nx] * *p + nxz * mua * d.dt; }
    if (zs) { float *p = ps + P_VX_Z * d.fsz + i; *p = cz[C_BH * d.nzA] * *p + nxz * mua * d.dt; }

    float accx = bb * -dxb(vx, i, d.c1x, d.c2x) * rKxh * d.dt;
    if (axh != 0.f) accx += axh * -dxb(ps + P_SXX_X * d.fsz, i, d.c1x, d.c2x);
    float accz = ba * -dzb(vz, i, ld, d.c1z, d.c2z) * rKzh * d.dt;
    if (azh != 0.f) accz += azh * -dzb(ps + P_SZZ_Z * d.fsz, i, ld, d.c1z, d.c2z);
    const float nxx = sxx[i] + accx, nzz = szz[i] + accz;
    sxx[i] = nxx; szz[i] = nzz;
    if (xs) { float *p = ps + P_VX_X * d.fsz + i; *p = cx[C_B * d.nx] * *p + lam * nzz * d.dt + l2u * nxx * d.dt; }
    if (zs) { float *p = ps + P_VZ_Z * d.fsz + i; *p = cz[C_B * d.nzA] * *p + l2u * nzz * d.dt + lam * nxx * d.dt; }
}

// ----------------------------------------------------------------------------------------
// Model preparation (once per sepfwi_set_model).  libCUFD.cu:71-77 (x MEGA), velInit /
// aveMuInit / aveBycInit utilities.cu:109-152 with the reference's double sub-expressions
// so the coefficient arrays are bit-identical to the reference's; defaults Model.cu:64-73.
// in: API arrays [nz][nx] dense.  maxcp: bit pattern of the largest Cp (positive floats order as ints).
template <bool SPONGE>
__global__ void k_model_prep(const Dims d, const float *lam_in, const float *mu_in, const float *rho_in,
                             float *model, int *maxcp)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, z = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= d.nx || z >= d.nz) return;
    const size_t k = (size_t)z * d.nx + x;
    const double scale = SPONGE ? 1.0 : 1e6;
    const float lam = (float)(lam_in[k] * scale), mu = (float)(mu_in[k] * scale);
    const float cp = (float)sqrt((lam + 2.0 * mu) / rho_in[k]);
    atomicMax(maxcp, __float_as_int(cp));
    if (z >= d.nzA) return;
    const size_t i = (size_t)z * d.ldx + x;
    model[M_LAM * d.fsz + i] = lam;
    model[M_MU * d.fsz + i] = mu;
    float mua = 0.f, bya = (float)(1.0 / 1000.0), byb = (float)(1.0 / 1000.0);
    if (z >= 2 && z <= d.nz - 3 && x >= 2 && x <= d.nx - 3) {
        const float m00 = mu, m10 = (float)(mu_in[k + d.nx] * scale), m01 = (float)(mu_in[k + 1] * scale),
                    m11 = (float)(mu_in[k + d.nx + 1] * scale);
        if (!(m00 == 0.0f || m10 == 0.0f || m01 == 0.0f || m11 == 0.0f))
            mua = (float)(4.0 / (1.0 / m00 + 1.0 / m10 + 1.0 / m01 + 1.0 / m11));
        const float rz = (float)(2.0 / (rho_in[k + d.nx] + rho_in[k]));   // z-neighbour average
        const float rx = (float)(2.0 / (rho_in[k + 1] + rho_in[k]));      // x-neighbour average
        // CPML flavour: vz uses the z average, vx the x average (utilities.cu:147-148);
        // sponge flavour has them swapped (elasticSolver.py:327-337).
        bya = SPONGE ? rx : rz;
        byb = SPONGE ? rz : rx;
    }
    model[M_MUAVE * d.fsz + i] = mua;
    model[M_BYCA * d.fsz + i] = bya;
    model[M_BYCB * d.fsz + i] = byb;
}

// out[z][x] (dense nz x nx) = sum over slots of grad[slot][which][z][x]; dead rows zero.
__global__ void k_grad_reduce(const Dims d, const float *grad, const int nslot, const int which, float *out)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, z = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= d.nx || z >= d.nz) return;
    float acc = 0.f;
    if (z < d.nzA)
        for (int s = 0; s < nslot; s++) acc += grad[((size_t)s * 3 + which) * d.fsz + (size_t)z * d.ldx + x];
    out[(size_t)z * d.nx + x] = acc;
}

}  // namespace sepfwi

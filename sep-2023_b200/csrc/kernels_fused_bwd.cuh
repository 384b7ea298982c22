// kernels_fused_bwd.cuh -- fused sm_100a kernels of the reverse-time sweep (default path).
//
// One backward time step = two launches that are independent of each other (they could share the
// GPU on two streams): both READ the adjoint state of buffer `pa` (the state after step it+1);
//
//   k_fused_recon  forward-field reconstruction  state(it+1) -> state(it)   [fwd buffer q -> q^1]
//                  v  -= D(sigma) b dt on the interior, ring restore of v      (el_velocity.cu:84-99, to_bnd)
//                  density imaging                                           (el_velocity.cu:100-110)
//                  sigma[src] -= amp ; sigma -= C D(v) dt, ring restore        (utilities.cu:541-550, el_stress.cu:89-103)
//                  lambda / mu imaging, four-point spray as a gather          (el_stress.cu:104-123)
//   k_fused_adj    adjoint sweep                    adj(it+1) -> adj(it)     [adj buffer pa -> pa^1]
//                  stf gradient                                              (utilities.cu:719-730)
//                  adjoint velocity update + its CPML memory                 (el_velocity_adj.cu:55-103)
//                  residual injection (deterministic gather per target cell) (utilities.cu:605-641)
//                  adjoint stress update + its CPML memory                   (el_stress_adj.cu:50-97)
//
// Same tiling as k_fused_fwd: a CTA owns a 16 x 64 tile, stages old fields + coefficients with
// cp.async, recomputes the first half step on a 2-cell halo, marches down columns with register
// windows.  All state is ping-pong; CPML memory variables are touched in the PML strips only and
// read straight from global memory there (3-16% of the cells).
#pragma once
#include "common.cuh"
#include "kernels_fused.cuh"

namespace sepfwi {

struct FusedBwdArgs {
    int it;       // reverse step: reconstructs state `it`, advances the adjoint to `it`
    int q;        // forward-field buffer that holds state it+1
    int pa;       // adjoint buffer that holds the adjoint state after step it+1
};

// ------------------------------------------------------------------------------------------------
// adjoint sweep
constexpr int A_SROWS = FTZ + 8;   // sigma^ tile rows  z0-4 .. z0+FTZ+3
constexpr int A_VROWS = FTZ + 4;   // v^ tile rows      z0-2 .. z0+FTZ+1   (also lambda, mu, mu_ave, buoyancies)
constexpr size_t A_SMEM = (size_t)(3 * A_SROWS + 7 * A_VROWS) * FW * sizeof(float);   // 61056 B

__global__ void __launch_bounds__(F_NT, 3) k_fused_adj(const KArgs a, const FusedBwdArgs fa)
{
    extern __shared__ __align__(16) float smem[];
    float *szz = smem, *sxz = szz + A_SROWS * FW, *sxx = sxz + A_SROWS * FW;
    float *svz = sxx + A_SROWS * FW, *svx = svz + A_VROWS * FW;
    float *slam = svx + A_VROWS * FW, *smu = slam + A_VROWS * FW, *smua = smu + A_VROWS * FW;
    float *sbya = smua + A_VROWS * FW, *sbyb = sbya + A_VROWS * FW;

    const Dims &d = a.d;
    const int tid = threadIdx.x, s = blockIdx.z;
    const int z0 = blockIdx.y * FTZ, x0 = blockIdx.x * FTX;
    const int ld = d.ldx;
    float *st = slot_state(a, s);
    const float *src = st + (size_t)(fa.pa ? S_ADJ1 : S_ADJ) * d.fsz;
    float *dst = st + (size_t)(fa.pa ? S_ADJ : S_ADJ1) * d.fsz;
    const float *psrc = st + (size_t)(fa.pa ? S_APSI1 : S_APSI) * d.fsz;
    float *pdst = st + (size_t)(fa.pa ? S_APSI : S_APSI1) * d.fsz;

    load_tile<A_SROWS, F_NT>(szz, src + (size_t)F_SZZ * d.fsz, z0 - 4, x0 - 4, d, tid);
    load_tile<A_SROWS, F_NT>(sxz, src + (size_t)F_SXZ * d.fsz, z0 - 4, x0 - 4, d, tid);
    load_tile<A_SROWS, F_NT>(sxx, src + (size_t)F_SXX * d.fsz, z0 - 4, x0 - 4, d, tid);
    load_tile<A_VROWS, F_NT>(svz, src + (size_t)F_VZ * d.fsz, z0 - 2, x0 - 4, d, tid);
    load_tile<A_VROWS, F_NT>(svx, src + (size_t)F_VX * d.fsz, z0 - 2, x0 - 4, d, tid);
    load_tile<A_VROWS, F_NT>(slam, a.model + (size_t)M_LAM * d.fsz, z0 - 2, x0 - 4, d, tid);
    load_tile<A_VROWS, F_NT>(smu, a.model + (size_t)M_MU * d.fsz, z0 - 2, x0 - 4, d, tid);
    load_tile<A_VROWS, F_NT>(smua, a.model + (size_t)M_MUAVE * d.fsz, z0 - 2, x0 - 4, d, tid);
    load_tile<A_VROWS, F_NT>(sbya, a.model + (size_t)M_BYCA * d.fsz, z0 - 2, x0 - 4, d, tid);
    load_tile<A_VROWS, F_NT>(sbyb, a.model + (size_t)M_BYCB * d.fsz, z0 - 2, x0 - 4, d, tid);
    cp_async_wait_all();
    __syncthreads();

    const int zs = a.t.zs[s], xs = a.t.xs[s];
    if (tid == 0 && zs >= z0 && zs < z0 + FTZ && xs >= x0 && xs < x0 + FTX) {
        const int si = (zs - z0 + 4) * FW + (xs - x0 + 4);
        a.gstf[(size_t)s * d.nSteps + fa.it] = -(szz[si] + a.t.rxz[s] * sxx[si]) * d.dt;
    }

    const int tx = tid % FW, g = tid / FW;

    // ---- A. adjoint velocity on the tile + 2-cell halo
    if (tx < FTX + 4) {
        const int cx = tx + 2, x = x0 - 2 + tx;
        constexpr int RP = A_VROWS / 4;   // 5
        const int j0 = g * RP;
        const bool xin = (x >= 2) && (x <= d.nx - 3);
        const bool xp = (x < d.nPml) || (x > d.nx - d.nPml - 1);
        float rKx = 1.f, ax = 0.f, rKxh = 1.f, axh = 0.f, bx = 0.f, bxh = 0.f;
        if (xin) {
            const float *c = a.cx + x;
            rKx = c[C_RK * d.nx]; ax = c[C_A * d.nx]; rKxh = c[C_RKH * d.nx]; axh = c[C_AH * d.nx];
            bx = c[C_B * d.nx]; bxh = c[C_BH * d.nx];
        }
        // sigma^-tile row of z is j+2: sxz needs rows j..j+3, szz / sxx rows j+1..j+4
        float xz0 = sxz[(j0 + 0) * FW + cx], xz1 = sxz[(j0 + 1) * FW + cx], xz2 = sxz[(j0 + 2) * FW + cx];
        float zz0 = szz[(j0 + 1) * FW + cx], zz1 = szz[(j0 + 2) * FW + cx], zz2 = szz[(j0 + 3) * FW + cx];
        float xx0 = sxx[(j0 + 1) * FW + cx], xx1 = sxx[(j0 + 2) * FW + cx], xx2 = sxx[(j0 + 3) * FW + cx];
#pragma unroll
        for (int k = 0; k < RP; k++) {
            const int j = j0 + k, z = z0 - 2 + j;
            const float xz3 = sxz[(j + 3) * FW + cx];
            const float zz3 = szz[(j + 4) * FW + cx], xx3 = sxx[(j + 4) * FW + cx];
            if (xin && z >= 2 && z <= d.nzA - 3) {
                const float *c = a.cz + z;
                const float rKz = c[C_RK * d.nzA], az = c[C_A * d.nzA], rKzh = c[C_RKH * d.nzA], azh = c[C_AH * d.nzA];
                const float *rzz = szz + (j + 2) * FW + cx, *rxx = sxx + (j + 2) * FW + cx, *rxz = sxz + (j + 2) * FW + cx;
                const int vi = j * FW + cx;
                const float lam = slam[vi], mu = smu[vi], mua = smua[vi];
                const float l2u = lam + 2.0f * mu;
                const size_t i = (size_t)z * ld + x;
                // -dxf(f) = -(c1 (f[x+1]-f[x]) - c2 (f[x+2]-f[x-1]))
                const float mdxf_zz = -(d.c1x * (rzz[1] - zz1) - d.c2x * (rzz[2] - rzz[-1]));
                const float mdxf_xx = -(d.c1x * (rxx[1] - xx1) - d.c2x * (rxx[2] - rxx[-1]));
                const float mdzb_xz = -(d.c1z * (xz2 - xz1) - d.c2z * (xz3 - xz0));
                float accx = (lam * mdxf_zz + l2u * mdxf_xx) * rKx * d.dt + mua * rKzh * mdzb_xz * d.dt;
                if (ax != 0.f) accx += ax * -dxf(psrc + (size_t)P_VX_X * d.fsz, i, d.c1x, d.c2x);
                if (azh != 0.f) accx += azh * -dzb(psrc + (size_t)P_VX_Z * d.fsz, i, ld, d.c1z, d.c2z);
                const float nvx = svx[vi] + accx;
                const float mdzf_zz = -(d.c1z * (zz2 - zz1) - d.c2z * (zz3 - zz0));
                const float mdzf_xx = -(d.c1z * (xx2 - xx1) - d.c2z * (xx3 - xx0));
                const float mdxb_xz = -(d.c1x * (xz2 - rxz[-1]) - d.c2x * (rxz[1] - rxz[-2]));
                float accz = (l2u * mdzf_zz + lam * mdzf_xx) * rKz * d.dt + mua * rKxh * mdxb_xz * d.dt;
                if (az != 0.f) accz += az * -dzf(psrc + (size_t)P_VZ_Z * d.fsz, i, ld, d.c1z, d.c2z);
                if (axh != 0.f) accz += axh * -dxb(psrc + (size_t)P_VZ_X * d.fsz, i, d.c1x, d.c2x);
                const float nvz = svz[vi] + accz;
                svx[vi] = nvx; svz[vi] = nvz;
                // CPML memory of the adjoint velocities (PML only).  Every CTA that recomputes a halo cell
                // writes the same value, so phase B may read them back from global memory after the barrier.
                const float bb = sbyb[vi], ba = sbya[vi];
                if (xp) {
                    pdst[(size_t)P_SXX_X * d.fsz + i] = bxh * psrc[(size_t)P_SXX_X * d.fsz + i] + bb * nvx * d.dt;
                    pdst[(size_t)P_SXZ_X * d.fsz + i] = bx * psrc[(size_t)P_SXZ_X * d.fsz + i] + ba * nvz * d.dt;
                }
                if ((z < d.nPml) || (z > d.nzA - d.nPml - 1)) {
                    pdst[(size_t)P_SXZ_Z * d.fsz + i] = c[C_B * d.nzA] * psrc[(size_t)P_SXZ_Z * d.fsz + i] + bb * nvx * d.dt;
                    pdst[(size_t)P_SZZ_Z * d.fsz + i] = c[C_BH * d.nzA] * psrc[(size_t)P_SZZ_Z * d.fsz + i] + ba * nvz * d.dt;
                }
            }
            xz0 = xz1; xz1 = xz2; xz2 = xz3;
            zz0 = zz1; zz1 = zz2; zz2 = zz3;
            xx0 = xx1; xx1 = xx2; xx2 = xx3;
        }
    }
    __syncthreads();

    // ---- A'. residual injection into the staged v^ (targets inside the tile + halo)
    {
        const int t = blockIdx.y * a.t.ntx + blockIdx.x;
        const int *tp = a.t.tileInjPtr + (size_t)s * (a.t.nTiles + 1);
        const int k0 = tp[t], k1 = tp[t + 1];
        if (k1 > k0) {
            const size_t tb = (size_t)s * a.t.maxInj, cb = (size_t)s * a.t.maxCon;
            const float *res = a.trace + ((size_t)s * d.nTrace + T_RES) * d.maxRec * d.nSteps + fa.it;
            for (int k = k0 + tid; k < k1; k += F_NT) {
                const int m = a.t.tileInj[(size_t)s * 4 * a.t.maxInj + k];
                const int cell = a.t.injCell[tb + m];
                const int z = cell / ld, x = cell - z * ld;
                const int p0 = a.t.injPtr[(size_t)s * (a.t.maxInj + 1) + m], p1 = a.t.injPtr[(size_t)s * (a.t.maxInj + 1) + m + 1];
                float *f = (a.t.injField[tb + m] == F_VZ ? svz : svx) + (z - z0 + 2) * FW + (x - x0 + 4);
                float v = *f;
                for (int p = p0; p < p1; p++) v += a.t.injCoef[cb + p] * res[(size_t)a.t.injRec[cb + p] * d.nSteps];
                *f = v;
            }
            __syncthreads();
        }
    }

    // ---- B. adjoint stress on the tile
    if (tx < FTX) {
        const int cx = tx + 4, x = x0 + tx;
        constexpr int RP = FTZ / 4;   // 4
        const int i0 = g * RP;
        const bool xin = (x >= 2) && (x <= d.nx - 3);
        const bool xst = (x < d.nPml + 2) || (x > d.nx - d.nPml - 3);
        float rKx = 1.f, ax = 0.f, rKxh = 1.f, axh = 0.f, bx = 0.f, bxh = 0.f;
        if (xin) {
            const float *c = a.cx + x;
            rKx = c[C_RK * d.nx]; ax = c[C_A * d.nx]; rKxh = c[C_RKH * d.nx]; axh = c[C_AH * d.nx];
            bx = c[C_B * d.nx]; bxh = c[C_BH * d.nx];
        }
        // v^-tile row of z is ii+2: vx needs rows ii+1..ii+4 (dzf), vz rows ii..ii+3 (dzb)
        float vx0 = svx[(i0 + 1) * FW + cx], vx1 = svx[(i0 + 2) * FW + cx], vx2 = svx[(i0 + 3) * FW + cx];
        float vz0 = svz[(i0 + 0) * FW + cx], vz1 = svz[(i0 + 1) * FW + cx], vz2 = svz[(i0 + 2) * FW + cx];
#pragma unroll
        for (int k = 0; k < RP; k++) {
            const int ii = i0 + k, z = z0 + ii;
            const float vx3 = svx[(ii + 4) * FW + cx];
            const float vz3 = svz[(ii + 3) * FW + cx];
            if (xin && z >= 2 && z <= d.nzA - 3) {
                const float *c = a.cz + z;
                const float rKz = c[C_RK * d.nzA], az = c[C_A * d.nzA], rKzh = c[C_RKH * d.nzA], azh = c[C_AH * d.nzA];
                const int vi = (ii + 2) * FW + cx, si = (ii + 4) * FW + cx;
                const float *rvz = svz + vi, *rvx = svx + vi;
                const float lam = slam[vi], mu = smu[vi], mua = smua[vi], ba = sbya[vi], bb = sbyb[vi];
                const float l2u = lam + 2.0f * mu;
                const size_t i = (size_t)z * ld + x;
                const float mdxf_vz = -(d.c1x * (rvz[1] - vz2) - d.c2x * (rvz[2] - rvz[-1]));
                const float mdzf_vx = -(d.c1z * (vx2 - vx1) - d.c2z * (vx3 - vx0));
                float acc = mdxf_vz * rKx * ba * d.dt + mdzf_vx * rKz * bb * d.dt;
                if (ax != 0.f) acc += ax * -dxf(pdst + (size_t)P_SXZ_X * d.fsz, i, d.c1x, d.c2x);
                if (az != 0.f) acc += az * -dzf(pdst + (size_t)P_SXZ_Z * d.fsz, i, ld, d.c1z, d.c2z);
                const float nxz = sxz[si] + acc;
                const float mdxb_vx = -(d.c1x * (vx1 - rvx[-1]) - d.c2x * (rvx[1] - rvx[-2]));
                const float mdzb_vz = -(d.c1z * (vz2 - vz1) - d.c2z * (vz3 - vz0));
                float accx = bb * mdxb_vx * rKxh * d.dt;
                if (axh != 0.f) accx += axh * -dxb(pdst + (size_t)P_SXX_X * d.fsz, i, d.c1x, d.c2x);
                float accz = ba * mdzb_vz * rKzh * d.dt;
                if (azh != 0.f) accz += azh * -dzb(pdst + (size_t)P_SZZ_Z * d.fsz, i, ld, d.c1z, d.c2z);
                const float nxx = sxx[si] + accx, nzz = szz[si] + accz;
                dst[(size_t)F_SXZ * d.fsz + i] = nxz; dst[(size_t)F_SXX * d.fsz + i] = nxx; dst[(size_t)F_SZZ * d.fsz + i] = nzz;
                dst[(size_t)F_VZ * d.fsz + i] = vz2; dst[(size_t)F_VX * d.fsz + i] = vx1;
                if (xst) {
                    pdst[(size_t)P_VZ_X * d.fsz + i] = bxh * psrc[(size_t)P_VZ_X * d.fsz + i] + nxz * mua * d.dt;
                    pdst[(size_t)P_VX_X * d.fsz + i] = bx * psrc[(size_t)P_VX_X * d.fsz + i] + lam * nzz * d.dt + l2u * nxx * d.dt;
                }
                if ((z < d.nPml + 2) || (z > d.nzA - d.nPml - 3)) {
                    pdst[(size_t)P_VX_Z * d.fsz + i] = c[C_BH * d.nzA] * psrc[(size_t)P_VX_Z * d.fsz + i] + nxz * mua * d.dt;
                    pdst[(size_t)P_VZ_Z * d.fsz + i] = c[C_B * d.nzA] * psrc[(size_t)P_VZ_Z * d.fsz + i] + l2u * nzz * d.dt + lam * nxx * d.dt;
                }
            }
            vx0 = vx1; vx1 = vx2; vx2 = vx3;
            vz0 = vz1; vz1 = vz2; vz2 = vz3;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// reverse-time reconstruction + imaging
constexpr int RW = FTX + 12;        // smem pitch: columns x0-8 .. x0+FTX+3
constexpr int R_SROWS = FTZ + 9;    // sigma tile rows  z0-5 .. z0+FTZ+3
constexpr int R_VROWS = FTZ + 5;    // v tile rows      z0-3 .. z0+FTZ+1   (also buoyancies)
constexpr int R_GROWS = FTZ + 1;    // rows z0-1 .. z0+FTZ-1: ga, gb, shear term, lambda, mu, mu_ave
constexpr size_t R_SMEM = (size_t)(3 * R_SROWS + 4 * R_VROWS + 6 * R_GROWS) * RW * sizeof(float);   // 79344 B

template <int ROWS, int NT>
__device__ __forceinline__ void load_tile_r(float *s, const float *g, int zfirst, int xfirst, const Dims &d, int tid)
{
    constexpr int V4 = RW / 4;
    for (int k = tid; k < ROWS * V4; k += NT) {
        const int r = k / V4, c4 = k - r * V4;
        const int z = zfirst + r, x = xfirst + c4 * 4;
        const bool ok = (z >= 0) && (z < d.nzA) && (x >= 0) && (x < d.ldx);
        cp_async16(s + r * RW + c4 * 4, ok ? g + (size_t)z * d.ldx + x : g, ok);
    }
}

__device__ __forceinline__ bool interior(const Dims &d, int z, int x)
{ return z >= d.nPml && z <= d.z1 && x >= d.nPml && x <= d.x1; }

__global__ void __launch_bounds__(F_NT, 2) k_fused_recon(const KArgs a, const FusedBwdArgs fa)
{
    extern __shared__ __align__(16) float smem[];
    float *szz = smem, *sxz = szz + R_SROWS * RW, *sxx = sxz + R_SROWS * RW;
    float *svz = sxx + R_SROWS * RW, *svx = svz + R_VROWS * RW;
    float *sbya = svx + R_VROWS * RW, *sbyb = sbya + R_VROWS * RW;
    float *sga = sbyb + R_VROWS * RW, *sgb = sga + R_GROWS * RW, *ssh = sgb + R_GROWS * RW;
    float *slam = ssh + R_GROWS * RW, *smu = slam + R_GROWS * RW, *smua = smu + R_GROWS * RW;

    const Dims &d = a.d;
    const int tid = threadIdx.x, s = blockIdx.z;
    // tiles cover the region interior U ring = [nPml-2, z1+2] x [nPml-2, x1+2]; x0 stays a multiple of 4
    const int zbase = d.nPml - 2, xbase = (d.nPml - 2) & ~3;
    const int z0 = zbase + blockIdx.y * FTZ, x0 = xbase + blockIdx.x * FTX;
    const int ld = d.ldx;
    float *st = slot_state(a, s);
    const float *src = st + (size_t)(fa.q ? S_FWD1 : S_FWD) * d.fsz;
    float *dst = st + (size_t)(fa.q ? S_FWD : S_FWD1) * d.fsz;
    const float *adj = st + (size_t)(fa.pa ? S_ADJ1 : S_ADJ) * d.fsz;
    float *grad = a.grad + (size_t)s * 3 * d.fsz;
    const float *ringb = a.ring + (((size_t)s * NFIELD) * d.nSteps + fa.it) * d.ringLen;
    const size_t rfs = (size_t)d.nSteps * d.ringLen;

    load_tile_r<R_SROWS, F_NT>(szz, src + (size_t)F_SZZ * d.fsz, z0 - 5, x0 - 8, d, tid);
    load_tile_r<R_SROWS, F_NT>(sxz, src + (size_t)F_SXZ * d.fsz, z0 - 5, x0 - 8, d, tid);
    load_tile_r<R_SROWS, F_NT>(sxx, src + (size_t)F_SXX * d.fsz, z0 - 5, x0 - 8, d, tid);
    load_tile_r<R_VROWS, F_NT>(svz, src + (size_t)F_VZ * d.fsz, z0 - 3, x0 - 8, d, tid);
    load_tile_r<R_VROWS, F_NT>(svx, src + (size_t)F_VX * d.fsz, z0 - 3, x0 - 8, d, tid);
    load_tile_r<R_VROWS, F_NT>(sbya, a.model + (size_t)M_BYCA * d.fsz, z0 - 3, x0 - 8, d, tid);
    load_tile_r<R_VROWS, F_NT>(sbyb, a.model + (size_t)M_BYCB * d.fsz, z0 - 3, x0 - 8, d, tid);
    load_tile_r<R_GROWS, F_NT>(slam, a.model + (size_t)M_LAM * d.fsz, z0 - 1, x0 - 8, d, tid);
    load_tile_r<R_GROWS, F_NT>(smu, a.model + (size_t)M_MU * d.fsz, z0 - 1, x0 - 8, d, tid);
    load_tile_r<R_GROWS, F_NT>(smua, a.model + (size_t)M_MUAVE * d.fsz, z0 - 1, x0 - 8, d, tid);
    cp_async_wait_all();
    __syncthreads();

    const int tx = tid % FW, g = tid / FW;
    const bool ring = tile_touches_ring_ext(d, z0 - 3, z0 + FTZ + 1, x0 - 3, x0 + FTX + 1);

    // ---- 1. velocities at time `it` on rows z0-3 .. z0+FTZ+1, columns x0-3 .. x0+FTX+1
    if (tx < FTX + 5) {
        const int cx = tx + 5, x = x0 - 3 + tx;       // smem column of x is x - (x0-8)
        constexpr int RP = (R_VROWS + 3) / 4;         // 6 rows per group (last group shorter)
        const int j0 = g * RP;
        // sigma-tile row of v-row j is j+2: szz needs rows j+1..j+4, sxz rows j..j+3
        float zz0 = szz[(j0 + 1) * RW + cx], zz1 = szz[(j0 + 2) * RW + cx], zz2 = szz[(j0 + 3) * RW + cx];
        float xz0 = sxz[(j0 + 0) * RW + cx], xz1 = sxz[(j0 + 1) * RW + cx], xz2 = sxz[(j0 + 2) * RW + cx];
#pragma unroll
        for (int k = 0; k < RP; k++) {
            const int j = j0 + k, z = z0 - 3 + j;
            if (j < R_VROWS) {
                const float zz3 = szz[(j + 4) * RW + cx];
                const float xz3 = sxz[(j + 3) * RW + cx];
                const int vi = j * RW + cx;
                float nvz = svz[vi], nvx = svx[vi];
                const bool in = interior(d, z, x);
                if (in) {
                    const float *rxz = sxz + (j + 2) * RW + cx, *rxx = sxx + (j + 2) * RW + cx;
                    const float A = (d.c1z * (zz2 - zz1) - d.c2z * (zz3 - zz0)) + (d.c1x * (xz2 - rxz[-1]) - d.c2x * (rxz[1] - rxz[-2]));
                    const float B = (d.c1z * (xz2 - xz1) - d.c2z * (xz3 - xz0)) + (d.c1x * (rxx[1] - rxx[0]) - d.c2x * (rxx[2] - rxx[-1]));
                    const float ba = sbya[vi], bb = sbyb[vi];
                    nvz -= A * ba * d.dt;
                    nvx -= B * bb * d.dt;
                    if (j >= 2 && j < R_GROWS + 2 && tx >= 2 && tx < FTX + 3) {     // rows z0-1.., columns x0-1..
                        const size_t i = (size_t)z * ld + x;
                        const int gi = (j - 2) * RW + cx;
                        sga[gi] = adj[(size_t)F_VZ * d.fsz + i] * A * d.dt * (0.5f * ba * ba);
                        sgb[gi] = adj[(size_t)F_VX * d.fsz + i] * B * d.dt * (0.5f * bb * bb);
                    }
                } else if (j >= 2 && j < R_GROWS + 2 && tx >= 2 && tx < FTX + 3) {
                    const int gi = (j - 2) * RW + cx;
                    sga[gi] = 0.f; sgb[gi] = 0.f;
                }
                if (ring) {
                    int ridx[4];
                    if (ring_indices(d, z, x, ridx) > 0) { nvz = ringb[F_VZ * rfs + ridx[0]]; nvx = ringb[F_VX * rfs + ridx[0]]; }
                }
                svz[vi] = nvz; svx[vi] = nvx;
                zz0 = zz1; zz1 = zz2; zz2 = zz3;
                xz0 = xz1; xz1 = xz2; xz2 = xz3;
            }
        }
    }
    __syncthreads();

    // ---- 2a. shear imaging term of the staggered cells rows z0-1 .. z0+FTZ-1, columns x0-1 .. x0+FTX-1
    if (tx < FTX + 1) {
        const int cx = tx + 7, x = x0 - 1 + tx;
        constexpr int RP = (R_GROWS + 3) / 4;         // 5
        const int j0 = g * RP;
        for (int k = 0; k < RP; k++) {
            const int jg = j0 + k, z = z0 - 1 + jg;   // v-tile row of z is jg+2
            if (jg < R_GROWS) {
                float sh = 0.f;
                if (interior(d, z, x)) {
                    const int gi = jg * RW + cx;
                    const float mua = smua[gi];
                    if (mua != 0.0f) {
                        const float *cvx = svx + (jg + 2) * RW + cx, *cvz = svz + (jg + 2) * RW + cx;
                        const float D3 = (d.c1z * (cvx[RW] - cvx[0]) - d.c2z * (cvx[2 * RW] - cvx[-RW])) +
                                         (d.c1x * (cvz[1] - cvz[0]) - d.c2x * (cvz[2] - cvz[-1]));
                        // mu_ave / sum(1/mu) = mu_ave^2 / 4  (mu_ave is the 4-point harmonic mean)
                        sh = -adj[(size_t)F_SXZ * d.fsz + (size_t)z * ld + x] * D3 * d.dt * (0.25f * mua * mua) * 1e6f;
                    }
                }
                ssh[jg * RW + cx] = sh;
            }
        }
    }
    __syncthreads();

    // ---- 2b. stresses at time `it` and the three gradients on the tile
    if (tx < FTX) {
        const int cx = tx + 8, x = x0 + tx;
        constexpr int RP = FTZ / 4;   // 4
        const int i0 = g * RP;
        const int zs = a.t.zs[s], xs = a.t.xs[s];
        for (int k = 0; k < RP; k++) {
            const int ii = i0 + k, z = z0 + ii;
            if (z <= d.z1 + 2 && x >= d.nPml - 2 && x <= d.x1 + 2) {
                const int si = (ii + 5) * RW + cx, vi = (ii + 3) * RW + cx, gi = (ii + 1) * RW + cx;
                const size_t i = (size_t)z * ld + x;
                float nzz = szz[si], nxx = sxx[si], nxz = sxz[si];
                if (z == zs && x == xs) {
                    const float amp = a.t.amp[(size_t)s * d.nSteps + fa.it];
                    nzz -= amp; nxx -= amp;
                }
                const bool in = interior(d, z, x);
                float gr = 0.f, gm = 0.f;
                if (in) {
                    const float *cvz = svz + vi, *cvx = svx + vi;
                    const float D1 = d.c1z * (cvz[0] - cvz[-RW]) - d.c2z * (cvz[RW] - cvz[-2 * RW]);
                    const float D2 = d.c1x * (cvx[0] - cvx[-1]) - d.c2x * (cvx[1] - cvx[-2]);
                    const float D3 = (d.c1z * (cvx[RW] - cvx[0]) - d.c2z * (cvx[2 * RW] - cvx[-RW])) +
                                     (d.c1x * (cvz[1] - cvz[0]) - d.c2x * (cvz[2] - cvz[-1]));
                    const float lam = slam[gi], mu = smu[gi], mua = smua[gi];
                    const float l2u = lam + 2.0f * mu;
                    nzz -= (l2u * D1 + lam * D2) * d.dt;
                    nxx -= (lam * D1 + l2u * D2) * d.dt;
                    nxz -= mua * D3 * d.dt;
                    const float za = adj[(size_t)F_SZZ * d.fsz + i], xa = adj[(size_t)F_SXX * d.fsz + i];
                    grad[0 * d.fsz + i] += -(za + xa) * (D1 + D2) * d.dt * 1e6f;
                    gm = (-2.0f * za * D1 * d.dt - 2.0f * xa * D2 * d.dt) * 1e6f;
                    gr = sga[gi] + sgb[gi];
                }
                // gathers of the reference's sprays (el_stress.cu:116-123, el_velocity.cu:104-110)
                float sh = ssh[gi] + ssh[gi - 1];
                if (z <= d.z1) { sh += ssh[gi - RW]; gr += sga[gi - RW]; if (x <= d.x1) sh += ssh[gi - RW - 1]; }
                gr += sgb[gi - 1];
                if (in || sh != 0.f) { const float mu = smu[gi]; grad[1 * d.fsz + i] += gm + sh / (mu * mu); }
                if (in || gr != 0.f) grad[2 * d.fsz + i] += gr;
                if (ring) {
                    int ridx[4];
                    if (ring_indices(d, z, x, ridx) > 0) { nzz = ringb[F_SZZ * rfs + ridx[0]]; nxz = ringb[F_SXZ * rfs + ridx[0]]; nxx = ringb[F_SXX * rfs + ridx[0]]; }
                }
                dst[(size_t)F_SZZ * d.fsz + i] = nzz; dst[(size_t)F_SXX * d.fsz + i] = nxx; dst[(size_t)F_SXZ * d.fsz + i] = nxz;
                dst[(size_t)F_VZ * d.fsz + i] = svz[vi]; dst[(size_t)F_VX * d.fsz + i] = svx[vi];
            }
        }
    }
}

}  // namespace sepfwi

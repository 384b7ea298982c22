// kernels_fused_bwd.cuh -- fused sm_100a kernels of the reverse-time sweep (default path).
//
// One backward time step = two launches that are independent of each other: both READ the adjoint
// state of buffer `pa` (the state after step it+1);
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
// Same structure as k_fused_fwd: a CTA owns a 16 x 64 tile, stages old fields + coefficients with
// cp.async, recomputes the first half step on a halo, and every thread owns a quad of four
// x-consecutive cells (128-bit LDS / STS / STG).  Tiles away from the PML, the rim and the ring
// take a branch-free path; the edge tiles run the same arithmetic per component plus the CPML /
// ring / spray-boundary handling.  All state is ping-pong; the adjoint CPML memory variables are
// touched in the PML strips only and read straight from global memory there.
#pragma once
#include "common.cuh"
#include "kernels_fused.cuh"

namespace sepfwi {

struct FusedBwdArgs {
    int it;       // reverse step: reconstructs state `it`, advances the adjoint to `it`
    int q;        // forward-field buffer that holds state it+1
    int pa;       // adjoint buffer that holds the adjoint state after step it+1
};

#define Q4(v) {v.x, v.y, v.z, v.w}
// 7-value x windows around a quad c given its left / right neighbour quads l, r
#define WIN_M2(l, c, r) {l.z, l.w, c.x, c.y, c.z, c.w, r.x}   // f[x-2 .. x+4]  -> backward difference: taps k .. k+3
#define WIN_M1(l, c, r) {l.w, c.x, c.y, c.z, c.w, r.x, r.y}   // f[x-1 .. x+5]  -> forward  difference: taps k .. k+3
// 4th-order difference from a 7-window at component k (same tap pattern for both window kinds)
#define DIFFX(w, k) (d.c1x * (w[(k) + 2] - w[(k) + 1]) - d.c2x * (w[(k) + 3] - w[k]))
// z differences from four row values: backward (z-2, z-1, z, z+1) and forward (z-1, z, z+1, z+2)
#define DIFFZ(m1, c0, p1, o) (d.c1z * ((p1) - (c0)) - d.c2z * ((o) - (m1)))

// ------------------------------------------------------------------------------------------------
// adjoint sweep
constexpr int A_SROWS = FTZ + 8;   // sigma^ tile rows  z0-4 .. z0+FTZ+3
constexpr int A_VROWS = FTZ + 4;   // v^ tile rows      z0-2 .. z0+FTZ+1   (also lambda, mu, mu_ave)
constexpr size_t A_SMEM = (size_t)(3 * A_SROWS + 5 * A_VROWS + 2 * FTZ) * FW * sizeof(float);   // 58752 B

template <bool EDGE>
__device__ __forceinline__ void fused_adj_body(const KArgs &a, const FusedBwdArgs &fa, float *smem, int z0, int x0, int s)
{
    float *szz = smem, *sxz = szz + A_SROWS * FW, *sxx = sxz + A_SROWS * FW;
    float *svz = sxx + A_SROWS * FW, *svx = svz + A_VROWS * FW;
    float *slam = svx + A_VROWS * FW, *smu = slam + A_VROWS * FW, *smua = smu + A_VROWS * FW;
    float *sbya = smua + A_VROWS * FW, *sbyb = sbya + FTZ * FW;

    const Dims &d = a.d;
    const int tid = threadIdx.x;
    const int ld = d.ldx;
    float *st = slot_state(a, s);
    const float *src = st + (size_t)(fa.pa ? S_ADJ1 : S_ADJ) * d.fsz;
    float *dst = st + (size_t)(fa.pa ? S_ADJ : S_ADJ1) * d.fsz;
    const float *psrc = st + (size_t)(fa.pa ? S_APSI1 : S_APSI) * d.fsz;
    float *pdst = st + (size_t)(fa.pa ? S_APSI : S_APSI1) * d.fsz;

    if (tid < F4_LD) {
        const int r = tid / FQ, c4 = tid - r * FQ;
        const int xq = x0 - 4 + 4 * c4;
        load_rows<A_SROWS, EDGE>(szz, src + (size_t)F_SZZ * d.fsz, z0 - 4, xq, r, c4, d);
        load_rows<A_SROWS, EDGE>(sxz, src + (size_t)F_SXZ * d.fsz, z0 - 4, xq, r, c4, d);
        load_rows<A_SROWS, EDGE>(sxx, src + (size_t)F_SXX * d.fsz, z0 - 4, xq, r, c4, d);
        load_rows<A_VROWS, EDGE>(svz, src + (size_t)F_VZ * d.fsz, z0 - 2, xq, r, c4, d);
        load_rows<A_VROWS, EDGE>(svx, src + (size_t)F_VX * d.fsz, z0 - 2, xq, r, c4, d);
        load_rows<A_VROWS, EDGE>(slam, a.model + (size_t)M_LAM * d.fsz, z0 - 2, xq, r, c4, d);
        load_rows<A_VROWS, EDGE>(smu, a.model + (size_t)M_MU * d.fsz, z0 - 2, xq, r, c4, d);
        load_rows<A_VROWS, EDGE>(smua, a.model + (size_t)M_MUAVE * d.fsz, z0 - 2, xq, r, c4, d);
        load_rows<FTZ, EDGE>(sbya, a.model + (size_t)M_BYCA * d.fsz, z0, xq, r, c4, d);
        load_rows<FTZ, EDGE>(sbyb, a.model + (size_t)M_BYCB * d.fsz, z0, xq, r, c4, d);
    }
    cp_async_wait_all();
    __syncthreads();

    const int zs = a.t.zs[s], xs = a.t.xs[s];
    if (tid == 0 && zs >= z0 && zs < z0 + FTZ && xs >= x0 && xs < x0 + FTX) {
        const int si = (zs - z0 + 4) * FW + (xs - x0 + 4);
        a.gstf[(size_t)s * d.nSteps + fa.it] = -(szz[si] + a.t.rxz[s] * sxx[si]) * d.dt;
    }

    // ---- A. adjoint velocity on rows z0-2 .. z0+FTZ+1, all 18 staged quads
    if (tid < A_VROWS * FQ) {
        const int j = tid / FQ, c4 = tid - j * FQ;
        const int z = z0 - 2 + j, xq = x0 - 4 + 4 * c4;
        const int cl = max(c4 - 1, 0), cr = min(c4 + 1, FQ - 1);
        const bool zin = !EDGE || (z >= 2 && z <= d.nzA - 3);
        const size_t iq = (size_t)z * ld + xq;
        float rKz = 1.f, az = 0.f, rKzh = 1.f, azh = 0.f, bz = 0.f, bzh = 0.f;
        const bool zp = EDGE && zin && ((z < d.nPml) || (z > d.nzA - d.nPml - 1));
        if (EDGE && zin) {
            const float *c = a.cz + z;
            rKz = c[C_RK * d.nzA]; az = c[C_A * d.nzA]; rKzh = c[C_RKH * d.nzA]; azh = c[C_AH * d.nzA];
            bz = c[C_B * d.nzA]; bzh = c[C_BH * d.nzA];
        }
        const float4 lam4 = ld4(slam, j, c4), mu4 = ld4(smu, j, c4), mua4 = ld4(smua, j, c4);
        const float lam[4] = Q4(lam4), mu[4] = Q4(mu4), mua[4] = Q4(mua4);
        float nvx[4], nvz[4];
        {   // v^x : -dxf(szz^), -dxf(sxx^), -dzb(sxz^)
            const float4 zc = ld4(szz, j + 2, c4), zl = ld4(szz, j + 2, cl), zr = ld4(szz, j + 2, cr);
            const float4 xc = ld4(sxx, j + 2, c4), xl = ld4(sxx, j + 2, cl), xr = ld4(sxx, j + 2, cr);
            const float4 q0 = ld4(sxz, j, c4), q1 = ld4(sxz, j + 1, c4), q2 = ld4(sxz, j + 2, c4), q3 = ld4(sxz, j + 3, c4);
            const float wzz[7] = WIN_M1(zl, zc, zr), wxx[7] = WIN_M1(xl, xc, xr);
            const float m2[4] = Q4(q0), m1[4] = Q4(q1), c0[4] = Q4(q2), p1[4] = Q4(q3);
            const float4 ov4 = ld4(svx, j, c4);
            const float ov[4] = Q4(ov4);
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int x = xq + c;
                float acc;
                if (!EDGE) {
                    acc = (lam[c] * -DIFFX(wzz, c) + (lam[c] + 2.0f * mu[c]) * -DIFFX(wxx, c)) * d.dt + mua[c] * -DIFFZ(m2[c], m1[c], c0[c], p1[c]) * d.dt;
                    nvx[c] = ov[c] + acc;
                } else {
                    nvx[c] = ov[c];
                    if (zin && x >= 2 && x <= d.nx - 3) {
                        const float *cxp = a.cx + x;
                        const float rKx = cxp[C_RK * d.nx], ax = cxp[C_A * d.nx];
                        const size_t i = iq + c;
                        acc = (lam[c] * -DIFFX(wzz, c) + (lam[c] + 2.0f * mu[c]) * -DIFFX(wxx, c)) * rKx * d.dt +
                              mua[c] * rKzh * -DIFFZ(m2[c], m1[c], c0[c], p1[c]) * d.dt;
                        if (ax != 0.f) acc += ax * -dxf(psrc + (size_t)P_VX_X * d.fsz, i, d.c1x, d.c2x);
                        if (azh != 0.f) acc += azh * -dzb(psrc + (size_t)P_VX_Z * d.fsz, i, ld, d.c1z, d.c2z);
                        nvx[c] = ov[c] + acc;
                    }
                }
            }
        }
        {   // v^z : -dzf(szz^), -dzf(sxx^), -dxb(sxz^)
            const float4 z0r = ld4(szz, j + 1, c4), z1r = ld4(szz, j + 2, c4), z2r = ld4(szz, j + 3, c4), z3r = ld4(szz, j + 4, c4);
            const float4 x0r = ld4(sxx, j + 1, c4), x1r = ld4(sxx, j + 2, c4), x2r = ld4(sxx, j + 3, c4), x3r = ld4(sxx, j + 4, c4);
            const float4 qc = ld4(sxz, j + 2, c4), ql = ld4(sxz, j + 2, cl), qr = ld4(sxz, j + 2, cr);
            const float wxz[7] = WIN_M2(ql, qc, qr);
            const float zm1[4] = Q4(z0r), zc0[4] = Q4(z1r), zp1[4] = Q4(z2r), zp2[4] = Q4(z3r);
            const float xm1[4] = Q4(x0r), xc0[4] = Q4(x1r), xp1[4] = Q4(x2r), xp2[4] = Q4(x3r);
            const float4 ov4 = ld4(svz, j, c4);
            const float ov[4] = Q4(ov4);
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int x = xq + c;
                float acc;
                if (!EDGE) {
                    acc = ((lam[c] + 2.0f * mu[c]) * -DIFFZ(zm1[c], zc0[c], zp1[c], zp2[c]) + lam[c] * -DIFFZ(xm1[c], xc0[c], xp1[c], xp2[c])) * d.dt +
                          mua[c] * -DIFFX(wxz, c) * d.dt;
                    nvz[c] = ov[c] + acc;
                } else {
                    nvz[c] = ov[c];
                    if (zin && x >= 2 && x <= d.nx - 3) {
                        const float *cxp = a.cx + x;
                        const float rKxh = cxp[C_RKH * d.nx], axh = cxp[C_AH * d.nx];
                        const size_t i = iq + c;
                        acc = ((lam[c] + 2.0f * mu[c]) * -DIFFZ(zm1[c], zc0[c], zp1[c], zp2[c]) + lam[c] * -DIFFZ(xm1[c], xc0[c], xp1[c], xp2[c])) * rKz * d.dt +
                              mua[c] * rKxh * -DIFFX(wxz, c) * d.dt;
                        if (az != 0.f) acc += az * -dzf(psrc + (size_t)P_VZ_Z * d.fsz, i, ld, d.c1z, d.c2z);
                        if (axh != 0.f) acc += axh * -dxb(psrc + (size_t)P_VZ_X * d.fsz, i, d.c1x, d.c2x);
                        nvz[c] = ov[c] + acc;
                        // CPML memory of the adjoint velocities (PML only).  Every CTA that recomputes a halo cell writes
                        // the same value, so phase B may read them back from global memory after the barrier.
                        // (only for the columns x0-2 .. x0+FTX+1 whose stencil windows are complete)
                        const bool xvalid = (x >= x0 - 2) && (x <= x0 + FTX + 1);
                        const float bb = a.model[(size_t)M_BYCB * d.fsz + i], ba = a.model[(size_t)M_BYCA * d.fsz + i];
                        if (xvalid && ((x < d.nPml) || (x > d.nx - d.nPml - 1))) {
                            pdst[(size_t)P_SXX_X * d.fsz + i] = cxp[C_BH * d.nx] * psrc[(size_t)P_SXX_X * d.fsz + i] + bb * nvx[c] * d.dt;
                            pdst[(size_t)P_SXZ_X * d.fsz + i] = cxp[C_B * d.nx] * psrc[(size_t)P_SXZ_X * d.fsz + i] + ba * nvz[c] * d.dt;
                        }
                        if (xvalid && zp) {
                            pdst[(size_t)P_SXZ_Z * d.fsz + i] = bz * psrc[(size_t)P_SXZ_Z * d.fsz + i] + bb * nvx[c] * d.dt;
                            pdst[(size_t)P_SZZ_Z * d.fsz + i] = bzh * psrc[(size_t)P_SZZ_Z * d.fsz + i] + ba * nvz[c] * d.dt;
                        }
                    }
                }
            }
        }
        st4(svx, j, c4, make_float4(nvx[0], nvx[1], nvx[2], nvx[3]));
        st4(svz, j, c4, make_float4(nvz[0], nvz[1], nvz[2], nvz[3]));
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
            for (int k = k0 + tid; k < k1; k += F4_NT) {
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

    // ---- B. adjoint stress on the owned quads
    if (tid < FTZ * (FTX / 4)) {
        const int ii = tid / (FTX / 4), c4 = tid - ii * (FTX / 4) + 1;
        const int z = z0 + ii, xq = x0 - 4 + 4 * c4;
        const size_t iq = (size_t)z * ld + xq;
        const bool zin = !EDGE || (z >= 2 && z <= d.nzA - 3);
        // v^-tile row of z is ii+2, sigma^-tile row ii+4
        const float4 v0 = ld4(svz, ii, c4), v1 = ld4(svz, ii + 1, c4), v2 = ld4(svz, ii + 2, c4), v3 = ld4(svz, ii + 3, c4);
        const float4 vl = ld4(svz, ii + 2, c4 - 1), vr = ld4(svz, ii + 2, c4 + 1);
        const float4 u0 = ld4(svx, ii + 1, c4), u1 = ld4(svx, ii + 2, c4), u2 = ld4(svx, ii + 3, c4), u3 = ld4(svx, ii + 4, c4);
        const float4 ul = ld4(svx, ii + 2, c4 - 1), ur = ld4(svx, ii + 2, c4 + 1);
        const float wvz[7] = WIN_M1(vl, v2, vr), wvx[7] = WIN_M2(ul, u1, ur);
        const float vzm2[4] = Q4(v0), vzm1[4] = Q4(v1), vzc[4] = Q4(v2), vzp1[4] = Q4(v3);
        const float vxm1[4] = Q4(u0), vxc[4] = Q4(u1), vxp1[4] = Q4(u2), vxp2[4] = Q4(u3);
        const float4 ozz4 = ld4(szz, ii + 4, c4), oxz4 = ld4(sxz, ii + 4, c4), oxx4 = ld4(sxx, ii + 4, c4);
        const float ozz[4] = Q4(ozz4), oxz[4] = Q4(oxz4), oxx[4] = Q4(oxx4);
        const float4 lam4 = ld4(slam, ii + 2, c4), mu4 = ld4(smu, ii + 2, c4), mua4 = ld4(smua, ii + 2, c4);
        const float4 ba4 = ld4(sbya, ii, c4), bb4 = ld4(sbyb, ii, c4);
        const float lam[4] = Q4(lam4), mu[4] = Q4(mu4), mua[4] = Q4(mua4), ba[4] = Q4(ba4), bb[4] = Q4(bb4);
        float rKz = 1.f, az = 0.f, rKzh = 1.f, azh = 0.f, bz = 0.f, bzh = 0.f;
        if (EDGE && zin) {
            const float *c = a.cz + z;
            rKz = c[C_RK * d.nzA]; az = c[C_A * d.nzA]; rKzh = c[C_RKH * d.nzA]; azh = c[C_AH * d.nzA];
            bz = c[C_B * d.nzA]; bzh = c[C_BH * d.nzA];
        }
        const bool zst = EDGE && ((z < d.nPml + 2) || (z > d.nzA - d.nPml - 3));
        float nzz[4], nxz[4], nxx[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int x = xq + c;
            const float mdxf_vz = -DIFFX(wvz, c), mdzf_vx = -DIFFZ(vxm1[c], vxc[c], vxp1[c], vxp2[c]);
            const float mdxb_vx = -DIFFX(wvx, c), mdzb_vz = -DIFFZ(vzm2[c], vzm1[c], vzc[c], vzp1[c]);
            if (!EDGE) {
                nxz[c] = oxz[c] + (mdxf_vz * ba[c] * d.dt + mdzf_vx * bb[c] * d.dt);
                nxx[c] = oxx[c] + bb[c] * mdxb_vx * d.dt;
                nzz[c] = ozz[c] + ba[c] * mdzb_vz * d.dt;
            } else {
                nxz[c] = oxz[c]; nxx[c] = oxx[c]; nzz[c] = ozz[c];
                if (zin && x >= 2 && x <= d.nx - 3) {
                    const float *cxp = a.cx + x;
                    const float rKx = cxp[C_RK * d.nx], ax = cxp[C_A * d.nx], rKxh = cxp[C_RKH * d.nx], axh = cxp[C_AH * d.nx];
                    const size_t i = iq + c;
                    float acc = mdxf_vz * rKx * ba[c] * d.dt + mdzf_vx * rKz * bb[c] * d.dt;
                    if (ax != 0.f) acc += ax * -dxf(pdst + (size_t)P_SXZ_X * d.fsz, i, d.c1x, d.c2x);
                    if (az != 0.f) acc += az * -dzf(pdst + (size_t)P_SXZ_Z * d.fsz, i, ld, d.c1z, d.c2z);
                    nxz[c] = oxz[c] + acc;
                    float accx = bb[c] * mdxb_vx * rKxh * d.dt;
                    if (axh != 0.f) accx += axh * -dxb(pdst + (size_t)P_SXX_X * d.fsz, i, d.c1x, d.c2x);
                    float accz = ba[c] * mdzb_vz * rKzh * d.dt;
                    if (azh != 0.f) accz += azh * -dzb(pdst + (size_t)P_SZZ_Z * d.fsz, i, ld, d.c1z, d.c2z);
                    nxx[c] = oxx[c] + accx; nzz[c] = ozz[c] + accz;
                    const float l2u = lam[c] + 2.0f * mu[c];
                    if ((x < d.nPml + 2) || (x > d.nx - d.nPml - 3)) {
                        pdst[(size_t)P_VZ_X * d.fsz + i] = cxp[C_BH * d.nx] * psrc[(size_t)P_VZ_X * d.fsz + i] + nxz[c] * mua[c] * d.dt;
                        pdst[(size_t)P_VX_X * d.fsz + i] = cxp[C_B * d.nx] * psrc[(size_t)P_VX_X * d.fsz + i] + lam[c] * nzz[c] * d.dt + l2u * nxx[c] * d.dt;
                    }
                    if (zst) {
                        pdst[(size_t)P_VX_Z * d.fsz + i] = bzh * psrc[(size_t)P_VX_Z * d.fsz + i] + nxz[c] * mua[c] * d.dt;
                        pdst[(size_t)P_VZ_Z * d.fsz + i] = bz * psrc[(size_t)P_VZ_Z * d.fsz + i] + l2u * nzz[c] * d.dt + lam[c] * nxx[c] * d.dt;
                    }
                }
            }
        }
        if (!EDGE || (z < d.nzA && xq < d.ldx)) {
            // cells outside the active region keep their (zero) value, so whole quads can be stored
            *reinterpret_cast<float4 *>(dst + (size_t)F_SZZ * d.fsz + iq) = make_float4(nzz[0], nzz[1], nzz[2], nzz[3]);
            *reinterpret_cast<float4 *>(dst + (size_t)F_SXZ * d.fsz + iq) = make_float4(nxz[0], nxz[1], nxz[2], nxz[3]);
            *reinterpret_cast<float4 *>(dst + (size_t)F_SXX * d.fsz + iq) = make_float4(nxx[0], nxx[1], nxx[2], nxx[3]);
            *reinterpret_cast<float4 *>(dst + (size_t)F_VZ * d.fsz + iq) = v2;
            *reinterpret_cast<float4 *>(dst + (size_t)F_VX * d.fsz + iq) = u1;
        }
    }
}

__global__ void __launch_bounds__(F4_NT, 3) k_fused_adj(const KArgs a, const FusedBwdArgs fa)
{
    extern __shared__ __align__(16) float smem[];
    const Dims &d = a.d;
    const int z0 = blockIdx.y * FTZ, x0 = blockIdx.x * FTX;
    // interior tile: the staged v^ region lies outside the PML strips (nPml + 2 wide) and the inactive rim
    const bool inner = (z0 - 2 >= d.nPml + 2) && (z0 + FTZ + 1 <= d.nzA - d.nPml - 3) && (x0 - 4 >= d.nPml + 2) &&
                       (x0 + FTX + 3 <= d.nx - d.nPml - 3);
    if (inner) fused_adj_body<false>(a, fa, smem, z0, x0, blockIdx.z);
    else fused_adj_body<true>(a, fa, smem, z0, x0, blockIdx.z);
}

// ------------------------------------------------------------------------------------------------
// reverse-time reconstruction + imaging
constexpr int RW = FTX + 16;        // smem pitch: columns x0-8 .. x0+FTX+7  (20 quads)
constexpr int RQ = RW / 4;
constexpr int R_SROWS = FTZ + 9;    // sigma tile rows  z0-5 .. z0+FTZ+3
constexpr int R_VROWS = FTZ + 5;    // v tile rows      z0-3 .. z0+FTZ+1   (also buoyancies)
constexpr int R_GROWS = FTZ + 1;    // rows z0-1 .. z0+FTZ-1: ga, gb, shear term, lambda, mu, mu_ave
constexpr size_t R_SMEM = (size_t)(3 * R_SROWS + 4 * R_VROWS + 6 * R_GROWS) * RW * sizeof(float);   // 83520 B
constexpr int R_LD = 19 * RQ;       // 380 loader threads: +380 in the quad index = +19 rows, same column

template <int ROWS, bool EDGE>
__device__ __forceinline__ void load_rows_r(float *s, const float *g, int zfirst, int xq, int r, int c4, const Dims &d)
{
#pragma unroll
    for (int rr = 0; rr < ROWS; rr += 19) {
        const int row = r + rr;
        if (row < ROWS) {
            const int z = zfirst + row;
            if (EDGE) {
                const bool ok = (z >= 0) && (z < d.nzA) && (xq >= 0) && (xq < d.ldx);
                cp_async16(s + row * RW + c4 * 4, ok ? g + (size_t)z * d.ldx + xq : g, ok);
            } else {
                cp_async16(s + row * RW + c4 * 4, g + (size_t)z * d.ldx + xq, true);
            }
        }
    }
}
__device__ __forceinline__ float4 ld4r(const float *s, int row, int c4) { return *reinterpret_cast<const float4 *>(s + row * RW + c4 * 4); }
__device__ __forceinline__ void st4r(float *s, int row, int c4, const float4 &v) { *reinterpret_cast<float4 *>(s + row * RW + c4 * 4) = v; }

__device__ __forceinline__ bool interior(const Dims &d, int z, int x)
{ return z >= d.nPml && z <= d.z1 && x >= d.nPml && x <= d.x1; }

template <bool EDGE>
__device__ __forceinline__ void fused_recon_body(const KArgs &a, const FusedBwdArgs &fa, float *smem, int z0, int x0, int s)
{
    float *szz = smem, *sxz = szz + R_SROWS * RW, *sxx = sxz + R_SROWS * RW;
    float *svz = sxx + R_SROWS * RW, *svx = svz + R_VROWS * RW;
    float *sbya = svx + R_VROWS * RW, *sbyb = sbya + R_VROWS * RW;
    float *sga = sbyb + R_VROWS * RW, *sgb = sga + R_GROWS * RW, *ssh = sgb + R_GROWS * RW;
    float *slam = ssh + R_GROWS * RW, *smu = slam + R_GROWS * RW, *smua = smu + R_GROWS * RW;

    const Dims &d = a.d;
    const int tid = threadIdx.x;
    const int ld = d.ldx;
    float *st = slot_state(a, s);
    const float *src = st + (size_t)(fa.q ? S_FWD1 : S_FWD) * d.fsz;
    float *dst = st + (size_t)(fa.q ? S_FWD : S_FWD1) * d.fsz;
    const float *adj = st + (size_t)(fa.pa ? S_ADJ1 : S_ADJ) * d.fsz;
    float *grad = a.grad + (size_t)s * 3 * d.fsz;
    const float *ringb = a.ring + (((size_t)s * NFIELD) * d.nSteps + fa.it) * d.ringLen;
    const size_t rfs = (size_t)d.nSteps * d.ringLen;

    if (tid < R_LD) {
        const int r = tid / RQ, c4 = tid - r * RQ;
        const int xq = x0 - 8 + 4 * c4;
        load_rows_r<R_SROWS, EDGE>(szz, src + (size_t)F_SZZ * d.fsz, z0 - 5, xq, r, c4, d);
        load_rows_r<R_SROWS, EDGE>(sxz, src + (size_t)F_SXZ * d.fsz, z0 - 5, xq, r, c4, d);
        load_rows_r<R_SROWS, EDGE>(sxx, src + (size_t)F_SXX * d.fsz, z0 - 5, xq, r, c4, d);
        load_rows_r<R_VROWS, EDGE>(svz, src + (size_t)F_VZ * d.fsz, z0 - 3, xq, r, c4, d);
        load_rows_r<R_VROWS, EDGE>(svx, src + (size_t)F_VX * d.fsz, z0 - 3, xq, r, c4, d);
        load_rows_r<R_VROWS, EDGE>(sbya, a.model + (size_t)M_BYCA * d.fsz, z0 - 3, xq, r, c4, d);
        load_rows_r<R_VROWS, EDGE>(sbyb, a.model + (size_t)M_BYCB * d.fsz, z0 - 3, xq, r, c4, d);
        load_rows_r<R_GROWS, EDGE>(slam, a.model + (size_t)M_LAM * d.fsz, z0 - 1, xq, r, c4, d);
        load_rows_r<R_GROWS, EDGE>(smu, a.model + (size_t)M_MU * d.fsz, z0 - 1, xq, r, c4, d);
        load_rows_r<R_GROWS, EDGE>(smua, a.model + (size_t)M_MUAVE * d.fsz, z0 - 1, xq, r, c4, d);
    }
    cp_async_wait_all();
    __syncthreads();

    const bool ring = EDGE && tile_touches_ring_ext(d, z0 - 3, z0 + FTZ + 1, x0 - 4, x0 + FTX + 3);

    // ---- 1. velocities at time `it` on rows z0-3 .. z0+FTZ+1, quads x0-4 .. x0+FTX+3 (smem quads 1..18)
    if (tid < R_VROWS * FQ) {
        const int j = tid / FQ, c4 = tid - j * FQ + 1;
        const int z = z0 - 3 + j, xq = x0 - 8 + 4 * c4;
        const size_t iq = (size_t)z * ld + xq;
        // sigma-tile row of z is j+2
        const float4 p0 = ld4r(szz, j + 1, c4), p1 = ld4r(szz, j + 2, c4), p2 = ld4r(szz, j + 3, c4), p3 = ld4r(szz, j + 4, c4);
        const float4 q0 = ld4r(sxz, j, c4), q1 = ld4r(sxz, j + 1, c4), q2 = ld4r(sxz, j + 2, c4), q3 = ld4r(sxz, j + 3, c4);
        const float4 ql = ld4r(sxz, j + 2, c4 - 1), qr = ld4r(sxz, j + 2, c4 + 1);
        const float4 xc = ld4r(sxx, j + 2, c4), xl = ld4r(sxx, j + 2, c4 - 1), xr = ld4r(sxx, j + 2, c4 + 1);
        const float wxz[7] = WIN_M2(ql, q2, qr), wxx[7] = WIN_M1(xl, xc, xr);
        const float zzm1[4] = Q4(p0), zzc[4] = Q4(p1), zzp1[4] = Q4(p2), zzp2[4] = Q4(p3);
        const float xzm2[4] = Q4(q0), xzm1[4] = Q4(q1), xzc[4] = Q4(q2), xzp1[4] = Q4(q3);
        const float4 ovz4 = ld4r(svz, j, c4), ovx4 = ld4r(svx, j, c4), ba4 = ld4r(sbya, j, c4), bb4 = ld4r(sbyb, j, c4);
        const float ovz[4] = Q4(ovz4), ovx[4] = Q4(ovx4), ba[4] = Q4(ba4), bb[4] = Q4(bb4);
        const bool grow = (j >= 2) && (j < R_GROWS + 2);        // rows z0-1 .. z0+FTZ-1 carry the imaging terms
        float az4[4] = {0.f, 0.f, 0.f, 0.f}, ax4[4] = {0.f, 0.f, 0.f, 0.f};
        if (grow && (!EDGE || (z >= 0 && z < d.nzA && xq >= 0 && xq < d.ldx))) {
            const float4 t0 = *reinterpret_cast<const float4 *>(adj + (size_t)F_VZ * d.fsz + iq);
            const float4 t1 = *reinterpret_cast<const float4 *>(adj + (size_t)F_VX * d.fsz + iq);
            az4[0] = t0.x; az4[1] = t0.y; az4[2] = t0.z; az4[3] = t0.w;
            ax4[0] = t1.x; ax4[1] = t1.y; ax4[2] = t1.z; ax4[3] = t1.w;
        }
        float nvz[4], nvx[4], ga[4], gb[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int x = xq + c;
            const float A = DIFFZ(zzm1[c], zzc[c], zzp1[c], zzp2[c]) + DIFFX(wxz, c);
            const float B = DIFFZ(xzm2[c], xzm1[c], xzc[c], xzp1[c]) + DIFFX(wxx, c);
            const bool in = !EDGE || interior(d, z, x);
            nvz[c] = in ? ovz[c] - A * ba[c] * d.dt : ovz[c];
            nvx[c] = in ? ovx[c] - B * bb[c] * d.dt : ovx[c];
            ga[c] = in ? az4[c] * A * d.dt * (0.5f * ba[c] * ba[c]) : 0.f;
            gb[c] = in ? ax4[c] * B * d.dt * (0.5f * bb[c] * bb[c]) : 0.f;
            if (EDGE && ring) {
                int r0, r1;
                ring_indices2(d, z, x, r0, r1);
                const int ri = r0 >= 0 ? r0 : r1;
                if (ri >= 0) { nvz[c] = ringb[F_VZ * rfs + ri]; nvx[c] = ringb[F_VX * rfs + ri]; }
            }
        }
        st4r(svz, j, c4, make_float4(nvz[0], nvz[1], nvz[2], nvz[3]));
        st4r(svx, j, c4, make_float4(nvx[0], nvx[1], nvx[2], nvx[3]));
        if (grow) {
            st4r(sga, j - 2, c4, make_float4(ga[0], ga[1], ga[2], ga[3]));
            st4r(sgb, j - 2, c4, make_float4(gb[0], gb[1], gb[2], gb[3]));
        }
    }
    __syncthreads();

    // ---- 2a. shear imaging term of the staggered cells, rows z0-1 .. z0+FTZ-1, quads x0-4 .. x0+FTX-1 (smem quads 1..17)
    if (tid < R_GROWS * (FTX / 4 + 1)) {
        const int jg = tid / (FTX / 4 + 1), c4 = tid - jg * (FTX / 4 + 1) + 1;
        const int z = z0 - 1 + jg, xq = x0 - 8 + 4 * c4;
        // v-tile row of z is jg+2
        const float4 u0 = ld4r(svx, jg + 1, c4), u1 = ld4r(svx, jg + 2, c4), u2 = ld4r(svx, jg + 3, c4), u3 = ld4r(svx, jg + 4, c4);
        const float4 vc = ld4r(svz, jg + 2, c4), vl = ld4r(svz, jg + 2, c4 - 1), vr = ld4r(svz, jg + 2, c4 + 1);
        const float wvz[7] = WIN_M1(vl, vc, vr);
        const float vxm1[4] = Q4(u0), vxc[4] = Q4(u1), vxp1[4] = Q4(u2), vxp2[4] = Q4(u3);
        const float4 mua4 = ld4r(smua, jg, c4);
        const float mua[4] = Q4(mua4);
        float sa[4] = {0.f, 0.f, 0.f, 0.f};
        if (!EDGE || (z >= 0 && z < d.nzA && xq >= 0 && xq < d.ldx)) {
            const float4 t = *reinterpret_cast<const float4 *>(adj + (size_t)F_SXZ * d.fsz + (size_t)z * ld + xq);
            sa[0] = t.x; sa[1] = t.y; sa[2] = t.z; sa[3] = t.w;
        }
        float sh[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const float D3 = DIFFZ(vxm1[c], vxc[c], vxp1[c], vxp2[c]) + DIFFX(wvz, c);
            // mu_ave / sum(1/mu) = mu_ave^2 / 4  (mu_ave is the 4-point harmonic mean; 0 when any corner is 0)
            const float v = -sa[c] * D3 * d.dt * (0.25f * mua[c] * mua[c]) * 1e6f;
            sh[c] = (!EDGE || interior(d, z, xq + c)) ? v : 0.f;
        }
        st4r(ssh, jg, c4, make_float4(sh[0], sh[1], sh[2], sh[3]));
    }
    __syncthreads();

    // ---- 2b. stresses at time `it` and the three gradients on the owned quads (smem quads 2..17)
    if (tid < FTZ * (FTX / 4)) {
        const int ii = tid / (FTX / 4), c4 = tid - ii * (FTX / 4) + 2;
        const int z = z0 + ii, xq = x0 - 8 + 4 * c4;
        if (!EDGE || (z <= d.z1 + 2 && xq + 3 >= d.nPml - 2 && xq <= d.x1 + 2)) {
            const size_t iq = (size_t)z * ld + xq;
            const int zs = a.t.zs[s], xs = a.t.xs[s];
            // v-tile row of z is ii+3, sigma-tile row ii+5, g-row ii+1
            const float4 v0 = ld4r(svz, ii + 1, c4), v1 = ld4r(svz, ii + 2, c4), v2 = ld4r(svz, ii + 3, c4), v3 = ld4r(svz, ii + 4, c4);
            const float4 vl = ld4r(svz, ii + 3, c4 - 1), vr = ld4r(svz, ii + 3, c4 + 1);
            const float4 u0 = ld4r(svx, ii + 2, c4), u1 = ld4r(svx, ii + 3, c4), u2 = ld4r(svx, ii + 4, c4), u3 = ld4r(svx, ii + 5, c4);
            const float4 ul = ld4r(svx, ii + 3, c4 - 1), ur = ld4r(svx, ii + 3, c4 + 1);
            const float wvz[7] = WIN_M1(vl, v2, vr), wvx[7] = WIN_M2(ul, u1, ur);
            const float vzm2[4] = Q4(v0), vzm1[4] = Q4(v1), vzc[4] = Q4(v2), vzp1[4] = Q4(v3);
            const float vxm1[4] = Q4(u0), vxc[4] = Q4(u1), vxp1[4] = Q4(u2), vxp2[4] = Q4(u3);
            const float4 ozz4 = ld4r(szz, ii + 5, c4), oxz4 = ld4r(sxz, ii + 5, c4), oxx4 = ld4r(sxx, ii + 5, c4);
            const float ozz[4] = Q4(ozz4), oxz[4] = Q4(oxz4), oxx[4] = Q4(oxx4);
            const float4 lam4 = ld4r(slam, ii + 1, c4), mu4 = ld4r(smu, ii + 1, c4), mua4 = ld4r(smua, ii + 1, c4);
            const float lam[4] = Q4(lam4), mu[4] = Q4(mu4), mua[4] = Q4(mua4);
            const float4 gac = ld4r(sga, ii + 1, c4), gau = ld4r(sga, ii, c4), gbc = ld4r(sgb, ii + 1, c4), gbl = ld4r(sgb, ii + 1, c4 - 1);
            const float4 shc = ld4r(ssh, ii + 1, c4), shl = ld4r(ssh, ii + 1, c4 - 1), shu = ld4r(ssh, ii, c4), shul = ld4r(ssh, ii, c4 - 1);
            const float gaC[4] = Q4(gac), gaU[4] = Q4(gau), gbW[5] = {gbl.w, gbc.x, gbc.y, gbc.z, gbc.w};
            const float shW[5] = {shl.w, shc.x, shc.y, shc.z, shc.w}, shUW[5] = {shul.w, shu.x, shu.y, shu.z, shu.w};
            const float4 za4 = *reinterpret_cast<const float4 *>(adj + (size_t)F_SZZ * d.fsz + iq);
            const float4 xa4 = *reinterpret_cast<const float4 *>(adj + (size_t)F_SXX * d.fsz + iq);
            const float za[4] = Q4(za4), xa[4] = Q4(xa4);
            float4 g0 = *reinterpret_cast<const float4 *>(grad + 0 * d.fsz + iq);
            float4 g1 = *reinterpret_cast<const float4 *>(grad + 1 * d.fsz + iq);
            float4 g2 = *reinterpret_cast<const float4 *>(grad + 2 * d.fsz + iq);
            float gl[4] = Q4(g0), gm[4] = Q4(g1), gr[4] = Q4(g2);
            float nzz[4], nxz[4], nxx[4];
            const bool zle = !EDGE || (z <= d.z1);
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int x = xq + c;
                const float D1 = DIFFZ(vzm2[c], vzm1[c], vzc[c], vzp1[c]), D2 = DIFFX(wvx, c);
                const float D3 = DIFFZ(vxm1[c], vxc[c], vxp1[c], vxp2[c]) + DIFFX(wvz, c);
                float tzz = ozz[c], txx = oxx[c], txz = oxz[c];
                if (z == zs && x == xs) {
                    const float amp = a.t.amp[(size_t)s * d.nSteps + fa.it];
                    tzz -= amp; txx -= amp;
                }
                const bool in = !EDGE || interior(d, z, x);
                if (in) {
                    const float l2u = lam[c] + 2.0f * mu[c];
                    tzz -= (l2u * D1 + lam[c] * D2) * d.dt;
                    txx -= (lam[c] * D1 + l2u * D2) * d.dt;
                    txz -= mua[c] * D3 * d.dt;
                    gl[c] += -(za[c] + xa[c]) * (D1 + D2) * d.dt * 1e6f;
                }
                // gathers of the reference's sprays (el_stress.cu:116-123, el_velocity.cu:104-110)
                float sh = shW[c + 1] + shW[c], grr = (in ? gaC[c] + gbW[c + 1] : 0.f) + gbW[c];
                if (zle) { sh += shUW[c + 1]; grr += gaU[c]; if (!EDGE || x <= d.x1) sh += shUW[c]; }
                const float gmn = in ? (-2.0f * za[c] * D1 * d.dt - 2.0f * xa[c] * D2 * d.dt) * 1e6f : 0.f;
                if (!EDGE || in || sh != 0.f) gm[c] += gmn + sh / (mu[c] * mu[c]);
                gr[c] += grr;
                if (EDGE && ring) {
                    int r0, r1;
                    ring_indices2(d, z, x, r0, r1);
                    const int ri = r0 >= 0 ? r0 : r1;
                    if (ri >= 0) { tzz = ringb[F_SZZ * rfs + ri]; txz = ringb[F_SXZ * rfs + ri]; txx = ringb[F_SXX * rfs + ri]; }
                }
                nzz[c] = tzz; nxx[c] = txx; nxz[c] = txz;
            }
            *reinterpret_cast<float4 *>(grad + 0 * d.fsz + iq) = make_float4(gl[0], gl[1], gl[2], gl[3]);
            *reinterpret_cast<float4 *>(grad + 1 * d.fsz + iq) = make_float4(gm[0], gm[1], gm[2], gm[3]);
            *reinterpret_cast<float4 *>(grad + 2 * d.fsz + iq) = make_float4(gr[0], gr[1], gr[2], gr[3]);
            *reinterpret_cast<float4 *>(dst + (size_t)F_SZZ * d.fsz + iq) = make_float4(nzz[0], nzz[1], nzz[2], nzz[3]);
            *reinterpret_cast<float4 *>(dst + (size_t)F_SXZ * d.fsz + iq) = make_float4(nxz[0], nxz[1], nxz[2], nxz[3]);
            *reinterpret_cast<float4 *>(dst + (size_t)F_SXX * d.fsz + iq) = make_float4(nxx[0], nxx[1], nxx[2], nxx[3]);
            *reinterpret_cast<float4 *>(dst + (size_t)F_VZ * d.fsz + iq) = v2;
            *reinterpret_cast<float4 *>(dst + (size_t)F_VX * d.fsz + iq) = u1;
        }
    }
}

__global__ void __launch_bounds__(F4_NT, 2) k_fused_recon(const KArgs a, const FusedBwdArgs fa)
{
    extern __shared__ __align__(16) float smem[];
    const Dims &d = a.d;
    // tiles cover the region interior U ring = [nPml-2, z1+2] x [nPml-2, x1+2]; x0 stays a multiple of 4
    const int z0 = d.nPml - 2 + blockIdx.y * FTZ, x0 = ((d.nPml - 2) & ~3) + blockIdx.x * FTX;
    // interior tile: every cell of the staged v region is an interior cell that is not part of the ring
    const bool inner = (z0 - 3 >= d.nPml + 3) && (z0 + FTZ + 1 <= d.z1 - 3) && (x0 - 4 >= d.nPml + 3) && (x0 + FTX + 3 <= d.x1 - 3);
    if (inner) fused_recon_body<false>(a, fa, smem, z0, x0, blockIdx.z);
    else fused_recon_body<true>(a, fa, smem, z0, x0, blockIdx.z);
}

}  // namespace sepfwi

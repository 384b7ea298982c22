// kernels_fused.cuh -- fused sm_100a time-step kernels (the default path).
//
// k_fused_fwd: ONE launch per forward time step.  A CTA owns a 16 x 64 tile of cells and
//   1. stages the five old fields with their halos and the five coefficient arrays in shared
//      memory (cp.async, 16-byte requests, zero-fill outside the grid): v with a 4-cell halo,
//      sigma with a 2-cell halo -- every global read of the step is in flight at once and the
//      compute phases touch shared memory only (CPML strips excepted);
//   2. records the traces of sample `it` for the receivers inside the tile and stores the
//      boundary ring of time `it` (both read the OLD state, exactly what the reference's
//      recording at it+1 of the previous step and from_bnd at `it` see);
//   3. updates sigma on the tile + 2-cell halo (halo recompute instead of a grid-wide
//      dependency), adds the explosive source, keeps the result in shared memory;
//   4. updates v on the tile from the new sigma.
// Each thread owns a quad of four x-consecutive cells: every shared-memory access is a 128-bit
// LDS/STS and every global store a 128-bit STG.  Tiles that touch neither the PML, the inactive
// rim, nor (in gradient mode) the boundary ring take a branch-free path with unpredicated loads;
// the few edge tiles run the same arithmetic per component with the CPML / ring / rim handling.
// State is ping-pong (read buffer p, write buffer p^1) because neighbouring tiles recompute
// each other's halo from the old state; the four velocity-side CPML memory variables are
// owner-only and stay in place.
//
// Arithmetic per cell is the same sequence as the baseline kernels (kernels_base.cuh), so the
// two paths agree to the last bit except where the compiler contracts differently.
// Reference lines: el_stress.cu:50-87, el_velocity.cu:45-82, utilities.cu:362-392,524-552,593-703.
#pragma once
#include "common.cuh"

namespace sepfwi {

constexpr int FTX = 64;            // tile width  (x)
constexpr int FTZ = 16;            // tile height (z)
constexpr int FW = FTX + 8;        // shared-memory row pitch: tile + 4-cell halo each side
constexpr int F_NT = 4 * FW;       // 288 threads: 72 columns x 4 row groups
constexpr int F_VROWS = FTZ + 8;   // v tile rows      z0-4 .. z0+FTZ+3
constexpr int F_SROWS = FTZ + 4;   // sigma tile rows  z0-2 .. z0+FTZ+1
// v (2) + sigma (3) + lambda, mu, mu_ave on the sigma rows (3) + buoyancies on the tile rows (2)
constexpr size_t F_SMEM = (size_t)(2 * F_VROWS + 6 * F_SROWS + 2 * FTZ) * FW * sizeof(float);   // 57600 B

__device__ __forceinline__ void cp_async16(float *smem_dst, const float *gsrc, bool valid)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 16 : 0;   // src-size 0 => 16 bytes of zeros, nothing is read
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc), "r"(sz));
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\n" ::);
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}

// rows [zfirst, zfirst+ROWS) x columns [xfirst, xfirst+FW) of a global [nzA][ldx] array -> smem tile
template <int ROWS, int NT>
__device__ __forceinline__ void load_tile(float *s, const float *g, int zfirst, int xfirst, const Dims &d, int tid)
{
    constexpr int V4 = FW / 4;
    for (int k = tid; k < ROWS * V4; k += NT) {
        const int r = k / V4, c4 = k - r * V4;
        const int z = zfirst + r, x = xfirst + c4 * 4;
        const bool ok = (z >= 0) && (z < d.nzA) && (x >= 0) && (x < d.ldx);
        cp_async16(s + r * FW + c4 * 4, ok ? g + (size_t)z * d.ldx + x : g, ok);
    }
}

// does the tile [z0,z0+FTZ) x [x0,x0+FTX) touch the 5-wide boundary ring frame?
__device__ __forceinline__ bool tile_touches_ring(const Dims &d, int z0, int x0)
{
    const int zlo = d.nPml - 2, zhi = d.z1 + 2, xlo = d.nPml - 2, xhi = d.x1 + 2;
    const int za = max(z0, zlo), zb = min(z0 + FTZ - 1, zhi), xa = max(x0, xlo), xb = min(x0 + FTX - 1, xhi);
    if (za > zb || xa > xb) return false;                                    // outside the frame's bounding box
    return !(za >= zlo + 5 && zb <= zhi - 5 && xa >= xlo + 5 && xb <= xhi - 5);   // not entirely in the hole
}

// same test for an arbitrary inclusive cell rectangle
__device__ __forceinline__ bool tile_touches_ring_ext(const Dims &d, int za_, int zb_, int xa_, int xb_)
{
    const int zlo = d.nPml - 2, zhi = d.z1 + 2, xlo = d.nPml - 2, xhi = d.x1 + 2;
    const int za = max(za_, zlo), zb = min(zb_, zhi), xa = max(xa_, xlo), xb = min(xb_, xhi);
    if (za > zb || xa > xb) return false;
    return !(za >= zlo + 5 && zb <= zhi - 5 && xa >= xlo + 5 && xb <= xhi - 5);
}

struct FusedFwdArgs {
    int it;          // time step: state `it` is read from buffer (it & 1), state it+1 written to the other
    int mask;        // trace components to record at sample `it` (0 = none)
    int fiber;
    int save_ring;   // store the boundary ring of time `it`
};

constexpr int FQ = FW / 4;                 // 18 quads per staged row
constexpr int F4_NT = 384;                 // threads of k_fused_fwd
constexpr int F4_LD = 21 * FQ;             // 378 loader threads: +378 in the quad index = +21 rows, same column

// one cp.async per (array, row): thread (r, c4) copies quad c4 of tile row r (and r+21 if the tile has more rows)
template <int ROWS, bool EDGE>
__device__ __forceinline__ void load_rows(float *s, const float *g, int zfirst, int xq, int r, int c4, const Dims &d)
{
#pragma unroll
    for (int rr = 0; rr < ROWS; rr += 21) {
        const int row = r + rr;
        if (row < ROWS) {
            const int z = zfirst + row;
            if (EDGE) {
                const bool ok = (z >= 0) && (z < d.nzA) && (xq >= 0) && (xq < d.ldx);
                cp_async16(s + row * FW + c4 * 4, ok ? g + (size_t)z * d.ldx + xq : g, ok);
            } else {
                cp_async16(s + row * FW + c4 * 4, g + (size_t)z * d.ldx + xq, true);
            }
        }
    }
}

__device__ __forceinline__ float4 ld4(const float *s, int row, int c4) { return *reinterpret_cast<const float4 *>(s + row * FW + c4 * 4); }
__device__ __forceinline__ void st4(float *s, int row, int c4, const float4 &v) { *reinterpret_cast<float4 *>(s + row * FW + c4 * 4) = v; }

template <bool EDGE>
__device__ __forceinline__ void fused_fwd_body(const KArgs &a, const FusedFwdArgs &fa, float *smem, int z0, int x0, int s)
{
    float *svz = smem, *svx = svz + F_VROWS * FW;
    float *szz = svx + F_VROWS * FW, *sxz = szz + F_SROWS * FW, *sxx = sxz + F_SROWS * FW;
    float *slam = sxx + F_SROWS * FW, *smu = slam + F_SROWS * FW, *smua = smu + F_SROWS * FW;
    float *sbya = smua + F_SROWS * FW, *sbyb = sbya + FTZ * FW;

    const Dims &d = a.d;
    const int tid = threadIdx.x;
    const int ld = d.ldx;
    const int p = fa.it & 1;
    float *st = slot_state(a, s);
    const float *src = st + (size_t)(p ? S_FWD1 : S_FWD) * d.fsz;
    float *dst = st + (size_t)(p ? S_FWD : S_FWD1) * d.fsz;
    const float *psiv_src = st + (size_t)(p ? S_FPSIV1 : S_FPSI) * d.fsz;
    float *psiv_dst = st + (size_t)(p ? S_FPSI : S_FPSIV1) * d.fsz;
    float *psis = st + (size_t)(S_FPSI + P_SZZ_Z) * d.fsz;

    // ---- 1. stage old fields and coefficients (every global read of the step is issued here)
    if (tid < F4_LD) {
        const int r = tid / FQ, c4 = tid - r * FQ;
        const int xq = x0 - 4 + 4 * c4;
        load_rows<F_VROWS, EDGE>(svz, src + (size_t)F_VZ * d.fsz, z0 - 4, xq, r, c4, d);
        load_rows<F_VROWS, EDGE>(svx, src + (size_t)F_VX * d.fsz, z0 - 4, xq, r, c4, d);
        load_rows<F_SROWS, EDGE>(szz, src + (size_t)F_SZZ * d.fsz, z0 - 2, xq, r, c4, d);
        load_rows<F_SROWS, EDGE>(sxz, src + (size_t)F_SXZ * d.fsz, z0 - 2, xq, r, c4, d);
        load_rows<F_SROWS, EDGE>(sxx, src + (size_t)F_SXX * d.fsz, z0 - 2, xq, r, c4, d);
        load_rows<F_SROWS, EDGE>(slam, a.model + (size_t)M_LAM * d.fsz, z0 - 2, xq, r, c4, d);
        load_rows<F_SROWS, EDGE>(smu, a.model + (size_t)M_MU * d.fsz, z0 - 2, xq, r, c4, d);
        load_rows<F_SROWS, EDGE>(smua, a.model + (size_t)M_MUAVE * d.fsz, z0 - 2, xq, r, c4, d);
        load_rows<FTZ, EDGE>(sbya, a.model + (size_t)M_BYCA * d.fsz, z0, xq, r, c4, d);
        load_rows<FTZ, EDGE>(sbyb, a.model + (size_t)M_BYCB * d.fsz, z0, xq, r, c4, d);
    }
    cp_async_wait_all();
    __syncthreads();

    // ---- 2. record sample `it` from the old state for the receivers that live in this tile
    if (fa.mask) {
        const int t = blockIdx.y * a.t.ntx + blockIdx.x;
        const int *tp = a.t.tilePtr + (size_t)s * (a.t.nTiles + 1);
        const int k0 = tp[t], k1 = tp[t + 1];
        if (k1 > k0) {
            const size_t cs = (size_t)d.maxRec * d.nSteps;
            for (int k = k0 + tid; k < k1; k += F4_NT) {
                const int r = a.t.tileRec[(size_t)s * d.maxRec + k];
                const int z = a.t.zrec[(size_t)s * d.maxRec + r], x = a.t.xrec[(size_t)s * d.maxRec + r];
                const int vi = (z - z0 + 4) * FW + (x - x0 + 4), si = (z - z0 + 2) * FW + (x - x0 + 4);
                float *tr = a.trace + (size_t)s * d.nTrace * cs + (size_t)r * d.nSteps + fa.it;
                const float exx = svx[vi] - svx[vi - 1], ezz = svz[vi] - svz[vi - FW];
                if (fa.mask & (1 << T_PR)) tr[T_PR * cs] = szz[si] + sxx[si];
                if (fa.mask & (1 << T_VX)) tr[T_VX * cs] = svx[vi];
                if (fa.mask & (1 << T_VZ)) tr[T_VZ * cs] = svz[vi];
                if (fa.mask & (1 << T_ETT)) {
                    const float *w = a.t.w ? a.t.w + ((size_t)s * d.maxRec + r) * 3 : nullptr;
                    float e = fa.fiber == 0 ? exx : ezz;
                    if (w) e = w[0] * exx + w[1] * ezz + w[2] * (0.5f * ((svx[vi + FW] - svx[vi]) + (svz[vi + 1] - svz[vi])));
                    tr[T_ETT * cs] = e;
                }
            }
            __syncthreads();   // phase 3 overwrites the sigma tile
        }
    }

    const int zs = a.t.zs[s], xs = a.t.xs[s];
    const bool ring = EDGE && fa.save_ring && tile_touches_ring(d, z0, x0);

    // ---- 3. stress update on rows z0-2 .. z0+FTZ+1, all 18 staged quads (the outer two columns of the
    //         outer quads lack a neighbour and are never consumed).  Two passes (normal stresses, then the
    //         shear stress) keep the live register set small.
    if (tid < F_SROWS * FQ) {
        const int j = tid / FQ, c4 = tid - j * FQ;
        const int z = z0 - 2 + j, xq = x0 - 4 + 4 * c4;
        const int cl = max(c4 - 1, 0), cr = min(c4 + 1, FQ - 1);
        const bool zin = !EDGE || (z >= 2 && z <= d.nzA - 3);
        const bool zp = EDGE && ((z < d.nPml) || (z > d.nzA - d.nPml - 1));
        const bool own = (j >= 2) && (j < FTZ + 2) && (c4 >= 1) && (c4 <= FTX / 4);
        const bool wr = own && (!EDGE || (z < d.nzA && xq < d.ldx));
        const size_t iq = (size_t)z * ld + xq;
        const float4 a2 = ld4(svz, j + 2, c4), b1 = ld4(svx, j + 2, c4);            // centre row of vz, vx
        const float vzq[4] = {a2.x, a2.y, a2.z, a2.w}, vxq[4] = {b1.x, b1.y, b1.z, b1.w};
        const float4 lam4 = ld4(slam, j, c4), mu4 = ld4(smu, j, c4);
        const float lam[4] = {lam4.x, lam4.y, lam4.z, lam4.w}, mu[4] = {mu4.x, mu4.y, mu4.z, mu4.w};
        {   // ---- szz, sxx from dvz/dz and dvx/dx
            const float4 a0 = ld4(svz, j, c4), a1 = ld4(svz, j + 1, c4), a3 = ld4(svz, j + 3, c4);
            const float4 bl = ld4(svx, j + 2, cl), br = ld4(svx, j + 2, cr);
            const float vxc[7] = {bl.z, bl.w, b1.x, b1.y, b1.z, b1.w, br.x};      // vx[x-2 .. x+4]
            const float vzm2[4] = {a0.x, a0.y, a0.z, a0.w}, vzm1[4] = {a1.x, a1.y, a1.z, a1.w}, vzp1[4] = {a3.x, a3.y, a3.z, a3.w};
            const float4 ozz4 = ld4(szz, j, c4), oxx4 = ld4(sxx, j, c4);
            const float ozz[4] = {ozz4.x, ozz4.y, ozz4.z, ozz4.w}, oxx[4] = {oxx4.x, oxx4.y, oxx4.z, oxx4.w};
            float nzz[4], nxx[4];
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int x = xq + c;
                float dvz_dz = d.c1z * (vzq[c] - vzm1[c]) - d.c2z * (vzp1[c] - vzm2[c]);
                float dvx_dx = d.c1x * (vxc[c + 2] - vxc[c + 1]) - d.c2x * (vxc[c + 3] - vxc[c]);
                bool act = true;
                if (EDGE) {
                    act = zin && (x >= 2) && (x <= d.nx - 3);
                    if (act) {
                        const size_t i = iq + c;
                        if (zp) {
                            const float *cz = a.cz + z;
                            const float m0 = cz[C_B * d.nzA] * psiv_src[(size_t)P_VZ_Z * d.fsz + i] + cz[C_A * d.nzA] * dvz_dz;
                            if (own) psiv_dst[(size_t)P_VZ_Z * d.fsz + i] = m0;
                            dvz_dz = dvz_dz * cz[C_RK * d.nzA] + m0;
                        }
                        if ((x < d.nPml) || (x > d.nx - d.nPml - 1)) {
                            const float *cxp = a.cx + x;
                            const float m0 = cxp[C_B * d.nx] * psiv_src[(size_t)P_VX_X * d.fsz + i] + cxp[C_A * d.nx] * dvx_dx;
                            if (own) psiv_dst[(size_t)P_VX_X * d.fsz + i] = m0;
                            dvx_dx = dvx_dx * cxp[C_RK * d.nx] + m0;
                        }
                        if (ring && own) {
                            int r0, r1;
                            ring_indices2(d, z, x, r0, r1);
                            const size_t fs = (size_t)d.nSteps * d.ringLen;
                            float *rb = a.ring + (((size_t)s * NFIELD) * d.nSteps + fa.it) * d.ringLen;
                            if (r0 >= 0) { rb[F_SZZ * fs + r0] = ozz[c]; rb[F_SXX * fs + r0] = oxx[c]; rb[F_VZ * fs + r0] = vzq[c]; rb[F_VX * fs + r0] = vxq[c]; }
                            if (r1 >= 0) { rb[F_SZZ * fs + r1] = ozz[c]; rb[F_SXX * fs + r1] = oxx[c]; rb[F_VZ * fs + r1] = vzq[c]; rb[F_VX * fs + r1] = vxq[c]; }
                        }
                    }
                }
                const float l2u = lam[c] + 2.0f * mu[c];
                float tzz = ozz[c] + (l2u * dvz_dz + lam[c] * dvx_dx) * d.dt;
                float txx = oxx[c] + (lam[c] * dvz_dz + l2u * dvx_dx) * d.dt;
                if (z == zs && x == xs) {
                    const float amp = a.t.amp[(size_t)s * d.nSteps + fa.it];
                    tzz += amp; txx += amp;
                }
                nzz[c] = act ? tzz : ozz[c]; nxx[c] = act ? txx : oxx[c];
            }
            const float4 rzz = make_float4(nzz[0], nzz[1], nzz[2], nzz[3]), rxx = make_float4(nxx[0], nxx[1], nxx[2], nxx[3]);
            st4(szz, j, c4, rzz); st4(sxx, j, c4, rxx);
            if (wr) {   // inactive cells keep their old value (zero by construction), so the whole quad can be stored
                *reinterpret_cast<float4 *>(dst + (size_t)F_SZZ * d.fsz + iq) = rzz;
                *reinterpret_cast<float4 *>(dst + (size_t)F_SXX * d.fsz + iq) = rxx;
            }
        }
        {   // ---- sxz from dvx/dz and dvz/dx
            const float4 b0 = ld4(svx, j + 1, c4), b2 = ld4(svx, j + 3, c4), b3 = ld4(svx, j + 4, c4);
            const float4 al = ld4(svz, j + 2, cl), ar = ld4(svz, j + 2, cr);
            const float vzc[7] = {al.w, a2.x, a2.y, a2.z, a2.w, ar.x, ar.y};      // vz[x-1 .. x+5]
            const float vxm1[4] = {b0.x, b0.y, b0.z, b0.w}, vxp1[4] = {b2.x, b2.y, b2.z, b2.w}, vxp2[4] = {b3.x, b3.y, b3.z, b3.w};
            const float4 oxz4 = ld4(sxz, j, c4), mua4 = ld4(smua, j, c4);
            const float oxz[4] = {oxz4.x, oxz4.y, oxz4.z, oxz4.w}, mua[4] = {mua4.x, mua4.y, mua4.z, mua4.w};
            float nxz[4];
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int x = xq + c;
                float dvx_dz = d.c1z * (vxp1[c] - vxq[c]) - d.c2z * (vxp2[c] - vxm1[c]);
                float dvz_dx = d.c1x * (vzc[c + 2] - vzc[c + 1]) - d.c2x * (vzc[c + 3] - vzc[c]);
                bool act = true;
                if (EDGE) {
                    act = zin && (x >= 2) && (x <= d.nx - 3);
                    if (act) {
                        const size_t i = iq + c;
                        if (zp) {
                            const float *cz = a.cz + z;
                            const float m1 = cz[C_BH * d.nzA] * psiv_src[(size_t)P_VX_Z * d.fsz + i] + cz[C_AH * d.nzA] * dvx_dz;
                            if (own) psiv_dst[(size_t)P_VX_Z * d.fsz + i] = m1;
                            dvx_dz = dvx_dz * cz[C_RKH * d.nzA] + m1;
                        }
                        if ((x < d.nPml) || (x > d.nx - d.nPml - 1)) {
                            const float *cxp = a.cx + x;
                            const float m1 = cxp[C_BH * d.nx] * psiv_src[(size_t)P_VZ_X * d.fsz + i] + cxp[C_AH * d.nx] * dvz_dx;
                            if (own) psiv_dst[(size_t)P_VZ_X * d.fsz + i] = m1;
                            dvz_dx = dvz_dx * cxp[C_RKH * d.nx] + m1;
                        }
                        if (ring && own) {
                            int r0, r1;
                            ring_indices2(d, z, x, r0, r1);
                            float *rb = a.ring + (((size_t)s * NFIELD + F_SXZ) * d.nSteps + fa.it) * d.ringLen;
                            if (r0 >= 0) rb[r0] = oxz[c];
                            if (r1 >= 0) rb[r1] = oxz[c];
                        }
                    }
                }
                const float txz = oxz[c] + mua[c] * (dvx_dz + dvz_dx) * d.dt;
                nxz[c] = act ? txz : oxz[c];
            }
            const float4 rxz = make_float4(nxz[0], nxz[1], nxz[2], nxz[3]);
            st4(sxz, j, c4, rxz);
            if (wr) *reinterpret_cast<float4 *>(dst + (size_t)F_SXZ * d.fsz + iq) = rxz;
        }
    }
    __syncthreads();

    // ---- 4. velocity update on the owned quads from the new stresses
    if (tid < FTZ * (FTX / 4)) {
        const int ii = tid / (FTX / 4), c4 = tid - ii * (FTX / 4) + 1;
        const int z = z0 + ii, xq = x0 - 4 + 4 * c4;
        // sigma-tile row of z is ii+2, v-tile row ii+4
        const float4 p0 = ld4(szz, ii + 1, c4), p1 = ld4(szz, ii + 2, c4), p2 = ld4(szz, ii + 3, c4), p3 = ld4(szz, ii + 4, c4);
        const float4 q0 = ld4(sxz, ii, c4), q1 = ld4(sxz, ii + 1, c4), q2 = ld4(sxz, ii + 2, c4), q3 = ld4(sxz, ii + 3, c4);
        const float4 ql = ld4(sxz, ii + 2, c4 - 1), qr = ld4(sxz, ii + 2, c4 + 1);
        const float4 xc = ld4(sxx, ii + 2, c4), xl = ld4(sxx, ii + 2, c4 - 1), xr = ld4(sxx, ii + 2, c4 + 1);
        const float4 ovz4 = ld4(svz, ii + 4, c4), ovx4 = ld4(svx, ii + 4, c4);
        const float4 bya4 = ld4(sbya, ii, c4), byb4 = ld4(sbyb, ii, c4);
        const float xzc[7] = {ql.z, ql.w, q2.x, q2.y, q2.z, q2.w, qr.x};          // sxz[x-2 .. x+4]
        const float xxc[7] = {xl.w, xc.x, xc.y, xc.z, xc.w, xr.x, xr.y};          // sxx[x-1 .. x+5]
        const float zzm1[4] = {p0.x, p0.y, p0.z, p0.w}, zzc[4] = {p1.x, p1.y, p1.z, p1.w}, zzp1[4] = {p2.x, p2.y, p2.z, p2.w}, zzp2[4] = {p3.x, p3.y, p3.z, p3.w};
        const float xzm2[4] = {q0.x, q0.y, q0.z, q0.w}, xzm1[4] = {q1.x, q1.y, q1.z, q1.w}, xzp1[4] = {q3.x, q3.y, q3.z, q3.w};
        const float ovz[4] = {ovz4.x, ovz4.y, ovz4.z, ovz4.w}, ovx[4] = {ovx4.x, ovx4.y, ovx4.z, ovx4.w};
        const float bya[4] = {bya4.x, bya4.y, bya4.z, bya4.w}, byb[4] = {byb4.x, byb4.y, byb4.z, byb4.w};
        float nvz[4], nvx[4];
        const bool zin = !EDGE || (z >= 2 && z <= d.nzA - 3);
        const bool zp = EDGE && ((z < d.nPml) || (z > d.nzA - d.nPml - 1));
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int x = xq + c;
            float dszz_dz = d.c1z * (zzp1[c] - zzc[c]) - d.c2z * (zzp2[c] - zzm1[c]);
            float dsxz_dz = d.c1z * (xzc[c + 2] - xzm1[c]) - d.c2z * (xzp1[c] - xzm2[c]);
            float dsxz_dx = d.c1x * (xzc[c + 2] - xzc[c + 1]) - d.c2x * (xzc[c + 3] - xzc[c]);
            float dsxx_dx = d.c1x * (xxc[c + 2] - xxc[c + 1]) - d.c2x * (xxc[c + 3] - xxc[c]);
            bool act = true;
            if (EDGE) {
                act = zin && (x >= 2) && (x <= d.nx - 3);
                if (act) {
                    const size_t i = (size_t)z * ld + x;
                    if (zp) {
                        const float *cz = a.cz + z;
                        float *m0p = psis + (size_t)(P_SZZ_Z - P_SZZ_Z) * d.fsz + i, *m1p = psis + (size_t)(P_SXZ_Z - P_SZZ_Z) * d.fsz + i;
                        const float m0 = cz[C_BH * d.nzA] * *m0p + cz[C_AH * d.nzA] * dszz_dz;
                        const float m1 = cz[C_B * d.nzA] * *m1p + cz[C_A * d.nzA] * dsxz_dz;
                        *m0p = m0; *m1p = m1;
                        dszz_dz = dszz_dz * cz[C_RKH * d.nzA] + m0;
                        dsxz_dz = dsxz_dz * cz[C_RK * d.nzA] + m1;
                    }
                    if ((x < d.nPml) || (x > d.nx - d.nPml)) {                  // el_velocity.cu:56,71
                        const float *cxp = a.cx + x;
                        float *m0p = psis + (size_t)(P_SXZ_X - P_SZZ_Z) * d.fsz + i, *m1p = psis + (size_t)(P_SXX_X - P_SZZ_Z) * d.fsz + i;
                        const float m0 = cxp[C_B * d.nx] * *m0p + cxp[C_A * d.nx] * dsxz_dx;
                        const float m1 = cxp[C_BH * d.nx] * *m1p + cxp[C_AH * d.nx] * dsxx_dx;
                        *m0p = m0; *m1p = m1;
                        dsxz_dx = dsxz_dx * cxp[C_RK * d.nx] + m0;
                        dsxx_dx = dsxx_dx * cxp[C_RKH * d.nx] + m1;
                    }
                }
            }
            const float tvz = ovz[c] + (dszz_dz + dsxz_dx) * bya[c] * d.dt;
            const float tvx = ovx[c] + (dsxz_dz + dsxx_dx) * byb[c] * d.dt;
            nvz[c] = act ? tvz : ovz[c]; nvx[c] = act ? tvx : ovx[c];
        }
        if (!EDGE || (z < d.nzA && xq < d.ldx)) {
            const size_t i = (size_t)z * ld + xq;
            *reinterpret_cast<float4 *>(dst + (size_t)F_VZ * d.fsz + i) = make_float4(nvz[0], nvz[1], nvz[2], nvz[3]);
            *reinterpret_cast<float4 *>(dst + (size_t)F_VX * d.fsz + i) = make_float4(nvx[0], nvx[1], nvx[2], nvx[3]);
        }
    }
}

#ifndef F4_MINB
#define F4_MINB 3
#endif
__global__ void __launch_bounds__(F4_NT, F4_MINB) k_fused_fwd(const KArgs a, const FusedFwdArgs fa)
{
    extern __shared__ __align__(16) float smem[];
    const Dims &d = a.d;
    const int z0 = blockIdx.y * FTZ, x0 = blockIdx.x * FTX;
    // interior tile: the whole staged sigma region is active and outside the PML, and (gradient mode) the tile does
    // not touch the boundary ring
    const bool inner = (z0 - 2 >= d.nPml) && (z0 + FTZ + 1 <= d.nzA - d.nPml - 1) && (x0 - 4 >= d.nPml) &&
                       (x0 + FTX + 3 <= d.nx - d.nPml - 1) && !(fa.save_ring && tile_touches_ring(d, z0, x0));
    if (inner) fused_fwd_body<false>(a, fa, smem, z0, x0, blockIdx.z);
    else fused_fwd_body<true>(a, fa, smem, z0, x0, blockIdx.z);
}

}  // namespace sepfwi

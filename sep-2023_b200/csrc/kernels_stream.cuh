// kernels_stream.cuh -- register-streaming sm_100a time-step kernels (the default path).
//
// Decomposition: the live grid is cut into x-strips of 120 columns and z-chunks of Lz rows; ONE WARP
// owns a (strip, chunk).  Lane l holds the quad of four x-consecutive cells at column
// strip*120 - 4 + 4*l, so lanes 1..30 own output and lanes 0 / 31 carry the 4-column halo that the
// second half step needs.  The warp marches down z one row per iteration:
//   * every global access is a 128-bit LDG / STG of 512 contiguous bytes per warp and array;
//   * the z-windows of the 4th-order stencil (5 rows of v, 4-5 rows of the new stresses) live in
//     registers and rotate by full unrolling (6 iterations per loop trip), nothing is re-read;
//   * x-neighbours come from the adjacent lanes with warp shuffles (3 per differentiated field);
//   * the loads of row r+1 are issued before row r is computed (register double buffering), so
//     every warp keeps ~5 KB of requests in flight and no barrier or shared memory is needed;
//   * the first half step (stress) is recomputed on a 2-row / 2-column halo instead of a grid-wide
//     dependency, the second (velocity) follows two rows behind out of the register windows --
//     hence ping-pong state (read buffer p, write buffer p^1).
// Warps whose footprint touches neither the CPML strips nor the inactive 2-cell rim run a
// branch-free instantiation; the others run the same arithmetic with the CPML memory variables
// (touched in the strips only) and the active-cell select.  Perimeter work that reads only the OLD
// state -- trace recording and the boundary-ring save -- is done by a few leading CTAs of the same
// launch straight from global memory.
//
// Arithmetic per cell is the sequence of the baseline kernels (kernels_base.cuh).
// Reference lines: el_stress.cu:50-87, el_velocity.cu:45-82, utilities.cu:362-392,524-552,593-703.
#pragma once
#include "common.cuh"
#include "kernels_base.cuh"

namespace sepfwi {

constexpr int SW_OWN = 120;          // columns owned by one warp (30 quads)
constexpr int SW_WPB = 4;            // warps per CTA
constexpr int SW_NT = SW_WPB * 32;

struct StreamArgs {
    int it;
    int mask, fiber, save_ring;      // forward: trace components recorded at sample `it`, ring save of time `it`
    int q, pa;                       // backward: forward buffer holding state it+1, adjoint buffer holding adj(it+1)
    const int4 *work;                // work list, one entry per warp: {x0 = first owned column, z0, z1 = owned rows [z0, z1), 1 if edge}
    int nWork;                       //   edge entries (CPML strips / inactive rim) come first: they are the slow ones
    int nAux;                        // leading CTAs doing the perimeter work
    int nrecMax;                     // largest receiver count over the slots of this batch
    int force;                       // debug/timing only: 1 = every warp takes the interior path (wrong at the edges), 2 = every warp the edge path
};

#define Q4(v) {v.x, v.y, v.z, v.w}
// streaming 128-bit load of data that is read-only for the whole launch (L1 allocation kept: the 8-column overlap
// of adjacent strips is served by L1 when the neighbouring warp sits in the same CTA)
__device__ __forceinline__ float4 ldq(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
// register-free look-ahead: pull the 128-byte line of the next row into L1 (edge warps have no registers to spare)
__device__ __forceinline__ void pf_l1(const float *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ float4 ldq_c(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }   // cached (small reused tables)
__device__ __forceinline__ float4 ldq_rw(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ void stq(float *p, const float4 &v) { *reinterpret_cast<float4 *>(p) = v; }
__device__ __forceinline__ float4 mk4(const float v[4]) { return make_float4(v[0], v[1], v[2], v[3]); }
__device__ __forceinline__ float sh_l(float v) { return __shfl_up_sync(0xffffffffu, v, 1); }     // value held by lane-1
__device__ __forceinline__ float sh_r(float v) { return __shfl_down_sync(0xffffffffu, v, 1); }   // value held by lane+1
// 7-value x windows of a quad q: backward differences need f[x-2 .. x+4], forward differences f[x-1 .. x+5]
#define XWIN_B(q) {sh_l(q.z), sh_l(q.w), q.x, q.y, q.z, q.w, sh_r(q.x)}
#define XWIN_F(q) {sh_l(q.w), q.x, q.y, q.z, q.w, sh_r(q.x), sh_r(q.y)}
// 4th-order staggered difference on a 7-window at component k (same taps for both window kinds)
#define DX7(w, k) (c1x * (w[(k) + 2] - w[(k) + 1]) - c2x * (w[(k) + 3] - w[k]))
// z difference from four consecutive rows (m1, c0, p1, p2): c1 (p1 - c0) - c2 (p2 - m1)
#define DZ4(m1, c0, p1, p2) (c1z * ((p1) - (c0)) - c2z * ((p2) - (m1)))

// ------------------------------------------------------------------------------------------------
// perimeter work of the forward step: traces of sample `it` and the ring of time `it`, both from the OLD state
__device__ __forceinline__ void stream_fwd_aux(const KArgs &a, const StreamArgs &sa, int s)
{
    const Dims &d = a.d;
    const int ld = d.ldx;
    const int p = sa.it & 1;
    const float *src = slot_state(a, s) + (size_t)(p ? S_FWD1 : S_FWD) * d.fsz;
    const int nth = sa.nAux * SW_NT, t0 = blockIdx.x * SW_NT + threadIdx.x;
    if (sa.mask) {
        const int nrec = a.t.nrec[s];
        const size_t cs = (size_t)d.maxRec * d.nSteps;
        const float *vz = src + (size_t)F_VZ * d.fsz, *vx = src + (size_t)F_VX * d.fsz;
        for (int r = t0; r < nrec; r += nth) {
            const int z = a.t.zrec[(size_t)s * d.maxRec + r], x = a.t.xrec[(size_t)s * d.maxRec + r];
            const size_t i = (size_t)z * ld + x;
            float *tr = a.trace + (size_t)s * d.nTrace * cs + (size_t)r * d.nSteps + sa.it;
            const float exx = vx[i] - vx[i - 1], ezz = vz[i] - vz[i - ld];
            if (sa.mask & (1 << T_PR)) tr[T_PR * cs] = src[(size_t)F_SZZ * d.fsz + i] + src[(size_t)F_SXX * d.fsz + i];
            if (sa.mask & (1 << T_VX)) tr[T_VX * cs] = vx[i];
            if (sa.mask & (1 << T_VZ)) tr[T_VZ * cs] = vz[i];
            if (sa.mask & (1 << T_ETT)) {
                float e = sa.fiber == 0 ? exx : ezz;
                if (a.t.w) {
                    const float *w = a.t.w + ((size_t)s * d.maxRec + r) * 3;
                    e = w[0] * exx + w[1] * ezz + w[2] * (0.5f * ((vx[i + ld] - vx[i]) + (vz[i + 1] - vz[i])));
                }
                tr[T_ETT * cs] = e;
            }
        }
    }
    if (sa.save_ring) {
        float *rb = a.ring + (((size_t)s * NFIELD) * d.nSteps + sa.it) * d.ringLen;
        const size_t fs = (size_t)d.nSteps * d.ringLen;
        for (int idx = t0; idx < d.ringLen; idx += nth) {
            int z, x;
            ring_cell(d, idx, z, x);
            const size_t i = (size_t)z * ld + x;
#pragma unroll
            for (int f = 0; f < NFIELD; f++) rb[f * fs + idx] = src[(size_t)f * d.fsz + i];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// forward step of one (strip, chunk)
struct FwdCtx {          // per-warp constants of the march
    const float *g, *m, *pv_src, *cxs, *cxv, *cz, *amp;
    float *o, *pv_dst, *ps;
    size_t fsz;
    int ld, nzA, nPml, zc0, zc1, zs, xs, xq0;
    unsigned amask;
    bool lown, xps, xpv;
    float c1z, c2z, c1x, c2x, dt;
};
struct FwdWin {          // register windows and the double-buffered row operands
    float4 vz[6], vx[6], zz[6], xz[6], xx[6];
    float4 ozz[2], oxz[2], oxx[2], lam[2], mu[2], mua[2], bya[2], byb[2];
};

// one row: prefetch row r+1's operands, stress at row r, velocity at row r-2.  U = r's phase in the 6-slot rotation.
template <bool EDGE, int U>
__device__ __forceinline__ void stream_fwd_row(const FwdCtx &k, FwdWin &w, const int r)
{
    const int ld = k.ld, nzA = k.nzA;
    const size_t fsz = k.fsz;
    const float c1z = k.c1z, c2z = k.c2z, c1x = k.c1x, c2x = k.c2x, dt = k.dt;
    constexpr int u = U;
    constexpr int cb = u & 1, nb = cb ^ 1;
    auto rowoff = [&](int row) { return (size_t)min(max(row, 0), nzA - 1) * ld; };   // clamped rows feed inactive / unowned cells only
    if (!EDGE) {   // ---- prefetch the rows of the next iteration (register double buffering)
        const size_t r3 = rowoff(r + 3), r1 = rowoff(r + 1), rq = rowoff(r - 1);
        w.vz[(u + 5) % 6] = ldq(k.g + F_VZ * fsz + r3); w.vx[(u + 5) % 6] = ldq(k.g + F_VX * fsz + r3);
        w.ozz[nb] = ldq(k.g + F_SZZ * fsz + r1); w.oxz[nb] = ldq(k.g + F_SXZ * fsz + r1); w.oxx[nb] = ldq(k.g + F_SXX * fsz + r1);
        w.lam[nb] = ldq(k.m + M_LAM * fsz + r1); w.mu[nb] = ldq(k.m + M_MU * fsz + r1); w.mua[nb] = ldq(k.m + M_MUAVE * fsz + r1);
        w.bya[nb] = ldq(k.m + M_BYCA * fsz + rq); w.byb[nb] = ldq(k.m + M_BYCB * fsz + rq);
    } else {       // ---- edge warps: this row's operands, no look-ahead (the extra CPML state needs the registers; these
                   //      few warps are latency-bound either way and overlap with the interior warps of the SM)
        const size_t r2 = rowoff(r + 2), r0 = rowoff(r), rq = rowoff(r - 2);
        w.vz[(u + 4) % 6] = ldq(k.g + F_VZ * fsz + r2); w.vx[(u + 4) % 6] = ldq(k.g + F_VX * fsz + r2);
        w.ozz[cb] = ldq(k.g + F_SZZ * fsz + r0); w.oxz[cb] = ldq(k.g + F_SXZ * fsz + r0); w.oxx[cb] = ldq(k.g + F_SXX * fsz + r0);
        w.lam[cb] = ldq(k.m + M_LAM * fsz + r0); w.mu[cb] = ldq(k.m + M_MU * fsz + r0); w.mua[cb] = ldq(k.m + M_MUAVE * fsz + r0);
        w.bya[cb] = ldq(k.m + M_BYCA * fsz + rq); w.byb[cb] = ldq(k.m + M_BYCB * fsz + rq);
        const size_t n2 = rowoff(r + 3), n0 = rowoff(r + 1), nq = rowoff(r - 1);
        pf_l1(k.g + F_VZ * fsz + n2); pf_l1(k.g + F_VX * fsz + n2);
        pf_l1(k.g + F_SZZ * fsz + n0); pf_l1(k.g + F_SXZ * fsz + n0); pf_l1(k.g + F_SXX * fsz + n0);
        pf_l1(k.m + M_LAM * fsz + n0); pf_l1(k.m + M_MU * fsz + n0); pf_l1(k.m + M_MUAVE * fsz + n0);
        pf_l1(k.m + M_BYCA * fsz + nq); pf_l1(k.m + M_BYCB * fsz + nq);
    }
    // ---- stress at row r from v rows r-2 .. r+2 (slots u .. u+4)
    {
        const float4 a0 = w.vz[u % 6], a1 = w.vz[(u + 1) % 6], a2 = w.vz[(u + 2) % 6], a3 = w.vz[(u + 3) % 6];
        const float4 b0 = w.vx[(u + 1) % 6], b1 = w.vx[(u + 2) % 6], b2 = w.vx[(u + 3) % 6], b3 = w.vx[(u + 4) % 6];
        const float vxc[7] = XWIN_B(b1), vzc[7] = XWIN_F(a2);
        const float vzm2[4] = Q4(a0), vzm1[4] = Q4(a1), vzq[4] = Q4(a2), vzp1[4] = Q4(a3);
        const float vxm1[4] = Q4(b0), vxq[4] = Q4(b1), vxp1[4] = Q4(b2), vxp2[4] = Q4(b3);
        const float l[4] = Q4(w.lam[cb]), mm[4] = Q4(w.mu[cb]), ma[4] = Q4(w.mua[cb]);
        const float pzz[4] = Q4(w.ozz[cb]), pxz[4] = Q4(w.oxz[cb]), pxx[4] = Q4(w.oxx[cb]);
        float nzz[4], nxz[4], nxx[4];
        const bool rowact = !EDGE || (r >= 2 && r <= nzA - 3);
        const bool zp = EDGE && rowact && ((r < k.nPml) || (r > nzA - k.nPml - 1));
        const bool xp = EDGE && rowact && k.xps;
        const bool rown = (r >= k.zc0) && (r < k.zc1);
        const size_t ro = (size_t)r * ld;       // only dereferenced for owned / active rows
        float pzq[4] = {0.f, 0.f, 0.f, 0.f}, pzh[4] = {0.f, 0.f, 0.f, 0.f}, pxq[4] = {0.f, 0.f, 0.f, 0.f}, pxh[4] = {0.f, 0.f, 0.f, 0.f};
        float bz = 0.f, az = 0.f, rkz = 1.f, bzh = 0.f, azh = 0.f, rkzh = 1.f;
        float bx[4], ax[4], rkx[4], bxh[4], axh[4], rkxh[4];
        if (EDGE) {
            if (zp) {
                const float4 t0 = ldq(k.pv_src + (size_t)P_VZ_Z * fsz + ro), t1 = ldq(k.pv_src + (size_t)P_VX_Z * fsz + ro);
                pzq[0] = t0.x; pzq[1] = t0.y; pzq[2] = t0.z; pzq[3] = t0.w;
                pzh[0] = t1.x; pzh[1] = t1.y; pzh[2] = t1.z; pzh[3] = t1.w;
                const float *cz = k.cz + r;
                bz = cz[C_B * nzA]; az = cz[C_A * nzA]; rkz = cz[C_RK * nzA];
                bzh = cz[C_BH * nzA]; azh = cz[C_AH * nzA]; rkzh = cz[C_RKH * nzA];
            }
            if (xp) {
                const float4 t0 = ldq(k.pv_src + (size_t)P_VX_X * fsz + ro), t1 = ldq(k.pv_src + (size_t)P_VZ_X * fsz + ro);
                pxq[0] = t0.x; pxq[1] = t0.y; pxq[2] = t0.z; pxq[3] = t0.w;
                pxh[0] = t1.x; pxh[1] = t1.y; pxh[2] = t1.z; pxh[3] = t1.w;
                const float *cx = k.cxs;
                const float4 q0 = ldq(cx + C_B * ld), q1 = ldq(cx + C_A * ld), q2 = ldq(cx + C_RK * ld);
                const float4 q3 = ldq(cx + C_BH * ld), q4 = ldq(cx + C_AH * ld), q5 = ldq(cx + C_RKH * ld);
                bx[0] = q0.x; bx[1] = q0.y; bx[2] = q0.z; bx[3] = q0.w; ax[0] = q1.x; ax[1] = q1.y; ax[2] = q1.z; ax[3] = q1.w;
                rkx[0] = q2.x; rkx[1] = q2.y; rkx[2] = q2.z; rkx[3] = q2.w; bxh[0] = q3.x; bxh[1] = q3.y; bxh[2] = q3.z; bxh[3] = q3.w;
                axh[0] = q4.x; axh[1] = q4.y; axh[2] = q4.z; axh[3] = q4.w; rkxh[0] = q5.x; rkxh[1] = q5.y; rkxh[2] = q5.z; rkxh[3] = q5.w;
            }
        }
#pragma unroll
        for (int c = 0; c < 4; c++) {
            float dvz_dz = c1z * (vzq[c] - vzm1[c]) - c2z * (vzp1[c] - vzm2[c]);
            float dvx_dx = DX7(vxc, c);
            float dvx_dz = c1z * (vxp1[c] - vxq[c]) - c2z * (vxp2[c] - vxm1[c]);
            float dvz_dx = DX7(vzc, c);
            const bool act = !EDGE || (rowact && ((k.amask >> c) & 1u));
            if (EDGE) {
                if (zp) {      // el_stress.cu:58-62
                    const float m0 = bz * pzq[c] + az * dvz_dz, m1 = bzh * pzh[c] + azh * dvx_dz;
                    dvz_dz = dvz_dz * rkz + m0; dvx_dz = dvx_dz * rkzh + m1;
                    if (act) { pzq[c] = m0; pzh[c] = m1; }
                }
                if (xp) {      // el_stress.cu:63-67 (table neutral -- 1/K = 1, a = b = 0 -- outside x < nPml || x > nx-nPml-1)
                    const float m0 = bx[c] * pxq[c] + ax[c] * dvx_dx, m1 = bxh[c] * pxh[c] + axh[c] * dvz_dx;
                    dvx_dx = dvx_dx * rkx[c] + m0; dvz_dx = dvz_dx * rkxh[c] + m1;
                    if (act) { pxq[c] = m0; pxh[c] = m1; }
                }
            }
            const float l2u = l[c] + 2.0f * mm[c];
            const float tzz = pzz[c] + (l2u * dvz_dz + l[c] * dvx_dx) * dt;
            const float txx = pxx[c] + (l[c] * dvz_dz + l2u * dvx_dx) * dt;
            const float txz = pxz[c] + ma[c] * (dvx_dz + dvz_dx) * dt;
            nzz[c] = act ? tzz : pzz[c]; nxx[c] = act ? txx : pxx[c]; nxz[c] = act ? txz : pxz[c];
        }
        if (r == k.zs) {      // explosive source, add_source utilities.cu:524-552 (after the update, in every warp that recomputes the cell)
            const float amp = *k.amp;
#pragma unroll
            for (int c = 0; c < 4; c++)
                if (k.xq0 + c == k.xs) { nzz[c] += amp; nxx[c] += amp; }
        }
        const float4 rzz = mk4(nzz), rxz = mk4(nxz), rxx = mk4(nxx);
        w.zz[u % 6] = rzz; w.xz[u % 6] = rxz; w.xx[u % 6] = rxx;          // stress row r lives in slot u
        if (k.lown && rown) {
            stq(k.o + F_SZZ * fsz + ro, rzz); stq(k.o + F_SXZ * fsz + ro, rxz); stq(k.o + F_SXX * fsz + ro, rxx);
            if (EDGE) {
                if (zp) { stq(k.pv_dst + (size_t)P_VZ_Z * fsz + ro, mk4(pzq)); stq(k.pv_dst + (size_t)P_VX_Z * fsz + ro, mk4(pzh)); }
                if (xp) { stq(k.pv_dst + (size_t)P_VX_X * fsz + ro, mk4(pxq)); stq(k.pv_dst + (size_t)P_VZ_X * fsz + ro, mk4(pxh)); }
            }
        }
    }
    // ---- velocity at row q = r-2 : szz rows q-1..q+2, sxz rows q-2..q+1, sxx row q (stress row r-j lives in slot u-j)
    {
        const int q = r - 2;
        const float4 p0 = w.zz[(u + 3) % 6], p1 = w.zz[(u + 4) % 6], p2 = w.zz[(u + 5) % 6], p3 = w.zz[u % 6];
        const float4 q0 = w.xz[(u + 2) % 6], q1 = w.xz[(u + 3) % 6], q2 = w.xz[(u + 4) % 6], q3 = w.xz[(u + 5) % 6];
        const float4 xc = w.xx[(u + 4) % 6];
        const float xzc[7] = XWIN_B(q2), xxc[7] = XWIN_F(xc);
        const float zzm1[4] = Q4(p0), zzc[4] = Q4(p1), zzp1[4] = Q4(p2), zzp2[4] = Q4(p3);
        const float xzm2[4] = Q4(q0), xzm1[4] = Q4(q1), xzp1[4] = Q4(q3);
        const float ovz[4] = Q4(w.vz[u % 6]), ovx[4] = Q4(w.vx[u % 6]);      // v row r-2 lives in slot u
        const float ba[4] = Q4(w.bya[cb]), bb[4] = Q4(w.byb[cb]);
        float nvz[4], nvx[4];
        const bool qown = (q >= k.zc0) && (q < k.zc1);
        const bool rowact = !EDGE || (q >= 2 && q <= nzA - 3);
        const bool zp = EDGE && rowact && qown && ((q < k.nPml) || (q > nzA - k.nPml - 1));
        const bool xp = EDGE && rowact && qown && k.xpv && k.lown;
        const size_t ro = (size_t)q * ld;
        float pzq[4] = {0.f, 0.f, 0.f, 0.f}, pzh[4] = {0.f, 0.f, 0.f, 0.f}, pxq[4] = {0.f, 0.f, 0.f, 0.f}, pxh[4] = {0.f, 0.f, 0.f, 0.f};
        float bz = 0.f, az = 0.f, rkz = 1.f, bzh = 0.f, azh = 0.f, rkzh = 1.f;
        float bx[4], ax[4], rkx[4], bxh[4], axh[4], rkxh[4];
        if (EDGE) {
            if (zp) {
                const float4 t0 = ldq_rw(k.ps + (size_t)P_SZZ_Z * fsz + ro), t1 = ldq_rw(k.ps + (size_t)P_SXZ_Z * fsz + ro);
                pzh[0] = t0.x; pzh[1] = t0.y; pzh[2] = t0.z; pzh[3] = t0.w;
                pzq[0] = t1.x; pzq[1] = t1.y; pzq[2] = t1.z; pzq[3] = t1.w;
                const float *cz = k.cz + q;
                bz = cz[C_B * nzA]; az = cz[C_A * nzA]; rkz = cz[C_RK * nzA];
                bzh = cz[C_BH * nzA]; azh = cz[C_AH * nzA]; rkzh = cz[C_RKH * nzA];
            }
            if (xp) {
                const float4 t0 = ldq_rw(k.ps + (size_t)P_SXZ_X * fsz + ro), t1 = ldq_rw(k.ps + (size_t)P_SXX_X * fsz + ro);
                pxq[0] = t0.x; pxq[1] = t0.y; pxq[2] = t0.z; pxq[3] = t0.w;
                pxh[0] = t1.x; pxh[1] = t1.y; pxh[2] = t1.z; pxh[3] = t1.w;
                const float *cx = k.cxv;
                const float4 e0 = ldq(cx + C_B * ld), e1 = ldq(cx + C_A * ld), e2 = ldq(cx + C_RK * ld);
                const float4 e3 = ldq(cx + C_BH * ld), e4 = ldq(cx + C_AH * ld), e5 = ldq(cx + C_RKH * ld);
                bx[0] = e0.x; bx[1] = e0.y; bx[2] = e0.z; bx[3] = e0.w; ax[0] = e1.x; ax[1] = e1.y; ax[2] = e1.z; ax[3] = e1.w;
                rkx[0] = e2.x; rkx[1] = e2.y; rkx[2] = e2.z; rkx[3] = e2.w; bxh[0] = e3.x; bxh[1] = e3.y; bxh[2] = e3.z; bxh[3] = e3.w;
                axh[0] = e4.x; axh[1] = e4.y; axh[2] = e4.z; axh[3] = e4.w; rkxh[0] = e5.x; rkxh[1] = e5.y; rkxh[2] = e5.z; rkxh[3] = e5.w;
            }
        }
#pragma unroll
        for (int c = 0; c < 4; c++) {
            float dszz_dz = c1z * (zzp1[c] - zzc[c]) - c2z * (zzp2[c] - zzm1[c]);
            float dsxz_dz = c1z * (xzc[c + 2] - xzm1[c]) - c2z * (xzp1[c] - xzm2[c]);
            float dsxz_dx = DX7(xzc, c);
            float dsxx_dx = DX7(xxc, c);
            const bool act = !EDGE || (rowact && ((k.amask >> c) & 1u));
            if (EDGE) {
                if (zp) {      // el_velocity.cu:51-55
                    const float m0 = bzh * pzh[c] + azh * dszz_dz, m1 = bz * pzq[c] + az * dsxz_dz;
                    dszz_dz = dszz_dz * rkzh + m0; dsxz_dz = dsxz_dz * rkz + m1;
                    if (act) { pzh[c] = m0; pzq[c] = m1; }
                }
                if (xp) {      // el_velocity.cu:56-60 (table neutral outside x < nPml || x > nx-nPml)
                    const float m0 = bx[c] * pxq[c] + ax[c] * dsxz_dx, m1 = bxh[c] * pxh[c] + axh[c] * dsxx_dx;
                    dsxz_dx = dsxz_dx * rkx[c] + m0; dsxx_dx = dsxx_dx * rkxh[c] + m1;
                    if (act) { pxq[c] = m0; pxh[c] = m1; }
                }
            }
            const float tvz = ovz[c] + (dszz_dz + dsxz_dx) * ba[c] * dt;
            const float tvx = ovx[c] + (dsxz_dz + dsxx_dx) * bb[c] * dt;
            nvz[c] = act ? tvz : ovz[c]; nvx[c] = act ? tvx : ovx[c];
        }
        if (k.lown && qown) {
            stq(k.o + F_VZ * fsz + ro, mk4(nvz)); stq(k.o + F_VX * fsz + ro, mk4(nvx));
            if (EDGE) {
                if (zp) { stq(k.ps + (size_t)P_SZZ_Z * fsz + ro, mk4(pzh)); stq(k.ps + (size_t)P_SXZ_Z * fsz + ro, mk4(pzq)); }
                if (xp) { stq(k.ps + (size_t)P_SXZ_X * fsz + ro, mk4(pxq)); stq(k.ps + (size_t)P_SXX_X * fsz + ro, mk4(pxh)); }
            }
        }
    }
}

template <bool EDGE>
__device__ __forceinline__ void stream_fwd_body(const KArgs &a, const StreamArgs &sa, const int s, const int4 wk, const int lane)
{
    const Dims &d = a.d;
    FwdCtx k;
    k.ld = d.ldx; k.nzA = d.nzA; k.nPml = d.nPml; k.fsz = d.fsz;
    const size_t fsz = d.fsz;
    const int p = sa.it & 1;
    float *st = slot_state(a, s);
    k.xq0 = wk.x - 4 + 4 * lane;                 // true first column of this lane's quad
    const bool colok = (k.xq0 >= 0) && (k.xq0 < d.ldx);
    const int xq = colok ? k.xq0 : 0;                      // lanes outside the array read column 0; their values feed inactive cells only
    k.g = st + (size_t)(p ? S_FWD1 : S_FWD) * fsz + xq;
    k.o = st + (size_t)(p ? S_FWD : S_FWD1) * fsz + xq;
    k.m = a.model + xq;
    k.pv_src = st + (size_t)(p ? S_FPSIV1 : S_FPSI) * fsz + xq;   // stress-side CPML memory: ping-pong (halo recompute reads the old one)
    k.pv_dst = st + (size_t)(p ? S_FPSI : S_FPSIV1) * fsz + xq;
    k.ps = st + (size_t)S_FPSI * fsz + xq;                         // velocity-side CPML memory: owner-only, in place
    k.cxs = a.cxs + xq; k.cxv = a.cxv + xq; k.cz = a.cz;
    k.amp = a.t.amp + (size_t)s * d.nSteps + sa.it;
    k.zc0 = wk.y; k.zc1 = wk.z;
    k.lown = (lane >= 1) && (lane <= 30) && colok;
    k.zs = a.t.zs[s]; k.xs = a.t.xs[s];
    k.c1z = d.c1z; k.c2z = d.c2z; k.c1x = d.c1x; k.c2x = d.c2x; k.dt = d.dt;
    // EDGE: active columns of the quad (x in [2, nx-3]) and whether the quad touches the x CPML strips
    k.amask = 0xf; k.xps = false; k.xpv = false;
    if (EDGE) {
        k.amask = 0;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int x = k.xq0 + c;
            if (x >= 2 && x <= d.nx - 3) {
                k.amask |= 1u << c;
                if ((x < d.nPml) || (x > d.nx - d.nPml - 1)) k.xps = true;     // el_stress.cu:63
                if ((x < d.nPml) || (x > d.nx - d.nPml)) k.xpv = true;         // el_velocity.cu:56,71
            }
        }
    }

    FwdWin w;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 6; j++) { w.vz[j] = w.vx[j] = w.zz[j] = w.xz[j] = w.xx[j] = zero; }
    auto rowoff = [&](int row) { return (size_t)min(max(row, 0), d.nzA - 1) * d.ldx; };
    // v rows zc0-4 .. zc0 -> slots 0..4 ; row rho lives in slot (rho - (zc0-4)) % 6   (edge warps load row zc0 in the first iteration)
#pragma unroll
    for (int j = 0; j < (EDGE ? 4 : 5); j++) {
        const size_t ro = rowoff(k.zc0 - 4 + j);
        w.vz[j] = ldq(k.g + F_VZ * fsz + ro); w.vx[j] = ldq(k.g + F_VX * fsz + ro);
    }
    w.ozz[0] = w.oxz[0] = w.oxx[0] = w.lam[0] = w.mu[0] = w.mua[0] = w.bya[0] = w.byb[0] = zero;
    w.ozz[1] = w.oxz[1] = w.oxx[1] = w.lam[1] = w.mu[1] = w.mua[1] = w.bya[1] = w.byb[1] = zero;
    if (!EDGE) {
        const size_t ro = rowoff(k.zc0 - 2);
        w.ozz[0] = ldq(k.g + F_SZZ * fsz + ro); w.oxz[0] = ldq(k.g + F_SXZ * fsz + ro); w.oxx[0] = ldq(k.g + F_SXX * fsz + ro);
        w.lam[0] = ldq(k.m + M_LAM * fsz + ro); w.mu[0] = ldq(k.m + M_MU * fsz + ro); w.mua[0] = ldq(k.m + M_MUAVE * fsz + ro);
    }
    const int niter = (k.zc1 - k.zc0) + 4;
    if (!EDGE) {
        // interior: rotate the windows by full unrolling (6 rows per trip); surplus rows of the last trip are computed and dropped
#pragma unroll 1
        for (int kk = 0; kk < niter; kk += 6) {
            const int r = k.zc0 - 2 + kk;
            stream_fwd_row<EDGE, 0>(k, w, r);     stream_fwd_row<EDGE, 1>(k, w, r + 1); stream_fwd_row<EDGE, 2>(k, w, r + 2);
            stream_fwd_row<EDGE, 3>(k, w, r + 3); stream_fwd_row<EDGE, 4>(k, w, r + 4); stream_fwd_row<EDGE, 5>(k, w, r + 5);
        }
    } else {
        // edge: one row per trip and explicit register moves, so the (much longer) body stays resident in the instruction cache
#pragma unroll 1
        for (int kk = 0; kk < niter; kk++) {
            stream_fwd_row<EDGE, 0>(k, w, k.zc0 - 2 + kk);
#pragma unroll
            for (int j = 0; j < 4; j++) { w.vz[j] = w.vz[j + 1]; w.vx[j] = w.vx[j + 1]; }
            w.zz[3] = w.zz[4]; w.zz[4] = w.zz[5]; w.zz[5] = w.zz[0];
            w.xz[2] = w.xz[3]; w.xz[3] = w.xz[4]; w.xz[4] = w.xz[5]; w.xz[5] = w.xz[0];
            w.xx[4] = w.xx[5]; w.xx[5] = w.xx[0];
        }
    }
}

// the CPML / rim variant is compiled out of line so that its extra live state cannot spill the interior loop
__device__ __forceinline__ void stream_fwd_edge(const KArgs &a, const StreamArgs &sa, const int s, const int4 wk, const int lane)
{ stream_fwd_body<true>(a, sa, s, wk, lane); }

#ifndef SW_MINB
#define SW_MINB 2
#endif
// grid: x = nAux + ceil(nWork / SW_WPB), y = slot
__global__ void __launch_bounds__(SW_NT, SW_MINB) k_stream_fwd(const KArgs a, const StreamArgs sa)
{
    const int s = blockIdx.y;
    if ((int)blockIdx.x < sa.nAux) { stream_fwd_aux(a, sa, s); return; }
    const int wg = ((int)blockIdx.x - sa.nAux) * SW_WPB + ((int)threadIdx.x >> 5);
    if (wg >= sa.nWork) return;      // whole warp leaves (the in-place CPML memory must not be updated twice)
    const int4 wk = __ldg(sa.work + wg);
    const int lane = threadIdx.x & 31;
    if ((wk.w == 0 || sa.force == 1) && sa.force != 2) stream_fwd_body<false>(a, sa, s, wk, lane);
    else stream_fwd_edge(a, sa, s, wk, lane);
}

}  // namespace sepfwi

// kernels_stream_bwd.cuh -- EXPERIMENT, NOT PART OF THE PRODUCT BUILD (round 2; see profiles/README.md "measured and not kept").
// The reverse-time step as ONE pass over the adjoint state (sm_100a).  Parity-green (60 / 60 -m gpu tests) but slower than the two
// launches it replaces: on the 8000 x 2000 grid it moves 2.17 GB per step instead of 2.94 GB, yet takes 614 us instead of 502 us --
// at 8 warps per SM the streaming kernels are bound by instruction issue latency (25 - 32 % issue utilisation), the one-pass row
// executes the sum of both kernels' instructions plus 184 register moves per row (its six-times-unrolled variant, without the
// moves, overflows the 32 KB L1.5 instruction cache: 2.0 no_instruction stalls per issue, 665 us).  To build it again:
// include it after kernels_stream.cuh and launch k_stream_bwd with StreamArgs.nEdge = number of edge items in the work list.
//
// The two halves of a reverse-time step -- reconstruction + imaging (k_stream_recon: fwd(it+1) -> fwd(it), gradients += ...)
// and the adjoint sweep (k_stream_adj: adj(it+1) -> adj(it)) -- are independent within a step, both read the adjoint state of
// time it+1 and the same five coefficient arrays, and the imaging needs exactly the adjoint values the adjoint sweep holds in
// its register windows at the same rows:
//     imaging of rho   at row r    uses  vz^, vx^ (it+1) at row r    = the "old" velocities of adjoint phase A at row r
//     imaging of l, mu at row r-2  uses  szz^, sxz^, sxx^ at row r-2 = slot u of the adjoint sweep's stress windows
// As two launches they move 104 + 60 = 164 B per cell (measured on the 8000 x 2000 grid: 1.81 + 1.13 GB per step).  Here ONE
// warp marches both sweeps down the same (strip, chunk): per row 18 arrays are read from HBM instead of 28 and 13 written --
// 124 B per cell against 116 B algorithmic.  The five coefficient rows are needed at two row offsets (lambda, mu, mu_ave: row r
// by the adjoint velocities, row r-2 by the stress reconstruction; the buoyancies the other way round); the second request of a
// row is issued two iterations after the first and is served by L2, it costs no HBM traffic.
//
// Interior items only (no CPML, no inactive rim, no boundary ring): the edge items of the work list keep the two specialised
// bodies of kernels_stream.cuh -- they run in the same launch, as reconstruction CTAs and adjoint CTAs next to each other.
// Arithmetic per cell is exactly that of stream_rec_row<false> / stream_adj_row<false>, in the same order.
// Reference lines: el_velocity.cu:84-117, el_stress.cu:89-128, el_velocity_adj.cu:22-108, el_stress_adj.cu:22-104,
// res_injection_exx/_ezz utilities.cu:605-641, add_source(isFor=false) utilities.cu:541-550.
#pragma once
#include "kernels_stream.cuh"

namespace sepfwi {

#ifndef SW_UNR_BWD
#define SW_UNR_BWD 1
#endif
constexpr int BF_NARR = 23, BF_NST = 2;
constexpr int BF_RING_BYTES = BF_NST * BF_NARR * 512;              // 23 KB per warp
constexpr int BF_WARP_BYTES = BF_RING_BYTES + 256 * 4;             // + the injection staging row
constexpr size_t BF_SMEM = (size_t)SW_WPB * BF_WARP_BYTES;         // 96 KB per CTA, 2 CTAs per SM
enum { BF_SZZ = 0, BF_SXZ, BF_SXX, BF_OVZ, BF_OVX,                 // forward state it+1: rows r+2, r+1, r, r, r
       BF_ASZ, BF_ASX, BF_ASXZ, BF_AVZ, BF_AVX,                    // adjoint state it+1: rows r+2 (stresses), r (velocities)
       BF_GR, BF_GL, BF_GM,                                        // gradients: row r (rho), r-2 (lambda, mu)
       BF_LAM, BF_MU, BF_MUA, BF_BA, BF_BB,                        // coefficients of row r
       BF_LAM2, BF_MU2, BF_MUA2, BF_BA2, BF_BB2 };                 // coefficients of row r-2 (second request: L2)

struct BwdCtx {
    const float *g, *adj, *m, *amp, *res;
    float *o, *ao, *grad;
    const int *injPtr, *injList;
    const SlotTab *t;
    float *stage;
    unsigned ring_s;
    const float4 *ring_p;
    size_t fsz, tb, cb;
    int ld, nzA, zc0, zc1, zs, xs, xq0, lane, s, nSteps;
    bool lown, anyinj;
    float c1z, c2z, c1x, c2x, dt;
};
struct BwdWin {
    float4 szz[6], sxz[6], sxx[6];     // forward stresses of time it+1: rows r-2 .. r+2 / r+1 / r
    float4 vz[6], vx[6];               // reconstructed forward velocities (time it): rows r-4 .. r
    float4 asz[6], asx[6], asxz[6];    // adjoint stresses of time it+1: rows r-2 .. r+2
    float4 avz[6], avx[6];             // new adjoint velocities: rows r-4 .. r
    float ga_prev[4], sh_prev[4];      // density term of row r-1, shear term of row r-3
};

// request the operands of the iteration whose row is r
__device__ __forceinline__ void stream_bwd_issue(const BwdCtx &k, const int r, const int stage)
{
    const int ld = k.ld, nzA = k.nzA;
    const size_t fsz = k.fsz;
    auto rowoff = [&](int row) { return (size_t)min(max(row, 0), nzA - 1) * ld; };
    const size_t r2 = rowoff(r + 2), r1 = rowoff(r + 1), r0 = rowoff(r), rq = rowoff(r - 2);
    const unsigned sb = k.ring_s + (unsigned)stage * (BF_NARR * 512);
    cp16(sb + BF_SZZ * 512, k.g + F_SZZ * fsz + r2); cp16(sb + BF_SXZ * 512, k.g + F_SXZ * fsz + r1); cp16(sb + BF_SXX * 512, k.g + F_SXX * fsz + r0);
    cp16(sb + BF_OVZ * 512, k.g + F_VZ * fsz + r0); cp16(sb + BF_OVX * 512, k.g + F_VX * fsz + r0);
    cp16(sb + BF_ASZ * 512, k.adj + F_SZZ * fsz + r2); cp16(sb + BF_ASX * 512, k.adj + F_SXX * fsz + r2); cp16(sb + BF_ASXZ * 512, k.adj + F_SXZ * fsz + r2);
    cp16(sb + BF_AVZ * 512, k.adj + F_VZ * fsz + r0); cp16(sb + BF_AVX * 512, k.adj + F_VX * fsz + r0);
    cp16(sb + BF_GR * 512, k.grad + 2 * fsz + r0); cp16(sb + BF_GL * 512, k.grad + 0 * fsz + rq); cp16(sb + BF_GM * 512, k.grad + 1 * fsz + rq);
    cp16(sb + BF_LAM * 512, k.m + M_LAM * fsz + r0); cp16(sb + BF_MU * 512, k.m + M_MU * fsz + r0); cp16(sb + BF_MUA * 512, k.m + M_MUAVE * fsz + r0);
    cp16(sb + BF_BA * 512, k.m + M_BYCA * fsz + r0); cp16(sb + BF_BB * 512, k.m + M_BYCB * fsz + r0);
    cp16(sb + BF_LAM2 * 512, k.m + M_LAM * fsz + rq); cp16(sb + BF_MU2 * 512, k.m + M_MU * fsz + rq); cp16(sb + BF_MUA2 * 512, k.m + M_MUAVE * fsz + rq);
    cp16(sb + BF_BA2 * 512, k.m + M_BYCA * fsz + rq); cp16(sb + BF_BB2 * 512, k.m + M_BYCB * fsz + rq);
    cp_commit();
}

// same table walk as stream_adj_inject (kernels_stream.cuh), on the fused context
__device__ __forceinline__ void stream_bwd_inject(const BwdCtx &k, const int r, float nvz[4], float nvx[4])
{
    if (r < 0 || r >= k.nzA) return;
    const int k0 = k.injPtr[r], k1 = k.injPtr[r + 1];
    if (k1 <= k0) return;
    const SlotTab &t = *k.t;
    float *sg = k.stage;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4 *>(sg + 4 * k.lane) = zero;
    *reinterpret_cast<float4 *>(sg + 128 + 4 * k.lane) = zero;
    __syncwarp();
    const int xbase = k.xq0 - 4 * k.lane;      // column of lane 0, component 0
    for (int j = k0 + k.lane; j < k1; j += 32) {
        const int mi = k.injList[j];
        const int cell = t.injCell[k.tb + mi];
        const int x = cell - r * k.ld;
        const int p0 = t.injPtr[(size_t)k.s * (t.maxInj + 1) + mi], p1 = t.injPtr[(size_t)k.s * (t.maxInj + 1) + mi + 1];
        float v = 0.f;
        for (int p = p0; p < p1; p++) v += t.injCoef[k.cb + p] * k.res[(size_t)t.injRec[k.cb + p] * k.nSteps];
        sg[(t.injField[k.tb + mi] == F_VZ ? 0 : 128) + (x - xbase)] = v;
    }
    __syncwarp();
    const float4 dz = *reinterpret_cast<const float4 *>(sg + 4 * k.lane), dx = *reinterpret_cast<const float4 *>(sg + 128 + 4 * k.lane);
    nvz[0] += dz.x; nvz[1] += dz.y; nvz[2] += dz.z; nvz[3] += dz.w;
    nvx[0] += dx.x; nvx[1] += dx.y; nvx[2] += dx.z; nvx[3] += dx.w;
    __syncwarp();
}

template <int U>
__device__ __forceinline__ void stream_bwd_row(const BwdCtx &k, BwdWin &w, const int r, const int stage)
{
    const int ld = k.ld;
    const size_t fsz = k.fsz;
    const float c1z = k.c1z, c2z = k.c2z, c1x = k.c1x, c2x = k.c2x, dt = k.dt;
    constexpr int u = U;
    stream_bwd_issue(k, r + (BF_NST - 1), stage == 0 ? BF_NST - 1 : stage - 1);
    cp_wait<BF_NST - 1>();
    const float4 *sb = k.ring_p + stage * (BF_NARR * 32);
    w.szz[(u + 4) % 6] = sb[BF_SZZ * 32];       // row r+2
    w.sxz[(u + 3) % 6] = sb[BF_SXZ * 32];       // row r+1
    w.sxx[(u + 2) % 6] = sb[BF_SXX * 32];       // row r
    w.asz[(u + 4) % 6] = sb[BF_ASZ * 32]; w.asx[(u + 4) % 6] = sb[BF_ASX * 32]; w.asxz[(u + 4) % 6] = sb[BF_ASXZ * 32];     // row r+2
    const bool rown = (r >= k.zc0) && (r < k.zc1);
    const size_t ro = (size_t)r * ld;
    const float4 avz4 = sb[BF_AVZ * 32], avx4 = sb[BF_AVX * 32];       // adjoint velocities of time it+1 at row r: imaging AND phase A
    const float avz[4] = Q4(avz4), avx[4] = Q4(avx4);
    // ---- reconstruction stage 1: forward velocities of time `it` at row r ; density imaging (stream_rec_row stage 1)
    {
        const float4 p0 = w.szz[(u + 1) % 6], p1 = w.szz[(u + 2) % 6], p2 = w.szz[(u + 3) % 6], p3 = w.szz[(u + 4) % 6];
        const float4 q0 = w.sxz[u % 6], q1 = w.sxz[(u + 1) % 6], q2 = w.sxz[(u + 2) % 6], q3 = w.sxz[(u + 3) % 6];
        const float4 xc = w.sxx[(u + 2) % 6];
        const float wxz[7] = XWIN_B(q2), wxx[7] = XWIN_F(xc);
        const float zzm1[4] = Q4(p0), zzc[4] = Q4(p1), zzp1[4] = Q4(p2), zzp2[4] = Q4(p3);
        const float xzm2[4] = Q4(q0), xzm1[4] = Q4(q1), xzc[4] = Q4(q2), xzp1[4] = Q4(q3);
        const float4 ovz4 = sb[BF_OVZ * 32], ovx4 = sb[BF_OVX * 32], ba4 = sb[BF_BA * 32], bb4 = sb[BF_BB * 32];
        const float ovz[4] = Q4(ovz4), ovx[4] = Q4(ovx4), ba[4] = Q4(ba4), bb[4] = Q4(bb4);
        float nvz[4], nvx[4], ga[4], gb[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const float A = DZ4(zzm1[c], zzc[c], zzp1[c], zzp2[c]) + DX7(wxz, c);
            const float B = DZ4(xzm2[c], xzm1[c], xzc[c], xzp1[c]) + DX7(wxx, c);
            nvz[c] = ovz[c] - A * ba[c] * dt;
            nvx[c] = ovx[c] - B * bb[c] * dt;
            ga[c] = avz[c] * A * dt * (0.5f * ba[c] * ba[c]);
            gb[c] = avx[c] * B * dt * (0.5f * bb[c] * bb[c]);
        }
        const float4 rvz = mk4(nvz), rvx = mk4(nvx);
        w.vz[(u + 4) % 6] = rvz; w.vx[(u + 4) % 6] = rvx;
        const float gbl = sh_l(gb[3]);
        if (k.lown && rown) {
            const float4 g4 = sb[BF_GR * 32];
            float gr[4] = Q4(g4);
            const float gbW[5] = {gbl, gb[0], gb[1], gb[2], gb[3]};
#pragma unroll
            for (int c = 0; c < 4; c++) gr[c] += ga[c] + gbW[c + 1] + gbW[c] + w.ga_prev[c];
            stq(k.grad + 2 * fsz + ro, mk4(gr));
            stq(k.o + F_VZ * fsz + ro, rvz); stq(k.o + F_VX * fsz + ro, rvx);
        }
#pragma unroll
        for (int c = 0; c < 4; c++) w.ga_prev[c] = ga[c];
    }
    // ---- adjoint phase A: adjoint velocities at row r from the adjoint stresses of time it+1 (stream_adj_row phase A)
    {
        const float4 z1 = w.asz[(u + 1) % 6], z2 = w.asz[(u + 2) % 6], z3 = w.asz[(u + 3) % 6], z4 = w.asz[(u + 4) % 6];
        const float4 x1 = w.asx[(u + 1) % 6], x2 = w.asx[(u + 2) % 6], x3 = w.asx[(u + 3) % 6], x4 = w.asx[(u + 4) % 6];
        const float4 q0 = w.asxz[u % 6], q1 = w.asxz[(u + 1) % 6], q2 = w.asxz[(u + 2) % 6], q3 = w.asxz[(u + 3) % 6];
        const float wzz[7] = XWIN_F(z2), wxx[7] = XWIN_F(x2), wxz[7] = XWIN_B(q2);
        const float zm1[4] = Q4(z1), zc0[4] = Q4(z2), zp1[4] = Q4(z3), zp2[4] = Q4(z4);
        const float xm1[4] = Q4(x1), xc0[4] = Q4(x2), xp1[4] = Q4(x3), xp2[4] = Q4(x4);
        const float m2[4] = Q4(q0), m1[4] = Q4(q1), c0[4] = Q4(q2), p1[4] = Q4(q3);
        const float4 lam4 = sb[BF_LAM * 32], mu4 = sb[BF_MU * 32], mua4 = sb[BF_MUA * 32];
        const float l[4] = Q4(lam4), mm[4] = Q4(mu4), ma[4] = Q4(mua4);
        float nvz[4], nvx[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const float l2u = l[c] + 2.0f * mm[c];
            const float accx = (l[c] * -DX7(wzz, c) + l2u * -DX7(wxx, c)) * dt + ma[c] * -DZ4(m2[c], m1[c], c0[c], p1[c]) * dt;
            const float accz = (l2u * -DZ4(zm1[c], zc0[c], zp1[c], zp2[c]) + l[c] * -DZ4(xm1[c], xc0[c], xp1[c], xp2[c])) * dt + ma[c] * -DX7(wxz, c) * dt;
            nvx[c] = avx[c] + accx; nvz[c] = avz[c] + accz;
        }
        if (k.anyinj) stream_bwd_inject(k, r, nvz, nvx);
        const float4 rvz = mk4(nvz), rvx = mk4(nvx);
        w.avz[(u + 4) % 6] = rvz; w.avx[(u + 4) % 6] = rvx;
        if (k.lown && rown) { stq(k.ao + F_VZ * fsz + ro, rvz); stq(k.ao + F_VX * fsz + ro, rvx); }
    }
    const int q = r - 2;
    const bool qown = (q >= k.zc0) && (q < k.zc1);
    const size_t rq = (size_t)q * ld;
    const float4 za4 = w.asz[u % 6], sa4 = w.asxz[u % 6], xa4 = w.asx[u % 6];      // adjoint stresses of time it+1 at row r-2: slot u
    const float za[4] = Q4(za4), sa[4] = Q4(sa4), xa[4] = Q4(xa4);
    // ---- reconstruction stage 2: forward stresses of time `it` at row q = r-2 ; lambda / mu imaging (stream_rec_row stage 2)
    {
        const float4 v0 = w.vz[u % 6], v1 = w.vz[(u + 1) % 6], v2 = w.vz[(u + 2) % 6], v3 = w.vz[(u + 3) % 6];
        const float4 u0 = w.vx[(u + 1) % 6], u1 = w.vx[(u + 2) % 6], u2 = w.vx[(u + 3) % 6], u3 = w.vx[(u + 4) % 6];
        const float wvz[7] = XWIN_F(v2), wvx[7] = XWIN_B(u1);
        const float vzm2[4] = Q4(v0), vzm1[4] = Q4(v1), vzc[4] = Q4(v2), vzp1[4] = Q4(v3);
        const float vxm1[4] = Q4(u0), vxc[4] = Q4(u1), vxp1[4] = Q4(u2), vxp2[4] = Q4(u3);
        const float ozz[4] = Q4(w.szz[u % 6]), oxz[4] = Q4(w.sxz[u % 6]), oxx[4] = Q4(w.sxx[u % 6]);     // forward stresses of row r-2: slot u
        const float4 lam4 = sb[BF_LAM2 * 32], mu4 = sb[BF_MU2 * 32], mua4 = sb[BF_MUA2 * 32];
        const float lam[4] = Q4(lam4), mu[4] = Q4(mu4), mua[4] = Q4(mua4);
        float D1[4], D2[4], D3[4], sh[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            D1[c] = DZ4(vzm2[c], vzm1[c], vzc[c], vzp1[c]); D2[c] = DX7(wvx, c);
            D3[c] = DZ4(vxm1[c], vxc[c], vxp1[c], vxp2[c]) + DX7(wvz, c);
            sh[c] = -sa[c] * D3[c] * dt * (0.25f * mua[c] * mua[c]) * 1e6f;
        }
        const float shl = sh_l(sh[3]), shul = sh_l(w.sh_prev[3]);
        if (k.lown && qown) {
            const float4 g0 = sb[BF_GL * 32], g1 = sb[BF_GM * 32];
            float gl[4] = Q4(g0), gm[4] = Q4(g1);
            const float shW[5] = {shl, sh[0], sh[1], sh[2], sh[3]}, shUW[5] = {shul, w.sh_prev[0], w.sh_prev[1], w.sh_prev[2], w.sh_prev[3]};
            float nzz[4], nxz[4], nxx[4];
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int x = k.xq0 + c;
                float tzz = ozz[c], txx = oxx[c], txz = oxz[c];
                if (q == k.zs && x == k.xs) { const float amp = *k.amp; tzz -= amp; txx -= amp; }
                const float l2u = lam[c] + 2.0f * mu[c];
                tzz -= (l2u * D1[c] + lam[c] * D2[c]) * dt;
                txx -= (lam[c] * D1[c] + l2u * D2[c]) * dt;
                txz -= mua[c] * D3[c] * dt;
                gl[c] += -(za[c] + xa[c]) * (D1[c] + D2[c]) * dt * 1e6f;
                const float shg = shW[c + 1] + shW[c] + shUW[c + 1] + shUW[c];
                const float gmn = (-2.0f * za[c] * D1[c] * dt - 2.0f * xa[c] * D2[c] * dt) * 1e6f;
                gm[c] += gmn + __fdividef(shg, mu[c] * mu[c]);
                nzz[c] = tzz; nxx[c] = txx; nxz[c] = txz;
            }
            stq(k.grad + 0 * fsz + rq, mk4(gl)); stq(k.grad + 1 * fsz + rq, mk4(gm));
            stq(k.o + F_SZZ * fsz + rq, mk4(nzz)); stq(k.o + F_SXZ * fsz + rq, mk4(nxz)); stq(k.o + F_SXX * fsz + rq, mk4(nxx));
        }
#pragma unroll
        for (int c = 0; c < 4; c++) w.sh_prev[c] = sh[c];
    }
    // ---- adjoint phase B: adjoint stresses at row q = r-2 from the new adjoint velocities (stream_adj_row phase B)
    {
        const float4 v0 = w.avz[u % 6], v1 = w.avz[(u + 1) % 6], v2 = w.avz[(u + 2) % 6], v3 = w.avz[(u + 3) % 6];
        const float4 u0 = w.avx[(u + 1) % 6], u1 = w.avx[(u + 2) % 6], u2 = w.avx[(u + 3) % 6], u3 = w.avx[(u + 4) % 6];
        const float wvz[7] = XWIN_F(v2), wvx[7] = XWIN_B(u1);
        const float vzm2[4] = Q4(v0), vzm1[4] = Q4(v1), vzc[4] = Q4(v2), vzp1[4] = Q4(v3);
        const float vxm1[4] = Q4(u0), vxc[4] = Q4(u1), vxp1[4] = Q4(u2), vxp2[4] = Q4(u3);
        const float4 bya4 = sb[BF_BA2 * 32], byb4 = sb[BF_BB2 * 32];
        const float ba[4] = Q4(bya4), bb[4] = Q4(byb4);
        float nzz[4], nxz[4], nxx[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const float mdxf_vz = -DX7(wvz, c), mdzf_vx = -DZ4(vxm1[c], vxc[c], vxp1[c], vxp2[c]);
            const float mdxb_vx = -DX7(wvx, c), mdzb_vz = -DZ4(vzm2[c], vzm1[c], vzc[c], vzp1[c]);
            nxz[c] = sa[c] + (mdxf_vz * ba[c] * dt + mdzf_vx * bb[c] * dt);
            nxx[c] = xa[c] + bb[c] * mdxb_vx * dt;
            nzz[c] = za[c] + ba[c] * mdzb_vz * dt;
        }
        if (k.lown && qown) {
            stq(k.ao + F_SZZ * fsz + rq, mk4(nzz)); stq(k.ao + F_SXZ * fsz + rq, mk4(nxz)); stq(k.ao + F_SXX * fsz + rq, mk4(nxx));
        }
    }
}

// interior (strip, chunk) item: both sweeps in one march
__device__ __forceinline__ void stream_bwd_body(const KArgs &a, const StreamArgs &sa, const int s, const int4 wk, const int lane,
                                                float *stage, const unsigned smem_warp, const float4 *ring_ptr)
{
    const Dims &d = a.d;
    BwdCtx k;
    k.ld = d.ldx; k.nzA = d.nzA; k.fsz = d.fsz; k.lane = lane; k.s = s; k.nSteps = d.nSteps;
    const size_t fsz = d.fsz;
    float *st = slot_state(a, s);
    k.xq0 = wk.x - 4 + 4 * lane;
    const bool colok = (k.xq0 >= 0) && (k.xq0 < d.ldx);
    const int xq = colok ? k.xq0 : 0;
    k.g = st + (size_t)(sa.q ? S_FWD1 : S_FWD) * fsz + xq;
    k.o = st + (size_t)(sa.q ? S_FWD : S_FWD1) * fsz + xq;
    k.adj = st + (size_t)(sa.pa ? S_ADJ1 : S_ADJ) * fsz + xq;
    k.ao = st + (size_t)(sa.pa ? S_ADJ : S_ADJ1) * fsz + xq;
    k.m = a.model + xq;
    k.grad = a.grad + (size_t)s * 3 * fsz + xq;
    k.amp = a.t.amp + (size_t)s * d.nSteps + sa.it;
    k.res = a.trace + ((size_t)s * d.nTrace + T_RES) * d.maxRec * d.nSteps + sa.it;
    k.zs = a.t.zs[s]; k.xs = a.t.xs[s];
    k.zc0 = wk.y; k.zc1 = wk.z;
    k.lown = (lane >= 1) && (lane <= 30) && colok;
    k.c1z = d.c1z; k.c2z = d.c2z; k.c1x = d.c1x; k.c2x = d.c2x; k.dt = d.dt;
    k.t = &a.t; k.stage = stage;
    k.tb = (size_t)s * a.t.maxInj; k.cb = (size_t)s * a.t.maxCon;
    {   // injection targets of this strip (halo columns included), rows of phase A
        const int strip = wk.x / SW_OWN;
        k.injPtr = a.t.sInjPtr + ((size_t)s * a.t.nStrips + strip) * (d.nzA + 1);
        k.injList = a.t.sInj + (size_t)s * 2 * a.t.maxInj;
        const int ra = max(k.zc0 - 2, 0), rb = min(k.zc1 + 2, d.nzA);
        k.anyinj = k.injPtr[rb] > k.injPtr[ra];
    }
    k.ring_s = smem_warp + lane * 16; k.ring_p = ring_ptr + lane;

    BwdWin w;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 6; j++) { w.szz[j] = w.sxz[j] = w.sxx[j] = w.vz[j] = w.vx[j] = w.asz[j] = w.asx[j] = w.asxz[j] = w.avz[j] = w.avx[j] = zero; }
#pragma unroll
    for (int c = 0; c < 4; c++) { w.ga_prev[c] = 0.f; w.sh_prev[c] = 0.f; }
    auto rowoff = [&](int row) { return (size_t)min(max(row, 0), d.nzA - 1) * d.ldx; };
    const int r0 = k.zc0 - 2;
    // windows before the first iteration -- forward: szz rows r0-1 .. r0+1 (slots 1..3), sxz rows r0-2 .. r0 (slots 0..2);
    // adjoint stresses: rows r0-2 .. r0+1 (slots 0..3); the newest row of each arrives through the ring
#pragma unroll
    for (int j = 0; j < 3; j++) {
        w.szz[j + 1] = ldq(k.g + F_SZZ * fsz + rowoff(r0 - 1 + j));
        w.sxz[j] = ldq(k.g + F_SXZ * fsz + rowoff(r0 - 2 + j));
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const size_t ro = rowoff(r0 - 2 + j);
        w.asz[j] = ldq(k.adj + F_SZZ * fsz + ro); w.asx[j] = ldq(k.adj + F_SXX * fsz + ro); w.asxz[j] = ldq(k.adj + F_SXZ * fsz + ro);
    }
#pragma unroll
    for (int j = 0; j < BF_NST - 1; j++) stream_bwd_issue(k, r0 + j, j);
    const int niter = (k.zc1 - k.zc0) + 4;
    // one row per trip: the row body (four phases) is ~24 KB of code and must stay inside the 32 KB L1.5 instruction cache --
    // unrolled six times it ran at 2.0 no_instruction stalls per issued instruction and 40 % of the DRAM bandwidth
    constexpr int UNR = SW_UNR_BWD;
    static_assert(UNR == 1 || UNR == 2 || UNR == 6, "6-slot windows: 1, 2 or 6 rows per trip");
    int stg = 0;
#pragma unroll 1
    for (int kk = 0; kk < niter; kk += UNR) {
        const int r = r0 + kk;
        stream_bwd_row<0>(k, w, r, stg); stg = stg == BF_NST - 1 ? 0 : stg + 1;
        if (UNR > 1) { stream_bwd_row<1 % UNR>(k, w, r + 1, stg); stg = stg == BF_NST - 1 ? 0 : stg + 1; }
        if (UNR > 2) {
            stream_bwd_row<2 % UNR>(k, w, r + 2, stg); stg = stg == BF_NST - 1 ? 0 : stg + 1;
            stream_bwd_row<3 % UNR>(k, w, r + 3, stg); stg = stg == BF_NST - 1 ? 0 : stg + 1;
            stream_bwd_row<4 % UNR>(k, w, r + 4, stg); stg = stg == BF_NST - 1 ? 0 : stg + 1;
            stream_bwd_row<5 % UNR>(k, w, r + 5, stg); stg = stg == BF_NST - 1 ? 0 : stg + 1;
        }
        if (UNR < 6) {
            win_shift<UNR>(w.szz); win_shift<UNR>(w.sxz); win_shift<UNR>(w.sxx); win_shift<UNR>(w.vz); win_shift<UNR>(w.vx);
            win_shift<UNR>(w.asz); win_shift<UNR>(w.asx); win_shift<UNR>(w.asxz); win_shift<UNR>(w.avz); win_shift<UNR>(w.avx);
        }
    }
    cp_wait<0>();
}

// ------------------------------------------------------------------------------------------------
// grid: x = 1 + 2 ceil(nEdge / SW_WPB) + ceil((nWork - nEdge) / SW_WPB), y = slot ; dynamic shared memory BWF_SMEM
//   CTA 0                         stf gradient (source_grad, utilities.cu:719-730)
//   CTAs 1 .. 2 ne                edge items [0, nEdge): reconstruction CTA and adjoint CTA of the same four items next to each other
//   the rest                      interior items [nEdge, nWork): one pass, both sweeps
constexpr size_t BWF_EDGE_BYTES = RC_SMEM > (AR_SMEM + (size_t)SW_WPB * 256 * sizeof(float)) ? RC_SMEM : (AR_SMEM + (size_t)SW_WPB * 256 * sizeof(float));
constexpr size_t BWF_SMEM = BF_SMEM > BWF_EDGE_BYTES ? BF_SMEM : BWF_EDGE_BYTES;
__global__ void __launch_bounds__(SW_NT, SW_MINB) k_stream_bwd(const KArgs a, const StreamArgs sa)
{
    extern __shared__ __align__(16) float smem[];
    pdl_launch_dependents();
    const int s = blockIdx.y;
    const Dims &d = a.d;
    if (blockIdx.x == 0) {
        pdl_wait();
        if (threadIdx.x == 0) {
            const float *src = slot_state(a, s) + (size_t)(sa.pa ? S_ADJ1 : S_ADJ) * d.fsz;
            const size_t i = (size_t)a.t.zs[s] * d.ldx + a.t.xs[s];
            a.gstf[(size_t)s * d.nSteps + sa.it] = -(src[(size_t)F_SZZ * d.fsz + i] + a.t.rxz[s] * src[(size_t)F_SXX * d.fsz + i]) * d.dt;
        }
        return;
    }
    const int wi = (int)threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ne = (sa.nEdge + SW_WPB - 1) / SW_WPB;
    const int b = (int)blockIdx.x - 1;
    if (b < 2 * ne) {
        const int wg = (b >> 1) * SW_WPB + wi;
        if (wg >= sa.nEdge) return;
        const int4 wk = __ldg(sa.work + wg);
        if ((b & 1) == 0) {
            const unsigned sw = (unsigned)__cvta_generic_to_shared(smem) + wi * RC_WARP_BYTES;
            const float4 *sp = reinterpret_cast<const float4 *>(smem) + wi * (RC_WARP_BYTES / 16);
            pdl_wait();
            stream_rec_body<true>(a, sa, s, wk, lane, sw, sp);
        } else {
            const unsigned sw = (unsigned)__cvta_generic_to_shared(smem) + wi * AR_WARP_BYTES;
            const float4 *sp = reinterpret_cast<const float4 *>(smem) + wi * (AR_WARP_BYTES / 16);
            float *stage = smem + AR_SMEM / sizeof(float) + wi * 256;
            pdl_wait();
            stream_adj_body<true>(a, sa, s, wk, lane, stage, sw, sp);
        }
        return;
    }
    const int wg = sa.nEdge + (b - 2 * ne) * SW_WPB + wi;
    if (wg >= sa.nWork) return;
    const int4 wk = __ldg(sa.work + wg);
    const unsigned sw = (unsigned)__cvta_generic_to_shared(smem) + wi * BF_WARP_BYTES;
    const float4 *sp = reinterpret_cast<const float4 *>(smem) + wi * (BF_WARP_BYTES / 16);
    float *stage = smem + (wi * BF_WARP_BYTES + BF_RING_BYTES) / sizeof(float);
    pdl_wait();
    stream_bwd_body(a, sa, s, wk, lane, stage, sw, sp);
}

}  // namespace sepfwi

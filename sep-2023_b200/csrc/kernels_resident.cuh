// kernels_resident.cuh -- shared-memory-resident forward time loop for grids that fit the SMs (sm_100a).
//
// On the reference-size grids (0.05 - 0.7 M cells) one forward time step is 2 - 3 us of HBM/L2 traffic, so a
// launch per step is bound by launch and memory latency, not bandwidth.  This kernel runs the WHOLE time loop
// of a shot in one cooperative launch:
//   * the live grid is cut into tiles of OR x 56 cells, one CTA (512 threads) per tile, at most one CTA per SM;
//     the five fields of the tile plus a 4-cell halo stay in shared memory for all time steps, the CPML memory
//     variables of the tile's PML cells too; the five model coefficients of a thread's cells stay in registers;
//   * thread (g, c) owns column c of the 64-column extended tile and the RPT consecutive rows of row group g:
//     z stencils come from register windows, x stencils from conflict-free shared-memory reads;
//   * a time step is  stress on own + 2 ring cells (recomputed instead of exchanged)  ->  barrier  ->  velocity on
//     the own cells  ->  publish the own cells within 4 of the tile edge to the global ping-pong velocity arrays
//     (st.cg)  ->  barrier  ->  lanes of warp 0 fence and write the step number into the eight neighbours' inbox lines
//     ->  every warp polls its OWN tile's inbox line (relaxed loads: no L1 invalidation, no two tiles on one line) and
//     pulls its share of the 4-cell velocity halo from L2 (ld.cg)  ->  barrier.  One neighbour exchange per step, no
//     grid-wide barrier, no launch.  Measured alternatives (acquire loads, a shared counter line, tagged 64-bit words,
//     overlapping the exchange with the inner stress rows) are listed in profiles/README.md;
//   * traces, the explosive source and the boundary-ring save (reference layout, bit-exact) run inside the same
//     phases from shared memory: v-derived samples and the v ring while v is read-only (stress phase), the
//     pressure sample and the sigma ring while sigma is read-only (velocity phase).
// Arithmetic per cell is the sequence of the baseline kernels (kernels_base.cuh).
// Reference lines: el_stress.cu:50-87, el_velocity.cu:45-82, utilities.cu:362-392,524-552,593-703, libCUFD.cu:268-332.
#pragma once
#include "common.cuh"

namespace sepfwi {

constexpr int RS_EW = 64;        // extended tile width (own 56 + 4-cell halo on each side)
constexpr int RS_OW = 56;
#ifndef RS_NG_
#define RS_NG_ 8
#endif
constexpr int RS_NG = RS_NG_;    // row groups (threads = 64 RS_NG)
constexpr int RS_NT = RS_NG * RS_EW;
constexpr int RS_PW = 32;        // rows / columns of CPML memory a tile can hold (>= nPml)
constexpr int RS_SPIN_MAX = 1 << 22;
constexpr int RS_FLAGW = 32;     // ints per tile inbox (one 128-byte line)

struct ResArgs {
    int ntx, ntz, slot0;         // tiles; first slot of this launch (blockIdx.y = slot - slot0)
    int orows;                   // own rows of a tile, <= RS_NG RPT - 8
    int mask, fiber, save_ring;
    int dbg;                     // timing experiments only (SEPFWI_RES_DEBUG): 1 no exchange wait, 2 no stress phase, 4 no velocity phase, 8 no fence + flag, 16 no v traces
    int *flags;                  // [gridDim.y][ntx*ntz][RS_FLAGW] inboxes: slot (dz+1)*3+(dx+1) = steps completed by the neighbour in direction
                                 //   (dz, dx); one 128-byte line per tile (no two tiles poll the same line); missing neighbours pre-set to INT_MAX
    int *err;                    // set to 1 when a neighbour never arrives (the launch then drains instead of hanging)
    const int *tilePtr;          // [gridDim.y][ntx*ntz+1] receivers bucketed by tile (CSR) ...
    const int *tileRec;          // [gridDim.y][maxRec]    ... receiver indices
    const int *ringPtr;          // [ntx*ntz+1] boundary-ring entries bucketed by owning tile (same for every slot)
    const int2 *ringEnt;         // {ring index, offset of the cell inside the extended tile}
};

__host__ __device__ constexpr size_t rs_smem_bytes(int RPT)
{ return sizeof(float) * ((size_t)5 * RS_NG * RPT * RS_EW + 4 * RS_PW * RS_EW + (size_t)4 * RS_NG * RPT * RS_PW + 6 * RS_NG * RPT); }

__device__ __forceinline__ int rs_ld_acquire(const int *p)
{ int v; asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ int rs_ld_relaxed(const int *p)
{ int v; asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void rs_st_relaxed(int *p, int v)
{ asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// traces derived from the velocities (sample `samp` = the state now in shared memory)
template <int ER>
__device__ __forceinline__ void rs_record_v(const KArgs &a, const ResArgs &ra, const float *F, int s, int n0, int n1, const int *trec,
                                            int tz0, int tx0, int samp)
{
    constexpr int FS = ER * RS_EW;
    const Dims &d = a.d;
    const size_t cs = (size_t)d.maxRec * d.nSteps;
    const float *vz = F + F_VZ * FS, *vx = F + F_VX * FS;
    for (int k = n0 + (int)threadIdx.x; k < n1; k += RS_NT) {
        const int r = trec[k];
        const int z = a.t.zrec[(size_t)s * d.maxRec + r], x = a.t.xrec[(size_t)s * d.maxRec + r];
        const int i = (z - tz0 + 4) * RS_EW + (x - tx0 + 4);
        float *tr = a.trace + (size_t)s * d.nTrace * cs + (size_t)r * d.nSteps + samp;
        if (ra.mask & (1 << T_VX)) tr[T_VX * cs] = vx[i];
        if (ra.mask & (1 << T_VZ)) tr[T_VZ * cs] = vz[i];
        if (ra.mask & (1 << T_ETT)) {
            const float exx = vx[i] - vx[i - 1], ezz = vz[i] - vz[i - RS_EW];
            float e = ra.fiber == 0 ? exx : ezz;
            if (a.t.w) {
                const float *w = a.t.w + ((size_t)s * d.maxRec + r) * 3;
                e = w[0] * exx + w[1] * ezz + w[2] * (0.5f * ((vx[i + RS_EW] - vx[i]) + (vz[i + 1] - vz[i])));
            }
            tr[T_ETT * cs] = e;
        }
    }
}

// pressure trace (sample `samp` = the stresses now in shared memory)
template <int ER>
__device__ __forceinline__ void rs_record_p(const KArgs &a, const float *F, int s, int n0, int n1, const int *trec, int tz0, int tx0, int samp)
{
    constexpr int FS = ER * RS_EW;
    const Dims &d = a.d;
    for (int k = n0 + (int)threadIdx.x; k < n1; k += RS_NT) {
        const int r = trec[k];
        const int z = a.t.zrec[(size_t)s * d.maxRec + r], x = a.t.xrec[(size_t)s * d.maxRec + r];
        const int i = (z - tz0 + 4) * RS_EW + (x - tx0 + 4);
        a.trace[((size_t)s * d.nTrace + T_PR) * d.maxRec * d.nSteps + (size_t)r * d.nSteps + samp] = F[F_SZZ * FS + i] + F[F_SXX * FS + i];
    }
}

// boundary ring of time `samp`, fields [f0, f1), reference layout ring[slot][field][it][idx] (utilities.cu:362-392)
template <int ER>
__device__ __forceinline__ void rs_ring_save(const KArgs &a, const ResArgs &ra, const float *F, int s, int e0, int e1, int samp, int f0, int f1)
{
    constexpr int FS = ER * RS_EW;
    const Dims &d = a.d;
    float *rb = a.ring + (((size_t)s * NFIELD) * d.nSteps + samp) * d.ringLen;
    const size_t fs = (size_t)d.nSteps * d.ringLen;
    for (int k = e0 + (int)threadIdx.x; k < e1; k += RS_NT) {
        const int2 e = __ldg(ra.ringEnt + k);
        for (int f = f0; f < f1; f++) rb[f * fs + e.x] = F[f * FS + e.y];
    }
}

// grid: x = ntx*ntz tiles, y = slots of this launch; cooperative launch (every CTA resident); dynamic smem rs_smem_bytes(RPT)
template <int RPT>
__global__ void __launch_bounds__(RS_NT, 1) k_resident_fwd(const KArgs a, const ResArgs ra)
{
    constexpr int ER = RS_NG * RPT, EW = RS_EW, FS = ER * EW;
    constexpr int NHALO = 8 * EW + 8 * (ER - 8);      // 4 rows above + 4 below, 4 + 4 columns beside the rows between
    constexpr int HPT = (NHALO + RS_NT - 1) / RS_NT;
    extern __shared__ __align__(16) float sm[];
    float *F = sm;                                    // [5][ER][EW] fields, order of common.cuh
    float *ZP = F + 5 * FS;                           // [4][RS_PW][EW] z CPML memory: vz_z, vx_z (stress side), szz_z, sxz_z (velocity side)
    float *XP = ZP + 4 * RS_PW * EW;                  // [4][ER][RS_PW] x CPML memory: vx_x, vz_x (stress side), sxz_x, sxx_x (velocity side)
    float *CZ = XP + 4 * ER * RS_PW;                  // [6][ER] z CPML profiles of the tile's rows
    const Dims &d = a.d;
    const int tid = threadIdx.x, g = tid >> 6, c = tid & 63;
    const int nT = ra.ntx * ra.ntz, tile = blockIdx.x, sub = blockIdx.y, s = ra.slot0 + sub;
    const int tzi = tile / ra.ntx, txi = tile - tzi * ra.ntx;
    const int OR = ra.orows;                          // own rows of a tile (<= ER - 8)
    const int tz0 = tzi * OR, tx0 = txi * RS_OW;      // first own cell
    const int ld = d.ldx, nzA = d.nzA, nx = d.nx, nPml = d.nPml, nSteps = d.nSteps;
    const size_t fsz = d.fsz;
    const float c1z = d.c1z, c2z = d.c2z, c1x = d.c1x, c2x = d.c2x, dt = d.dt;
    const int x = tx0 - 4 + c;
    const int r0 = g * RPT;                           // first extended-tile row of this thread
    const int zb = tz0 - 4 + r0;                      // its global row
    float *st = slot_state(a, s);

    for (int i = tid; i < (int)(rs_smem_bytes(RPT) / sizeof(float)); i += RS_NT) sm[i] = 0.f;
    __syncthreads();
    for (int i = tid; i < 6 * ER; i += RS_NT) {
        const int k = i / ER, r = i - k * ER;
        CZ[i] = __ldg(a.cz + (size_t)k * nzA + min(max(tz0 - 4 + r, 0), nzA - 1));
    }
    // model coefficients of this thread's cells; row masks: stress rows (own + 2), velocity rows (own), rows inside the
    // z CPML strips, own rows a neighbour's halo reads
    float lam[RPT], l2u[RPT], mua[RPT], bya[RPT], byb[RPT];
    unsigned smask = 0, vmask = 0, zpm = 0, pubm = 0;
    const bool xin = x >= 0 && x < nx;
#pragma unroll
    for (int j = 0; j < RPT; j++) {
        const int z = zb + j, r = r0 + j;
        const bool in = xin && z >= 0 && z < nzA;
        const size_t i = in ? (size_t)z * ld + x : 0;
        const float l_ = in ? __ldg(a.model + M_LAM * fsz + i) : 0.f, m_ = in ? __ldg(a.model + M_MU * fsz + i) : 0.f;
        lam[j] = l_; l2u[j] = l_ + 2.0f * m_;
        mua[j] = in ? __ldg(a.model + M_MUAVE * fsz + i) : 0.f;
        bya[j] = in ? __ldg(a.model + M_BYCA * fsz + i) : 0.f;
        byb[j] = in ? __ldg(a.model + M_BYCB * fsz + i) : 0.f;
        const bool zact = z >= 2 && z <= nzA - 3;
        if (zact && r >= 2 && r < OR + 6) smask |= 1u << j;
        if (zact && r >= 4 && r < OR + 4) vmask |= 1u << j;
        if (zact && ((z < nPml) || (z > nzA - nPml - 1))) zpm |= 1u << j;
        if ((r >= 4 && r < 8) || (r >= OR && r < OR + 4)) pubm |= 1u << j;
    }
    // x CPML profile of this thread's column (el_stress.cu:63, el_velocity.cu:56,71)
    const bool colact = x >= 2 && x <= nx - 3;
    const bool xps = colact && ((x < nPml) || (x > nx - nPml - 1)), xpv = colact && ((x < nPml) || (x > nx - nPml));
    float bx = 0.f, ax = 0.f, rkx = 1.f, bxh = 0.f, axh = 0.f, rkxh = 1.f;
    if (xps) {
        const float *cx = a.cx + x;
        bx = __ldg(cx + C_B * nx); ax = __ldg(cx + C_A * nx); rkx = __ldg(cx + C_RK * nx);
        bxh = __ldg(cx + C_BH * nx); axh = __ldg(cx + C_AH * nx); rkxh = __ldg(cx + C_RKH * nx);
    }
    const int xp0 = (tx0 - 4 < nPml) ? max(tx0 - 4, 0) : nx - nPml;     // first column kept in XP (the host plan guarantees one side per tile)
    const int zp0 = (tz0 - 4 < nPml) ? max(tz0 - 4, 0) : nzA - nPml;    // first row kept in ZP
    const int xpi = min(max(x - xp0, 0), RS_PW - 1);
    const float *ampv = a.t.amp + (size_t)s * nSteps;
    const bool scol = c >= 2 && c < EW - 2 && colact, vcol = c >= 4 && c < EW - 4 && colact;
    const bool pubcol = c < 8 || c >= EW - 8;
    // explosive source: the thread whose stress rows hold the cell adds the amplitude after its update (every tile that
    // recomputes the cell does)
    int src_i = -1;
    {
        const int js = a.t.zs[s] - zb;
        if (scol && x == a.t.xs[s] && js >= 0 && js < RPT && ((smask >> js) & 1u)) src_i = (r0 + js) * EW + c;
    }
    float *Fc = F + r0 * EW + c;                       // this thread's first cell in field 0
    float *ZPc = ZP + (zb - zp0) * EW + c;             // ... in the z CPML block (only dereferenced on rows of zpm)
    float *XPc = XP + r0 * RS_PW + xpi;
    const float *CZr = CZ + r0;

    // halo cells this thread pulls in after every step
    int hsm[HPT], hgl[HPT], hfl[HPT];
#pragma unroll
    for (int k = 0; k < HPT; k++) {
        const int id = tid + k * RS_NT;
        int r = 0, cc = 0;
        bool ok = id < NHALO;
        if (id < 4 * EW) { r = id / EW; cc = id - r * EW; }
        else if (id < 8 * EW) { const int q = id - 4 * EW; r = q / EW; cc = q - r * EW; r += OR + 4; }
        else { const int q = id - 8 * EW; r = 4 + q / 8; const int e = q & 7; cc = e < 4 ? e : EW - 8 + e; ok = ok && r < OR + 4; }
        const int z = tz0 - 4 + r, xx = tx0 - 4 + cc;
        ok = ok && z >= 0 && z < nzA && xx >= 0 && xx < nx;
        hsm[k] = ok ? r * EW + cc : 0; hgl[k] = ok ? z * ld + xx : 0;
        hfl[k] = ok ? 1 : -1;
    }
    int *inbox = ra.flags + (size_t)(sub * nT + tile) * RS_FLAGW;
    // lane t < 9 (t != 4) of warp 0 tells the neighbour in direction (t/3-1, t%3-1); that neighbour sees it as coming from the opposite direction
    int *outbox = nullptr;
    if (tid < 9 && tid != 4) {
        const int nz_ = tzi + tid / 3 - 1, nx_ = txi + tid % 3 - 1;
        if (nz_ >= 0 && nz_ < ra.ntz && nx_ >= 0 && nx_ < ra.ntx) outbox = ra.flags + (size_t)(sub * nT + nz_ * ra.ntx + nx_) * RS_FLAGW + (8 - tid);
    }
    const int lane = tid & 31;
    const int n0 = ra.tilePtr[sub * (nT + 1) + tile], n1 = ra.tilePtr[sub * (nT + 1) + tile + 1];
    const int *trec = ra.tileRec + (size_t)sub * d.maxRec;
    const int e0 = ra.save_ring ? ra.ringPtr[tile] : 0, e1 = ra.save_ring ? ra.ringPtr[tile + 1] : 0;
    __syncthreads();

    int dead = 0;
    for (int it = 0; it <= nSteps - 2; it++) {
        // ================= stress phase: v(it) is read-only =================
        if (ra.mask && it >= 1 && n1 > n0 && !(ra.dbg & 16)) rs_record_v<ER>(a, ra, F, s, n0, n1, trec, tz0, tx0, it);
        if (e1 > e0) rs_ring_save<ER>(a, ra, F, s, e0, e1, it, F_VZ, F_VX + 1);
        if (scol && !(ra.dbg & 2)) {
            const float *VZ = Fc + F_VZ * FS, *VX = Fc + F_VX * FS;
            float wz[RPT + 4], wx[RPT + 5];            // rows r0 - 2 .. r0 + RPT + 1 (clamped: the clamped rows only feed masked cells)
            auto woff = [&](int k) {
                int o = (k - 2) * EW;
                if (k < 2) o = (max(r0 + k - 2, 0) - r0) * EW;
                if (k >= RPT + 2) o = (min(r0 + k - 2, ER - 1) - r0) * EW;
                return o;
            };
            // rolling windows: row j uses wz[j .. j+3], wx[j+1 .. j+4]; the entries of row j+1 are requested one row ahead
#pragma unroll
            for (int k = 0; k < 4; k++) { wz[k] = VZ[woff(k)]; wx[k + 1] = VX[woff(k + 1)]; }
#pragma unroll
            for (int j = 0; j < RPT; j++) {
                if (j + 1 < RPT) { wz[j + 4] = VZ[woff(j + 4)]; wx[j + 5] = VX[woff(j + 5)]; }
                if (!((smask >> j) & 1u)) continue;
                const float *vxr = VX + j * EW, *vzr = VZ + j * EW;
                float dvz_dz = c1z * (wz[j + 2] - wz[j + 1]) - c2z * (wz[j + 3] - wz[j]);
                float dvx_dz = c1z * (wx[j + 3] - wx[j + 2]) - c2z * (wx[j + 4] - wx[j + 1]);
                float dvx_dx = c1x * (wx[j + 2] - vxr[-1]) - c2x * (vxr[1] - vxr[-2]);
                float dvz_dx = c1x * (vzr[1] - wz[j + 2]) - c2x * (vzr[2] - vzr[-1]);
                if ((zpm >> j) & 1u) {                          // el_stress.cu:58-62
                    float *p0 = ZPc + j * EW, *p1 = p0 + RS_PW * EW;
                    const float m0 = CZr[C_B * ER + j] * *p0 + CZr[C_A * ER + j] * dvz_dz;
                    const float m1 = CZr[C_BH * ER + j] * *p1 + CZr[C_AH * ER + j] * dvx_dz;
                    *p0 = m0; *p1 = m1;
                    dvz_dz = dvz_dz * CZr[C_RK * ER + j] + m0;
                    dvx_dz = dvx_dz * CZr[C_RKH * ER + j] + m1;
                }
                if (xps) {                                       // el_stress.cu:63-67
                    float *p0 = XPc + j * RS_PW, *p1 = p0 + ER * RS_PW;
                    const float m0 = bx * *p0 + ax * dvx_dx;
                    const float m1 = bxh * *p1 + axh * dvz_dx;
                    *p0 = m0; *p1 = m1;
                    dvx_dx = dvx_dx * rkx + m0;
                    dvz_dx = dvz_dx * rkxh + m1;
                }
                float *pzz = Fc + F_SZZ * FS + j * EW, *pxz = Fc + F_SXZ * FS + j * EW, *pxx = Fc + F_SXX * FS + j * EW;
                *pzz = *pzz + (l2u[j] * dvz_dz + lam[j] * dvx_dx) * dt;
                *pxx = *pxx + (lam[j] * dvz_dz + l2u[j] * dvx_dx) * dt;
                *pxz = *pxz + mua[j] * (dvx_dz + dvz_dx) * dt;
            }
            if (src_i >= 0) {                                    // add_source, utilities.cu:524-552
                const float amp = ampv[it];
                F[F_SZZ * FS + src_i] += amp; F[F_SXX * FS + src_i] += amp;
            }
        }
        __syncthreads();
        // ================= velocity phase: sigma(it+1) is read-only =================
        if ((ra.mask & (1 << T_PR)) && n1 > n0) rs_record_p<ER>(a, F, s, n0, n1, trec, tz0, tx0, it + 1);
        if (e1 > e0 && it + 1 <= nSteps - 2) rs_ring_save<ER>(a, ra, F, s, e0, e1, it + 1, F_SZZ, F_SXX + 1);
        if (vcol && !(ra.dbg & 4)) {
            float *vout = st + (size_t)(((it + 1) & 1) ? S_FWD1 : S_FWD) * fsz + ((ptrdiff_t)zb * ld + x);   // only dereferenced for in-grid rows
            const float *ZZ = Fc + F_SZZ * FS, *XZ = Fc + F_SXZ * FS, *XX = Fc + F_SXX * FS;
            float wzz[RPT + 4], wxz[RPT + 4];          // szz rows r0 - 1 .. r0 + RPT + 1 ; sxz rows r0 - 2 .. r0 + RPT
            auto zoff = [&](int k) { return ((k < 2 ? max(r0 + k - 1, 0) : k >= RPT + 1 ? min(r0 + k - 1, ER - 1) : r0 + k - 1) - r0) * EW; };
            auto xoff = [&](int k) { return ((k < 2 ? max(r0 + k - 2, 0) : k >= RPT + 1 ? min(r0 + k - 2, ER - 1) : r0 + k - 2) - r0) * EW; };
#pragma unroll
            for (int k = 0; k < 4; k++) { wzz[k] = ZZ[zoff(k)]; wxz[k] = XZ[xoff(k)]; }
#pragma unroll
            for (int j = 0; j < RPT; j++) {
                if (j + 1 < RPT) { wzz[j + 4] = ZZ[zoff(j + 4)]; wxz[j + 4] = XZ[xoff(j + 4)]; }
                if (!((vmask >> j) & 1u)) continue;
                const float *xzr = XZ + j * EW, *xxr = XX + j * EW;
                float dszz_dz = c1z * (wzz[j + 2] - wzz[j + 1]) - c2z * (wzz[j + 3] - wzz[j]);
                float dsxz_dz = c1z * (wxz[j + 2] - wxz[j + 1]) - c2z * (wxz[j + 3] - wxz[j]);
                float dsxz_dx = c1x * (wxz[j + 2] - xzr[-1]) - c2x * (xzr[1] - xzr[-2]);
                float dsxx_dx = c1x * (xxr[1] - xxr[0]) - c2x * (xxr[2] - xxr[-1]);
                if ((zpm >> j) & 1u) {                          // el_velocity.cu:51-55
                    float *p0 = ZPc + (2 * RS_PW + j) * EW, *p1 = p0 + RS_PW * EW;
                    const float m0 = CZr[C_BH * ER + j] * *p0 + CZr[C_AH * ER + j] * dszz_dz;
                    const float m1 = CZr[C_B * ER + j] * *p1 + CZr[C_A * ER + j] * dsxz_dz;
                    *p0 = m0; *p1 = m1;
                    dszz_dz = dszz_dz * CZr[C_RKH * ER + j] + m0;
                    dsxz_dz = dsxz_dz * CZr[C_RK * ER + j] + m1;
                }
                if (xpv) {                                       // el_velocity.cu:56-60
                    float *p0 = XPc + (2 * ER + j) * RS_PW, *p1 = p0 + ER * RS_PW;
                    const float m0 = bx * *p0 + ax * dsxz_dx;
                    const float m1 = bxh * *p1 + axh * dsxx_dx;
                    *p0 = m0; *p1 = m1;
                    dsxz_dx = dsxz_dx * rkx + m0;
                    dsxx_dx = dsxx_dx * rkxh + m1;
                }
                float *pvz = Fc + F_VZ * FS + j * EW, *pvx = Fc + F_VX * FS + j * EW;
                const float nvz = *pvz + (dszz_dz + dsxz_dx) * bya[j] * dt;
                const float nvx = *pvx + (dsxz_dz + dsxx_dx) * byb[j] * dt;
                *pvz = nvz; *pvx = nvx;
                if (pubcol || ((pubm >> j) & 1u)) {              // a neighbour's halo reads this cell
                    __stcg(vout + F_VZ * fsz + (size_t)j * ld, nvz); __stcg(vout + F_VX * fsz + (size_t)j * ld, nvx);
                }
            }
        }
        __syncthreads();
        // ================= neighbour exchange of v(it+1) =================
        // producer: stores above -> barrier -> fence + release by one thread.  consumer: relaxed polls of the step counter
        // (no L1 invalidation), then L2 loads (ld.cg) issued after the poll's branch.
        if (tid < 32 && !(ra.dbg & 8)) {
            asm volatile("fence.acq_rel.gpu;" ::: "memory");      // release pattern: fence + relaxed store (MEMBAR.ALL.GPU; __threadfence() would be the sequentially consistent MEMBAR.SC.GPU)
            if (outbox) rs_st_relaxed(outbox, it + 1);
        }
        if (!(ra.dbg & 1)) {
            // every warp waits on the tile's own inbox line: lanes 0..8 watch one neighbour each
            int spins = 0;
            for (;;) {
                const int v = (lane < 9 && lane != 4) ? rs_ld_relaxed(inbox + lane) : 0x7fffffff;
                if (__all_sync(0xffffffffu, v >= it + 1)) break;
                if (++spins > RS_SPIN_MAX || ((spins & 255) == 0 && *(volatile int *)ra.err)) { *ra.err = 1; dead = 1; break; }
            }
            // acquire side of the exchange: the polls above are relaxed loads, and a control dependency does not order the halo
            // loads after them in the PTX memory model.  Once every flag has arrived, the polling lanes re-read theirs with ONE
            // ld.acquire (pairs with the producer's fence + relaxed store) and bar.warp.sync extends the order to the lanes
            // that did not poll.  -DRS_RELAXED_EXCHANGE drops it (what round 1 shipped: works on sm_100a, not guaranteed).
#ifndef RS_RELAXED_EXCHANGE
            if (lane < 9 && lane != 4) (void)rs_ld_acquire(inbox + lane);
            __syncwarp();
#endif
            const float *vin = st + (size_t)(((it + 1) & 1) ? S_FWD1 : S_FWD) * fsz;
#pragma unroll
            for (int k = 0; k < HPT; k++) {
                if (hfl[k] < 0) continue;
                F[F_VZ * FS + hsm[k]] = __ldcg(vin + F_VZ * fsz + hgl[k]);
                F[F_VX * FS + hsm[k]] = __ldcg(vin + F_VX * fsz + hgl[k]);
            }
        }
        if (__syncthreads_or(dead)) break;
    }
    // final sample of the v-derived traces and the final state (the reverse-time loop starts from it)
    if (ra.mask && n1 > n0) rs_record_v<ER>(a, ra, F, s, n0, n1, trec, tz0, tx0, nSteps - 1);
    {
        float *fo = st + (size_t)(((nSteps - 1) & 1) ? S_FWD1 : S_FWD) * fsz;
        if (c >= 4 && c < EW - 4 && xin) {
#pragma unroll
            for (int j = 0; j < RPT; j++) {
                const int r = r0 + j, z = zb + j;
                if (r < 4 || r >= OR + 4 || z < 0 || z >= nzA) continue;
#pragma unroll
                for (int f = 0; f < NFIELD; f++) fo[f * fsz + (size_t)z * ld + x] = F[f * FS + r * EW + c];
            }
        }
    }
}

}  // namespace sepfwi

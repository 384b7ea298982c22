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
#include <cuda.h>      // CUtensorMap (the type only; the encoder is fetched through cudaGetDriverEntryPoint, nothing links libcuda)
#include "common.cuh"
#include "kernels_base.cuh"

namespace sepfwi {

constexpr int SW_OWN = 120;          // columns owned by one warp (30 quads)
#ifndef SW_WPB_
#define SW_WPB_ 4
#endif
constexpr int SW_WPB = SW_WPB_;      // warps per CTA
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
    int tma;                         // interior warps fetch their operand rows with TMA tensor copies (the maps are valid)
};

// ------------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor) operand path of the interior warps.  The state block [slot][array][z][x] and the model block
// [array][z][x] are described by two tensor maps; one elected lane requests a whole operand row (128 floats = 512 B) per
// array with ONE instruction and four integer coordinates -- instead of 32 lanes x (16-byte cp.async + 64-bit address
// arithmetic + row clamping) -- and the bytes land in the same per-warp ring, signalled through one mbarrier per stage.
// Out-of-range rows / columns are zero-filled by the TMA unit (they only feed unowned or inactive cells).  The copies go
// L2 -> shared memory without touching L1.
struct __align__(64) TmaMaps { CUtensorMap state, model, grad; };
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, int count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_row4(unsigned dst, const CUtensorMap *map, int x, int z, int arr, int slot, unsigned bar)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(dst), "l"(map), "r"(x), "r"(z), "r"(arr), "r"(slot), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_row3(unsigned dst, const CUtensorMap *map, int x, int z, int arr, unsigned bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(map), "r"(x), "r"(z), "r"(arr), "r"(bar) : "memory");
}

#define Q4(v) {v.x, v.y, v.z, v.w}
// Register windows are LINEAR: slot j of a 6-slot window holds row (first row of the trip) - 2 + j for the fields that are read
// ahead (rows r-2 .. r+2 of the iteration at row r sit in slots u .. u+4) and row (first row of the trip) - 4 + j for the fields
// the iteration produces (rows r-4 .. r in slots u .. u+4), u = the row's position inside the loop trip.  A trip of UNR rows ends
// by moving every window down UNR slots.  UNR = 6 needs no moves (the indices wrap), but its loop body (6 x ~0.6-1.5 k
// instructions) overflows the 32 KB L1.5 instruction cache: ncu showed 0.6 - 2.0 no_instruction stalls per issued instruction.
template <int S> __device__ __forceinline__ void win_shift(float4 (&a)[6])
{
#pragma unroll
    for (int j = 0; j + S < 6; j++) a[j] = a[j + S];
}
// (6-slot windows: 1, 2 or 6 rows per trip)
#ifndef SW_UNR_FWD
#define SW_UNR_FWD 2
#endif
#ifndef SW_UNR_ADJ
#define SW_UNR_ADJ 2
#endif
#ifndef SW_UNR_REC
#define SW_UNR_REC 2
#endif
// Programmatic dependent launch: every kernel of a time loop lets its successor start launching right away and
// waits for its predecessor's results only after its own set-up -- removes the ~2-4 us launch gap per time step.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// streaming 128-bit load of data that is read-only for the whole launch (L1 allocation kept: the 8-column overlap
// of adjacent strips is served by L1 when the neighbouring warp sits in the same CTA)
__device__ __forceinline__ float4 ldq(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
// register-free look-ahead: pull the 128-byte line of the next row into L1 (edge warps have no registers to spare)
__device__ __forceinline__ void pf_l1(const float *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ float4 ldq_c(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }   // cached (small reused tables)
__device__ __forceinline__ float4 ldq_rw(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ void stq(float *p, const float4 &v) { *reinterpret_cast<float4 *>(p) = v; }
__device__ __forceinline__ float4 mk4(const float v[4]) { return make_float4(v[0], v[1], v[2], v[3]); }
// store only the columns of a halo lane whose stencil window is complete (lane 0: comps 2,3 ; lane 31: comps 0,1):
// the other half of the quad belongs to the neighbouring warp, which may be writing different (newer) values
__device__ __forceinline__ void stq_halo(float *p, const float v[4], int lane)
{
    if (lane == 0) *reinterpret_cast<float2 *>(p + 2) = make_float2(v[2], v[3]);
    else if (lane == 31) *reinterpret_cast<float2 *>(p) = make_float2(v[0], v[1]);
    else *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
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

#ifndef RC_CP_OP
#define RC_CP_OP "cp.async.ca.shared.global"
#endif
__device__ __forceinline__ void cp16(unsigned saddr, const float *g)
{ asm volatile(RC_CP_OP " [%0], [%1], 16;" ::"r"(saddr), "l"(g)); }
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ------------------------------------------------------------------------------------------------
// forward step of one (strip, chunk)
//
// Operand ring: the ten (interior) or eighteen (edge: + eight CPML memory variables) row quads an iteration needs are
// requested FR_NST-1 rows ahead with cp.async into a per-warp shared-memory ring; each lane copies and later reads only
// its own 16 bytes, so cp.async.wait_group is the only synchronisation and the registers hold just the stencil windows.
constexpr int FR_NARR_I = 10, FR_NST_I = 3;                 // interior: 3 stages x 10 arrays x 512 B = 15 KB per warp
constexpr int FR_NARR_E = 18, FR_NST_E = 2;                 // edge:     2 stages x 18 arrays x 512 B = 18 KB per warp
#ifdef SW_INTERIOR_ONLY
constexpr int FR_WARP_BYTES = FR_NARR_I * FR_NST_I * 512;
#else
constexpr int FR_WARP_BYTES = FR_NARR_E * FR_NST_E * 512;
#endif
constexpr size_t FR_SMEM = (size_t)SW_WPB * FR_WARP_BYTES;  // 72 KB per CTA, 2 CTAs per SM
enum { FA_VZ = 0, FA_VX, FA_OZZ, FA_OXZ, FA_OXX, FA_LAM, FA_MU, FA_MUA, FA_BYA, FA_BYB,       // rows r+2 (v), r, r-2 (buoyancies)
       FA_PVZZ, FA_PVXZ, FA_PVXX, FA_PVZX, FA_PSZZ, FA_PSXZZ, FA_PSXZX, FA_PSXX };              // CPML memory: rows r (stress side), r-2 (velocity side)

struct FwdCtx {          // per-warp constants of the march
    const float *g, *m, *pv_src, *cxs, *cxv, *cz, *amp;
    float *o, *pv_dst, *ps;
    const float4 *ring_p;    // this lane's 16 bytes of stage 0, array 0
    unsigned ring_s;         // the same as a shared-memory address
    size_t fsz;
    int ld, nzA, nPml, zc0, zc1, zs, xs, xq0;
    unsigned amask;
    bool lown, xps, xpv;
    float c1z, c2z, c1x, c2x, dt;
    // TMA path (interior warps): tensor maps, the warp's ring base and mbarriers, first column / input buffer / slot coordinates
    const TmaMaps *tm;
    unsigned ring_w, bar_s;
    int x0, cin, slot, lane;
    bool tma;
};
struct FwdWin {          // register windows: old velocities rows r-2 .. r+2, new stresses rows r-4 .. r
    float4 vz[6], vx[6], zz[6], xz[6], xx[6];
};

template <bool EDGE>
__device__ __forceinline__ bool fwd_zp(const FwdCtx &k, int row)      // row inside the z CPML strips and active
{ return EDGE && row >= 2 && row <= k.nzA - 3 && ((row < k.nPml) || (row > k.nzA - k.nPml - 1)); }

// request the operands of the iteration whose stress row is r
template <bool EDGE>
__device__ __forceinline__ void stream_fwd_issue(const FwdCtx &k, const int r, const int stage)
{
    constexpr int NARR = EDGE ? FR_NARR_E : FR_NARR_I;
    const int ld = k.ld, nzA = k.nzA;
    const size_t fsz = k.fsz;
    if (!EDGE && k.tma) {
        // every lane has read the stage that is re-armed here (row r - NST) before lane 0 asks the TMA unit to overwrite it
        __syncwarp();
        if (k.lane == 0) {
            const unsigned bar = k.bar_s + 8u * stage, dst = k.ring_w + (unsigned)stage * (NARR * 512);
            mbar_expect_tx(bar, NARR * 512);
            tma_row4(dst + FA_VZ * 512, &k.tm->state, k.x0, r + 2, k.cin + F_VZ, k.slot, bar);
            tma_row4(dst + FA_VX * 512, &k.tm->state, k.x0, r + 2, k.cin + F_VX, k.slot, bar);
            tma_row4(dst + FA_OZZ * 512, &k.tm->state, k.x0, r, k.cin + F_SZZ, k.slot, bar);
            tma_row4(dst + FA_OXZ * 512, &k.tm->state, k.x0, r, k.cin + F_SXZ, k.slot, bar);
            tma_row4(dst + FA_OXX * 512, &k.tm->state, k.x0, r, k.cin + F_SXX, k.slot, bar);
            tma_row3(dst + FA_LAM * 512, &k.tm->model, k.x0, r, M_LAM, bar);
            tma_row3(dst + FA_MU * 512, &k.tm->model, k.x0, r, M_MU, bar);
            tma_row3(dst + FA_MUA * 512, &k.tm->model, k.x0, r, M_MUAVE, bar);
            tma_row3(dst + FA_BYA * 512, &k.tm->model, k.x0, r - 2, M_BYCA, bar);
            tma_row3(dst + FA_BYB * 512, &k.tm->model, k.x0, r - 2, M_BYCB, bar);
        }
        return;
    }
    auto rowoff = [&](int row) { return (size_t)min(max(row, 0), nzA - 1) * ld; };   // clamped rows feed inactive / unowned cells only
    const size_t r2 = rowoff(r + 2), r0 = rowoff(r), rq = rowoff(r - 2);
    const unsigned sb = k.ring_s + (unsigned)stage * (NARR * 512);
    cp16(sb + FA_VZ * 512, k.g + F_VZ * fsz + r2); cp16(sb + FA_VX * 512, k.g + F_VX * fsz + r2);
    cp16(sb + FA_OZZ * 512, k.g + F_SZZ * fsz + r0); cp16(sb + FA_OXZ * 512, k.g + F_SXZ * fsz + r0); cp16(sb + FA_OXX * 512, k.g + F_SXX * fsz + r0);
    cp16(sb + FA_LAM * 512, k.m + M_LAM * fsz + r0); cp16(sb + FA_MU * 512, k.m + M_MU * fsz + r0); cp16(sb + FA_MUA * 512, k.m + M_MUAVE * fsz + r0);
    cp16(sb + FA_BYA * 512, k.m + M_BYCA * fsz + rq); cp16(sb + FA_BYB * 512, k.m + M_BYCB * fsz + rq);
    if (EDGE) {
        const int q = r - 2;
        const bool ract = (r >= 2 && r <= nzA - 3), qact = (q >= 2 && q <= nzA - 3) && (q >= k.zc0) && (q < k.zc1);
        if (fwd_zp<EDGE>(k, r)) { cp16(sb + FA_PVZZ * 512, k.pv_src + (size_t)P_VZ_Z * fsz + r0); cp16(sb + FA_PVXZ * 512, k.pv_src + (size_t)P_VX_Z * fsz + r0); }
        if (ract && k.xps) { cp16(sb + FA_PVXX * 512, k.pv_src + (size_t)P_VX_X * fsz + r0); cp16(sb + FA_PVZX * 512, k.pv_src + (size_t)P_VZ_X * fsz + r0); }
        if (qact && fwd_zp<EDGE>(k, q)) { cp16(sb + FA_PSZZ * 512, k.ps + (size_t)P_SZZ_Z * fsz + rq); cp16(sb + FA_PSXZZ * 512, k.ps + (size_t)P_SXZ_Z * fsz + rq); }
        if (qact && k.xpv && k.lown) { cp16(sb + FA_PSXZX * 512, k.ps + (size_t)P_SXZ_X * fsz + rq); cp16(sb + FA_PSXX * 512, k.ps + (size_t)P_SXX_X * fsz + rq); }
    }
    cp_commit();
}

// one row: request row r + (NST-1), then stress at row r and velocity at row r-2.  U = r's phase in the 6-slot rotation.
// PB: second phase of the row -- 0 always, 1 never (the chunk's lead-in rows: nothing of it would be kept), 2 warp-uniform test
template <bool EDGE, int U, int PB>
__device__ __forceinline__ void stream_fwd_row(const FwdCtx &k, FwdWin &w, const int r, const int stage, const unsigned phase = 0)
{
    constexpr int NARR = EDGE ? FR_NARR_E : FR_NARR_I;
    constexpr int NST = EDGE ? FR_NST_E : FR_NST_I;
    const int ld = k.ld, nzA = k.nzA;
    const size_t fsz = k.fsz;
    const float c1z = k.c1z, c2z = k.c2z, c1x = k.c1x, c2x = k.c2x, dt = k.dt;
    constexpr int u = U;
    stream_fwd_issue<EDGE>(k, r + (NST - 1), stage == 0 ? NST - 1 : stage - 1);
    if (!EDGE && k.tma) mbar_wait(k.bar_s + 8u * stage, phase);     // the TMA unit has delivered this stage's ten rows
    else cp_wait<NST - 1>();
    const float4 *sb = k.ring_p + stage * (NARR * 32);
    w.vz[(u + 4) % 6] = sb[FA_VZ * 32]; w.vx[(u + 4) % 6] = sb[FA_VX * 32];       // row r+2
    // ---- stress at row r from v rows r-2 .. r+2 (slots u .. u+4)
    {
        const float4 a0 = w.vz[u % 6], a1 = w.vz[(u + 1) % 6], a2 = w.vz[(u + 2) % 6], a3 = w.vz[(u + 3) % 6];
        const float4 b0 = w.vx[(u + 1) % 6], b1 = w.vx[(u + 2) % 6], b2 = w.vx[(u + 3) % 6], b3 = w.vx[(u + 4) % 6];
        const float vxc[7] = XWIN_B(b1), vzc[7] = XWIN_F(a2);
        const float vzm2[4] = Q4(a0), vzm1[4] = Q4(a1), vzq[4] = Q4(a2), vzp1[4] = Q4(a3);
        const float vxm1[4] = Q4(b0), vxq[4] = Q4(b1), vxp1[4] = Q4(b2), vxp2[4] = Q4(b3);
        const float4 lam4 = sb[FA_LAM * 32], mu4 = sb[FA_MU * 32], mua4 = sb[FA_MUA * 32];
        const float4 ozz4 = sb[FA_OZZ * 32], oxz4 = sb[FA_OXZ * 32], oxx4 = sb[FA_OXX * 32];
        const float l[4] = Q4(lam4), mm[4] = Q4(mu4), ma[4] = Q4(mua4);
        const float pzz[4] = Q4(ozz4), pxz[4] = Q4(oxz4), pxx[4] = Q4(oxx4);
        float nzz[4], nxz[4], nxx[4];
        const bool rowact = !EDGE || (r >= 2 && r <= nzA - 3);
        const bool zp = fwd_zp<EDGE>(k, r);
        const bool xp = EDGE && rowact && k.xps;
        const bool rown = (r >= k.zc0) && (r < k.zc1);
        const size_t ro = (size_t)r * ld;       // only dereferenced for owned / active rows
        float pzq[4] = {0.f, 0.f, 0.f, 0.f}, pzh[4] = {0.f, 0.f, 0.f, 0.f}, pxq[4] = {0.f, 0.f, 0.f, 0.f}, pxh[4] = {0.f, 0.f, 0.f, 0.f};
        float bz = 0.f, az = 0.f, rkz = 1.f, bzh = 0.f, azh = 0.f, rkzh = 1.f;
        float bx[4], ax[4], rkx[4], bxh[4], axh[4], rkxh[4];
        if (EDGE) {
            if (zp) {
                const float4 t0 = sb[FA_PVZZ * 32], t1 = sb[FA_PVXZ * 32];
                pzq[0] = t0.x; pzq[1] = t0.y; pzq[2] = t0.z; pzq[3] = t0.w;
                pzh[0] = t1.x; pzh[1] = t1.y; pzh[2] = t1.z; pzh[3] = t1.w;
                const float *cz = k.cz + r;
                bz = cz[C_B * nzA]; az = cz[C_A * nzA]; rkz = cz[C_RK * nzA];
                bzh = cz[C_BH * nzA]; azh = cz[C_AH * nzA]; rkzh = cz[C_RKH * nzA];
            }
            if (xp) {
                const float4 t0 = sb[FA_PVXX * 32], t1 = sb[FA_PVZX * 32];
                pxq[0] = t0.x; pxq[1] = t0.y; pxq[2] = t0.z; pxq[3] = t0.w;
                pxh[0] = t1.x; pxh[1] = t1.y; pxh[2] = t1.z; pxh[3] = t1.w;
                const float *cx = k.cxs;
                const float4 q0 = ldq_c(cx + C_B * ld), q1 = ldq_c(cx + C_A * ld), q2 = ldq_c(cx + C_RK * ld);
                const float4 q3 = ldq_c(cx + C_BH * ld), q4 = ldq_c(cx + C_AH * ld), q5 = ldq_c(cx + C_RKH * ld);
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
        w.zz[(u + 4) % 6] = rzz; w.xz[(u + 4) % 6] = rxz; w.xx[(u + 4) % 6] = rxx;          // new stress row r-j lives in slot u+4-j
        if (k.lown && rown) {
            stq(k.o + F_SZZ * fsz + ro, rzz); stq(k.o + F_SXZ * fsz + ro, rxz); stq(k.o + F_SXX * fsz + ro, rxx);
            if (EDGE) {
                if (zp) { stq(k.pv_dst + (size_t)P_VZ_Z * fsz + ro, mk4(pzq)); stq(k.pv_dst + (size_t)P_VX_Z * fsz + ro, mk4(pzh)); }
                if (xp) { stq(k.pv_dst + (size_t)P_VX_X * fsz + ro, mk4(pxq)); stq(k.pv_dst + (size_t)P_VZ_X * fsz + ro, mk4(pxh)); }
            }
        }
    }
    // ---- velocity at row q = r-2 : szz rows q-1..q+2, sxz rows q-2..q+1, sxx row q (stress row r-j lives in slot u+4-j)
    if (PB == 0 || (PB == 2 && (r - 2 >= k.zc0) && (r - 2 < k.zc1))) {      // warp-uniform: on the chunk's four lead-in rows (and the surplus row of the last trip) nothing of this phase is kept
        const int q = r - 2;
        const float4 p0 = w.zz[(u + 1) % 6], p1 = w.zz[(u + 2) % 6], p2 = w.zz[(u + 3) % 6], p3 = w.zz[(u + 4) % 6];
        const float4 q0 = w.xz[u % 6], q1 = w.xz[(u + 1) % 6], q2 = w.xz[(u + 2) % 6], q3 = w.xz[(u + 3) % 6];
        const float4 xc = w.xx[(u + 2) % 6];
        const float xzc[7] = XWIN_B(q2), xxc[7] = XWIN_F(xc);
        const float zzm1[4] = Q4(p0), zzc[4] = Q4(p1), zzp1[4] = Q4(p2), zzp2[4] = Q4(p3);
        const float xzm2[4] = Q4(q0), xzm1[4] = Q4(q1), xzp1[4] = Q4(q3);
        const float ovz[4] = Q4(w.vz[u % 6]), ovx[4] = Q4(w.vx[u % 6]);      // v row r-2 lives in slot u
        const float4 bya4 = sb[FA_BYA * 32], byb4 = sb[FA_BYB * 32];
        const float ba[4] = Q4(bya4), bb[4] = Q4(byb4);
        float nvz[4], nvx[4];
        const bool qown = (q >= k.zc0) && (q < k.zc1);
        const bool rowact = !EDGE || (q >= 2 && q <= nzA - 3);
        const bool zp = qown && fwd_zp<EDGE>(k, q);
        const bool xp = EDGE && rowact && qown && k.xpv && k.lown;
        const size_t ro = (size_t)q * ld;
        float pzq[4] = {0.f, 0.f, 0.f, 0.f}, pzh[4] = {0.f, 0.f, 0.f, 0.f}, pxq[4] = {0.f, 0.f, 0.f, 0.f}, pxh[4] = {0.f, 0.f, 0.f, 0.f};
        float bz = 0.f, az = 0.f, rkz = 1.f, bzh = 0.f, azh = 0.f, rkzh = 1.f;
        float bx[4], ax[4], rkx[4], bxh[4], axh[4], rkxh[4];
        if (EDGE) {
            if (zp) {
                const float4 t0 = sb[FA_PSZZ * 32], t1 = sb[FA_PSXZZ * 32];
                pzh[0] = t0.x; pzh[1] = t0.y; pzh[2] = t0.z; pzh[3] = t0.w;
                pzq[0] = t1.x; pzq[1] = t1.y; pzq[2] = t1.z; pzq[3] = t1.w;
                const float *cz = k.cz + q;
                bz = cz[C_B * nzA]; az = cz[C_A * nzA]; rkz = cz[C_RK * nzA];
                bzh = cz[C_BH * nzA]; azh = cz[C_AH * nzA]; rkzh = cz[C_RKH * nzA];
            }
            if (xp) {
                const float4 t0 = sb[FA_PSXZX * 32], t1 = sb[FA_PSXX * 32];
                pxq[0] = t0.x; pxq[1] = t0.y; pxq[2] = t0.z; pxq[3] = t0.w;
                pxh[0] = t1.x; pxh[1] = t1.y; pxh[2] = t1.z; pxh[3] = t1.w;
                const float *cx = k.cxv;
                const float4 e0 = ldq_c(cx + C_B * ld), e1 = ldq_c(cx + C_A * ld), e2 = ldq_c(cx + C_RK * ld);
                const float4 e3 = ldq_c(cx + C_BH * ld), e4 = ldq_c(cx + C_AH * ld), e5 = ldq_c(cx + C_RKH * ld);
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
__device__ __forceinline__ void stream_fwd_body(const KArgs &a, const StreamArgs &sa, const int s, const int4 wk, const int lane,
                                                const unsigned smem_warp, const float4 *ring_ptr, const TmaMaps *tm = nullptr,
                                                const unsigned bar_s = 0)
{
    constexpr int NST = EDGE ? FR_NST_E : FR_NST_I;
    const Dims &d = a.d;
    FwdCtx k;
    k.ld = d.ldx; k.nzA = d.nzA; k.nPml = d.nPml; k.fsz = d.fsz;
    const size_t fsz = d.fsz;
    const int p = sa.it & 1;
    float *st = slot_state(a, s);
    k.xq0 = wk.x - 4 + 4 * lane;                           // true first column of this lane's quad
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
    k.ring_s = smem_warp + lane * 16; k.ring_p = ring_ptr + lane;
    k.tm = tm; k.ring_w = smem_warp; k.bar_s = bar_s; k.x0 = wk.x - 4; k.cin = p ? S_FWD1 : S_FWD; k.slot = s; k.lane = lane;
    k.tma = !EDGE && sa.tma != 0 && tm != nullptr;
    if (k.tma) {      // one mbarrier per ring stage, armed by lane 0 (count 1) and completed by the TMA unit's byte count
        if (lane == 0) {
#pragma unroll
            for (int j = 0; j < NST; j++) mbar_init(bar_s + 8u * j, 1);
            mbar_fence_init();
        }
        __syncwarp();
    }
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
    const int r0 = k.zc0 - 2;
    // v rows r0-2 .. r0+1 -> slots 0..3 ; row rho lives in slot (rho - (zc0-4)) % 6 ; row r+2 arrives through the ring
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const size_t ro = rowoff(r0 - 2 + j);
        w.vz[j] = ldq(k.g + F_VZ * fsz + ro); w.vx[j] = ldq(k.g + F_VX * fsz + ro);
    }
#pragma unroll
    for (int j = 0; j < NST - 1; j++) stream_fwd_issue<EDGE>(k, r0 + j, j);
    const int niter = (k.zc1 - k.zc0) + 4;
    // UNR rows per trip, then the windows move down UNR slots; surplus rows of the last trip are computed and dropped.
    // Edge warps: one row per trip (their body is much longer).
    constexpr int UNR = EDGE ? 1 : SW_UNR_FWD;
    static_assert(UNR == 1 || UNR == 2, "one or two rows per trip");
    // the first four rows only fill the windows of the second phase (its row q = r - 2 is not owned yet): interior warps run them
    // in a lead-in loop without that phase, edge warps (one long body) test per row
    constexpr int PBM = EDGE ? 2 : 0, PBL = EDGE ? 2 : 1;
    int stage = 0;
    unsigned phase = 0;      // parity of the stage's mbarrier: flips each time the ring wraps
#define FWD_NEXT_STAGE() do { if (stage == NST - 1) { stage = 0; phase ^= 1u; } else stage++; } while (0)
    int kk = 0;
    if (!EDGE) {
#pragma unroll 1
        for (; kk < 4; kk += UNR) {
            const int r = r0 + kk;
            stream_fwd_row<EDGE, 0, PBL>(k, w, r, stage, phase); FWD_NEXT_STAGE();
            if (UNR > 1) { stream_fwd_row<EDGE, 1 % UNR, PBL>(k, w, r + 1, stage, phase); FWD_NEXT_STAGE(); }
            win_shift<UNR>(w.vz); win_shift<UNR>(w.vx); win_shift<UNR>(w.zz); win_shift<UNR>(w.xz); win_shift<UNR>(w.xx);
        }
    }
#pragma unroll 1
    for (; kk < niter; kk += UNR) {
        const int r = r0 + kk;
        stream_fwd_row<EDGE, 0, PBM>(k, w, r, stage, phase); FWD_NEXT_STAGE();
        if (UNR > 1) { stream_fwd_row<EDGE, 1 % UNR, PBM>(k, w, r + 1, stage, phase); FWD_NEXT_STAGE(); }
        win_shift<UNR>(w.vz); win_shift<UNR>(w.vx); win_shift<UNR>(w.zz); win_shift<UNR>(w.xz); win_shift<UNR>(w.xx);
    }
    if (!EDGE && k.tma) {
        // the NST - 1 rows requested ahead are still in flight: the warp must not leave (and free its shared memory) before they land
#pragma unroll
        for (int j = 0; j < NST - 1; j++) { mbar_wait(k.bar_s + 8u * stage, phase); FWD_NEXT_STAGE(); }
    } else cp_wait<0>();
#undef FWD_NEXT_STAGE
}

#ifndef SW_MINB
#define SW_MINB 2
#endif
// grid: x = nAux + ceil(nWork / SW_WPB), y = slot ; dynamic shared memory FR_SMEM
__global__ void __launch_bounds__(SW_NT, SW_MINB) k_stream_fwd(const KArgs a, const StreamArgs sa, const __grid_constant__ TmaMaps tm)
{
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) unsigned long long bars[SW_WPB][4];
    pdl_launch_dependents();
    const int s = blockIdx.y;
    if ((int)blockIdx.x < sa.nAux) { pdl_wait(); stream_fwd_aux(a, sa, s); return; }
    const int wg = ((int)blockIdx.x - sa.nAux) * SW_WPB + ((int)threadIdx.x >> 5);
    if (wg >= sa.nWork) return;      // whole warp leaves (the in-place CPML memory must not be updated twice)
    const int4 wk = __ldg(sa.work + wg);
    const int lane = threadIdx.x & 31;
    const unsigned sw = (unsigned)__cvta_generic_to_shared(smem) + (threadIdx.x >> 5) * FR_WARP_BYTES;
    const float4 *sp = reinterpret_cast<const float4 *>(smem) + (threadIdx.x >> 5) * (FR_WARP_BYTES / 16);
    pdl_wait();
    const unsigned bar_s = smem_u32(&bars[threadIdx.x >> 5][0]);
#ifdef SW_INTERIOR_ONLY      // timing experiment: the edge body is not even compiled (wrong results at the edges)
    stream_fwd_body<false>(a, sa, s, wk, lane, sw, sp, &tm, bar_s);
#else
    if ((wk.w == 0 || sa.force == 1) && sa.force != 2) stream_fwd_body<false>(a, sa, s, wk, lane, sw, sp, &tm, bar_s);
    else stream_fwd_body<true>(a, sa, s, wk, lane, sw, sp);
#endif
}

// ================================================================================================
// adjoint sweep  adj(it+1) -> adj(it)   [adjoint buffer pa -> pa^1]
//   phase A (row r)   adjoint velocities from the old adjoint stresses, their CPML memory, residual injection
//   phase B (row r-2) adjoint stresses from the new adjoint velocities (register windows), their CPML memory
// el_velocity_adj.cu:22-108, el_stress_adj.cu:22-104, res_injection_exx/_ezz utilities.cu:605-641,
// source_grad utilities.cu:719-730 (stf gradient: done by the leading CTA from the old state).
struct AdjCtx {
    const float *g, *m, *psrc, *cxa, *cz, *res;
    float *o, *pdst;
    const int *injPtr;       // this strip's row CSR of injection targets: injPtr[z] .. injPtr[z+1]
    const int *injList;      // -> index into injCell / injField / injPtr of the slot tables
    const SlotTab *t;
    float *stage;            // per-warp shared staging row: [2][128] floats (vz, vx increments by column)
    const float4 *ring_p;    // this lane's 16 bytes of ring stage 0, array 0
    unsigned ring_s;
    size_t fsz, tb, cb;
    int ld, nzA, nx, nPml, zc0, zc1, xq0, lane, s, nSteps;
    unsigned amask;
    bool lown, colok, anyinj;
    bool xpl, xsl, xany;     // lane's quad touches the x CPML strip (nPml wide) / the kept strip (nPml + 2 wide); any lane of the warp does
    float c1z, c2z, c1x, c2x, dt;
};
struct AdjWin {
    float4 sz[6], sx[6], sxz[6];      // old adjoint stresses, rows r-2 .. r+2
    float4 vz[6], vx[6];              // new adjoint velocities, rows r-4 .. r
    float4 qxx[3], qxz[3];            // edge warps: new x CPML memory (P_SXX_X, P_SXZ_X) of rows r-2 .. r, phase B's x windows
};
// operand ring (same scheme as the forward kernel): ten row quads per iteration, requested AR_NST-1 rows ahead
constexpr int AR_NARR = 10, AR_NST = 3;
constexpr int AR_WARP_BYTES = AR_NARR * AR_NST * 512;          // 15 KB per warp
constexpr size_t AR_SMEM = (size_t)SW_WPB * AR_WARP_BYTES;     // 60 KB per CTA
// edge warps use the same 15 KB as 2 stages x 15 arrays: the five extra slots carry the old-time-level CPML memory that phase A
// of the row consumes right away -- ncu showed 35 % of the edge warps' stall samples on the shuffles waiting for those in-line
// loads.  A warp that touches the x strips takes its four x arrays, any other edge warp the z arrays of the row.
constexpr int AR_NARR_E = 15, AR_NST_E = 2;
static_assert(AR_NARR_E * AR_NST_E * 512 <= AR_WARP_BYTES, "edge ring must fit the interior ring");
enum { AA_SZ = 0, AA_SX, AA_SXZ, AA_OVZ, AA_OVX, AA_LAM, AA_MU, AA_MUA, AA_BYA, AA_BYB,    // rows r+2 (stresses), r, r-2 (buoyancies)
       AE_0, AE_1, AE_2, AE_3, AE_4 };
// x-type warp: psrc P_VX_X, P_VZ_X (x windows of phase A), psrc P_SXX_X, P_SXZ_X (old values of the update), row r
enum { AX_PVXX = AE_0, AX_PVZX = AE_1, AX_PSXXX = AE_2, AX_PSXZX = AE_3 };
// z-type warp: psrc P_VX_Z row r+1, P_VZ_Z row r+2 (newest rows of the z windows), psrc P_SXZ_Z, P_SZZ_Z row r (old values)
enum { AZ_PVXZ = AE_0, AZ_PVZZ = AE_1, AZ_PSXZZ = AE_2, AZ_PSZZZ = AE_3 };

// residual injection into the freshly updated adjoint velocities of row r (all 128 columns of the warp)
__device__ __forceinline__ void stream_adj_inject(const AdjCtx &k, const int r, float nvz[4], float nvx[4])
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

// request the operands of the iteration whose phase-A row is r
template <bool EDGE>
__device__ __forceinline__ void stream_adj_issue(const AdjCtx &k, const int r, const int stage)
{
    constexpr int NARR = EDGE ? AR_NARR_E : AR_NARR;
    const int ld = k.ld, nzA = k.nzA;
    const size_t fsz = k.fsz;
    auto rowoff = [&](int row) { return (size_t)min(max(row, 0), nzA - 1) * ld; };
    const size_t r2 = rowoff(r + 2), r0 = rowoff(r), rq = rowoff(r - 2);
    const unsigned sb = k.ring_s + (unsigned)stage * (NARR * 512);
    if (EDGE) {
        const bool rowact = (r >= 2 && r <= nzA - 3);
        if (k.xany) {
            if (rowact && k.xsl) { cp16(sb + AX_PVXX * 512, k.psrc + (size_t)P_VX_X * fsz + r0); cp16(sb + AX_PVZX * 512, k.psrc + (size_t)P_VZ_X * fsz + r0); }
            if (rowact && k.xpl) { cp16(sb + AX_PSXXX * 512, k.psrc + (size_t)P_SXX_X * fsz + r0); cp16(sb + AX_PSXZX * 512, k.psrc + (size_t)P_SXZ_X * fsz + r0); }
        } else if (rowact && ((r < k.nPml) || (r > nzA - k.nPml - 1))) {
            cp16(sb + AZ_PVXZ * 512, k.psrc + (size_t)P_VX_Z * fsz + r0 + ld); cp16(sb + AZ_PVZZ * 512, k.psrc + (size_t)P_VZ_Z * fsz + r0 + 2 * ld);
            cp16(sb + AZ_PSXZZ * 512, k.psrc + (size_t)P_SXZ_Z * fsz + r0); cp16(sb + AZ_PSZZZ * 512, k.psrc + (size_t)P_SZZ_Z * fsz + r0);
        }
    }
    cp16(sb + AA_SZ * 512, k.g + F_SZZ * fsz + r2); cp16(sb + AA_SX * 512, k.g + F_SXX * fsz + r2); cp16(sb + AA_SXZ * 512, k.g + F_SXZ * fsz + r2);
    cp16(sb + AA_OVZ * 512, k.g + F_VZ * fsz + r0); cp16(sb + AA_OVX * 512, k.g + F_VX * fsz + r0);
    cp16(sb + AA_LAM * 512, k.m + M_LAM * fsz + r0); cp16(sb + AA_MU * 512, k.m + M_MU * fsz + r0); cp16(sb + AA_MUA * 512, k.m + M_MUAVE * fsz + r0);
    // buoyancies: interior warps need row r-2 (phase B); edge warps take row r for the CPML memory update of phase A -- phase B then
    // re-reads row r-2 in line, an L1 hit (it was requested two iterations earlier), where row r in line was an L2 round trip per row
    const size_t rb = EDGE ? r0 : rq;
    cp16(sb + AA_BYA * 512, k.m + M_BYCA * fsz + rb); cp16(sb + AA_BYB * 512, k.m + M_BYCB * fsz + rb);
    cp_commit();
}

template <bool EDGE, int U, int PB>
__device__ __forceinline__ void stream_adj_row(const AdjCtx &k, AdjWin &w, const int r, const int stage)
{
    const int ld = k.ld, nzA = k.nzA;
    const size_t fsz = k.fsz;
    const float c1z = k.c1z, c2z = k.c2z, c1x = k.c1x, c2x = k.c2x, dt = k.dt;
    constexpr int u = U;
    constexpr int NST = EDGE ? AR_NST_E : AR_NST, NARR = EDGE ? AR_NARR_E : AR_NARR;
    stream_adj_issue<EDGE>(k, r + (NST - 1), stage == 0 ? NST - 1 : stage - 1);
    cp_wait<NST - 1>();
    const float4 *sb = k.ring_p + stage * (NARR * 32);
    w.sz[(u + 4) % 6] = sb[AA_SZ * 32]; w.sx[(u + 4) % 6] = sb[AA_SX * 32]; w.sxz[(u + 4) % 6] = sb[AA_SXZ * 32];     // row r+2
    // ---- phase A: adjoint velocities at row r from the old adjoint stresses, rows r-2 .. r+2 (slots u .. u+4)
    {
        const float4 z1 = w.sz[(u + 1) % 6], z2 = w.sz[(u + 2) % 6], z3 = w.sz[(u + 3) % 6], z4 = w.sz[(u + 4) % 6];
        const float4 x1 = w.sx[(u + 1) % 6], x2 = w.sx[(u + 2) % 6], x3 = w.sx[(u + 3) % 6], x4 = w.sx[(u + 4) % 6];
        const float4 q0 = w.sxz[u % 6], q1 = w.sxz[(u + 1) % 6], q2 = w.sxz[(u + 2) % 6], q3 = w.sxz[(u + 3) % 6];
        const float wzz[7] = XWIN_F(z2), wxx[7] = XWIN_F(x2), wxz[7] = XWIN_B(q2);
        const float zm1[4] = Q4(z1), zc0[4] = Q4(z2), zp1[4] = Q4(z3), zp2[4] = Q4(z4);
        const float xm1[4] = Q4(x1), xc0[4] = Q4(x2), xp1[4] = Q4(x3), xp2[4] = Q4(x4);
        const float m2[4] = Q4(q0), m1[4] = Q4(q1), c0[4] = Q4(q2), p1[4] = Q4(q3);
        const float4 lam4 = sb[AA_LAM * 32], mu4 = sb[AA_MU * 32], mua4 = sb[AA_MUA * 32], ovz4 = sb[AA_OVZ * 32], ovx4 = sb[AA_OVX * 32];
        const float l[4] = Q4(lam4), mm[4] = Q4(mu4), ma[4] = Q4(mua4);
        const float ovz[4] = Q4(ovz4), ovx[4] = Q4(ovx4);
        float nvz[4], nvx[4];
        const bool rown = (r >= k.zc0) && (r < k.zc1);
        const size_t ro = (size_t)r * ld;
        if (!EDGE) {
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const float l2u = l[c] + 2.0f * mm[c];
                const float accx = (l[c] * -DX7(wzz, c) + l2u * -DX7(wxx, c)) * dt + ma[c] * -DZ4(m2[c], m1[c], c0[c], p1[c]) * dt;
                const float accz = (l2u * -DZ4(zm1[c], zc0[c], zp1[c], zp2[c]) + l[c] * -DZ4(xm1[c], xc0[c], xp1[c], xp2[c])) * dt + ma[c] * -DX7(wxz, c) * dt;
                nvx[c] = ovx[c] + accx; nvz[c] = ovz[c] + accz;
            }
        } else {
            // CPML rows / strips, vectorised: quads + shuffles for the x stencils of the memory variables, four row quads
            // for the z stencils (re-read through L1).  a = 0 and 1/K = 1 outside the strips, so no per-cell branches.
            const bool rowact = (r >= 2 && r <= nzA - 3);
            float rKz = 1.f, az = 0.f, rKzh = 1.f, azh = 0.f, bz = 0.f, bzh = 0.f;
            const bool zp = rowact && ((r < k.nPml) || (r > nzA - k.nPml - 1));
            if (rowact) {
                const float *cz = k.cz + r;
                rKz = cz[C_RK * nzA]; az = cz[C_A * nzA]; rKzh = cz[C_RKH * nzA]; azh = cz[C_AH * nzA]; bz = cz[C_B * nzA]; bzh = cz[C_BH * nzA];
            }
            const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
            const bool xl = rowact && k.xpl;                     // this lane's quad touches the x strip (nPml wide)
            const float4 e0 = ldq_c(k.cxa + C_RK * ld), e1 = ldq_c(k.cxa + C_A * ld), e2 = ldq_c(k.cxa + C_RKH * ld), e3 = ldq_c(k.cxa + C_AH * ld);
            const float rKx[4] = Q4(e0), ax[4] = Q4(e1), rKxh[4] = Q4(e2), axh[4] = Q4(e3);
            float dpx[4] = {0.f, 0.f, 0.f, 0.f}, dpz[4] = {0.f, 0.f, 0.f, 0.f};     // a-weighted stencils of the old memory variables
            if (k.xany) {      // warp-uniform: shuffles inside
                const bool xld = rowact && k.xsl;      // the stencils reach two columns into the kept strip (nPml + 2 wide)
                const float4 pa = xld ? sb[AX_PVXX * 32] : zero4, pb = xld ? sb[AX_PVZX * 32] : zero4;     // prefetched with the row's operands
                const float wa[7] = XWIN_F(pa), wb[7] = XWIN_B(pb);
#pragma unroll
                for (int c = 0; c < 4; c++) { dpx[c] = ax[c] * -DX7(wa, c); dpz[c] = axh[c] * -DX7(wb, c); }
            }
            if (zp) {
                const float *pz0 = k.psrc + (size_t)P_VX_Z * fsz + ro, *pz1 = k.psrc + (size_t)P_VZ_Z * fsz + ro;
                const float4 t0 = ldq(pz0 - 2 * ld), t1 = ldq(pz0 - ld), t2 = ldq(pz0), t3 = k.xany ? ldq(pz0 + ld) : sb[AZ_PVXZ * 32];
                const float4 s0 = ldq(pz1 - ld), s1 = ldq(pz1), s2 = ldq(pz1 + ld), s3 = k.xany ? ldq(pz1 + 2 * ld) : sb[AZ_PVZZ * 32];
                const float a0[4] = Q4(t0), a1[4] = Q4(t1), a2[4] = Q4(t2), a3[4] = Q4(t3), b0[4] = Q4(s0), b1[4] = Q4(s1), b2[4] = Q4(s2), b3[4] = Q4(s3);
#pragma unroll
                for (int c = 0; c < 4; c++) { dpx[c] += azh * -DZ4(a0[c], a1[c], a2[c], a3[c]); dpz[c] += az * -DZ4(b0[c], b1[c], b2[c], b3[c]); }
            }
#pragma unroll
            for (int c = 0; c < 4; c++) {
                nvx[c] = ovx[c]; nvz[c] = ovz[c];
                if (rowact && ((k.amask >> c) & 1u)) {
                    const float l2u = l[c] + 2.0f * mm[c];
                    const float accx = (l[c] * -DX7(wzz, c) + l2u * -DX7(wxx, c)) * rKx[c] * dt + ma[c] * rKzh * -DZ4(m2[c], m1[c], c0[c], p1[c]) * dt + dpx[c];
                    const float accz = (l2u * -DZ4(zm1[c], zc0[c], zp1[c], zp2[c]) + l[c] * -DZ4(xm1[c], xc0[c], xp1[c], xp2[c])) * rKz * dt + ma[c] * rKxh[c] * -DX7(wxz, c) * dt + dpz[c];
                    nvx[c] = ovx[c] + accx; nvz[c] = ovz[c] + accz;
                }
            }
            // CPML memory of the adjoint velocities (strips only, from the velocities BEFORE the injection).  Written for every
            // column whose stencil window is complete (lane 0 comps 2,3 .. lane 31 comps 0,1): phase B reads them back, and the
            // neighbouring warps that recompute the same halo cells write the same values.
            w.qxx[2] = zero4; w.qxz[2] = zero4;      // what phase B would read back from pdst two rows later (zero outside the x strips)
            if (xl || zp) {
                const float4 bar4 = sb[AA_BYA * 32], bbr4 = sb[AA_BYB * 32];      // row r (edge ring)
                const float bar[4] = Q4(bar4), bbr[4] = Q4(bbr4);
                unsigned vm = k.amask;                               // columns with a complete window that are active
                if (k.lane == 0) vm &= 0xcu;
                if (k.lane == 31) vm &= 0x3u;
                if (!k.colok) vm = 0;
                if (xl && vm) {
                    const float4 f0 = ldq_c(k.cxa + C_BH * ld), f1 = ldq_c(k.cxa + C_B * ld);
                    const float bxh[4] = Q4(f0), bx[4] = Q4(f1);
                    const float4 o0 = sb[AX_PSXXX * 32], o1 = sb[AX_PSXZX * 32];      // xl implies an x-type warp
                    float n0[4] = Q4(o0), n1[4] = Q4(o1);
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        const int x = k.xq0 + c;
                        if (((vm >> c) & 1u) && ((x < k.nPml) || (x > k.nx - k.nPml - 1))) {
                            n0[c] = bxh[c] * n0[c] + bbr[c] * nvx[c] * dt;
                            n1[c] = bx[c] * n1[c] + bar[c] * nvz[c] * dt;
                        }
                    }
                    stq_halo(k.pdst + (size_t)P_SXX_X * fsz + ro, n0, k.lane); stq_halo(k.pdst + (size_t)P_SXZ_X * fsz + ro, n1, k.lane);
                    w.qxx[2] = mk4(n0); w.qxz[2] = mk4(n1);
                }
                if (zp && vm) {
                    const float4 o0 = k.xany ? ldq(k.psrc + (size_t)P_SXZ_Z * fsz + ro) : sb[AZ_PSXZZ * 32],
                                 o1 = k.xany ? ldq(k.psrc + (size_t)P_SZZ_Z * fsz + ro) : sb[AZ_PSZZZ * 32];
                    float n0[4] = Q4(o0), n1[4] = Q4(o1);
#pragma unroll
                    for (int c = 0; c < 4; c++)
                        if ((vm >> c) & 1u) {
                            n0[c] = bz * n0[c] + bbr[c] * nvx[c] * dt;
                            n1[c] = bzh * n1[c] + bar[c] * nvz[c] * dt;
                        }
                    stq_halo(k.pdst + (size_t)P_SXZ_Z * fsz + ro, n0, k.lane); stq_halo(k.pdst + (size_t)P_SZZ_Z * fsz + ro, n1, k.lane);
                    // phase B reads these rows back over the next four iterations (z stencils of the new memory variables): start
                    // the L1 fill now instead of at the first dependent load
                    pf_l1(k.pdst + (size_t)P_SXZ_Z * fsz + ro); pf_l1(k.pdst + (size_t)P_SZZ_Z * fsz + ro);
                }
            }
        }
        if (k.anyinj) stream_adj_inject(k, r, nvz, nvx);
        const float4 rvz = mk4(nvz), rvx = mk4(nvx);
        w.vz[(u + 4) % 6] = rvz; w.vx[(u + 4) % 6] = rvx;              // new adjoint velocity row r-j lives in slot u+4-j
        if (k.lown && rown) { stq(k.o + F_VZ * fsz + ro, rvz); stq(k.o + F_VX * fsz + ro, rvx); }
    }
    if (EDGE) __syncwarp();      // phase B reads CPML memory written by other lanes of this warp
    // ---- phase B: adjoint stresses at row q = r-2 from v^z rows q-2..q+1, v^x rows q-1..q+2 (row r-j lives in slot u+4-j)
    if (PB == 0 || (PB == 2 && (r - 2 >= k.zc0) && (r - 2 < k.zc1))) {      // warp-uniform: nothing of this phase is kept for a row the chunk does not own
        const int q = r - 2;
        const float4 v0 = w.vz[u % 6], v1 = w.vz[(u + 1) % 6], v2 = w.vz[(u + 2) % 6], v3 = w.vz[(u + 3) % 6];
        const float4 u0 = w.vx[(u + 1) % 6], u1 = w.vx[(u + 2) % 6], u2 = w.vx[(u + 3) % 6], u3 = w.vx[(u + 4) % 6];
        const float wvz[7] = XWIN_F(v2), wvx[7] = XWIN_B(u1);
        const float vzm2[4] = Q4(v0), vzm1[4] = Q4(v1), vzc[4] = Q4(v2), vzp1[4] = Q4(v3);
        const float vxm1[4] = Q4(u0), vxc[4] = Q4(u1), vxp1[4] = Q4(u2), vxp2[4] = Q4(u3);
        const float ozz[4] = Q4(w.sz[u % 6]), oxx[4] = Q4(w.sx[u % 6]), oxz[4] = Q4(w.sxz[u % 6]);     // old stresses of row r-2 live in slot u
        const size_t ro = (size_t)q * ld;      // q is an owned row here
        const float4 bya4 = EDGE ? ldq(k.m + M_BYCA * fsz + ro) : sb[AA_BYA * 32], byb4 = EDGE ? ldq(k.m + M_BYCB * fsz + ro) : sb[AA_BYB * 32];
        const float ba[4] = Q4(bya4), bb[4] = Q4(byb4);
        float nzz[4], nxz[4], nxx[4];
        const bool qown = (q >= k.zc0) && (q < k.zc1);
        if (!EDGE) {
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const float mdxf_vz = -DX7(wvz, c), mdzf_vx = -DZ4(vxm1[c], vxc[c], vxp1[c], vxp2[c]);
                const float mdxb_vx = -DX7(wvx, c), mdzb_vz = -DZ4(vzm2[c], vzm1[c], vzc[c], vzp1[c]);
                nxz[c] = oxz[c] + (mdxf_vz * ba[c] * dt + mdzf_vx * bb[c] * dt);
                nxx[c] = oxx[c] + bb[c] * mdxb_vx * dt;
                nzz[c] = ozz[c] + ba[c] * mdzb_vz * dt;
            }
        } else {
            const bool rowact = qown && (q >= 2 && q <= nzA - 3);
            float rKz = 1.f, az = 0.f, rKzh = 1.f, azh = 0.f, bz = 0.f, bzh = 0.f;
            if (rowact) {
                const float *cz = k.cz + q;
                rKz = cz[C_RK * nzA]; az = cz[C_A * nzA]; rKzh = cz[C_RKH * nzA]; azh = cz[C_AH * nzA]; bz = cz[C_B * nzA]; bzh = cz[C_BH * nzA];
            }
            const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
            const bool zpq = rowact && ((q < k.nPml) || (q > nzA - k.nPml - 1));       // rows where a_z can be non-zero
            const bool zst = rowact && ((q < k.nPml + 2) || (q > nzA - k.nPml - 3));   // rows whose memory variables are kept
            const bool xl2 = rowact && k.xsl && k.lown;
            const float4 e0 = ldq_c(k.cxa + C_RK * ld), e1 = ldq_c(k.cxa + C_A * ld), e2 = ldq_c(k.cxa + C_RKH * ld), e3 = ldq_c(k.cxa + C_AH * ld);
            const float rKx[4] = Q4(e0), ax[4] = Q4(e1), rKxh[4] = Q4(e2), axh[4] = Q4(e3);
            float dxz[4] = {0.f, 0.f, 0.f, 0.f}, dxx[4] = {0.f, 0.f, 0.f, 0.f}, dzz[4] = {0.f, 0.f, 0.f, 0.f};
            if (k.xany) {      // x stencils of the memory variables written in phase A of this and earlier rows (plain loads: same-warp data)
                const bool xld = rowact && k.xsl;
                const float4 pa = xld ? w.qxz[0] : zero4, pb = xld ? w.qxx[0] : zero4;      // this warp's own phase-A values of row q = r-2
                const float wa[7] = XWIN_F(pa), wb[7] = XWIN_B(pb);
#pragma unroll
                for (int c = 0; c < 4; c++) { dxz[c] = ax[c] * -DX7(wa, c); dxx[c] = axh[c] * -DX7(wb, c); }
            }
            if (zpq) {
                const float *pz0 = k.pdst + (size_t)P_SXZ_Z * fsz + ro, *pz1 = k.pdst + (size_t)P_SZZ_Z * fsz + ro;
                const float4 t0 = ldq_rw(pz0 - ld), t1 = ldq_rw(pz0), t2 = ldq_rw(pz0 + ld), t3 = ldq_rw(pz0 + 2 * ld);
                const float4 s0 = ldq_rw(pz1 - 2 * ld), s1 = ldq_rw(pz1 - ld), s2 = ldq_rw(pz1), s3 = ldq_rw(pz1 + ld);
                const float a0[4] = Q4(t0), a1[4] = Q4(t1), a2[4] = Q4(t2), a3[4] = Q4(t3), b0[4] = Q4(s0), b1[4] = Q4(s1), b2[4] = Q4(s2), b3[4] = Q4(s3);
#pragma unroll
                for (int c = 0; c < 4; c++) { dxz[c] += az * -DZ4(a0[c], a1[c], a2[c], a3[c]); dzz[c] = azh * -DZ4(b0[c], b1[c], b2[c], b3[c]); }
            }
#pragma unroll
            for (int c = 0; c < 4; c++) {
                nxz[c] = oxz[c]; nxx[c] = oxx[c]; nzz[c] = ozz[c];
                if (rowact && ((k.amask >> c) & 1u)) {
                    const float mdxf_vz = -DX7(wvz, c), mdzf_vx = -DZ4(vxm1[c], vxc[c], vxp1[c], vxp2[c]);
                    const float mdxb_vx = -DX7(wvx, c), mdzb_vz = -DZ4(vzm2[c], vzm1[c], vzc[c], vzp1[c]);
                    nxz[c] = oxz[c] + (mdxf_vz * rKx[c] * ba[c] * dt + mdzf_vx * rKz * bb[c] * dt + dxz[c]);
                    nxx[c] = oxx[c] + (bb[c] * mdxb_vx * rKxh[c] * dt + dxx[c]);
                    nzz[c] = ozz[c] + (ba[c] * mdzb_vz * rKzh * dt + dzz[c]);
                }
            }
            // CPML memory of the adjoint stresses: owner-only, strips nPml + 2 wide (el_stress_adj.cu:67-72,88-95)
            if ((xl2 || zst) && k.lown) {
                const float4 l4 = ldq(k.m + M_LAM * fsz + ro), m4 = ldq(k.m + M_MU * fsz + ro), a4 = ldq(k.m + M_MUAVE * fsz + ro);
                const float lq[4] = Q4(l4), mq[4] = Q4(m4), maq[4] = Q4(a4);
                if (xl2) {
                    const float4 f0 = ldq_c(k.cxa + C_BH * ld), f1 = ldq_c(k.cxa + C_B * ld);
                    const float bxh[4] = Q4(f0), bx[4] = Q4(f1);
                    const float4 o0 = ldq(k.psrc + (size_t)P_VZ_X * fsz + ro), o1 = ldq(k.psrc + (size_t)P_VX_X * fsz + ro);
                    float n0[4] = Q4(o0), n1[4] = Q4(o1);
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        const int x = k.xq0 + c;
                        if (((k.amask >> c) & 1u) && ((x < k.nPml + 2) || (x > k.nx - k.nPml - 3))) {
                            n0[c] = bxh[c] * n0[c] + nxz[c] * maq[c] * dt;
                            n1[c] = bx[c] * n1[c] + lq[c] * nzz[c] * dt + (lq[c] + 2.0f * mq[c]) * nxx[c] * dt;
                        }
                    }
                    stq(k.pdst + (size_t)P_VZ_X * fsz + ro, mk4(n0)); stq(k.pdst + (size_t)P_VX_X * fsz + ro, mk4(n1));
                }
                if (zst) {
                    const float4 o0 = ldq(k.psrc + (size_t)P_VX_Z * fsz + ro), o1 = ldq(k.psrc + (size_t)P_VZ_Z * fsz + ro);
                    float n0[4] = Q4(o0), n1[4] = Q4(o1);
#pragma unroll
                    for (int c = 0; c < 4; c++)
                        if ((k.amask >> c) & 1u) {
                            n0[c] = bzh * n0[c] + nxz[c] * maq[c] * dt;
                            n1[c] = bz * n1[c] + (lq[c] + 2.0f * mq[c]) * nzz[c] * dt + lq[c] * nxx[c] * dt;
                        }
                    stq(k.pdst + (size_t)P_VX_Z * fsz + ro, mk4(n0)); stq(k.pdst + (size_t)P_VZ_Z * fsz + ro, mk4(n1));
                }
            }
        }
        if (k.lown && qown) {
            stq(k.o + F_SZZ * fsz + ro, mk4(nzz)); stq(k.o + F_SXZ * fsz + ro, mk4(nxz)); stq(k.o + F_SXX * fsz + ro, mk4(nxx));
        }
    }
}

template <bool EDGE>
__device__ __forceinline__ void stream_adj_body(const KArgs &a, const StreamArgs &sa, const int s, const int4 wk, const int lane, float *stage,
                                                const unsigned smem_warp, const float4 *ring_ptr)
{
    const Dims &d = a.d;
    AdjCtx k;
    k.ring_s = smem_warp + lane * 16; k.ring_p = ring_ptr + lane;
    k.ld = d.ldx; k.nzA = d.nzA; k.nx = d.nx; k.nPml = d.nPml; k.fsz = d.fsz; k.lane = lane; k.s = s; k.nSteps = d.nSteps;
    const size_t fsz = d.fsz;
    float *st = slot_state(a, s);
    k.xq0 = wk.x - 4 + 4 * lane;
    k.colok = (k.xq0 >= 0) && (k.xq0 < d.ldx);
    const int xq = k.colok ? k.xq0 : 0;
    k.g = st + (size_t)(sa.pa ? S_ADJ1 : S_ADJ) * fsz + xq;
    k.o = st + (size_t)(sa.pa ? S_ADJ : S_ADJ1) * fsz + xq;
    k.psrc = st + (size_t)(sa.pa ? S_APSI1 : S_APSI) * fsz + xq;
    k.pdst = st + (size_t)(sa.pa ? S_APSI : S_APSI1) * fsz + xq;
    k.m = a.model + xq;
    k.cxa = a.cxa + xq; k.cz = a.cz;
    k.zc0 = wk.y; k.zc1 = wk.z;
    k.lown = (lane >= 1) && (lane <= 30) && k.colok;
    k.c1z = d.c1z; k.c2z = d.c2z; k.c1x = d.c1x; k.c2x = d.c2x; k.dt = d.dt;
    k.t = &a.t; k.stage = stage;
    k.tb = (size_t)s * a.t.maxInj; k.cb = (size_t)s * a.t.maxCon;
    k.res = a.trace + ((size_t)s * d.nTrace + T_RES) * d.maxRec * d.nSteps + sa.it;
    {   // injection targets of this strip (halo columns included), rows of phase A
        const int strip = wk.x / SW_OWN;
        k.injPtr = a.t.sInjPtr + ((size_t)s * a.t.nStrips + strip) * (d.nzA + 1);
        k.injList = a.t.sInj + (size_t)s * 2 * a.t.maxInj;
        const int ra = max(k.zc0 - 2, 0), rb = min(k.zc1 + 2, d.nzA);
        k.anyinj = k.injPtr[rb] > k.injPtr[ra];
    }
    k.amask = 0xf; k.xpl = false; k.xsl = false; k.xany = false;
    if (EDGE) {
        k.amask = 0;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int x = k.xq0 + c;
            if (x >= 2 && x <= d.nx - 3) {
                k.amask |= 1u << c;
                if ((x < d.nPml) || (x > d.nx - d.nPml - 1)) k.xpl = true;
                if ((x < d.nPml + 2) || (x > d.nx - d.nPml - 3)) k.xsl = true;
            }
        }
        k.xany = __any_sync(0xffffffffu, k.xsl);
    }
    AdjWin w;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 6; j++) { w.sz[j] = w.sx[j] = w.sxz[j] = w.vz[j] = w.vx[j] = zero; }
#pragma unroll
    for (int j = 0; j < 3; j++) { w.qxx[j] = w.qxz[j] = zero; }
    auto rowoff = [&](int row) { return (size_t)min(max(row, 0), d.nzA - 1) * d.ldx; };
    const int r0 = k.zc0 - 2;
#pragma unroll
    for (int j = 0; j < 4; j++) {       // rows r0-2 .. r0+1 ; row r+2 arrives through the ring
        const size_t ro = rowoff(r0 - 2 + j);
        w.sz[j] = ldq(k.g + F_SZZ * fsz + ro); w.sx[j] = ldq(k.g + F_SXX * fsz + ro); w.sxz[j] = ldq(k.g + F_SXZ * fsz + ro);
    }
#pragma unroll
    for (int j = 0; j < (EDGE ? AR_NST_E : AR_NST) - 1; j++) stream_adj_issue<EDGE>(k, r0 + j, j);
    const int niter = (k.zc1 - k.zc0) + 4;
    constexpr int UNR = EDGE ? 1 : SW_UNR_ADJ, NST = EDGE ? AR_NST_E : AR_NST;
    static_assert(UNR == 1 || UNR == 2, "one or two rows per trip");
    constexpr int PBM = EDGE ? 2 : 0, PBL = EDGE ? 2 : 1;      // see stream_fwd_body
    int stg = 0, kk = 0;
    if (!EDGE) {
#pragma unroll 1
        for (; kk < 4; kk += UNR) {
            const int r = r0 + kk;
            stream_adj_row<EDGE, 0, PBL>(k, w, r, stg); stg = stg == NST - 1 ? 0 : stg + 1;
            if (UNR > 1) { stream_adj_row<EDGE, 1 % UNR, PBL>(k, w, r + 1, stg); stg = stg == NST - 1 ? 0 : stg + 1; }
            win_shift<UNR>(w.sz); win_shift<UNR>(w.sx); win_shift<UNR>(w.sxz); win_shift<UNR>(w.vz); win_shift<UNR>(w.vx);
        }
    }
#pragma unroll 1
    for (; kk < niter; kk += UNR) {
        const int r = r0 + kk;
        stream_adj_row<EDGE, 0, PBM>(k, w, r, stg); stg = stg == NST - 1 ? 0 : stg + 1;
        if (UNR > 1) { stream_adj_row<EDGE, 1 % UNR, PBM>(k, w, r + 1, stg); stg = stg == NST - 1 ? 0 : stg + 1; }
        win_shift<UNR>(w.sz); win_shift<UNR>(w.sx); win_shift<UNR>(w.sxz); win_shift<UNR>(w.vz); win_shift<UNR>(w.vx);
        if (EDGE) { w.qxx[0] = w.qxx[1]; w.qxx[1] = w.qxx[2]; w.qxz[0] = w.qxz[1]; w.qxz[1] = w.qxz[2]; }
    }
    cp_wait<0>();
}


// grid: x = 1 + ceil(nWork / SW_WPB), y = slot ; CTA 0 writes the stf gradient
__global__ void __launch_bounds__(SW_NT, SW_MINB) k_stream_adj(const KArgs a, const StreamArgs sa)
{
    __shared__ __align__(16) float stage[SW_WPB][256];
    extern __shared__ __align__(128) float smem[];
    pdl_launch_dependents();
    const int s = blockIdx.y;
    const Dims &d = a.d;
    if (blockIdx.x == 0) {
        pdl_wait();
        if (threadIdx.x == 0) {      // source_grad, utilities.cu:719-730, from the adjoint state after step it+1
            const float *src = slot_state(a, s) + (size_t)(sa.pa ? S_ADJ1 : S_ADJ) * d.fsz;
            const size_t i = (size_t)a.t.zs[s] * d.ldx + a.t.xs[s];
            a.gstf[(size_t)s * d.nSteps + sa.it] = -(src[(size_t)F_SZZ * d.fsz + i] + a.t.rxz[s] * src[(size_t)F_SXX * d.fsz + i]) * d.dt;
        }
        return;
    }
    const int wg = ((int)blockIdx.x - 1) * SW_WPB + ((int)threadIdx.x >> 5);
    if (wg >= sa.nWork) return;
    const int4 wk = __ldg(sa.work + wg);
    const int lane = threadIdx.x & 31;
    const unsigned sw = (unsigned)__cvta_generic_to_shared(smem) + (threadIdx.x >> 5) * AR_WARP_BYTES;
    const float4 *sp = reinterpret_cast<const float4 *>(smem) + (threadIdx.x >> 5) * (AR_WARP_BYTES / 16);
    pdl_wait();
#ifdef SW_INTERIOR_ONLY
    stream_adj_body<false>(a, sa, s, wk, lane, stage[threadIdx.x >> 5], sw, sp);
#else
    if ((wk.w == 0 || sa.force == 1) && sa.force != 2) stream_adj_body<false>(a, sa, s, wk, lane, stage[threadIdx.x >> 5], sw, sp);
    else stream_adj_body<true>(a, sa, s, wk, lane, stage[threadIdx.x >> 5], sw, sp);
#endif
}

// ================================================================================================
// reverse-time reconstruction + imaging   fwd(it+1) -> fwd(it)   [forward buffer q -> q^1], gradients += ...
//   stage 1 (row r)    v(it) = v - D(sigma) b dt on the interior, ring restore; density imaging terms ga, gb and the
//                      grho gather (rows r, r-1 / columns x, x-1)                       el_velocity.cu:84-117, to_bnd
//   stage 2 (row r-2)  sigma(it) = sigma - amp[src] - C D(v(it)) dt, ring restore; glam, gmu (normal part + the four-point
//                      shear spray of el_stress.cu:112-123 evaluated as a gather over rows r-2, r-3 / columns x, x-1)
// Same march as the other two kernels, but 18 operand quads per row: they are prefetched through a per-warp
// shared-memory ring with cp.async (each lane copies and later reads only its own 16 bytes, so a cp.async.wait_group
// is the only synchronisation) -- the registers hold just the stencil windows.
constexpr int RC_NARR = 18;
#ifndef RC_NST_
#define RC_NST_ 2
#endif
#ifndef RC_MINB
#define RC_MINB 2
#endif
constexpr int RC_NST = RC_NST_;                             // ring stages: operands are requested RC_NST-1 rows ahead
constexpr int RC_WARP_BYTES = RC_NST * RC_NARR * 512;       // 27 KB per warp
constexpr size_t RC_SMEM = (size_t)SW_WPB * RC_WARP_BYTES;  // 108 KB per CTA, 2 CTAs per SM
enum { RA_SZZ = 0, RA_SXZ, RA_SXX, RA_OVZ, RA_OVX, RA_AVZ, RA_AVX, RA_BA, RA_BB, RA_GR,      // rows r+2, r+1, r, then row r
       RA_ASZZ, RA_ASXZ, RA_ASXX, RA_LAM, RA_MU, RA_MUA, RA_GL, RA_GM };                        // row r-2

struct RecCtx {
    const float *g, *adj, *m, *ringb, *amp;
    float *o, *grad;
    unsigned ring_s;          // shared-memory address of this lane's 16 bytes in stage 0, array 0
    const float4 *ring_p;     // the same location as a pointer (plain loads: the compiler schedules them; cp_wait's memory clobber orders them)
    size_t fsz, rfs;
    int ld, nzA, nx, nPml, z1, x1, zc0, zc1, zs, xs, xq0;
    bool lown, ring;
    float c1z, c2z, c1x, c2x, dt;
    Dims d;
};
struct RecWin {
    float4 szz[6], sxz[6], sxx[6];     // old stresses (state it+1): rows r-2 .. r+2 / r+1 / r
    float4 vz[6], vx[6];               // reconstructed velocities (state it): rows r-4 .. r
    float ga_prev[4], sh_prev[4];      // density term of row r-1, shear term of row r-3
};

// request the operands of the iteration whose stage-1 row is r
__device__ __forceinline__ void stream_rec_issue(const RecCtx &k, const int r, const int stage)
{
    const int ld = k.ld, nzA = k.nzA;
    const size_t fsz = k.fsz;
    auto rowoff = [&](int row) { return (size_t)min(max(row, 0), nzA - 1) * ld; };
    const size_t r2 = rowoff(r + 2), r1 = rowoff(r + 1), r0 = rowoff(r), rq = rowoff(r - 2);
    const unsigned sb = k.ring_s + (unsigned)stage * (RC_NARR * 512);
    cp16(sb + RA_SZZ * 512, k.g + F_SZZ * fsz + r2); cp16(sb + RA_SXZ * 512, k.g + F_SXZ * fsz + r1); cp16(sb + RA_SXX * 512, k.g + F_SXX * fsz + r0);
    cp16(sb + RA_OVZ * 512, k.g + F_VZ * fsz + r0); cp16(sb + RA_OVX * 512, k.g + F_VX * fsz + r0);
    cp16(sb + RA_AVZ * 512, k.adj + F_VZ * fsz + r0); cp16(sb + RA_AVX * 512, k.adj + F_VX * fsz + r0);
    cp16(sb + RA_BA * 512, k.m + M_BYCA * fsz + r0); cp16(sb + RA_BB * 512, k.m + M_BYCB * fsz + r0);
    cp16(sb + RA_GR * 512, k.grad + 2 * fsz + r0);
    cp16(sb + RA_ASZZ * 512, k.adj + F_SZZ * fsz + rq); cp16(sb + RA_ASXZ * 512, k.adj + F_SXZ * fsz + rq); cp16(sb + RA_ASXX * 512, k.adj + F_SXX * fsz + rq);
    cp16(sb + RA_LAM * 512, k.m + M_LAM * fsz + rq); cp16(sb + RA_MU * 512, k.m + M_MU * fsz + rq); cp16(sb + RA_MUA * 512, k.m + M_MUAVE * fsz + rq);
    cp16(sb + RA_GL * 512, k.grad + 0 * fsz + rq); cp16(sb + RA_GM * 512, k.grad + 1 * fsz + rq);
    cp_commit();
}

template <bool EDGE, int U, int PB>
__device__ __forceinline__ void stream_rec_row(const RecCtx &k, RecWin &w, const int r, const int stage)
{
    const int ld = k.ld;
    const size_t fsz = k.fsz;
    const float c1z = k.c1z, c2z = k.c2z, c1x = k.c1x, c2x = k.c2x, dt = k.dt;
    constexpr int u = U;
    stream_rec_issue(k, r + (RC_NST - 1), stage == 0 ? RC_NST - 1 : stage - 1);
    cp_wait<RC_NST - 1>();
    const float4 *sb = k.ring_p + stage * (RC_NARR * 32);
    w.szz[(u + 4) % 6] = sb[RA_SZZ * 32];       // row r+2
    w.sxz[(u + 3) % 6] = sb[RA_SXZ * 32];       // row r+1
    w.sxx[(u + 2) % 6] = sb[RA_SXX * 32];       // row r
    // ---- stage 1: velocities of time `it` at row r ; density imaging
    {
        const float4 p0 = w.szz[(u + 1) % 6], p1 = w.szz[(u + 2) % 6], p2 = w.szz[(u + 3) % 6], p3 = w.szz[(u + 4) % 6];
        const float4 q0 = w.sxz[u % 6], q1 = w.sxz[(u + 1) % 6], q2 = w.sxz[(u + 2) % 6], q3 = w.sxz[(u + 3) % 6];
        const float4 xc = w.sxx[(u + 2) % 6];
        const float wxz[7] = XWIN_B(q2), wxx[7] = XWIN_F(xc);
        const float zzm1[4] = Q4(p0), zzc[4] = Q4(p1), zzp1[4] = Q4(p2), zzp2[4] = Q4(p3);
        const float xzm2[4] = Q4(q0), xzm1[4] = Q4(q1), xzc[4] = Q4(q2), xzp1[4] = Q4(q3);
        const float4 ovz4 = sb[RA_OVZ * 32], ovx4 = sb[RA_OVX * 32], avz4 = sb[RA_AVZ * 32], avx4 = sb[RA_AVX * 32];
        const float4 ba4 = sb[RA_BA * 32], bb4 = sb[RA_BB * 32];
        const float ovz[4] = Q4(ovz4), ovx[4] = Q4(ovx4), avz[4] = Q4(avz4), avx[4] = Q4(avx4), ba[4] = Q4(ba4), bb[4] = Q4(bb4);
        float nvz[4], nvx[4], ga[4], gb[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const float A = DZ4(zzm1[c], zzc[c], zzp1[c], zzp2[c]) + DX7(wxz, c);
            const float B = DZ4(xzm2[c], xzm1[c], xzc[c], xzp1[c]) + DX7(wxx, c);
            const bool in = !EDGE || interior(k.d, r, k.xq0 + c);
            nvz[c] = in ? ovz[c] - A * ba[c] * dt : ovz[c];
            nvx[c] = in ? ovx[c] - B * bb[c] * dt : ovx[c];
            ga[c] = in ? avz[c] * A * dt * (0.5f * ba[c] * ba[c]) : 0.f;
            gb[c] = in ? avx[c] * B * dt * (0.5f * bb[c] * bb[c]) : 0.f;
            if (EDGE && k.ring) {
                int r0, r1;
                ring_indices2(k.d, r, k.xq0 + c, r0, r1);
                const int ri = r0 >= 0 ? r0 : r1;
                if (ri >= 0) { nvz[c] = k.ringb[F_VZ * k.rfs + ri]; nvx[c] = k.ringb[F_VX * k.rfs + ri]; }
            }
        }
        const float4 rvz = mk4(nvz), rvx = mk4(nvx);
        w.vz[(u + 4) % 6] = rvz; w.vx[(u + 4) % 6] = rvx;       // reconstructed velocity row r-j lives in slot u+4-j
        // grho gather (el_velocity.cu:104-110 sprays ga to (z,x),(z+1,x) and gb to (z,x),(z,x+1))
        const float gbl = sh_l(gb[3]);
        const bool rown = (r >= k.zc0) && (r < k.zc1);
        const bool rreg = !EDGE || (r >= k.nPml - 2 && r <= k.z1 + 2 && k.xq0 + 3 >= k.nPml - 2 && k.xq0 <= k.x1 + 2);
        if (k.lown && rown && rreg) {
            const size_t ro = (size_t)r * ld;
            const float4 g4 = sb[RA_GR * 32];
            float gr[4] = Q4(g4);
            const float gbW[5] = {gbl, gb[0], gb[1], gb[2], gb[3]};
            const bool zle = !EDGE || (r <= k.z1);
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const bool in = !EDGE || interior(k.d, r, k.xq0 + c);
                float grr = (in ? ga[c] + gbW[c + 1] : 0.f) + gbW[c];
                if (zle) grr += w.ga_prev[c];
                gr[c] += grr;
            }
            stq(k.grad + 2 * fsz + ro, mk4(gr));
            stq(k.o + F_VZ * fsz + ro, rvz); stq(k.o + F_VX * fsz + ro, rvx);
        }
#pragma unroll
        for (int c = 0; c < 4; c++) w.ga_prev[c] = ga[c];
    }
    // ---- stage 2: stresses of time `it` at row q = r-2 ; lambda / mu imaging
    if (PB == 0 || (PB == 2 && (r - 2 >= k.zc0 - 1) && (r - 2 < k.zc1))) {      // warp-uniform: owned rows, and the row above them for its shear term (sh_prev)
        const int q = r - 2;
        const float4 v0 = w.vz[u % 6], v1 = w.vz[(u + 1) % 6], v2 = w.vz[(u + 2) % 6], v3 = w.vz[(u + 3) % 6];
        const float4 u0 = w.vx[(u + 1) % 6], u1 = w.vx[(u + 2) % 6], u2 = w.vx[(u + 3) % 6], u3 = w.vx[(u + 4) % 6];
        const float wvz[7] = XWIN_F(v2), wvx[7] = XWIN_B(u1);
        const float vzm2[4] = Q4(v0), vzm1[4] = Q4(v1), vzc[4] = Q4(v2), vzp1[4] = Q4(v3);
        const float vxm1[4] = Q4(u0), vxc[4] = Q4(u1), vxp1[4] = Q4(u2), vxp2[4] = Q4(u3);
        const float ozz[4] = Q4(w.szz[u % 6]), oxz[4] = Q4(w.sxz[u % 6]), oxx[4] = Q4(w.sxx[u % 6]);     // old stresses of row r-2: slot u
        const float4 za4 = sb[RA_ASZZ * 32], sa4 = sb[RA_ASXZ * 32], xa4 = sb[RA_ASXX * 32];
        const float4 lam4 = sb[RA_LAM * 32], mu4 = sb[RA_MU * 32], mua4 = sb[RA_MUA * 32];
        const float za[4] = Q4(za4), sa[4] = Q4(sa4), xa[4] = Q4(xa4), lam[4] = Q4(lam4), mu[4] = Q4(mu4), mua[4] = Q4(mua4);
        float D1[4], D2[4], D3[4], sh[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            D1[c] = DZ4(vzm2[c], vzm1[c], vzc[c], vzp1[c]); D2[c] = DX7(wvx, c);
            D3[c] = DZ4(vxm1[c], vxc[c], vxp1[c], vxp2[c]) + DX7(wvz, c);
            // mu_ave / sum(1/mu) = mu_ave^2 / 4  (mu_ave is the 4-point harmonic mean; 0 when any corner is 0)
            const float v = -sa[c] * D3[c] * dt * (0.25f * mua[c] * mua[c]) * 1e6f;
            sh[c] = (!EDGE || interior(k.d, q, k.xq0 + c)) ? v : 0.f;
        }
        const float shl = sh_l(sh[3]), shul = sh_l(w.sh_prev[3]);
        const bool qown = (q >= k.zc0) && (q < k.zc1);
        const bool qreg = !EDGE || (q >= k.nPml - 2 && q <= k.z1 + 2 && k.xq0 + 3 >= k.nPml - 2 && k.xq0 <= k.x1 + 2);
        if (k.lown && qown && qreg) {
            const size_t ro = (size_t)q * ld;
            const float4 g0 = sb[RA_GL * 32], g1 = sb[RA_GM * 32];
            float gl[4] = Q4(g0), gm[4] = Q4(g1);
            const float shW[5] = {shl, sh[0], sh[1], sh[2], sh[3]}, shUW[5] = {shul, w.sh_prev[0], w.sh_prev[1], w.sh_prev[2], w.sh_prev[3]};
            float nzz[4], nxz[4], nxx[4];
            const bool zle = !EDGE || (q <= k.z1);
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int x = k.xq0 + c;
                float tzz = ozz[c], txx = oxx[c], txz = oxz[c];
                if (q == k.zs && x == k.xs) { const float amp = *k.amp; tzz -= amp; txx -= amp; }
                const bool in = !EDGE || interior(k.d, q, x);
                if (in) {
                    const float l2u = lam[c] + 2.0f * mu[c];
                    tzz -= (l2u * D1[c] + lam[c] * D2[c]) * dt;
                    txx -= (lam[c] * D1[c] + l2u * D2[c]) * dt;
                    txz -= mua[c] * D3[c] * dt;
                    gl[c] += -(za[c] + xa[c]) * (D1[c] + D2[c]) * dt * 1e6f;
                }
                float shg = shW[c + 1] + shW[c];
                if (zle) { shg += shUW[c + 1]; if (!EDGE || x <= k.x1) shg += shUW[c]; }
                const float gmn = in ? (-2.0f * za[c] * D1[c] * dt - 2.0f * xa[c] * D2[c] * dt) * 1e6f : 0.f;
                // fast division (2 ulp): the IEEE routine's slow path is taken for the zero / denormal numerators that fill most
                // of the grid early in the reverse sweep and then costs more than the rest of the kernel
                if (!EDGE || in || shg != 0.f) gm[c] += gmn + __fdividef(shg, mu[c] * mu[c]);
                if (EDGE && k.ring) {
                    int r0, r1;
                    ring_indices2(k.d, q, x, r0, r1);
                    const int ri = r0 >= 0 ? r0 : r1;
                    if (ri >= 0) { tzz = k.ringb[F_SZZ * k.rfs + ri]; txz = k.ringb[F_SXZ * k.rfs + ri]; txx = k.ringb[F_SXX * k.rfs + ri]; }
                }
                nzz[c] = tzz; nxx[c] = txx; nxz[c] = txz;
            }
            stq(k.grad + 0 * fsz + ro, mk4(gl)); stq(k.grad + 1 * fsz + ro, mk4(gm));
            stq(k.o + F_SZZ * fsz + ro, mk4(nzz)); stq(k.o + F_SXZ * fsz + ro, mk4(nxz)); stq(k.o + F_SXX * fsz + ro, mk4(nxx));
        }
#pragma unroll
        for (int c = 0; c < 4; c++) w.sh_prev[c] = sh[c];
    }
}

template <bool EDGE>
__device__ __forceinline__ void stream_rec_body(const KArgs &a, const StreamArgs &sa, const int s, const int4 wk, const int lane, const unsigned smem_warp, const float4 *ring_ptr)
{
    const Dims &d = a.d;
    RecCtx k;
    k.d = d;
    k.ld = d.ldx; k.nzA = d.nzA; k.nx = d.nx; k.nPml = d.nPml; k.z1 = d.z1; k.x1 = d.x1; k.fsz = d.fsz;
    const size_t fsz = d.fsz;
    float *st = slot_state(a, s);
    k.xq0 = wk.x - 4 + 4 * lane;
    const bool colok = (k.xq0 >= 0) && (k.xq0 < d.ldx);
    const int xq = colok ? k.xq0 : 0;
    k.g = st + (size_t)(sa.q ? S_FWD1 : S_FWD) * fsz + xq;
    k.o = st + (size_t)(sa.q ? S_FWD : S_FWD1) * fsz + xq;
    k.adj = st + (size_t)(sa.pa ? S_ADJ1 : S_ADJ) * fsz + xq;
    k.m = a.model + xq;
    k.grad = a.grad + (size_t)s * 3 * fsz + xq;
    k.ringb = a.ring + (((size_t)s * NFIELD) * d.nSteps + sa.it) * d.ringLen;
    k.rfs = (size_t)d.nSteps * d.ringLen;
    k.amp = a.t.amp + (size_t)s * d.nSteps + sa.it;
    k.zs = a.t.zs[s]; k.xs = a.t.xs[s];
    // rows outside the region interior U ring = [nPml-2, z1+2] are never touched by the reverse sweep
    k.zc0 = max(wk.y, d.nPml - 2); k.zc1 = min(wk.z, d.z1 + 3);
    if (k.zc1 <= k.zc0) return;
    if (wk.x + SW_OWN <= d.nPml - 2 || wk.x > d.x1 + 2) return;
    k.lown = (lane >= 1) && (lane <= 30) && colok;
    k.c1z = d.c1z; k.c2z = d.c2z; k.c1x = d.c1x; k.c2x = d.c2x; k.dt = d.dt;
    k.ring = EDGE && tile_touches_ring_ext(d, k.zc0 - 2, k.zc1 + 1, wk.x - 4, wk.x + SW_OWN + 3);
    k.ring_s = smem_warp + lane * 16;
    k.ring_p = ring_ptr + lane;

    RecWin w;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 6; j++) { w.szz[j] = w.sxz[j] = w.sxx[j] = w.vz[j] = w.vx[j] = zero; }
#pragma unroll
    for (int c = 0; c < 4; c++) { w.ga_prev[c] = 0.f; w.sh_prev[c] = 0.f; }
    auto rowoff = [&](int row) { return (size_t)min(max(row, 0), d.nzA - 1) * d.ldx; };
    const int r0 = k.zc0 - 2;
    // windows before the first iteration: szz rows r0-1 .. r0+1 (slots 1..3), sxz rows r0-2 .. r0 (slots 0..2)
#pragma unroll
    for (int j = 0; j < 3; j++) {
        w.szz[j + 1] = ldq(k.g + F_SZZ * fsz + rowoff(r0 - 1 + j));
        w.sxz[j] = ldq(k.g + F_SXZ * fsz + rowoff(r0 - 2 + j));
    }
#pragma unroll
    for (int j = 0; j < RC_NST - 1; j++) stream_rec_issue(k, r0 + j, j);
    const int niter = (k.zc1 - k.zc0) + 4;
    constexpr int UNR = EDGE ? 1 : SW_UNR_REC;
    static_assert(UNR == 1 || UNR == 2, "one or two rows per trip");
    // stage 2 is needed from row q = zc0 - 1 on (its shear term enters row zc0): two lead-in rows without it for interior warps
    constexpr int PBM = EDGE ? 2 : 0, PBL = EDGE ? 2 : 1;
    int stage = 0, kk = 0;
    if (!EDGE) {
#pragma unroll 1
        for (; kk < 2; kk += UNR) {
            const int r = r0 + kk;
            stream_rec_row<EDGE, 0, PBL>(k, w, r, stage); stage = stage == RC_NST - 1 ? 0 : stage + 1;
            if (UNR > 1) { stream_rec_row<EDGE, 1 % UNR, PBL>(k, w, r + 1, stage); stage = stage == RC_NST - 1 ? 0 : stage + 1; }
            win_shift<UNR>(w.szz); win_shift<UNR>(w.sxz); win_shift<UNR>(w.sxx); win_shift<UNR>(w.vz); win_shift<UNR>(w.vx);
        }
    }
#pragma unroll 1
    for (; kk < niter; kk += UNR) {
        const int r = r0 + kk;
        stream_rec_row<EDGE, 0, PBM>(k, w, r, stage); stage = stage == RC_NST - 1 ? 0 : stage + 1;
        if (UNR > 1) { stream_rec_row<EDGE, 1 % UNR, PBM>(k, w, r + 1, stage); stage = stage == RC_NST - 1 ? 0 : stage + 1; }
        win_shift<UNR>(w.szz); win_shift<UNR>(w.sxz); win_shift<UNR>(w.sxx); win_shift<UNR>(w.vz); win_shift<UNR>(w.vx);
    }
    cp_wait<0>();
}

// grid: x = ceil(nWork / SW_WPB), y = slot ; dynamic shared memory RC_SMEM
__global__ void __launch_bounds__(SW_NT, RC_MINB) k_stream_recon(const KArgs a, const StreamArgs sa)
{
    extern __shared__ __align__(128) float smem[];
    pdl_launch_dependents();
    const int s = blockIdx.y;
    const int wg = (int)blockIdx.x * SW_WPB + ((int)threadIdx.x >> 5);
    if (wg >= sa.nWork) return;
    const int4 wk = __ldg(sa.work + wg);
    const int lane = threadIdx.x & 31;
    pdl_wait();
    const unsigned sw = (unsigned)__cvta_generic_to_shared(smem) + (threadIdx.x >> 5) * RC_WARP_BYTES;
    const float4 *sp = reinterpret_cast<const float4 *>(smem) + (threadIdx.x >> 5) * (RC_WARP_BYTES / 16);
#ifdef SW_INTERIOR_ONLY
    stream_rec_body<false>(a, sa, s, wk, lane, sw, sp);
#else
    if ((wk.w == 0 || sa.force == 1) && sa.force != 2) stream_rec_body<false>(a, sa, s, wk, lane, sw, sp);
    else stream_rec_body<true>(a, sa, s, wk, lane, sw, sp);
#endif
}

// EXPERIMENT (SEPFWI_MERGE=1): the reverse-time step as ONE launch -- reconstruction and adjoint sweep of a step only read the
// adjoint buffer `pa` and write disjoint arrays, so their CTAs can share a grid: the adjoint CTAs (slower items) come first.
__global__ void __launch_bounds__(SW_NT, RC_MINB) k_stream_bwd(const KArgs a, const StreamArgs sr, const StreamArgs sa, const int nAdjCta)
{
    __shared__ __align__(16) float stage[SW_WPB][256];
    extern __shared__ __align__(128) float smem[];
    pdl_launch_dependents();
    const int s = blockIdx.y;
    const Dims &d = a.d;
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    if ((int)blockIdx.x < nAdjCta) {
        if (blockIdx.x == 0) {
            pdl_wait();
            if (threadIdx.x == 0) {
                const float *src = slot_state(a, s) + (size_t)(sa.pa ? S_ADJ1 : S_ADJ) * d.fsz;
                const size_t i = (size_t)a.t.zs[s] * d.ldx + a.t.xs[s];
                a.gstf[(size_t)s * d.nSteps + sa.it] = -(src[(size_t)F_SZZ * d.fsz + i] + a.t.rxz[s] * src[(size_t)F_SXX * d.fsz + i]) * d.dt;
            }
            return;
        }
        const int wg = ((int)blockIdx.x - 1) * SW_WPB + wp;
        if (wg >= sa.nWork) return;
        const int4 wk = __ldg(sa.work + wg);
        const unsigned sw = (unsigned)__cvta_generic_to_shared(smem) + wp * AR_WARP_BYTES;
        const float4 *sp = reinterpret_cast<const float4 *>(smem) + wp * (AR_WARP_BYTES / 16);
        pdl_wait();
        if (wk.w == 0) stream_adj_body<false>(a, sa, s, wk, lane, stage[wp], sw, sp);
        else stream_adj_body<true>(a, sa, s, wk, lane, stage[wp], sw, sp);
    } else {
        const int wg = ((int)blockIdx.x - nAdjCta) * SW_WPB + wp;
        if (wg >= sr.nWork) return;
        const int4 wk = __ldg(sr.work + wg);
        pdl_wait();
        const unsigned sw = (unsigned)__cvta_generic_to_shared(smem) + wp * RC_WARP_BYTES;
        const float4 *sp = reinterpret_cast<const float4 *>(smem) + wp * (RC_WARP_BYTES / 16);
        if (wk.w == 0) stream_rec_body<false>(a, sr, s, wk, lane, sw, sp);
        else stream_rec_body<true>(a, sr, s, wk, lane, sw, sp);
    }
}

// ================================================================================================
// Sponge flavour (DAS_Waveform_Modeling/src/elasticSolver.py:241-276, the Numba CPU propagator's scheme) as ONE launch per time step:
//   phase A (row r)    v  = damp (v + D(sigma) b dt)            update_velocity :310-345 + the sponge :247-248
//   phase B (row r-2)  sigma = damp (sigma + C D(v_new) dt) + stf dt / 2 at the source   update_stress :348-386, :253-260
// the order (velocity first, stresses from the NEW velocities) is the adjoint sweep's march; no CPML memory, the multiplicative
// sponge is one more operand row.  The 2-cell rim never moves (update_* loop over [2, n-2)), so every warp runs the same body
// with an active-cell select.  Traces of sample it-1 (recorded AFTER step it-1, :263-276: pressure halved, strain rates divided by
// the spacing) are written by the leading CTAs from the old state; the host records the last sample after the loop.
constexpr int SP_NARR = 12, SP_NST = 3;
constexpr int SP_WARP_BYTES = SP_NARR * SP_NST * 512;          // 18 KB per warp
constexpr size_t SP_SMEM = (size_t)SW_WPB * SP_WARP_BYTES;     // 72 KB per CTA
enum { SA_SZ = 0, SA_SX, SA_SXZ, SA_OVZ, SA_OVX, SA_BYA, SA_BYB, SA_DV,      // rows r+2 (stresses), r
       SA_LAM, SA_MU, SA_MUA, SA_DS };                                        // row r-2

struct SpCtx {
    const float *g, *m, *damp, *amp;
    float *o;
    const float4 *ring_p;
    unsigned ring_s;
    size_t fsz;
    int ld, nzA, zc0, zc1, zs, xs, xq0;
    unsigned amask;
    bool lown;
    float c1z, c2z, c1x, c2x, dt;
};
struct SpWin { float4 sz[6], sx[6], sxz[6], vz[6], vx[6]; };

__device__ __forceinline__ void stream_sp_issue(const SpCtx &k, const int r, const int stage)
{
    const int ld = k.ld, nzA = k.nzA;
    const size_t fsz = k.fsz;
    auto rowoff = [&](int row) { return (size_t)min(max(row, 0), nzA - 1) * ld; };
    const size_t r2 = rowoff(r + 2), r0 = rowoff(r), rq = rowoff(r - 2);
    const unsigned sb = k.ring_s + (unsigned)stage * (SP_NARR * 512);
    cp16(sb + SA_SZ * 512, k.g + F_SZZ * fsz + r2); cp16(sb + SA_SX * 512, k.g + F_SXX * fsz + r2); cp16(sb + SA_SXZ * 512, k.g + F_SXZ * fsz + r2);
    cp16(sb + SA_OVZ * 512, k.g + F_VZ * fsz + r0); cp16(sb + SA_OVX * 512, k.g + F_VX * fsz + r0);
    cp16(sb + SA_BYA * 512, k.m + M_BYCA * fsz + r0); cp16(sb + SA_BYB * 512, k.m + M_BYCB * fsz + r0); cp16(sb + SA_DV * 512, k.damp + r0);
    cp16(sb + SA_LAM * 512, k.m + M_LAM * fsz + rq); cp16(sb + SA_MU * 512, k.m + M_MU * fsz + rq); cp16(sb + SA_MUA * 512, k.m + M_MUAVE * fsz + rq);
    cp16(sb + SA_DS * 512, k.damp + rq);
    cp_commit();
}

template <int U, int PB>
__device__ __forceinline__ void stream_sp_row(const SpCtx &k, SpWin &w, const int r, const int stage)
{
    const int ld = k.ld, nzA = k.nzA;
    const size_t fsz = k.fsz;
    const float c1z = k.c1z, c2z = k.c2z, c1x = k.c1x, c2x = k.c2x, dt = k.dt;
    constexpr int u = U;
    stream_sp_issue(k, r + (SP_NST - 1), stage == 0 ? SP_NST - 1 : stage - 1);
    cp_wait<SP_NST - 1>();
    const float4 *sb = k.ring_p + stage * (SP_NARR * 32);
    w.sz[(u + 4) % 6] = sb[SA_SZ * 32]; w.sx[(u + 4) % 6] = sb[SA_SX * 32]; w.sxz[(u + 4) % 6] = sb[SA_SXZ * 32];     // row r+2
    // ---- phase A: velocities at row r from the old stresses, rows r-2 .. r+2 (slots u .. u+4)
    {
        const float4 z1 = w.sz[(u + 2) % 6], z2 = w.sz[(u + 3) % 6], z3 = w.sz[(u + 4) % 6], z0 = w.sz[(u + 1) % 6];
        const float4 q0 = w.sxz[u % 6], q1 = w.sxz[(u + 1) % 6], q2 = w.sxz[(u + 2) % 6], q3 = w.sxz[(u + 3) % 6];
        const float4 xc = w.sx[(u + 2) % 6];
        const float wxz[7] = XWIN_B(q2), wxx[7] = XWIN_F(xc);
        const float zzm1[4] = Q4(z0), zzc[4] = Q4(z1), zzp1[4] = Q4(z2), zzp2[4] = Q4(z3);
        const float xzm2[4] = Q4(q0), xzm1[4] = Q4(q1), xzc[4] = Q4(q2), xzp1[4] = Q4(q3);
        const float4 ovz4 = sb[SA_OVZ * 32], ovx4 = sb[SA_OVX * 32], ba4 = sb[SA_BYA * 32], bb4 = sb[SA_BYB * 32], dv4 = sb[SA_DV * 32];
        const float ovz[4] = Q4(ovz4), ovx[4] = Q4(ovx4), ba[4] = Q4(ba4), bb[4] = Q4(bb4), dv[4] = Q4(dv4);
        float nvz[4], nvx[4];
        const bool rowact = (r >= 2 && r <= nzA - 3);
#pragma unroll
        for (int c = 0; c < 4; c++) {
            // dszz/dz forward, dsxz/dx backward -> vz ; dsxz/dz backward, dsxx/dx forward -> vx   (k_velocity_fwd<true>)
            const float dszz_dz = DZ4(zzm1[c], zzc[c], zzp1[c], zzp2[c]);
            const float dsxz_dx = DX7(wxz, c);
            const float dsxz_dz = DZ4(xzm2[c], xzm1[c], xzc[c], xzp1[c]);
            const float dsxx_dx = DX7(wxx, c);
            const bool act = rowact && ((k.amask >> c) & 1u);
            const float tvz = (ovz[c] + (dszz_dz + dsxz_dx) * ba[c] * dt) * dv[c];
            const float tvx = (ovx[c] + (dsxz_dz + dsxx_dx) * bb[c] * dt) * dv[c];
            nvz[c] = act ? tvz : ovz[c]; nvx[c] = act ? tvx : ovx[c];
        }
        const float4 rvz = mk4(nvz), rvx = mk4(nvx);
        w.vz[(u + 4) % 6] = rvz; w.vx[(u + 4) % 6] = rvx;              // new velocity row r-j lives in slot u+4-j
        const bool rown = (r >= k.zc0) && (r < k.zc1);
        if (k.lown && rown) { const size_t ro = (size_t)r * ld; stq(k.o + F_VZ * fsz + ro, rvz); stq(k.o + F_VX * fsz + ro, rvx); }
    }
    // ---- phase B: stresses at row q = r-2 from the new velocities: vz rows q-2..q+1, vx rows q-1..q+2
    if (PB == 0 || (PB == 2 && (r - 2 >= k.zc0) && (r - 2 < k.zc1))) {      // warp-uniform: nothing of this phase is kept for a row the chunk does not own
        const int q = r - 2;
        const float4 v0 = w.vz[u % 6], v1 = w.vz[(u + 1) % 6], v2 = w.vz[(u + 2) % 6], v3 = w.vz[(u + 3) % 6];
        const float4 u0 = w.vx[(u + 1) % 6], u1 = w.vx[(u + 2) % 6], u2 = w.vx[(u + 3) % 6], u3 = w.vx[(u + 4) % 6];
        const float wvz[7] = XWIN_F(v2), wvx[7] = XWIN_B(u1);
        const float vzm2[4] = Q4(v0), vzm1[4] = Q4(v1), vzc[4] = Q4(v2), vzp1[4] = Q4(v3);
        const float vxm1[4] = Q4(u0), vxc[4] = Q4(u1), vxp1[4] = Q4(u2), vxp2[4] = Q4(u3);
        const float ozz[4] = Q4(w.sz[u % 6]), oxx[4] = Q4(w.sx[u % 6]), oxz[4] = Q4(w.sxz[u % 6]);     // old stresses of row r-2: slot u
        const float4 lam4 = sb[SA_LAM * 32], mu4 = sb[SA_MU * 32], mua4 = sb[SA_MUA * 32], ds4 = sb[SA_DS * 32];
        const float l[4] = Q4(lam4), mm[4] = Q4(mu4), ma[4] = Q4(mua4), ds[4] = Q4(ds4);
        float nzz[4], nxz[4], nxx[4];
        const bool qown = (q >= k.zc0) && (q < k.zc1);
        const bool rowact = (q >= 2 && q <= nzA - 3);
#pragma unroll
        for (int c = 0; c < 4; c++) {
            // dvz/dz backward, dvx/dx backward -> szz, sxx ; dvx/dz forward, dvz/dx forward -> sxz   (k_stress_fwd<true>)
            const float dvz_dz = DZ4(vzm2[c], vzm1[c], vzc[c], vzp1[c]);
            const float dvx_dx = DX7(wvx, c);
            const float dvx_dz = DZ4(vxm1[c], vxc[c], vxp1[c], vxp2[c]);
            const float dvz_dx = DX7(wvz, c);
            const bool act = rowact && ((k.amask >> c) & 1u);
            const float l2u = l[c] + 2.0f * mm[c];
            float tzz = (ozz[c] + (l2u * dvz_dz + l[c] * dvx_dx) * dt) * ds[c];
            float txx = (oxx[c] + (l[c] * dvz_dz + l2u * dvx_dx) * dt) * ds[c];
            const float txz = (oxz[c] + ma[c] * (dvx_dz + dvz_dx) * dt) * ds[c];
            if (q == k.zs && k.xq0 + c == k.xs) { const float amp = *k.amp; tzz += amp; txx += amp; }      // after the sponge, :259-260
            nzz[c] = act ? tzz : ozz[c]; nxx[c] = act ? txx : oxx[c]; nxz[c] = act ? txz : oxz[c];
        }
        if (k.lown && qown) {
            const size_t ro = (size_t)q * ld;
            stq(k.o + F_SZZ * fsz + ro, mk4(nzz)); stq(k.o + F_SXZ * fsz + ro, mk4(nxz)); stq(k.o + F_SXX * fsz + ro, mk4(nxx));
        }
    }
}

// traces of sample `samp` from the state in buffer `par` (k_record<true> arithmetic)
__device__ __forceinline__ void stream_sp_aux(const KArgs &a, const StreamArgs &sa, int s, int samp, int par)
{
    const Dims &d = a.d;
    const int ld = d.ldx;
    const float *src = slot_state(a, s) + (size_t)(par ? S_FWD1 : S_FWD) * d.fsz;
    const int nth = sa.nAux * SW_NT, t0 = blockIdx.x * SW_NT + threadIdx.x;
    const int nrec = a.t.nrec[s];
    const size_t cs = (size_t)d.maxRec * d.nSteps;
    const float *vz = src + (size_t)F_VZ * d.fsz, *vx = src + (size_t)F_VX * d.fsz;
    for (int r = t0; r < nrec; r += nth) {
        const int z = a.t.zrec[(size_t)s * d.maxRec + r], x = a.t.xrec[(size_t)s * d.maxRec + r];
        const size_t i = (size_t)z * ld + x;
        float *tr = a.trace + (size_t)s * d.nTrace * cs + (size_t)r * d.nSteps + samp;
        const float exx = (vx[i] - vx[i - 1]) * d.rdx, ezz = (vz[i] - vz[i - ld]) * d.rdz;
        const float exz = 0.5f * ((vx[i + ld] - vx[i]) * d.rdz + (vz[i + 1] - vz[i]) * d.rdx);
        if (sa.mask & (1 << T_PR)) tr[T_PR * cs] = 0.5f * (src[(size_t)F_SXX * d.fsz + i] + src[(size_t)F_SZZ * d.fsz + i]);
        if (sa.mask & (1 << T_VX)) tr[T_VX * cs] = vx[i];
        if (sa.mask & (1 << T_VZ)) tr[T_VZ * cs] = vz[i];
        if (sa.mask & (1 << T_ETT)) {
            const float *w = a.t.w ? a.t.w + ((size_t)s * d.maxRec + r) * 3 : nullptr;
            tr[T_ETT * cs] = w ? w[0] * exx + w[1] * ezz + w[2] * exz : (sa.fiber == 0 ? exx : ezz);
        }
        if (sa.mask & (1 << T_EXX)) tr[T_EXX * cs] = exx;
        if (sa.mask & (1 << T_EZZ)) tr[T_EZZ * cs] = ezz;
        if (sa.mask & (1 << T_EXZ)) tr[T_EXZ * cs] = exz;
    }
}

// grid: x = nAux + ceil(nWork / SW_WPB), y = slot ; dynamic shared memory SP_SMEM ; time step sa.it reads buffer it & 1
__global__ void __launch_bounds__(SW_NT, SW_MINB) k_stream_sponge(const KArgs a, const StreamArgs sa)
{
    extern __shared__ __align__(128) float smem[];
    pdl_launch_dependents();
    const int s = blockIdx.y;
    const int p = sa.it & 1;
    if ((int)blockIdx.x < sa.nAux) { pdl_wait(); if (sa.mask) stream_sp_aux(a, sa, s, sa.it - 1, p); return; }
    const int wg = ((int)blockIdx.x - sa.nAux) * SW_WPB + ((int)threadIdx.x >> 5);
    if (wg >= sa.nWork) return;
    const int4 wk = __ldg(sa.work + wg);
    const int lane = threadIdx.x & 31;
    const Dims &d = a.d;
    SpCtx k;
    k.ld = d.ldx; k.nzA = d.nzA; k.fsz = d.fsz;
    const size_t fsz = d.fsz;
    float *st = slot_state(a, s);
    k.xq0 = wk.x - 4 + 4 * lane;
    const bool colok = (k.xq0 >= 0) && (k.xq0 < d.ldx);
    const int xq = colok ? k.xq0 : 0;
    k.g = st + (size_t)(p ? S_FWD1 : S_FWD) * fsz + xq;
    k.o = st + (size_t)(p ? S_FWD : S_FWD1) * fsz + xq;
    k.m = a.model + xq; k.damp = a.damp + xq;
    k.amp = a.t.amp + (size_t)s * d.nSteps + sa.it;
    k.zc0 = wk.y; k.zc1 = wk.z;
    k.lown = (lane >= 1) && (lane <= 30) && colok;
    k.zs = a.t.zs[s]; k.xs = a.t.xs[s];
    k.c1z = d.c1z; k.c2z = d.c2z; k.c1x = d.c1x; k.c2x = d.c2x; k.dt = d.dt;
    k.ring_s = (unsigned)__cvta_generic_to_shared(smem) + (threadIdx.x >> 5) * SP_WARP_BYTES + lane * 16;
    k.ring_p = reinterpret_cast<const float4 *>(smem) + (threadIdx.x >> 5) * (SP_WARP_BYTES / 16) + lane;
    k.amask = 0;
#pragma unroll
    for (int c = 0; c < 4; c++) { const int x = k.xq0 + c; if (x >= 2 && x <= d.nx - 3) k.amask |= 1u << c; }
    SpWin w;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 6; j++) { w.sz[j] = w.sx[j] = w.sxz[j] = w.vz[j] = w.vx[j] = zero; }
    auto rowoff = [&](int row) { return (size_t)min(max(row, 0), d.nzA - 1) * d.ldx; };
    const int r0 = k.zc0 - 2;
    pdl_wait();
#pragma unroll
    for (int j = 0; j < 4; j++) {       // old stress rows r0-2 .. r0+1 ; row r+2 arrives through the ring
        const size_t ro = rowoff(r0 - 2 + j);
        w.sz[j] = ldq(k.g + F_SZZ * fsz + ro); w.sx[j] = ldq(k.g + F_SXX * fsz + ro); w.sxz[j] = ldq(k.g + F_SXZ * fsz + ro);
    }
#pragma unroll
    for (int j = 0; j < SP_NST - 1; j++) stream_sp_issue(k, r0 + j, j);
    const int niter = (k.zc1 - k.zc0) + 4;
    int stg = 0, kk = 0;
#pragma unroll 1
    for (; kk < 4; kk += 2) {      // lead-in rows: velocities only (their stress row q = r - 2 is not owned)
        stream_sp_row<0, 1>(k, w, r0 + kk, stg); stg = stg == SP_NST - 1 ? 0 : stg + 1;
        stream_sp_row<1, 1>(k, w, r0 + kk + 1, stg); stg = stg == SP_NST - 1 ? 0 : stg + 1;
        win_shift<2>(w.sz); win_shift<2>(w.sx); win_shift<2>(w.sxz); win_shift<2>(w.vz); win_shift<2>(w.vx);
    }
#pragma unroll 1
    for (; kk < niter; kk += 2) {
        stream_sp_row<0, 0>(k, w, r0 + kk, stg); stg = stg == SP_NST - 1 ? 0 : stg + 1;
        stream_sp_row<1, 0>(k, w, r0 + kk + 1, stg); stg = stg == SP_NST - 1 ? 0 : stg + 1;
        win_shift<2>(w.sz); win_shift<2>(w.sx); win_shift<2>(w.sxz); win_shift<2>(w.vz); win_shift<2>(w.vx);
    }
    cp_wait<0>();
}

}  // namespace sepfwi

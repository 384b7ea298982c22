// common.cuh -- device-side data model shared by every kernel of libsepfwi.
//
// Layout in HBM:
//   * every 2-D array is row-major [z][x] with x fastest and a row pitch `ldx`
//     (multiple of 32 floats, rows 128-byte aligned).  Only the nzA = nz - nPad
//     live rows are stored; the reference's nPad alignment rows are never
//     updated (active range ends at nz-nPad-3, el_stress.cu:52) and exist only in
//     the API-side (nz, nx) arrays.
//   * per concurrent shot ("slot") one state block of NSTATE arrays:
//     forward fields, forward CPML memory, adjoint fields, adjoint CPML memory.
//   * the boundary ring store keeps the reference's linear layout bit-exactly:
//     ring[slot][field][it][idx], idx as in utilities.cu:362-392.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sepfwi {

// field order = ring-buffer order of Boundary.cu:30-41
enum { F_SZZ = 0, F_SXZ = 1, F_SXX = 2, F_VZ = 3, F_VX = 4, NFIELD = 5 };
// CPML memory variables; names follow the derivative they damp.
//   forward stress update uses the four P_V*, forward velocity update the four P_S*;
//   the adjoint sweep reuses the same slots for its own memory variables
//   (el_stress_adj.cu / el_velocity_adj.cu use the forward arrays, libCUFD.cu:503-517).
enum { P_VZ_Z = 0, P_VZ_X = 1, P_VX_Z = 2, P_VX_X = 3, P_SZZ_Z = 4, P_SXZ_X = 5, P_SXZ_Z = 6, P_SXX_X = 7, NPSI = 8 };
// State block of one slot (array index; every array has fsz floats):
//   S_FWD / S_FWD1    forward fields, ping-pong pair (the baseline kernels update S_FWD in place)
//   S_FPSI            forward CPML memory: P_V* (stress update) and P_S* (velocity update)
//   S_FPSIV1          second copy of the four P_V* for the streaming forward kernel's out-of-place update (halo recompute reads the old one)
//   S_ADJ / S_ADJ1    adjoint fields, ping-pong pair
//   S_APSI / S_APSI1  adjoint CPML memory, ping-pong pair
enum { S_FWD = 0, S_FWD1 = NFIELD, S_FPSI = 2 * NFIELD, S_FPSIV1 = 2 * NFIELD + NPSI, NSTATE_FWD = 2 * NFIELD + NPSI + 4,
       S_ADJ = NSTATE_FWD, S_ADJ1 = S_ADJ + NFIELD, S_APSI = S_ADJ + 2 * NFIELD, S_APSI1 = S_APSI + NPSI,
       NSTATE = S_APSI1 + NPSI };
enum { M_LAM = 0, M_MU = 1, M_MUAVE = 2, M_BYCA = 3, M_BYCB = 4, NMODEL = 5 };
// 1-D CPML profiles, stored as six rows: 1/K, a, b at integer points then at half points
enum { C_RK = 0, C_A = 1, C_B = 2, C_RKH = 3, C_AH = 4, C_BH = 5, NCOEF = 6 };
// trace components
enum { T_PR = 0, T_VX = 1, T_VZ = 2, T_ETT = 3, T_EXX = 4, T_EZZ = 5, T_EXZ = 6, NTRACE = 7 };
// in gradient mode the pr/vx slots of the trace block hold obs and residual
enum { T_OBS = 0, T_RES = 1 };

struct Dims {
    int nz, nx;        // API grid
    int nzA;           // live rows = nz - nPad
    int ldx;           // row pitch in floats
    int nPml, nPad, nSteps;
    int z1, x1;        // last interior row / column: nzA-1-nPml, nx-1-nPml
    int nzB, nxB;      // ring strip lengths (Boundary.cu:20-21)
    int ringLen;       // 2*5*(nzB+nxB)
    int maxRec;
    int nTrace;        // trace components allocated per slot
    float dt;
    float c1z, c2z, c1x, c2x;   // 9/8/dz, 1/24/dz, 9/8/dx, 1/24/dx
    float rdz, rdx;
    size_t fsz;        // floats per 2-D array = nzA*ldx
    size_t sstride;    // floats between the state blocks of consecutive slots
};

// Per-slot tables (device pointers; slot s at [s] / [s*stride]).
struct SlotTab {
    const int *zs, *xs;        // source cell
    const int *nrec;
    const int *zrec, *xrec;    // [slot][maxRec]
    const float *amp;          // [slot][nSteps] source amplitude per step, already scaled
    const float *rxz;          // [slot]
    const float *w;            // [slot][maxRec][3] ett weights (exx, ezz, exz)
    // deterministic adjoint injection: unique (field, cell) targets, CSR over contributions
    const int *injN;           // [slot] number of targets
    const int *injCell;        // [slot][maxInj]  linear cell index z*ldx+x
    const int *injField;       // [slot][maxInj]  F_VZ or F_VX
    const int *injPtr;         // [slot][maxInj+1]
    const int *injRec;         // [slot][maxCon]
    const float *injCoef;      // [slot][maxCon]
    int maxInj, maxCon;
    // injection targets bucketed for the streaming kernels: per 120-column strip (4-column halo included, so a target
    // may sit in two strips) a row CSR  sInjPtr[slot][strip][nzA+1] into  sInj[slot][2*maxInj] -> target index
    const int *sInjPtr, *sInj;
    int nStrips;
};

struct KArgs {
    Dims d;
    float *state;        // [slot][NSTATE][fsz]
    const float *model;  // [NMODEL][fsz]
    const float *cz;     // [NCOEF][nzA]
    const float *cx;     // [NCOEF][nx]
    const float *cxs;    // [NCOEF][ldx] x profiles for the stress-side CPML, neutral (1/K = 1, a = b = 0) outside x < nPml || x > nx-nPml-1
    const float *cxa;    // [NCOEF][ldx] the plain x profiles at pitch ldx (adjoint sweep; 1/K = 1, a = b = 0 beyond nx)
    const float *cxv;    // [NCOEF][ldx] same for the velocity side: neutral outside x < nPml || x > nx-nPml (el_velocity.cu:56,71)
    const float *damp;   // sponge flavour: [fsz] multiplicative profile (NULL otherwise)
    float *ring;         // [slot][NFIELD][nSteps][ringLen]
    float *trace;        // [slot][nTrace][maxRec*nSteps]
    float *grad;         // [slot][3][fsz]   glam, gmu, grho
    float *gstf;         // [slot][nSteps]
    SlotTab t;
};

__device__ __forceinline__ float *slot_state(const KArgs &a, int s) { return a.state + (size_t)s * a.d.sstride; }

// 4th-order staggered differences; B = backward (x, x-1 | x+1, x-2), F = forward (x+1, x | x+2, x-1)
__device__ __forceinline__ float dzb(const float *f, size_t i, int ld, float c1, float c2)
{ return c1 * (f[i] - f[i - ld]) - c2 * (f[i + ld] - f[i - 2 * ld]); }
__device__ __forceinline__ float dzf(const float *f, size_t i, int ld, float c1, float c2)
{ return c1 * (f[i + ld] - f[i]) - c2 * (f[i + 2 * ld] - f[i - ld]); }
__device__ __forceinline__ float dxb(const float *f, size_t i, float c1, float c2)
{ return c1 * (f[i] - f[i - 1]) - c2 * (f[i + 1] - f[i - 2]); }
__device__ __forceinline__ float dxf(const float *f, size_t i, float c1, float c2)
{ return c1 * (f[i + 1] - f[i]) - c2 * (f[i + 2] - f[i - 1]); }

// Inverse of the reference ring map (utilities.cu:362-392): for cell (z,x) return up to two
// ring indices (side strip and top/bottom strip overlap at the corners).  Returns the count.
__device__ __forceinline__ int ring_indices(const Dims &d, int z, int x, int idx[4])
{
    const int L = 5;
    int n = 0;
    int i = z - (d.nPml - 2), j = x - (d.nPml - 2);
    if (i >= 0 && i < d.nzB) {
        if (j >= 0 && j < L) idx[n++] = j * d.nzB + i;                              // left
        int jr = d.nx - d.nPml + 1 - x;
        if (jr >= 0 && jr < L) idx[n++] = L * d.nzB + jr * d.nzB + i;               // right
    }
    if (j >= 0 && j < d.nxB) {
        if (i >= 0 && i < L) idx[n++] = 2 * L * d.nzB + i * d.nxB + j;              // top
        int ib = d.nzA - d.nPml + 1 - z;
        if (ib >= 0 && ib < L) idx[n++] = L * (2 * d.nzB + d.nxB) + ib * d.nxB + j; // bottom
    }
    return n;
}

// register-only variant: a cell sits in at most one side strip and one top/bottom strip (interior >= 8 cells);
// i0 / i1 = ring index or -1
__device__ __forceinline__ void ring_indices2(const Dims &d, int z, int x, int &i0, int &i1)
{
    const int L = 5;
    i0 = -1; i1 = -1;
    const int i = z - (d.nPml - 2), j = x - (d.nPml - 2);
    if (i >= 0 && i < d.nzB) {
        const int jr = d.nx - d.nPml + 1 - x;
        if (j >= 0 && j < L) i0 = j * d.nzB + i;
        else if (jr >= 0 && jr < L) i0 = L * d.nzB + jr * d.nzB + i;
    }
    if (j >= 0 && j < d.nxB) {
        const int ib = d.nzA - d.nPml + 1 - z;
        if (i >= 0 && i < L) i1 = 2 * L * d.nzB + i * d.nxB + j;
        else if (ib >= 0 && ib < L) i1 = L * (2 * d.nzB + d.nxB) + ib * d.nxB + j;
    }
}

// cell inside the physical domain (the region the reverse-time sweep reconstructs, el_velocity.cu:88 / el_stress.cu:93)
__device__ __forceinline__ bool interior(const Dims &d, int z, int x)
{ return z >= d.nPml && z <= d.z1 && x >= d.nPml && x <= d.x1; }

// does the inclusive cell rectangle [za_, zb_] x [xa_, xb_] touch the 5-wide boundary ring frame?
__device__ __forceinline__ bool tile_touches_ring_ext(const Dims &d, int za_, int zb_, int xa_, int xb_)
{
    const int zlo = d.nPml - 2, zhi = d.z1 + 2, xlo = d.nPml - 2, xhi = d.x1 + 2;
    const int za = max(za_, zlo), zb = min(zb_, zhi), xa = max(xa_, xlo), xb = min(xb_, xhi);
    if (za > zb || xa > xb) return false;                                         // outside the frame's bounding box
    return !(za >= zlo + 5 && zb <= zhi - 5 && xa >= xlo + 5 && xb <= xhi - 5);   // not entirely in the hole
}

}  // namespace sepfwi

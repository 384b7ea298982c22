/*
 * oracle.c -- CPU restatement of the seisfwi/SEP-2023 elastic FWI hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke test and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (sep-2023_b200/csrc, libsepfwi.so) never links, loads or calls anything here.
 *
 * Parity pin: the Numba flavour (ora_numba_forward) is checked against the
 * reference's own DAS_Waveform_Modeling/src/elasticSolver.py imported in the
 * build container (tests/golden/make_numba_golden.py -> tests/golden/numba_*.npz);
 * the TorchFWI flavour is checked against the reference CUDA sources compiled
 * in place into oracle/_ref/ and run on the B200 box
 * (tests/golden/make_cufd_golden.py -> tests/golden/cufd_*.npz).
 *
 * Every function cites the reference file:line it follows.  Abbreviations:
 *   SRC/ = DAS_Waveform_Inversion/Ops/FWI/Src/     MOD/ = DAS_Waveform_Modeling/
 *
 * Array convention of this file: 2-D fields are row-major [z][x] (x fastest),
 * i.e. the layout of the torch tensors the reference API receives.  The
 * reference CUDA code transposes to z-fastest internally (SRC/libCUFD.cu:67-77);
 * the arithmetic does not depend on that, and the one place where a linear
 * layout is part of the contract -- the boundary ring buffer -- is reproduced
 * index-for-index (ora_ring_cell).
 *
 * Arithmetic: the reference kernels are fp32 with a few sub-expressions that C
 * promotes to double (literals 2.0, 1.0, pow()).  `mixed` != 0 reproduces those
 * promotions; `mixed` == 0 evaluates everything in fp32.  Compile with
 * -ffp-contract=off so results do not depend on the host's FMA support.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define ORA_PI 3.141592653589793238462643383279502884197169
#define ORA_MEGA 1e6

typedef struct {
    int nz, nx;        /* padded grid (nz includes the nPad dead rows) */
    int nPml, nPad, nSteps;
    float dz, dx, dt, f0;
    int mixed;         /* 1: reproduce the reference's accidental fp64 sub-expressions */
    int fiber;         /* 0: horizontal fiber, exx  (recording_exx, SRC/utilities.cu:593-602)
                          1: vertical fiber,  ezz  (recording_ezz, SRC/utilities.cu:620-628) */
    float src_rxz;     /* sxx/szz ratio used by source_grad only (SRC/utilities.cu:719-730) */
    int race;          /* emulate the reference's res_injection race (SRC/utilities.cu:613-614,
                          launched <<<(nrec+31)/32, 32>>>): receivers r=32k-1 and r=32k sit in
                          different thread blocks; when they touch the same cell both do a plain
                          load-add-store and one update is lost.  0: race-free (sequential);
                          1: the `+=` of the lower block is lost; 2: the `-=` of the upper block is lost */
} ora_par;

#define F(a, z, x) (a)[(size_t)(z) * nx + (x)]

/* ------------------------------------------------------------------------- */
/* CPML profiles: SRC/utilities.cu:243-359 (cpmlInit), called by SRC/Cpml.cu:46-52
 * with N = nz-nPad for z and N = nx for x.  CpAve is overridden to 3000.     */
void ora_cpml(int N, int nPml, float dh, float f0, float dt,
              float *K, float *a, float *b, float *Kh, float *ah, float *bh)
{
    const float K_MAX = 2.0f;
    const float ALPHA_MAX = (float)(2.0 * ORA_PI * (f0 / 2.0));
    const float NPOWER = 8.0f;
    const float c1 = 0.25f, c2 = 0.75f, c3 = 0.0f;
    const float Rcoef = 0.0008f;
    float thick = nPml * dh;
    float CpAve = 3000.0f;
    float d0 = (float)(-(NPOWER + 1) * CpAve * log(Rcoef) / (2.0 * thick));
    for (int i = 0; i < N; i++) {
        float damp = 0.f, damph = 0.f, alpha = 0.f, alphah = 0.f;
        float dn;
        K[i] = 1.f; Kh[i] = 1.f; a[i] = 0.f; ah[i] = 0.f;
        float depth = (nPml - i) * dh;                       /* left, integer points */
        if (depth >= 0.0f) {
            dn = depth / thick;
            damp = (float)(d0 * (c1 * dn + c2 * pow(dn, NPOWER) + c3 * pow(dn, 2 * NPOWER)));
            K[i] = (float)(1.0 + (K_MAX - 1.0) * pow(dn, NPOWER));
            alpha = (float)(ALPHA_MAX * (1.0 - dn));
        }
        depth = (float)((nPml - i - 0.5) * dh);              /* left, half points */
        if (depth >= 0.0f) {
            dn = depth / thick;
            damph = (float)(d0 * (c1 * dn + c2 * pow(dn, NPOWER) + c3 * pow(dn, 2 * NPOWER)));
            Kh[i] = (float)(1.0 + (K_MAX - 1.0) * pow(dn, NPOWER));
            alphah = (float)(ALPHA_MAX * (1.0 - dn));
        }
        depth = (nPml - N + i) * dh;                         /* right, integer points */
        if (depth >= 0.0f) {
            dn = depth / thick;
            damp = (float)(d0 * (c1 * dn + c2 * pow(dn, NPOWER) + c3 * pow(dn, 2 * NPOWER)));
            K[i] = (float)(1.0 + (K_MAX - 1.0) * pow(dn, NPOWER));
            alpha = (float)(ALPHA_MAX * (1.0 - dn));
        }
        depth = (float)((nPml - N + i + 0.5) * dh);          /* right, half points */
        if (depth >= 0.0f) {
            dn = depth / thick;
            damph = (float)(d0 * (c1 * dn + c2 * pow(dn, NPOWER) + c3 * pow(dn, 2 * NPOWER)));
            Kh[i] = (float)(1.0 + (K_MAX - 1.0) * powf(dn, NPOWER));
            alphah = (float)(ALPHA_MAX * (1.0 - dn));
        }
        if (alpha < 0.f) alpha = 0.f;
        if (alphah < 0.f) alphah = 0.f;
        b[i] = expf(-(damp / K[i] + alpha) * dt);
        bh[i] = expf(-(damph / Kh[i] + alphah) * dt);
        if (fabs(damp) > 1.0e-6)
            a[i] = (float)(damp * (b[i] - 1.0) / (K[i] * (damp + K[i] * alpha)));
        if (fabs(damph) > 1.0e-6)
            ah[i] = (float)(damph * (bh[i] - 1.0) / (Kh[i] * (damph + Kh[i] * alphah)));
    }
}

/* ------------------------------------------------------------------------- */
/* Model preparation: MPa -> Pa (SRC/libCUFD.cu:71-77), Cp (velInit,
 * SRC/utilities.cu:109-123), harmonic 4-point shear modulus (aveMuInit,
 * :125-140), buoyancies (aveBycInit, :142-152).  Defaults outside [2,n-3]:
 * ave_Mu = 0, ave_Byc = 1/1000 (SRC/Model.cu:64-73).                         */
void ora_model_prep(int nz, int nx, const float *lam_mpa, const float *mu_mpa,
                    const float *den, float *lam, float *mu, float *muave,
                    float *byca, float *bycb, float *cp)
{
    size_t n = (size_t)nz * nx;
    for (size_t i = 0; i < n; i++) {
        lam[i] = (float)(lam_mpa[i] * ORA_MEGA);
        mu[i] = (float)(mu_mpa[i] * ORA_MEGA);
        cp[i] = (float)sqrt((lam[i] + 2.0 * mu[i]) / den[i]);
        muave[i] = 0.f;
        byca[i] = (float)(1.0 / 1000.0);
        bycb[i] = (float)(1.0 / 1000.0);
    }
    for (int z = 2; z <= nz - 3; z++)
        for (int x = 2; x <= nx - 3; x++) {
            float a = F(mu, z, x), b = F(mu, z + 1, x), c = F(mu, z, x + 1), d = F(mu, z + 1, x + 1);
            if (a == 0.0f || b == 0.0f || c == 0.0f || d == 0.0f)
                F(muave, z, x) = 0.f;
            else
                F(muave, z, x) = (float)(4.0 / (1.0 / a + 1.0 / b + 1.0 / c + 1.0 / d));
            F(byca, z, x) = (float)(2.0 / (F(den, z + 1, x) + F(den, z, x)));
            F(bycb, z, x) = (float)(2.0 / (F(den, z, x + 1) + F(den, z, x)));
        }
}

/* Courant number, SRC/utilities.cu:225-241.  Returns the number; caller errors if > 1. */
float ora_courant(const float *cp, size_t n, float dt, float dz, float dx)
{
    float mx = cp[0];
    for (size_t i = 0; i < n; i++) if (cp[i] > mx) mx = cp[i];
    float dh = dz < dx ? dz : dx;
    return (float)(mx * dt * sqrtf(2.0f) * (1.0 / 24.0 + 9.0 / 8.0) / dh);
}

/* Source-time-function end taper: cuda_window without weights,
 * SRC/utilities.cu:844-884, called with ratio 0.001 (SRC/Src_Rec.cu:137).    */
void ora_stf_taper(int nt, float dt, float ratio, float *stf)
{
    float t0 = 0.f, t3 = nt * dt;
    float off = nt * dt * ratio;
    if (2.0 * off >= t3 - t0) return;
    float t1 = t0 + off, t2 = t3 - off;
    for (int i = 0; i < nt; i++) {
        float t = i * dt, w;
        if (t >= t0 && t < t1) w = (float)sin(ORA_PI / 2.0 * (t - t0) / (t1 - t0));
        else if (t >= t1 && t < t2) w = 1.0f;
        else if (t >= t2 && t < t3) w = (float)cos(ORA_PI / 2.0 * (t - t2) / (t3 - t2));
        else w = 0.0f;
        stf[i] *= w * w;
    }
}

/* ------------------------------------------------------------------------- */
/* Boundary ring: SRC/Boundary.cu:17-27 (sizes) and SRC/utilities.cu:362-392
 * (index map).  Five layers, two of them inside the PML.                     */
int ora_ring_len(const ora_par *p)
{
    int nzB = p->nz - 2 * p->nPml - p->nPad + 4;
    int nxB = p->nx - 2 * p->nPml + 4;
    return 2 * 5 * (nzB + nxB);
}

void ora_ring_cell(const ora_par *p, int idx, int *z, int *x)
{
    const int L = 5;
    int nPml = p->nPml;
    int nzB = p->nz - 2 * nPml - p->nPad + 4;
    int nxB = p->nx - 2 * nPml + 4;
    if (idx < L * nzB) {                          /* left */
        int j = idx / nzB, i = idx - j * nzB;
        *z = i + nPml - 2; *x = j + nPml - 2;
    } else if (idx < 2 * L * nzB) {               /* right */
        int q = idx - L * nzB;
        int j = q / nzB, i = q - j * nzB;
        *z = i + nPml - 2; *x = p->nx - nPml - j - 1 + 2;
    } else if (idx < L * (2 * nzB + nxB)) {       /* top */
        int q = idx - 2 * L * nzB;
        int i = q / nxB, j = q - i * nxB;
        *z = i + nPml - 2; *x = j + nPml - 2;
    } else {                                      /* bottom */
        int q = idx - L * (2 * nzB + nxB);
        int i = q / nxB, j = q - i * nxB;
        *z = p->nz - nPml - p->nPad - i - 1 + 2; *x = j + nPml - 2;
    }
}

static void ring_save(const ora_par *p, const float *f, float *bnd, int it)
{
    int len = ora_ring_len(p), nx = p->nx;
    for (int idx = 0; idx < len; idx++) {
        int z, x; ora_ring_cell(p, idx, &z, &x);
        bnd[(size_t)it * len + idx] = F(f, z, x);
    }
}
static void ring_restore(const ora_par *p, float *f, const float *bnd, int it)
{
    int len = ora_ring_len(p), nx = p->nx;
    for (int idx = 0; idx < len; idx++) {
        int z, x; ora_ring_cell(p, idx, &z, &x);
        F(f, z, x) = bnd[(size_t)it * len + idx];
    }
}

/* ------------------------------------------------------------------------- */
typedef struct {
    const float *lam, *mu, *muave, *byca, *bycb;           /* Pa, prepared */
    const float *Kz, *az, *bz, *Kzh, *azh, *bzh;           /* length nz-nPad */
    const float *Kx, *ax, *bx, *Kxh, *axh, *bxh;           /* length nx */
} ora_model;

static const float C1 = (float)(9.0 / 8.0);
static const float C2 = (float)(1.0 / 24.0);

#define DZB(f, z, x) ((C1 * (F(f, z, x) - F(f, (z) - 1, x)) - C2 * (F(f, (z) + 1, x) - F(f, (z) - 2, x))) / dz)
#define DZF(f, z, x) ((C1 * (F(f, (z) + 1, x) - F(f, z, x)) - C2 * (F(f, (z) + 2, x) - F(f, (z) - 1, x))) / dz)
#define DXB(f, z, x) ((C1 * (F(f, z, x) - F(f, z, (x) - 1)) - C2 * (F(f, z, (x) + 1) - F(f, z, (x) - 2))) / dx)
#define DXF(f, z, x) ((C1 * (F(f, z, (x) + 1) - F(f, z, x)) - C2 * (F(f, z, (x) + 2) - F(f, z, (x) - 1))) / dx)

/* Forward stress update, SRC/el_stress.cu:50-87. */
static void stress_fwd(const ora_par *p, const ora_model *m, const float *vz, const float *vx,
                       float *szz, float *sxx, float *sxz,
                       float *m_vz_z, float *m_vz_x, float *m_vx_z, float *m_vx_x)
{
    int nz = p->nz, nx = p->nx, nPml = p->nPml, nPad = p->nPad;
    float dz = p->dz, dx = p->dx, dt = p->dt;
#pragma omp parallel for schedule(static)
    for (int z = 2; z <= nz - nPad - 3; z++) {
        int zp = (z < nPml) || (z > nz - nPml - nPad - 1);
        for (int x = 2; x <= nx - 3; x++) {
            int xp = (x < nPml) || (x > nx - nPml - 1);
            float dvz_dz = DZB(vz, z, x);
            float dvx_dx = DXB(vx, z, x);
            if (zp) {
                F(m_vz_z, z, x) = m->bz[z] * F(m_vz_z, z, x) + m->az[z] * dvz_dz;
                dvz_dz = dvz_dz / m->Kz[z] + F(m_vz_z, z, x);
            }
            if (xp) {
                F(m_vx_x, z, x) = m->bx[x] * F(m_vx_x, z, x) + m->ax[x] * dvx_dx;
                dvx_dx = dvx_dx / m->Kx[x] + F(m_vx_x, z, x);
            }
            float l = F(m->lam, z, x), u = F(m->mu, z, x);
            if (p->mixed) {
                F(szz, z, x) = (float)(F(szz, z, x) + ((l + 2.0 * u) * dvz_dz + l * dvx_dx) * dt);
                F(sxx, z, x) = (float)(F(sxx, z, x) + (l * dvz_dz + (l + 2.0 * u) * dvx_dx) * dt);
            } else {
                float l2u = l + 2.0f * u;
                F(szz, z, x) += (l2u * dvz_dz + l * dvx_dx) * dt;
                F(sxx, z, x) += (l * dvz_dz + l2u * dvx_dx) * dt;
            }
            float dvx_dz = DZF(vx, z, x);
            float dvz_dx = DXF(vz, z, x);
            if (zp) {
                F(m_vx_z, z, x) = m->bzh[z] * F(m_vx_z, z, x) + m->azh[z] * dvx_dz;
                dvx_dz = dvx_dz / m->Kzh[z] + F(m_vx_z, z, x);
            }
            if (xp) {
                F(m_vz_x, z, x) = m->bxh[x] * F(m_vz_x, z, x) + m->axh[x] * dvz_dx;
                dvz_dx = dvz_dx / m->Kxh[x] + F(m_vz_x, z, x);
            }
            F(sxz, z, x) += F(m->muave, z, x) * (dvx_dz + dvz_dx) * dt;
        }
    }
}

/* Forward velocity update, SRC/el_velocity.cu:45-82 (x-strip test is x > nx-nPml). */
static void velocity_fwd(const ora_par *p, const ora_model *m, float *vz, float *vx,
                         const float *szz, const float *sxx, const float *sxz,
                         float *m_szz_z, float *m_sxz_x, float *m_sxz_z, float *m_sxx_x)
{
    int nz = p->nz, nx = p->nx, nPml = p->nPml, nPad = p->nPad;
    float dz = p->dz, dx = p->dx, dt = p->dt;
#pragma omp parallel for schedule(static)
    for (int z = 2; z <= nz - nPad - 3; z++) {
        int zp = (z < nPml) || (z > nz - nPml - nPad - 1);
        for (int x = 2; x <= nx - 3; x++) {
            int xp = (x < nPml) || (x > nx - nPml);
            float dszz_dz = DZF(szz, z, x);
            float dsxz_dx = DXB(sxz, z, x);
            if (zp) {
                F(m_szz_z, z, x) = m->bzh[z] * F(m_szz_z, z, x) + m->azh[z] * dszz_dz;
                dszz_dz = dszz_dz / m->Kzh[z] + F(m_szz_z, z, x);
            }
            if (xp) {
                F(m_sxz_x, z, x) = m->bx[x] * F(m_sxz_x, z, x) + m->ax[x] * dsxz_dx;
                dsxz_dx = dsxz_dx / m->Kx[x] + F(m_sxz_x, z, x);
            }
            F(vz, z, x) += (dszz_dz + dsxz_dx) * F(m->byca, z, x) * dt;
            float dsxz_dz = DZB(sxz, z, x);
            float dsxx_dx = DXF(sxx, z, x);
            if (zp) {
                F(m_sxz_z, z, x) = m->bz[z] * F(m_sxz_z, z, x) + m->az[z] * dsxz_dz;
                dsxz_dz = dsxz_dz / m->Kz[z] + F(m_sxz_z, z, x);
            }
            if (xp) {
                F(m_sxx_x, z, x) = m->bxh[x] * F(m_sxx_x, z, x) + m->axh[x] * dsxx_dx;
                dsxx_dx = dsxx_dx / m->Kxh[x] + F(m_sxx_x, z, x);
            }
            F(vx, z, x) += (dsxz_dz + dsxx_dx) * F(m->bycb, z, x) * dt;
        }
    }
}

/* Reverse-time velocity reconstruction + density imaging, SRC/el_velocity.cu:84-117.
 * The spray is written sequentially here (the reference uses atomicAdd); the
 * x+1 bounds check of the reference is a chained comparison that is always true
 * (SRC/el_velocity.cu:109), so that spray is unconditional.                  */
static void velocity_bwd(const ora_par *p, const ora_model *m, float *vz, float *vx,
                         const float *szz, const float *sxx, const float *sxz,
                         const float *vz_adj, const float *vx_adj, float *gden)
{
    int nz = p->nz, nx = p->nx, nPml = p->nPml, nPad = p->nPad;
    float dz = p->dz, dx = p->dx, dt = p->dt;
    int z1 = nz - nPad - 1 - nPml, x1 = nx - 1 - nPml;
    for (int z = nPml; z <= z1; z++)
        for (int x = nPml; x <= x1; x++) {
            float dszz_dz = DZF(szz, z, x);
            float dsxz_dx = DXB(sxz, z, x);
            float ba = F(m->byca, z, x), bb = F(m->bycb, z, x);
            F(vz, z, x) -= (dszz_dz + dsxz_dx) * ba * dt;
            float dsxz_dz = DZB(sxz, z, x);
            float dsxx_dx = DXF(sxx, z, x);
            F(vx, z, x) -= (dsxz_dz + dsxx_dx) * bb * dt;
            float ga, gb;
            if (p->mixed) {
                ga = (float)(-F(vz_adj, z, x) * (dszz_dz + dsxz_dx) * dt * (-pow(ba, 2) / 2.0));
                gb = (float)(-F(vx_adj, z, x) * (dsxz_dz + dsxx_dx) * dt * (-pow(bb, 2) / 2.0));
            } else {
                ga = -F(vz_adj, z, x) * (dszz_dz + dsxz_dx) * dt * (-(ba * ba) / 2.0f);
                gb = -F(vx_adj, z, x) * (dsxz_dz + dsxx_dx) * dt * (-(bb * bb) / 2.0f);
            }
            F(gden, z, x) += ga;
            F(gden, z, x) += gb;
            if (z + 1 <= z1) F(gden, z + 1, x) += ga;
            F(gden, z, x + 1) += gb;
        }
}

/* Reverse-time stress reconstruction + lambda/mu imaging, SRC/el_stress.cu:89-128. */
static void stress_bwd(const ora_par *p, const ora_model *m, const float *vz, const float *vx,
                       float *szz, float *sxx, float *sxz,
                       const float *szz_adj, const float *sxx_adj, const float *sxz_adj,
                       float *glam, float *gmu)
{
    int nz = p->nz, nx = p->nx, nPml = p->nPml, nPad = p->nPad;
    float dz = p->dz, dx = p->dx, dt = p->dt;
    int z1 = nz - nPad - 1 - nPml, x1 = nx - 1 - nPml;
    for (int z = nPml; z <= z1; z++)
        for (int x = nPml; x <= x1; x++) {
            float dvz_dz = DZB(vz, z, x);
            float dvx_dx = DXB(vx, z, x);
            float l = F(m->lam, z, x), u = F(m->mu, z, x), ua = F(m->muave, z, x);
            float dvx_dz = DZF(vx, z, x);
            float dvz_dx = DXF(vz, z, x);
            if (p->mixed) {
                F(szz, z, x) = (float)(F(szz, z, x) - ((l + 2.0 * u) * dvz_dz + l * dvx_dx) * dt);
                F(sxx, z, x) = (float)(F(sxx, z, x) - (l * dvz_dz + (l + 2.0 * u) * dvx_dx) * dt);
            } else {
                float l2u = l + 2.0f * u;
                F(szz, z, x) -= (l2u * dvz_dz + l * dvx_dx) * dt;
                F(sxx, z, x) -= (l * dvz_dz + l2u * dvx_dx) * dt;
            }
            F(sxz, z, x) -= ua * (dvx_dz + dvz_dx) * dt;
            float sz = F(szz_adj, z, x), sx = F(sxx_adj, z, x);
            if (p->mixed) {
                F(glam, z, x) = (float)(F(glam, z, x) + -(sz + sx) * (dvz_dz + dvx_dx) * dt * ORA_MEGA);
                F(gmu, z, x) = (float)(F(gmu, z, x) + (-2.0 * sz * dvz_dz * dt - 2.0 * sx * dvx_dx * dt) * ORA_MEGA);
            } else {
                F(glam, z, x) += -(sz + sx) * (dvz_dz + dvx_dx) * dt * 1e6f;
                F(gmu, z, x) += (-2.0f * sz * dvz_dz * dt - 2.0f * sx * dvx_dx * dt) * 1e6f;
            }
            if (ua != 0.0f) {
                float m00 = F(m->mu, z, x), m10 = F(m->mu, z + 1, x);
                float m01 = F(m->mu, z, x + 1), m11 = F(m->mu, z + 1, x + 1);
                if (p->mixed) {
                    float scale = (float)(-F(sxz_adj, z, x) * (dvx_dz + dvz_dx) * dt * ua /
                                          (1.0 / m00 + 1.0 / m10 + 1.0 / m01 + 1.0 / m11) * ORA_MEGA);
                    F(gmu, z, x) += (float)(1.0 / pow(m00, 2) * scale);
                    if (z + 1 <= z1) F(gmu, z + 1, x) += (float)(1.0 / pow(m10, 2) * scale);
                    F(gmu, z, x + 1) += (float)(1.0 / pow(m01, 2) * scale);
                    if (z + 1 <= z1 && x + 1 <= x1) F(gmu, z + 1, x + 1) += (float)(1.0 / pow(m11, 2) * scale);
                } else {
                    float scale = -F(sxz_adj, z, x) * (dvx_dz + dvz_dx) * dt * ua /
                                  (1.0f / m00 + 1.0f / m10 + 1.0f / m01 + 1.0f / m11) * 1e6f;
                    F(gmu, z, x) += 1.0f / (m00 * m00) * scale;
                    if (z + 1 <= z1) F(gmu, z + 1, x) += 1.0f / (m10 * m10) * scale;
                    F(gmu, z, x + 1) += 1.0f / (m01 * m01) * scale;
                    if (z + 1 <= z1 && x + 1 <= x1) F(gmu, z + 1, x + 1) += 1.0f / (m11 * m11) * scale;
                }
            }
        }
}

/* Adjoint derivative stencils carry the sign pattern -c1 ... + c2. */
#define AZB(f, z, x) ((-C1 * (F(f, z, x) - F(f, (z) - 1, x)) + C2 * (F(f, (z) + 1, x) - F(f, (z) - 2, x))) / dz)
#define AZF(f, z, x) ((-C1 * (F(f, (z) + 1, x) - F(f, z, x)) + C2 * (F(f, (z) + 2, x) - F(f, (z) - 1, x))) / dz)
#define AXB(f, z, x) ((-C1 * (F(f, z, x) - F(f, z, (x) - 1)) + C2 * (F(f, z, (x) + 1) - F(f, z, (x) - 2))) / dx)
#define AXF(f, z, x) ((-C1 * (F(f, z, (x) + 1) - F(f, z, x)) + C2 * (F(f, z, (x) + 2) - F(f, z, (x) - 1))) / dx)

/* Adjoint velocity update, SRC/el_velocity_adj.cu:55-103.  Memory arrays named
 * after the reference: ms_* = d_mem_ds*_d*, mv_* = d_mem_dv*_d*.             */
static void velocity_adj(const ora_par *p, const ora_model *m, float *vz, float *vx,
                         const float *szz, const float *sxx, const float *sxz,
                         float *ms_zz_z, float *ms_xz_x, float *ms_xz_z, float *ms_xx_x,
                         const float *mv_z_z, const float *mv_z_x, const float *mv_x_z, const float *mv_x_x)
{
    int nz = p->nz, nx = p->nx, nPml = p->nPml, nPad = p->nPad;
    float dz = p->dz, dx = p->dx, dt = p->dt;
#pragma omp parallel for schedule(static)
    for (int z = 2; z <= nz - nPad - 3; z++) {
        int zp = (z < nPml) || (z > nz - nPml - nPad - 1);
        for (int x = 2; x <= nx - 3; x++) {
            int xp = (x < nPml) || (x > nx - nPml - 1);
            float l = F(m->lam, z, x), u = F(m->mu, z, x), ua = F(m->muave, z, x);
            float dpsixx_dx = AXF(mv_x_x, z, x);
            float dszz_dx = AXF(szz, z, x);
            float dsxx_dx = AXF(sxx, z, x);
            float dpsixz_dz = AZB(mv_x_z, z, x);
            float dsxz_dz = AZB(sxz, z, x);
            if (p->mixed)
                F(vx, z, x) = (float)(F(vx, z, x) + (m->ax[x] * dpsixx_dx + l * dszz_dx / m->Kx[x] * dt +
                                                      (l + 2.0 * u) * dsxx_dx / m->Kx[x] * dt +
                                                      m->azh[z] * dpsixz_dz + ua / m->Kzh[z] * dsxz_dz * dt));
            else
                F(vx, z, x) += (m->ax[x] * dpsixx_dx + l * dszz_dx / m->Kx[x] * dt +
                                (l + 2.0f * u) * dsxx_dx / m->Kx[x] * dt +
                                m->azh[z] * dpsixz_dz + ua / m->Kzh[z] * dsxz_dz * dt);
            float bb = F(m->bycb, z, x), ba = F(m->byca, z, x);
            if (xp) F(ms_xx_x, z, x) = m->bxh[x] * F(ms_xx_x, z, x) + bb * F(vx, z, x) * dt;
            if (zp) F(ms_xz_z, z, x) = m->bz[z] * F(ms_xz_z, z, x) + bb * F(vx, z, x) * dt;

            float dpsizz_dz = AZF(mv_z_z, z, x);
            float dszz_dz = AZF(szz, z, x);
            float dsxx_dz = AZF(sxx, z, x);
            float dpsizx_dx = AXB(mv_z_x, z, x);
            float dsxz_dx = AXB(sxz, z, x);
            if (p->mixed)
                F(vz, z, x) = (float)(F(vz, z, x) + (m->az[z] * dpsizz_dz + (l + 2.0 * u) * dszz_dz / m->Kz[z] * dt +
                                                      l * dsxx_dz / m->Kz[z] * dt + m->axh[x] * dpsizx_dx +
                                                      ua / m->Kxh[x] * dsxz_dx * dt));
            else
                F(vz, z, x) += (m->az[z] * dpsizz_dz + (l + 2.0f * u) * dszz_dz / m->Kz[z] * dt +
                                l * dsxx_dz / m->Kz[z] * dt + m->axh[x] * dpsizx_dx +
                                ua / m->Kxh[x] * dsxz_dx * dt);
            if (xp) F(ms_xz_x, z, x) = m->bx[x] * F(ms_xz_x, z, x) + ba * F(vz, z, x) * dt;
            if (zp) F(ms_zz_z, z, x) = m->bzh[z] * F(ms_zz_z, z, x) + ba * F(vz, z, x) * dt;
        }
    }
}

/* Adjoint stress update, SRC/el_stress_adj.cu:50-97 (memory updates unconditional). */
static void stress_adj(const ora_par *p, const ora_model *m, const float *vz, const float *vx,
                       float *szz, float *sxx, float *sxz,
                       const float *ms_zz_z, const float *ms_xz_x, const float *ms_xz_z, const float *ms_xx_x,
                       float *mv_z_z, float *mv_z_x, float *mv_x_z, float *mv_x_x)
{
    int nz = p->nz, nx = p->nx, nPad = p->nPad;
    float dz = p->dz, dx = p->dx, dt = p->dt;
#pragma omp parallel for schedule(static)
    for (int z = 2; z <= nz - nPad - 3; z++)
        for (int x = 2; x <= nx - 3; x++) {
            float l = F(m->lam, z, x), u = F(m->mu, z, x), ua = F(m->muave, z, x);
            float ba = F(m->byca, z, x), bb = F(m->bycb, z, x);
            float dphi_xz_x_dx = AXF(ms_xz_x, z, x);
            float dvz_dx = AXF(vz, z, x);
            float dphi_xz_z_dz = AZF(ms_xz_z, z, x);
            float dvx_dz = AZF(vx, z, x);
            F(sxz, z, x) += m->ax[x] * dphi_xz_x_dx + dvz_dx / m->Kx[x] * ba * dt +
                            m->az[z] * dphi_xz_z_dz + dvx_dz / m->Kz[z] * bb * dt;
            F(mv_z_x, z, x) = m->bxh[x] * F(mv_z_x, z, x) + F(sxz, z, x) * ua * dt;
            F(mv_x_z, z, x) = m->bzh[z] * F(mv_x_z, z, x) + F(sxz, z, x) * ua * dt;

            float dphi_xx_x_dx = AXB(ms_xx_x, z, x);
            float dvx_dx = AXB(vx, z, x);
            float dphi_zz_z_dz = AZB(ms_zz_z, z, x);
            float dvz_dz = AZB(vz, z, x);
            F(sxx, z, x) += m->axh[x] * dphi_xx_x_dx + bb * dvx_dx / m->Kxh[x] * dt;
            F(szz, z, x) += m->azh[z] * dphi_zz_z_dz + ba * dvz_dz / m->Kzh[z] * dt;
            if (p->mixed) {
                F(mv_x_x, z, x) = (float)(m->bx[x] * F(mv_x_x, z, x) + l * F(szz, z, x) * dt +
                                          (l + 2.0 * u) * F(sxx, z, x) * dt);
                F(mv_z_z, z, x) = (float)(m->bz[z] * F(mv_z_z, z, x) + (l + 2.0 * u) * F(szz, z, x) * dt +
                                          l * F(sxx, z, x) * dt);
            } else {
                float l2u = l + 2.0f * u;
                F(mv_x_x, z, x) = m->bx[x] * F(mv_x_x, z, x) + l * F(szz, z, x) * dt + l2u * F(sxx, z, x) * dt;
                F(mv_z_z, z, x) = m->bz[z] * F(mv_z_z, z, x) + l2u * F(szz, z, x) * dt + l * F(sxx, z, x) * dt;
            }
        }
}

/* ------------------------------------------------------------------------- */
typedef struct {
    float *vz, *vx, *szz, *sxx, *sxz;
    float *mv_z_z, *mv_z_x, *mv_x_z, *mv_x_x;   /* d_mem_dv*_d* */
    float *ms_zz_z, *ms_xz_x, *ms_xz_z, *ms_xx_x; /* d_mem_ds*_d* */
} ora_state;

static float *zalloc(size_t n) { return (float *)calloc(n, sizeof(float)); }

/* One shot of the forward time loop, SRC/libCUFD.cu:268-332 (A.3 of SURVEY.md).
 * stf must already be tapered.  zs/xs/zrec/xrec are padded-grid indices
 * (interior index + nPml, SRC/Src_Rec.cu:87-115).  Traces are [rec][nSteps].
 * bnd (optional) = 5 buffers of len*nSteps in the order szz,sxz,sxx,vz,vx
 * (SRC/Boundary.cu:30-41).  state_out (optional) receives the final 5 fields.  */
static void forward_loop(const ora_par *p, const ora_model *m, const float *stf,
                         int zs, int xs, int nrec, const int *zrec, const int *xrec,
                         float *d_pr, float *d_vx, float *d_vz, float *d_ett,
                         float *bnd[5], ora_state *s)
{
    int nx = p->nx, nSteps = p->nSteps;
    float dt = p->dt;
    float scale = p->mixed ? (float)pow(1500.0, 2) : 1500.0f * 1500.0f;
    for (int it = 0; it <= nSteps - 2; it++) {
        if (bnd) {
            ring_save(p, s->szz, bnd[0], it);
            ring_save(p, s->sxz, bnd[1], it);
            ring_save(p, s->sxx, bnd[2], it);
            ring_save(p, s->vz, bnd[3], it);
            ring_save(p, s->vx, bnd[4], it);
        }
        stress_fwd(p, m, s->vz, s->vx, s->szz, s->sxx, s->sxz, s->mv_z_z, s->mv_z_x, s->mv_x_z, s->mv_x_x);
        F(s->szz, zs, xs) += scale * stf[it] * dt;       /* add_source, SRC/utilities.cu:524-552 */
        F(s->sxx, zs, xs) += scale * stf[it] * dt;
        velocity_fwd(p, m, s->vz, s->vx, s->szz, s->sxx, s->sxz, s->ms_zz_z, s->ms_xz_x, s->ms_xz_z, s->ms_xx_x);
        for (int r = 0; r < nrec; r++) {                 /* recording*, SRC/utilities.cu:593-703 */
            int z = zrec[r], x = xrec[r];
            size_t o = (size_t)r * nSteps + it + 1;
            if (d_pr) d_pr[o] = F(s->szz, z, x) + F(s->sxx, z, x);
            if (d_vx) d_vx[o] = F(s->vx, z, x);
            if (d_vz) d_vz[o] = F(s->vz, z, x);
            if (d_ett) d_ett[o] = p->fiber == 0 ? F(s->vx, z, x) - F(s->vx, z, x - 1)
                                                : F(s->vz, z, x) - F(s->vz, z - 1, x);
        }
    }
}

static void state_alloc(ora_state *s, size_t n)
{
    float **f = (float **)s;
    for (int i = 0; i < 13; i++) f[i] = zalloc(n);
}
static void state_free(ora_state *s)
{
    float **f = (float **)s;
    for (int i = 0; i < 13; i++) free(f[i]);
}

typedef struct {
    const float *lam_mpa, *mu_mpa, *den;   /* [nz][nx] as the Python op receives them */
} ora_inputs;

/* Build everything derived from the model (Model + Cpml constructors). */
typedef struct {
    float *lam, *mu, *muave, *byca, *bycb, *cp;
    float *cz[6], *cx[6];
    ora_model m;
} ora_prepared;

static void prepare(const ora_par *p, const float *lam_mpa, const float *mu_mpa, const float *den, ora_prepared *q)
{
    size_t n = (size_t)p->nz * p->nx;
    q->lam = zalloc(n); q->mu = zalloc(n); q->muave = zalloc(n);
    q->byca = zalloc(n); q->bycb = zalloc(n); q->cp = zalloc(n);
    ora_model_prep(p->nz, p->nx, lam_mpa, mu_mpa, den, q->lam, q->mu, q->muave, q->byca, q->bycb, q->cp);
    int nza = p->nz - p->nPad;
    for (int i = 0; i < 6; i++) { q->cz[i] = zalloc(nza); q->cx[i] = zalloc(p->nx); }
    ora_cpml(nza, p->nPml, p->dz, p->f0, p->dt, q->cz[0], q->cz[1], q->cz[2], q->cz[3], q->cz[4], q->cz[5]);
    ora_cpml(p->nx, p->nPml, p->dx, p->f0, p->dt, q->cx[0], q->cx[1], q->cx[2], q->cx[3], q->cx[4], q->cx[5]);
    ora_model *m = &q->m;
    m->lam = q->lam; m->mu = q->mu; m->muave = q->muave; m->byca = q->byca; m->bycb = q->bycb;
    m->Kz = q->cz[0]; m->az = q->cz[1]; m->bz = q->cz[2]; m->Kzh = q->cz[3]; m->azh = q->cz[4]; m->bzh = q->cz[5];
    m->Kx = q->cx[0]; m->ax = q->cx[1]; m->bx = q->cx[2]; m->Kxh = q->cx[3]; m->axh = q->cx[4]; m->bxh = q->cx[5];
}
static void prepared_free(ora_prepared *q)
{
    free(q->lam); free(q->mu); free(q->muave); free(q->byca); free(q->bycb); free(q->cp);
    for (int i = 0; i < 6; i++) { free(q->cz[i]); free(q->cx[i]); }
}

/* Public: Courant number of a model given in MPa (compCourantNumber on h_Cp). */
float ora_courant_model(const ora_par *p, const float *lam_mpa, const float *mu_mpa, const float *den)
{
    ora_prepared q; prepare(p, lam_mpa, mu_mpa, den, &q);
    float c = ora_courant(q.cp, (size_t)p->nz * p->nx, p->dt, p->dz, p->dx);
    prepared_free(&q);
    return c;
}

/* Public: forward modelling of one shot (calc_id = 2 path).  stf_raw is the
 * untapered row of the stf matrix; the taper of SRC/Src_Rec.cu:137 is applied
 * here.  Any of the four outputs may be NULL.  bnd5 (optional) is one buffer of
 * 5*len*nSteps floats; fields_out (optional) 5*nz*nx floats (szz,sxz,sxx,vz,vx). */
void ora_forward(const ora_par *p, const float *lam_mpa, const float *mu_mpa, const float *den,
                 const float *stf_raw, int zs, int xs, int nrec, const int *zrec, const int *xrec,
                 float *d_pr, float *d_vx, float *d_vz, float *d_ett, float *bnd5, float *fields_out)
{
    size_t n = (size_t)p->nz * p->nx;
    ora_prepared q; prepare(p, lam_mpa, mu_mpa, den, &q);
    float *stf = (float *)malloc(sizeof(float) * p->nSteps);
    memcpy(stf, stf_raw, sizeof(float) * p->nSteps);
    ora_stf_taper(p->nSteps, p->dt, 0.001f, stf);
    ora_state s; state_alloc(&s, n);
    float *outs[4] = {d_pr, d_vx, d_vz, d_ett};
    for (int k = 0; k < 4; k++)
        if (outs[k]) memset(outs[k], 0, sizeof(float) * (size_t)nrec * p->nSteps);
    float *bnd[5];
    if (bnd5) {
        size_t l = (size_t)ora_ring_len(p) * p->nSteps;
        for (int k = 0; k < 5; k++) bnd[k] = bnd5 + k * l;
    }
    forward_loop(p, &q.m, stf, zs, xs, nrec, zrec, xrec, d_pr, d_vx, d_vz, d_ett, bnd5 ? bnd : NULL, &s);
    if (fields_out) {
        memcpy(fields_out + 0 * n, s.szz, n * sizeof(float));
        memcpy(fields_out + 1 * n, s.sxz, n * sizeof(float));
        memcpy(fields_out + 2 * n, s.sxx, n * sizeof(float));
        memcpy(fields_out + 3 * n, s.vz, n * sizeof(float));
        memcpy(fields_out + 4 * n, s.vx, n * sizeof(float));
    }
    state_free(&s); free(stf); prepared_free(&q);
}

/* Public: misfit and gradient of one shot, accumulated INTO glam/gmu/gden
 * (the reference accumulates over the shots of one cufd call,
 * SRC/Model.cu:74-78 + SRC/libCUFD.cu:170-708) and written to gstf[nSteps].
 * obs_ett is [nrec][nSteps].  Returns sum(res_ett^2) (without the 1/2 of
 * SRC/libCUFD.cu:776).  syn_ett / res_ett (optional) receive the traces.
 * with_adj = 0 gives the calc_id = 0 path (misfit only).                      */
double ora_gradient(const ora_par *p, const float *lam_mpa, const float *mu_mpa, const float *den,
                    const float *stf_raw, int zs, int xs, int nrec, const int *zrec, const int *xrec,
                    const float *obs_ett, int with_adj,
                    float *glam, float *gmu, float *gden, float *gstf, float *syn_ett, float *res_ett)
{
    int nx = p->nx, nSteps = p->nSteps;
    size_t n = (size_t)p->nz * p->nx;
    float dt = p->dt;
    ora_prepared q; prepare(p, lam_mpa, mu_mpa, den, &q);
    float *stf = (float *)malloc(sizeof(float) * nSteps);
    memcpy(stf, stf_raw, sizeof(float) * nSteps);
    ora_stf_taper(nSteps, dt, 0.001f, stf);
    ora_state s; state_alloc(&s, n);
    size_t nd = (size_t)nrec * nSteps;
    float *ett = zalloc(nd), *res = zalloc(nd);
    float *bnd[5] = {0, 0, 0, 0, 0};
    if (with_adj) {
        size_t l = (size_t)ora_ring_len(p) * nSteps;
        for (int k = 0; k < 5; k++) bnd[k] = zalloc(l);
    }
    forward_loop(p, &q.m, stf, zs, xs, nrec, zrec, xrec, NULL, NULL, NULL, ett, with_adj ? bnd : NULL, &s);

    /* residual and misfit: gpuMinus + cuda_cal_objective, SRC/utilities.cu:154-205 */
    double J = 0.0;
    for (int r = 0; r < nrec; r++)
        for (int it = 0; it < nSteps; it++) {
            size_t o = (size_t)r * nSteps + it;
            res[o] = it > 0 ? obs_ett[o] - ett[o] : 0.0f;
            J += (double)res[o] * res[o];
        }
    if (syn_ett) memcpy(syn_ett, ett, nd * sizeof(float));
    if (res_ett) memcpy(res_ett, res, nd * sizeof(float));

    if (with_adj) {
        /* backward loop, SRC/libCUFD.cu:500-653 (A.6 of SURVEY.md).  The adjoint
         * state and all eight memory arrays restart from zero (:503-517); the
         * pre-loop adjoint launches on the zero state (:520-542) have no effect. */
        ora_state a; state_alloc(&a, n);
        float scale = p->mixed ? (float)pow(1500.0, 2) : 1500.0f * 1500.0f;
        for (int it = nSteps - 2; it >= 0; it--) {
            /* source_grad, SRC/utilities.cu:719-730 */
            gstf[it] = p->mixed ? (float)(-(F(a.szz, zs, xs) + (double)p->src_rxz * F(a.sxx, zs, xs)) * dt)
                                : -(F(a.szz, zs, xs) + p->src_rxz * F(a.sxx, zs, xs)) * dt;
            velocity_bwd(p, &q.m, s.vz, s.vx, s.szz, s.sxx, s.sxz, a.vz, a.vx, gden);
            ring_restore(p, s.vz, bnd[3], it);
            ring_restore(p, s.vx, bnd[4], it);
            F(s.szz, zs, xs) -= scale * stf[it] * dt;
            F(s.sxx, zs, xs) -= scale * stf[it] * dt;
            stress_bwd(p, &q.m, s.vz, s.vx, s.szz, s.sxx, s.sxz, a.szz, a.sxx, a.sxz, glam, gmu);
            ring_restore(p, s.szz, bnd[0], it);
            ring_restore(p, s.sxz, bnd[1], it);
            ring_restore(p, s.sxx, bnd[2], it);
            velocity_adj(p, &q.m, a.vz, a.vx, a.szz, a.sxx, a.sxz, a.ms_zz_z, a.ms_xz_x, a.ms_xz_z, a.ms_xx_x,
                         a.mv_z_z, a.mv_z_x, a.mv_x_z, a.mv_x_x);
            /* res_injection_exx / _ezz, SRC/utilities.cu:605-641 (sequential = race-free) */
            for (int r = 0; r < nrec; r++) {
                int z = zrec[r], x = xrec[r];
                float rv = res[(size_t)r * nSteps + it];
                /* does the first receiver of the next block subtract at the cell this one adds to? */
                int seam_hi = p->race && (r % 32 == 31) && r + 1 < nrec &&
                              (p->fiber == 0 ? (zrec[r + 1] == z && xrec[r + 1] - 1 == x) : (xrec[r + 1] == x && zrec[r + 1] - 1 == z));
                int seam_lo = p->race && (r % 32 == 0) && r > 0 &&
                              (p->fiber == 0 ? (zrec[r - 1] == z && xrec[r - 1] == x - 1) : (xrec[r - 1] == x && zrec[r - 1] == z - 1));
                int skip_add = seam_hi && p->race == 1, skip_sub = seam_lo && p->race == 2;
                if (p->fiber == 0) { if (!skip_add) F(a.vx, z, x) += rv; if (!skip_sub) F(a.vx, z, x - 1) -= rv; }
                else               { if (!skip_add) F(a.vz, z, x) += rv; if (!skip_sub) F(a.vz, z - 1, x) -= rv; }
            }
            stress_adj(p, &q.m, a.vz, a.vx, a.szz, a.sxx, a.sxz, a.ms_zz_z, a.ms_xz_x, a.ms_xz_z, a.ms_xx_x,
                       a.mv_z_z, a.mv_z_x, a.mv_x_z, a.mv_x_x);
        }
        gstf[nSteps - 1] = 0.0f;
        state_free(&a);
        for (int k = 0; k < 5; k++) free(bnd[k]);
    }
    state_free(&s); free(stf); free(ett); free(res); prepared_free(&q);
    return J;
}

/* ------------------------------------------------------------------------- */
/* Numba flavour: MOD/src/elasticSolver.py:185-386 (forward_it, update_velocity,
 * update_stress), fp64, arrays (nx,nz) C-order => [x][z], z fastest.  All
 * inputs already padded by ndamp (np.pad 'edge', :37-47); sx/sz, gx/gz, dxg/dzg
 * are padded-grid indices.  Outputs are [n][nt].  sens is [ndas][6].          */
#define G(a, i, j) (a)[(size_t)(i) * nz + (j)]
void ora_numba_forward(int nx, int nz, double dx, double dz, double dt, int nt,
                       const double *lam, const double *mu, const double *rho, const double *damp,
                       const double *stf, int sx, int sz,
                       int ngeo, const int *gx, const int *gz,
                       int ndas, const int *dxg, const int *dzg, const double *sens,
                       double *geoVx, double *geoVz, double *geoPr,
                       double *dasExx, double *dasEzz, double *dasExz, double *dasEtt)
{
    size_t n = (size_t)nx * nz;
    double *vx = (double *)calloc(n, 8), *vz = (double *)calloc(n, 8);
    double *sxx = (double *)calloc(n, 8), *szz = (double *)calloc(n, 8), *sxz = (double *)calloc(n, 8);
    const double c1 = 9.0 / 8.0, c2 = 1.0 / 24.0;
    for (int it = 0; it < nt; it++) {
        /* update_velocity, MOD/src/elasticSolver.py:310-345, then damping :247-248 */
#pragma omp parallel for schedule(static)
        for (int i = 2; i < nx - 2; i++)
            for (int j = 2; j < nz - 2; j++) {
                double rhox = 0.5 * (G(rho, i, j) + G(rho, i + 1, j));
                double rhoz = 0.5 * (G(rho, i, j) + G(rho, i, j + 1));
                double szz_z = (c1 * (G(szz, i, j + 1) - G(szz, i, j)) - c2 * (G(szz, i, j + 2) - G(szz, i, j - 1))) / dz;
                double sxz_x = (c1 * (G(sxz, i, j) - G(sxz, i - 1, j)) - c2 * (G(sxz, i + 1, j) - G(sxz, i - 2, j))) / dx;
                double sxz_z = (c1 * (G(sxz, i, j) - G(sxz, i, j - 1)) - c2 * (G(sxz, i, j + 1) - G(sxz, i, j - 2))) / dz;
                double sxx_x = (c1 * (G(sxx, i + 1, j) - G(sxx, i, j)) - c2 * (G(sxx, i + 2, j) - G(sxx, i - 1, j))) / dx;
                G(vx, i, j) += (sxz_z + sxx_x) * dt / rhoz;
                G(vz, i, j) += (szz_z + sxz_x) * dt / rhox;
            }
#pragma omp parallel for schedule(static)
        for (size_t k = 0; k < n; k++) { vx[k] *= damp[k]; vz[k] *= damp[k]; }
        /* update_stress, MOD/src/elasticSolver.py:348-386, then damping :254-256 */
#pragma omp parallel for schedule(static)
        for (int i = 2; i < nx - 2; i++)
            for (int j = 2; j < nz - 2; j++) {
                double muxz = 0.0;
                if (G(mu, i, j) != 0.0 && G(mu, i + 1, j) != 0.0 && G(mu, i, j + 1) != 0.0 && G(mu, i + 1, j + 1) != 0.0)
                    muxz = 4.0 / (1 / G(mu, i, j) + 1 / G(mu, i + 1, j) + 1 / G(mu, i, j + 1) + 1 / G(mu, i + 1, j + 1));
                double vzz = (c1 * (G(vz, i, j) - G(vz, i, j - 1)) - c2 * (G(vz, i, j + 1) - G(vz, i, j - 2))) / dz;
                double vxx = (c1 * (G(vx, i, j) - G(vx, i - 1, j)) - c2 * (G(vx, i + 1, j) - G(vx, i - 2, j))) / dx;
                double vxz = (c1 * (G(vx, i, j + 1) - G(vx, i, j)) - c2 * (G(vx, i, j + 2) - G(vx, i, j - 1))) / dz;
                double vzx = (c1 * (G(vz, i + 1, j) - G(vz, i, j)) - c2 * (G(vz, i + 2, j) - G(vz, i - 1, j))) / dx;
                double l = G(lam, i, j), u = G(mu, i, j);
                G(szz, i, j) += ((l + 2 * u) * vzz + l * vxx) * dt;
                G(sxx, i, j) += (l * vzz + (l + 2 * u) * vxx) * dt;
                G(sxz, i, j) += (vxz + vzx) * muxz * dt;
            }
#pragma omp parallel for schedule(static)
        for (size_t k = 0; k < n; k++) { sxx[k] *= damp[k]; szz[k] *= damp[k]; sxz[k] *= damp[k]; }
        /* explosive source :259-260 */
        G(sxx, sx, sz) += stf[it] * dt / 2.0;
        G(szz, sx, sz) += stf[it] * dt / 2.0;
        /* geophones :263-266, DAS :269-276 */
        for (int r = 0; r < ngeo; r++) {
            size_t o = (size_t)r * nt + it;
            geoVx[o] = G(vx, gx[r], gz[r]);
            geoVz[o] = G(vz, gx[r], gz[r]);
            geoPr[o] = (G(sxx, gx[r], gz[r]) + G(szz, gx[r], gz[r])) * 0.5;
        }
        for (int r = 0; r < ndas; r++) {
            size_t o = (size_t)r * nt + it;
            int i = dxg[r], j = dzg[r];
            dasExx[o] = (G(vx, i, j) - G(vx, i - 1, j)) / dx;
            dasEzz[o] = (G(vz, i, j) - G(vz, i, j - 1)) / dz;
            dasExz[o] = 0.5 * ((G(vx, i, j + 1) - G(vx, i, j)) / dz + (G(vz, i + 1, j) - G(vz, i, j)) / dx);
            dasEtt[o] = sens[r * 6 + 0] * dasExx[o] + sens[r * 6 + 3] * dasEzz[o] + sens[r * 6 + 1] * dasExz[o];
        }
    }
    free(vx); free(vz); free(sxx); free(szz); free(sxz);
}

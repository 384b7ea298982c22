// exp_stream.cu -- design experiment: register-streaming forward step (interior cells only).
// A warp owns a 128-column strip (30 owned quads + 1 halo quad each side) and marches down z,
// keeping the z-windows of the stencil in registers and fetching x-neighbours with warp shuffles.
// Compares against a naive two-kernel step and times both.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o /tmp/exp_stream tools/exp_stream.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

struct P {
    int nzA, nx, ldx, Lz, nStrips, nChunks;
    size_t fsz;
    float dt, c1z, c2z, c1x, c2x;
    const float *src;   // [5][fsz]  szz sxz sxx vz vx
    float *dst;
    const float *model; // [5][fsz] lam mu mua bya byb
};
enum { F_SZZ = 0, F_SXZ = 1, F_SXX = 2, F_VZ = 3, F_VX = 4 };

// ---------------- naive reference
__global__ void k_stress(P p, const float *s, float *o)
{
    int x = blockIdx.x * 32 + threadIdx.x, z = blockIdx.y * 8 + threadIdx.y;
    if (z < 2 || z > p.nzA - 3 || x < 2 || x > p.nx - 3) return;
    size_t i = (size_t)z * p.ldx + x; int ld = p.ldx;
    const float *vz = s + F_VZ * p.fsz, *vx = s + F_VX * p.fsz;
    float dvz_dz = p.c1z * (vz[i] - vz[i - ld]) - p.c2z * (vz[i + ld] - vz[i - 2 * ld]);
    float dvx_dx = p.c1x * (vx[i] - vx[i - 1]) - p.c2x * (vx[i + 1] - vx[i - 2]);
    float dvx_dz = p.c1z * (vx[i + ld] - vx[i]) - p.c2z * (vx[i + 2 * ld] - vx[i - ld]);
    float dvz_dx = p.c1x * (vz[i + 1] - vz[i]) - p.c2x * (vz[i + 2] - vz[i - 1]);
    float lam = p.model[i], mu = p.model[p.fsz + i], mua = p.model[2 * p.fsz + i];
    float l2u = lam + 2.0f * mu;
    o[F_SZZ * p.fsz + i] = s[F_SZZ * p.fsz + i] + (l2u * dvz_dz + lam * dvx_dx) * p.dt;
    o[F_SXX * p.fsz + i] = s[F_SXX * p.fsz + i] + (lam * dvz_dz + l2u * dvx_dx) * p.dt;
    o[F_SXZ * p.fsz + i] = s[F_SXZ * p.fsz + i] + mua * (dvx_dz + dvz_dx) * p.dt;
}
__global__ void k_velocity(P p, const float *s, float *o)
{
    int x = blockIdx.x * 32 + threadIdx.x, z = blockIdx.y * 8 + threadIdx.y;
    if (z < 4 || z > p.nzA - 5 || x < 4 || x > p.nx - 5) return;
    size_t i = (size_t)z * p.ldx + x; int ld = p.ldx;
    const float *szz = o + F_SZZ * p.fsz, *sxz = o + F_SXZ * p.fsz, *sxx = o + F_SXX * p.fsz;
    float dszz_dz = p.c1z * (szz[i + ld] - szz[i]) - p.c2z * (szz[i + 2 * ld] - szz[i - ld]);
    float dsxz_dx = p.c1x * (sxz[i] - sxz[i - 1]) - p.c2x * (sxz[i + 1] - sxz[i - 2]);
    float dsxz_dz = p.c1z * (sxz[i] - sxz[i - ld]) - p.c2z * (sxz[i + ld] - sxz[i - 2 * ld]);
    float dsxx_dx = p.c1x * (sxx[i + 1] - sxx[i]) - p.c2x * (sxx[i + 2] - sxx[i - 1]);
    o[F_VZ * p.fsz + i] = s[F_VZ * p.fsz + i] + (dszz_dz + dsxz_dx) * p.model[3 * p.fsz + i] * p.dt;
    o[F_VX * p.fsz + i] = s[F_VX * p.fsz + i] + (dsxz_dz + dsxx_dx) * p.model[4 * p.fsz + i] * p.dt;
}

// ---------------- streaming kernel
#define Q4(v) {v.x, v.y, v.z, v.w}
__device__ __forceinline__ float4 ldq(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ void stq(float *p, const float4 &v) { *reinterpret_cast<float4 *>(p) = v; }
__device__ __forceinline__ float shl_(float v) { return __shfl_up_sync(0xffffffffu, v, 1); }    // value of lane-1
__device__ __forceinline__ float shr_(float v) { return __shfl_down_sync(0xffffffffu, v, 1); }  // value of lane+1

constexpr int OWN = 120;
#ifndef WPB
#define WPB 4
#endif
#ifndef MINB
#define MINB 3
#endif

__global__ void __launch_bounds__(WPB * 32, MINB) k_stream(const P p)
{
    const int lane = threadIdx.x & 31;
    const int wg = min(blockIdx.x * WPB + (threadIdx.x >> 5), p.nStrips * p.nChunks - 1);   // duplicates recompute the last chunk
    const int strip = wg % p.nStrips, chunk = wg / p.nStrips;
    const int xq = 4 + strip * OWN - 4 + 4 * lane;            // interior experiment: first owned column is 4
    const int zc0 = 4 + chunk * p.Lz, zc1 = min(zc0 + p.Lz, p.nzA - 4);
    const bool colok = (xq >= 0) && (xq + 3 < p.ldx);
    const bool lown = (lane >= 1) && (lane <= 30) && (xq + 3 <= p.nx - 5);
    const int ld = p.ldx;
    const size_t fsz = p.fsz;
    const float *g = p.src + (colok ? xq : 0);
    const float *m = p.model + (colok ? xq : 0);
    float *o = p.dst + (colok ? xq : 0);
    const float c1z = p.c1z, c2z = p.c2z, c1x = p.c1x, c2x = p.c2x, dt = p.dt;

    float4 vz[6], vx[6], zz[6], xz[6], xx[6];
    float4 ozz[2], oxz[2], oxx[2], lam[2], mu[2], mua[2], bya[2], byb[2];
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 6; j++) { vz[j] = vx[j] = zz[j] = xz[j] = xx[j] = zero; }
    // preload v rows zc0-4 .. zc0 (slots 0..4) ; row rho <-> slot (rho - (zc0-4)) % 6
#pragma unroll
    for (int j = 0; j < 5; j++) {
        const size_t ro = (size_t)(zc0 - 4 + j) * ld;
        vz[j] = ldq(g + F_VZ * fsz + ro); vx[j] = ldq(g + F_VX * fsz + ro);
    }
    {
        const size_t ro = (size_t)(zc0 - 2) * ld;
        ozz[0] = ldq(g + F_SZZ * fsz + ro); oxz[0] = ldq(g + F_SXZ * fsz + ro); oxx[0] = ldq(g + F_SXX * fsz + ro);
        lam[0] = ldq(m + 0 * fsz + ro); mu[0] = ldq(m + 1 * fsz + ro); mua[0] = ldq(m + 2 * fsz + ro);
        bya[0] = zero; byb[0] = zero;
    }
    const int niter = (zc1 - zc0) + 4;
#pragma unroll 1
    for (int kk = 0; kk < niter; kk += 6) {
#pragma unroll
        for (int u = 0; u < 6; u++) {
            const int k = kk + u;
            {
                const int r = zc0 - 2 + k;            // stress row of this iteration
                const int cb = u & 1, nb = cb ^ 1;
                // ---- prefetch for the next iteration
                {
                    const int rn = min(r + 3, p.nzA - 1);
                    const size_t ro = (size_t)rn * ld;
                    vz[(u + 5) % 6] = ldq(g + F_VZ * fsz + ro); vx[(u + 5) % 6] = ldq(g + F_VX * fsz + ro);
                    const size_t r1 = (size_t)min(r + 1, p.nzA - 1) * ld;
                    ozz[nb] = ldq(g + F_SZZ * fsz + r1); oxz[nb] = ldq(g + F_SXZ * fsz + r1); oxx[nb] = ldq(g + F_SXX * fsz + r1);
                    lam[nb] = ldq(m + 0 * fsz + r1); mu[nb] = ldq(m + 1 * fsz + r1); mua[nb] = ldq(m + 2 * fsz + r1);
                    const size_t rq = (size_t)max(r - 1, 0) * ld;
                    bya[nb] = ldq(m + 3 * fsz + rq); byb[nb] = ldq(m + 4 * fsz + rq);
                }
                // ---- stress at row r : v rows r-2..r+2 <-> slots u .. u+4
                {
                    const float4 a0 = vz[u % 6], a1 = vz[(u + 1) % 6], a2 = vz[(u + 2) % 6], a3 = vz[(u + 3) % 6];
                    const float4 b0 = vx[(u + 1) % 6], b1 = vx[(u + 2) % 6], b2 = vx[(u + 3) % 6], b3 = vx[(u + 4) % 6];
                    const float vxc[7] = {shl_(b1.z), shl_(b1.w), b1.x, b1.y, b1.z, b1.w, shr_(b1.x)};   // vx[x-2..x+4]
                    const float vzc[7] = {shl_(a2.w), a2.x, a2.y, a2.z, a2.w, shr_(a2.x), shr_(a2.y)};   // vz[x-1..x+5]
                    const float vzm2[4] = Q4(a0), vzm1[4] = Q4(a1), vzq[4] = Q4(a2), vzp1[4] = Q4(a3);
                    const float vxm1[4] = Q4(b0), vxq[4] = Q4(b1), vxp1[4] = Q4(b2), vxp2[4] = Q4(b3);
                    const float l[4] = Q4(lam[cb]), mm[4] = Q4(mu[cb]), ma[4] = Q4(mua[cb]);
                    const float pzz[4] = Q4(ozz[cb]), pxz[4] = Q4(oxz[cb]), pxx[4] = Q4(oxx[cb]);
                    float nzz[4], nxz[4], nxx[4];
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        const float dvz_dz = c1z * (vzq[c] - vzm1[c]) - c2z * (vzp1[c] - vzm2[c]);
                        const float dvx_dx = c1x * (vxc[c + 2] - vxc[c + 1]) - c2x * (vxc[c + 3] - vxc[c]);
                        const float dvx_dz = c1z * (vxp1[c] - vxq[c]) - c2z * (vxp2[c] - vxm1[c]);
                        const float dvz_dx = c1x * (vzc[c + 2] - vzc[c + 1]) - c2x * (vzc[c + 3] - vzc[c]);
                        const float l2u = l[c] + 2.0f * mm[c];
                        nzz[c] = pzz[c] + (l2u * dvz_dz + l[c] * dvx_dx) * dt;
                        nxx[c] = pxx[c] + (l[c] * dvz_dz + l2u * dvx_dx) * dt;
                        nxz[c] = pxz[c] + ma[c] * (dvx_dz + dvz_dx) * dt;
                    }
                    const float4 rzz = make_float4(nzz[0], nzz[1], nzz[2], nzz[3]), rxz = make_float4(nxz[0], nxz[1], nxz[2], nxz[3]),
                                 rxx = make_float4(nxx[0], nxx[1], nxx[2], nxx[3]);
                    zz[u % 6] = rzz; xz[u % 6] = rxz; xx[u % 6] = rxx;      // stress row r <-> slot u
                    if (lown && r >= zc0 && r < zc1) {
                        const size_t ro = (size_t)r * ld;
                        stq(o + F_SZZ * fsz + ro, rzz); stq(o + F_SXZ * fsz + ro, rxz); stq(o + F_SXX * fsz + ro, rxx);
                    }
                }
                // ---- velocity at row q = r-2 : szz rows q-1..q+2 = r-3..r ; sxz rows q-2..q+1 = r-4..r-1 ; sxx row q
                {
                    const int q = r - 2;
                    const float4 p0 = zz[(u + 3) % 6], p1 = zz[(u + 4) % 6], p2 = zz[(u + 5) % 6], p3 = zz[u % 6];
                    const float4 q0 = xz[(u + 2) % 6], q1 = xz[(u + 3) % 6], q2 = xz[(u + 4) % 6], q3 = xz[(u + 5) % 6];
                    const float4 xc = xx[(u + 4) % 6];
                    const float xzc[7] = {shl_(q2.z), shl_(q2.w), q2.x, q2.y, q2.z, q2.w, shr_(q2.x)};
                    const float xxc[7] = {shl_(xc.w), xc.x, xc.y, xc.z, xc.w, shr_(xc.x), shr_(xc.y)};
                    const float zzm1[4] = Q4(p0), zzc[4] = Q4(p1), zzp1[4] = Q4(p2), zzp2[4] = Q4(p3);
                    const float xzm2[4] = Q4(q0), xzm1[4] = Q4(q1), xzp1[4] = Q4(q3);
                    const float ovz[4] = Q4(vz[u % 6]), ovx[4] = Q4(vx[u % 6]);       // v row r-2 <-> slot u
                    const float ba[4] = Q4(bya[cb]), bb[4] = Q4(byb[cb]);
                    float nvz[4], nvx[4];
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        const float dszz_dz = c1z * (zzp1[c] - zzc[c]) - c2z * (zzp2[c] - zzm1[c]);
                        const float dsxz_dz = c1z * (xzc[c + 2] - xzm1[c]) - c2z * (xzp1[c] - xzm2[c]);
                        const float dsxz_dx = c1x * (xzc[c + 2] - xzc[c + 1]) - c2x * (xzc[c + 3] - xzc[c]);
                        const float dsxx_dx = c1x * (xxc[c + 2] - xxc[c + 1]) - c2x * (xxc[c + 3] - xxc[c]);
                        nvz[c] = ovz[c] + (dszz_dz + dsxz_dx) * ba[c] * dt;
                        nvx[c] = ovx[c] + (dsxz_dz + dsxx_dx) * bb[c] * dt;
                    }
                    if (lown && q >= zc0 && q < zc1) {
                        const size_t ro = (size_t)q * ld;
                        stq(o + F_VZ * fsz + ro, make_float4(nvz[0], nvz[1], nvz[2], nvz[3]));
                        stq(o + F_VX * fsz + ro, make_float4(nvx[0], nvx[1], nvx[2], nvx[3]));
                    }
                }
            }
        }
    }
}

int main(int argc, char **argv)
{
    int nzA = argc > 1 ? atoi(argv[1]) : 2080, nx = argc > 2 ? atoi(argv[2]) : 8064, Lz = argc > 3 ? atoi(argv[3]) : 40;
    int reps = 50;
    P p;
    p.nzA = nzA; p.nx = nx; p.ldx = (nx + 31) / 32 * 32; p.fsz = (size_t)nzA * p.ldx; p.Lz = Lz;
    p.dt = 1e-3f; p.c1z = 9.f / 8 / 10; p.c2z = 1.f / 24 / 10; p.c1x = p.c1z; p.c2x = p.c2z;
    p.nStrips = (nx - 8 + OWN - 1) / OWN; p.nChunks = (nzA - 8 + Lz - 1) / Lz;
    std::vector<float> h(5 * p.fsz), hm(5 * p.fsz);
    srand(1);
    for (auto &v : h) v = (rand() / (float)RAND_MAX - 0.5f);
    for (size_t i = 0; i < p.fsz; i++) {
        hm[i] = 1e3f * (1 + rand() / (float)RAND_MAX); hm[p.fsz + i] = 5e2f * (1 + rand() / (float)RAND_MAX); hm[2 * p.fsz + i] = hm[p.fsz + i];
        hm[3 * p.fsz + i] = 1e-3f * (1 + rand() / (float)RAND_MAX); hm[4 * p.fsz + i] = hm[3 * p.fsz + i];
    }
    float *src, *d1, *d2, *model;
    CK(cudaMalloc(&src, 5 * p.fsz * 4)); CK(cudaMalloc(&d1, 5 * p.fsz * 4)); CK(cudaMalloc(&d2, 5 * p.fsz * 4)); CK(cudaMalloc(&model, 5 * p.fsz * 4));
    CK(cudaMemcpy(src, h.data(), 5 * p.fsz * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(model, hm.data(), 5 * p.fsz * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(d1, 0, 5 * p.fsz * 4)); CK(cudaMemset(d2, 0, 5 * p.fsz * 4));
    p.src = src; p.model = model;
    dim3 nb((nx + 31) / 32, (nzA + 7) / 8), nt(32, 8);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float ms;
    // naive
    for (int w = 0; w < 3; w++) { k_stress<<<nb, nt>>>(p, src, d1); k_velocity<<<nb, nt>>>(p, src, d1); }
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; i++) { k_stress<<<nb, nt>>>(p, src, d1); k_velocity<<<nb, nt>>>(p, src, d1); }
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    const double cells = (double)nzA * nx;
    printf("grid %d x %d: naive 2 kernels %.2f us/step -> %.0f GB/s (52 B/cell)\n", nzA, nx, 1e3 * ms / reps, 52 * cells / (ms / reps * 1e-3) / 1e9);
    // streaming
    p.dst = d2;
    const int nw = p.nStrips * p.nChunks, nblk = (nw + WPB - 1) / WPB;
    for (int w = 0; w < 3; w++) k_stream<<<nblk, WPB * 32>>>(p);
    CK(cudaGetLastError());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; i++) k_stream<<<nblk, WPB * 32>>>(p);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("   streaming Lz=%d (%d warps, %d CTAs): %.2f us/step -> %.0f GB/s (52 B/cell), frac of 6451 = %.3f\n", Lz, nw, nblk, 1e3 * ms / reps,
           52 * cells / (ms / reps * 1e-3) / 1e9, 52 * cells / (ms / reps * 1e-3) / 1e9 / 6451.2);
    // compare on the velocity region (z,x in [4, n-5])
    std::vector<float> r1(5 * p.fsz), r2(5 * p.fsz);
    CK(cudaMemcpy(r1.data(), d1, 5 * p.fsz * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(r2.data(), d2, 5 * p.fsz * 4, cudaMemcpyDeviceToHost));
    double maxd = 0, maxv = 0; size_t nbad = 0;
    for (int f = 0; f < 5; f++)
        for (int z = 4; z <= nzA - 5; z++)
            for (int x = 4; x <= nx - 5; x++) {
                size_t i = f * p.fsz + (size_t)z * p.ldx + x;
                double dd = fabs((double)r1[i] - r2[i]);
                if (dd > maxd) maxd = dd;
                if (fabs(r1[i]) > maxv) maxv = fabs(r1[i]);
                if (dd > 1e-4 * (1 + fabs(r1[i]))) nbad++;
            }
    printf("   max |diff| = %.3e (max |ref| = %.3e), mismatches %zu\n", maxd, maxv, nbad);
    return 0;
}

// dmma_probe.cu -- does the FP64 tensor path (mma.sync ... f64, SASS DMMA) run beside the DFMA pipe on B200?
// Times (a) DFMA only, (b) DMMA only (m8n8k4, m16n8k8, m16n8k16), (c) both interleaved in one warp,
// (d) DFMA in half the warps and DMMA in the other half.  Prints TFLOP/s of each pipe in each mix.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/dmma_probe tools/probes/dmma_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ void dmma884(double (&d)[2], double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&d)[4], const double (&a)[4], const double (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma1684(double (&d)[4], const double (&a)[2], double b)
{
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}

// mode 0: DFMA only (NF chains); 1: DMMA m8n8k4 only (NM chains); 2: interleaved; 3: split by warp parity;
// 4: m16n8k8 only; 5: m16n8k4 only; 6: m16n8k4 interleaved with DFMA
template <int MODE, int NF, int NM>
__global__ void __launch_bounds__(256) probe(int iters, double seed, double* out)
{
    double f[NF];
    double m[NM][4];
    double a4[4] = {seed, seed * 0.5, seed * 0.25, seed * 0.125}, b2[2] = {seed, -seed};
    double a2[2] = {seed, seed * 0.5};
#pragma unroll
    for (int k = 0; k < NF; ++k) f[k] = seed + k + threadIdx.x;
#pragma unroll
    for (int k = 0; k < NM; ++k) { m[k][0] = seed + k; m[k][1] = seed - k; m[k][2] = seed * k; m[k][3] = 1.0; }
    const double x = seed * 1e-3, y = 1.0 - 1e-9;
    const bool wdf = (MODE != 3) || ((threadIdx.x >> 5) & 1) == 0;
    const bool wmm = (MODE != 3) || ((threadIdx.x >> 5) & 1) == 1;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 2 || MODE == 6 || (MODE == 3 && wdf)) {
#pragma unroll
            for (int k = 0; k < NF; ++k) f[k] = fma(f[k], y, x);
        }
        if (MODE == 1 || MODE == 2 || (MODE == 3 && wmm)) {
#pragma unroll
            for (int k = 0; k < NM; ++k) { double (&d)[2] = reinterpret_cast<double (&)[2]>(m[k]); dmma884(d, x, y); }
        }
        if (MODE == 4) {
#pragma unroll
            for (int k = 0; k < NM; ++k) dmma1688(m[k], a4, b2);
        }
        if (MODE == 5 || MODE == 6) {
#pragma unroll
            for (int k = 0; k < NM; ++k) dmma1684(m[k], a2, y);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < NF; ++k) s += f[k];
#pragma unroll
    for (int k = 0; k < NM; ++k) s += m[k][0] + m[k][1] + m[k][2] + m[k][3];
    if (s == 12345.678) out[0] = s;
}

template <int MODE, int NF, int NM>
void run(const char* name, double fma_per_thread_iter, double mma_flop_per_warp_iter)
{
    int dev = 0, sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double* out; cudaMalloc(&out, 8);
    const int iters = 20000, blocks = sms * 4, threads = 256;
    probe<MODE, NF, NM><<<blocks, threads>>>(100, 1.0, out);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<MODE, NF, NM><<<blocks, threads>>>(iters, 1.0, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double nthreads = (double)blocks * threads, nwarps = nthreads / 32;
    double dfma = 0, dmma = 0;
    if (MODE == 0 || MODE == 2 || MODE == 6) dfma = nthreads * iters * fma_per_thread_iter * 2;
    if (MODE == 3) dfma = nthreads / 2 * iters * fma_per_thread_iter * 2;
    if (MODE == 1 || MODE == 2 || MODE == 4 || MODE == 5 || MODE == 6) dmma = nwarps * iters * mma_flop_per_warp_iter;
    if (MODE == 3) dmma = nwarps / 2 * iters * mma_flop_per_warp_iter;
    printf("%-44s %8.3f ms  DFMA %6.2f TF/s  DMMA %6.2f TF/s  sum %6.2f  (%s)\n", name, ms, dfma / ms / 1e9, dmma / ms / 1e9,
           (dfma + dmma) / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

// One m8n8k4 on random data: which summation order does the hardware use?
__global__ void layout_kernel(const double* A, const double* B, const double* Cin, double* D)
{
    const int lane = threadIdx.x;
    double d[2] = {Cin[(lane >> 2) * 8 + (lane & 3) * 2], Cin[(lane >> 2) * 8 + (lane & 3) * 2 + 1]};
    dmma884(d, A[(lane >> 2) * 4 + (lane & 3)], B[(lane & 3) * 8 + (lane >> 2)]);      // A row-major 8x4, B[k][n] 4x8
    D[(lane >> 2) * 8 + (lane & 3) * 2] = d[0];
    D[(lane >> 2) * 8 + (lane & 3) * 2 + 1] = d[1];
}
#include <cmath>
static void layout_check()
{
    double A[32], B[32], C[64], D[64], *dA, *dB, *dC, *dD;
    srand(7);
    auto rnd = []() { return (rand() / (double)RAND_MAX - 0.5) * 2.0; };
    int match_fwd = 0, match_rev = 0, match_other = 0;
    cudaMalloc(&dA, sizeof A); cudaMalloc(&dB, sizeof B); cudaMalloc(&dC, sizeof C); cudaMalloc(&dD, sizeof D);
    for (int trial = 0; trial < 50; ++trial) {
        for (double& v : A) v = rnd();
        for (double& v : B) v = rnd();
        for (double& v : C) v = rnd();
        cudaMemcpy(dA, A, sizeof A, cudaMemcpyHostToDevice); cudaMemcpy(dB, B, sizeof B, cudaMemcpyHostToDevice);
        cudaMemcpy(dC, C, sizeof C, cudaMemcpyHostToDevice);
        layout_kernel<<<1, 32>>>(dA, dB, dC, dD);
        cudaMemcpy(D, dD, sizeof D, cudaMemcpyDeviceToHost);
        for (int m = 0; m < 8; ++m)
            for (int n = 0; n < 8; ++n) {
                double f = C[m * 8 + n], r = C[m * 8 + n];
                for (int k = 0; k < 4; ++k) f = fma(A[m * 4 + k], B[k * 8 + n], f);
                for (int k = 3; k >= 0; --k) r = fma(A[m * 4 + k], B[k * 8 + n], r);
                if (D[m * 8 + n] == f) ++match_fwd;
                else if (D[m * 8 + n] == r) ++match_rev;
                else ++match_other;
            }
    }
    printf("m8n8k4 layout/order check over 3200 outputs: == fma chain k=0..3 from c: %d, == k=3..0: %d, neither: %d\n",
           match_fwd, match_rev, match_other);
}

int main()
{
    layout_check();
    run<0, 8, 1>("DFMA only, 8 chains", 8, 0);
    run<1, 1, 8>("DMMA m8n8k4 only, 8 chains", 0, 8 * 2.0 * 8 * 8 * 4);
    run<1, 1, 4>("DMMA m8n8k4 only, 4 chains", 0, 4 * 2.0 * 8 * 8 * 4);
    run<4, 1, 4>("DMMA m16n8k8 only, 4 chains", 0, 4 * 2.0 * 16 * 8 * 8);
    run<5, 1, 4>("DMMA m16n8k4 only, 4 chains", 0, 4 * 2.0 * 16 * 8 * 4);
    run<2, 8, 8>("interleaved 8 DFMA + 8 DMMA884 per iter", 8, 8 * 2.0 * 8 * 8 * 4);
    run<2, 8, 1>("interleaved 8 DFMA + 1 DMMA884 per iter", 8, 1 * 2.0 * 8 * 8 * 4);
    run<2, 8, 2>("interleaved 8 DFMA + 2 DMMA884 per iter", 8, 2 * 2.0 * 8 * 8 * 4);
    run<6, 8, 1>("interleaved 8 DFMA + 1 DMMA16x8x4 per iter", 8, 1 * 2.0 * 16 * 8 * 4);
    run<6, 8, 2>("interleaved 8 DFMA + 2 DMMA16x8x4 per iter", 8, 2 * 2.0 * 16 * 8 * 4);
    run<3, 8, 8>("warp-split: even warps DFMA, odd warps DMMA884", 8, 8 * 2.0 * 8 * 8 * 4);
    return 0;
}

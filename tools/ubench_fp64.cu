// Micro-benchmark of the fp64 pipe on one SM sub-partition (B200): dependent-issue latency of DFMA / DADD /
// DSETP+FSEL chains and the throughput of N independent DFMA chains, for 1..4 warps per sub-partition.
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench_fp64 tools/ubench_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void dfma_chains(double *out, const double *in, int iters, long long *cycles)
{
    double a = in[0], b = in[1];
    double acc[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) acc[c] = in[2 + c] + threadIdx.x;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int c = 0; c < CHAINS; c++) acc[c] = __fma_rn(acc[c], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) s += acc[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// the DTW cost pattern: 4 row chains x 12 features, a[r][k] and b[k] all in distinct registers
__global__ void dot4x12(double *out, const double *in, int iters, long long *cycles)
{
    double a[4][12], b[12], acc[4];
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int k = 0; k < 12; k++) a[r][k] = in[(r * 12 + k) & 15] + threadIdx.x * 1e-3 + r + k;
#pragma unroll
    for (int k = 0; k < 12; k++) b[k] = in[k] + 1e-3 * k + 1e-6 * threadIdx.x;
#pragma unroll
    for (int r = 0; r < 4; r++) acc[r] = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 12; k++)
#pragma unroll
            for (int r = 0; r < 4; r++) acc[r] = __fma_rn(a[r][k], b[k], acc[r]);
#pragma unroll
        for (int k = 0; k < 12; k++) b[k] = __longlong_as_double(__double_as_longlong(b[k]) ^ (long long)((i ^ threadIdx.x) & 1));   // cheap int op: new b each round
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc[0] + acc[1] + acc[2] + acc[3];
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// throughput of the other fp64-pipe instructions, 8 independent chains each
template <int OP>
__global__ void op_tp(double *out, const double *in, int iters, long long *cycles)
{
    double v[8], w[8];
#pragma unroll
    for (int c = 0; c < 8; c++) { v[c] = in[c] + threadIdx.x; w[c] = in[c + 1] * 3.0; }
    double k1 = in[1] + 1.0;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int c = 0; c < 8; c++) {
                if (OP == 0) v[c] = __dadd_rn(v[c], k1);
                if (OP == 1) v[c] = __dmul_rn(v[c], k1);
                if (OP == 2) v[c] = (v[c] < w[c]) ? v[c] + 0.0 * w[c] : w[c];     // placeholder, see OP 3
                if (OP == 3) { double t = w[c]; w[c] = (v[c] < t) ? v[c] : t; v[c] = t; }   // DSETP + 2 FSEL (+ moves)
            }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < 8; c++) s += v[c] + w[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// the DP chain of one DTW row: y = up + c; m = min(y, x) (setp + select); v = min(z, m); up = v
__global__ void dp_chain(double *out, const double *in, int iters, long long *cycles)
{
    double up = in[0] + threadIdx.x, c = in[1], x = in[2], z = in[3];
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            double y = __dadd_rn(up, c);
            double m = y < x ? y : x;
            double v = z < m ? z : m;
            up = v;
            x += 1e-9;      // keeps the compiler from collapsing the chain (independent of up)
            z += 2e-9;
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = up;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <typename K>
static void run(const char *name, K kern, int threads, int iters, int ops_per_iter_per_thread, double *d_out, double *d_in, long long *d_cyc)
{
    kern<<<1, threads>>>(d_out, d_in, iters, d_cyc);
    cudaDeviceSynchronize();
    kern<<<1, threads>>>(d_out, d_in, iters, d_cyc);
    cudaDeviceSynchronize();
    long long cyc = 0;
    cudaMemcpy(&cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost);
    const double per_op = (double)cyc / ((double)iters * ops_per_iter_per_thread);
    printf("%-34s warps/SM %2d  cycles per (warp-)op %.2f   ops/clk/SM %.2f\n", name, threads / 32, per_op,
           (double)iters * ops_per_iter_per_thread * (threads / 32) * 32 / (double)cyc);
}

int main()
{
    double *d_out, *d_in;
    long long *d_cyc;
    cudaMalloc(&d_out, 1 << 20);
    cudaMalloc(&d_in, 1024);
    cudaMalloc(&d_cyc, 1024);
    double h[16] = {1.0000001, 1e-9, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0};
    h[2] = 1e300;
    h[3] = 1e300;
    double hd[4] = {0.0, 1e-3, 1e300, 1e300};
    cudaMemcpy(d_in, h, sizeof(h), cudaMemcpyHostToDevice);
    const int it = 20000;
    for (int threads : {32, 128, 256, 384, 512}) {
        run("DFMA 1 chain  (dependent latency)", dfma_chains<1>, threads, it, 8 * 1, d_out, d_in, d_cyc);
        run("DFMA 2 chains", dfma_chains<2>, threads, it, 8 * 2, d_out, d_in, d_cyc);
        run("DFMA 4 chains", dfma_chains<4>, threads, it, 8 * 4, d_out, d_in, d_cyc);
        run("DFMA 8 chains", dfma_chains<8>, threads, it, 8 * 8, d_out, d_in, d_cyc);
    }
    for (int threads : {128, 256, 384}) {
        run("dot 4 rows x 12 (distinct regs)", dot4x12, threads, it, 48, d_out, d_in, d_cyc);
        run("DADD x8 chains", op_tp<0>, threads, it, 32, d_out, d_in, d_cyc);
        run("DMUL x8 chains", op_tp<1>, threads, it, 32, d_out, d_in, d_cyc);
        run("DSETP+2FSEL x8 chains", op_tp<3>, threads, it, 32, d_out, d_in, d_cyc);
    }
    cudaMemcpy(d_in, hd, sizeof(hd), cudaMemcpyHostToDevice);
    for (int threads : {32, 128, 384})
        run("DP row chain (DADD+2x(DSETP,FSEL))", dp_chain, threads, it, 8, d_out, d_in, d_cyc);
    return 0;
}

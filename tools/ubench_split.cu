// Micro-benchmark for DESIGN.md K2: does splitting a DTW step into a cost-producer warp (48 FMAs in four independent
// chains) and a DP-consumer warp (the add / compare / select chain through four rows) let the hardware overlap the two
// better than one warp that runs both, statically interleaved by the compiler?
//   fused : W warps per SM, each does [cost(col s+1) ; dp(col s)] per step            (the shipped kernel's structure)
//   split : W/2 producer warps (cost -> shared memory) + W/2 consumer warps (shared memory -> dp), no handshake
//           (upper bound for the split design: synchronisation is free)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench_split tools/ubench_split.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void cost4(const double (&a)[4][12], const double (&b)[12], double (&c)[4])
{
#pragma unroll
    for (int r = 0; r < 4; r++) c[r] = __dmul_rn(a[r][0], b[0]);
#pragma unroll
    for (int k = 1; k < 12; k++)
#pragma unroll
        for (int r = 0; r < 4; r++) c[r] = __fma_rn(a[r][k], b[k], c[r]);
#pragma unroll
    for (int r = 0; r < 4; r++) c[r] = __dsub_rn(1.0, c[r]);
}

__device__ __forceinline__ void dp4(double (&left)[4], double &up_prev, double up, const double (&c)[4], unsigned &bits)
{
    double diag = up_prev, upv = up;
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const double x = __dadd_rn(left[r], c[r]), y = __dadd_rn(upv, c[r]), z = __fma_rn(2.0, c[r], diag);
        const bool yx = y < x;
        const double m = yx ? y : x;
        const bool zm = z < m;
        const double v = zm ? z : m;
        bits += (yx ? 1u : 0u) + (zm ? 2u : 0u);
        diag = left[r];
        left[r] = v;
        upv = v;
    }
    up_prev = up;
}

__global__ void __launch_bounds__(512) fused(double *out, const double *in, int steps, long long *cycles)
{
    __shared__ double sb[132 * 14];
    for (int i = threadIdx.x; i < 132 * 14; i += blockDim.x) sb[i] = in[i % 64] + 1e-3 * i;
    const int lane = threadIdx.x & 31;
    double a[4][12], left[4], c[4], up_prev = 0.5, bottom = 0.25;
    for (int r = 0; r < 4; r++) { left[r] = 1.0 + r; for (int k = 0; k < 12; k++) a[r][k] = in[(r * 12 + k) & 63] * 1e-2 + 1e-4 * threadIdx.x; }
    unsigned bits = 0;
    __syncthreads();
    { double b[12]; for (int k = 0; k < 12; k++) b[k] = sb[(lane & 127) * 14 + k]; cost4(a, b, c); }
    long long t0 = clock64();
    for (int s = 0; s < steps; s++) {
        double b[12], cn[4];
        const double *col = sb + ((s + 1 - lane) & 127) * 14;
#pragma unroll
        for (int k = 0; k < 12; k += 2) { const double2 v = *reinterpret_cast<const double2 *>(col + k); b[k] = v.x; b[k + 1] = v.y; }
        cost4(a, b, cn);
        const double up = __shfl_up_sync(0xffffffffu, bottom, 1);
        dp4(left, up_prev, up, c, bits);
        bottom = left[3];
#pragma unroll
        for (int r = 0; r < 4; r++) c[r] = cn[r];
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = left[0] + left[3] + bits;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

__global__ void __launch_bounds__(512) split(double *out, const double *in, int steps, long long *cycles)
{
    __shared__ double sb[132 * 14];
    __shared__ double sc[8][2][32 * 4 + 8];          // per pair: 2-deep ring of 128 costs (padded)
    for (int i = threadIdx.x; i < 132 * 14; i += blockDim.x) sb[i] = in[i % 64] + 1e-3 * i;
    for (int i = threadIdx.x; i < 8 * 2 * 136; i += blockDim.x) (&sc[0][0][0])[i] = 0.5 + 1e-3 * (i & 63);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int npairs = blockDim.x / 64;
    const bool producer = warp < npairs;              // warps 0..npairs-1 produce, the rest consume
    const int pair = producer ? warp : warp - npairs;
    __syncthreads();
    long long t0 = clock64();
    double res = 0;
    if (producer) {
        double a[4][12], c[4];
        for (int r = 0; r < 4; r++) for (int k = 0; k < 12; k++) a[r][k] = in[(r * 12 + k) & 63] * 1e-2 + 1e-4 * threadIdx.x;
        for (int s = 0; s < steps; s++) {
            double b[12];
            const double *col = sb + ((s + 1 - lane) & 127) * 14;
#pragma unroll
            for (int k = 0; k < 12; k += 2) { const double2 v = *reinterpret_cast<const double2 *>(col + k); b[k] = v.x; b[k + 1] = v.y; }
            cost4(a, b, c);
            double *dst = &sc[pair][s & 1][lane * 4];
            *reinterpret_cast<double2 *>(dst) = make_double2(c[0], c[1]);
            *reinterpret_cast<double2 *>(dst + 2) = make_double2(c[2], c[3]);
        }
        res = c[0];
    } else {
        double left[4], up_prev = 0.5, bottom = 0.25;
        for (int r = 0; r < 4; r++) left[r] = 1.0 + r;
        unsigned bits = 0;
        for (int s = 0; s < steps; s++) {
            const double *src = &sc[pair][s & 1][lane * 4];
            double c[4];
            asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(c[0]), "=d"(c[1]) : "r"((unsigned)__cvta_generic_to_shared(src)));
            asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(c[2]), "=d"(c[3]) : "r"((unsigned)__cvta_generic_to_shared(src + 2)));
            const double up = __shfl_up_sync(0xffffffffu, bottom, 1);
            dp4(left, up_prev, up, c, bits);
            bottom = left[3];
        }
        res = left[0] + left[3] + bits;
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = res;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;        // warp 0 = a producer
    if (threadIdx.x == blockDim.x - 32) cycles[blockIdx.x + 64] = t1 - t0;   // last warp = a consumer
}

int main()
{
    double *d_out, *d_in;
    long long *d_cyc;
    cudaMalloc(&d_out, 1 << 20);
    cudaMalloc(&d_in, 1024);
    cudaMalloc(&d_cyc, 2048);
    double h[64];
    for (int i = 0; i < 64; i++) h[i] = 0.1 + 0.01 * i;
    cudaMemcpy(d_in, h, sizeof(h), cudaMemcpyHostToDevice);
    const int steps = 20000;
    long long c[128];
    for (int warps : {4, 8, 12, 16}) {
        for (int rep = 0; rep < 2; rep++) fused<<<1, warps * 32>>>(d_out, d_in, steps, d_cyc);
        cudaDeviceSynchronize();
        cudaMemcpy(c, d_cyc, sizeof(c), cudaMemcpyDeviceToHost);
        printf("fused  %2d warps/SM (%d per sub-partition): %.1f cycles per band-step per sub-partition\n", warps, warps / 4,
               (double)c[0] / steps / (warps / 4.0));
    }
    for (int warps : {8, 16}) {
        for (int rep = 0; rep < 2; rep++) split<<<1, warps * 32>>>(d_out, d_in, steps, d_cyc);
        cudaDeviceSynchronize();
        cudaMemcpy(c, d_cyc, sizeof(c), cudaMemcpyDeviceToHost);
        const double worst = (double)(c[0] > c[64] ? c[0] : c[64]);
        printf("split  %2d warps/SM (%d producer + %d consumer per sub-partition): producer %.1f, consumer %.1f -> %.1f cycles per band-step per sub-partition\n",
               warps, warps / 8, warps / 8, (double)c[0] / steps, (double)c[64] / steps, worst / steps / (warps / 8.0));
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

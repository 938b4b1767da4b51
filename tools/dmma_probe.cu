// Does mma.sync.aligned.m8n8k4.f64 reproduce the sequential fp64 FMA chain of K2's local cost bit for bit?
// K2 computes s = a0*b0; s = fma(a_k, b_k, s), k = 1..11 (reference: 1 - np.dot(a, b), numpy's sequential order for 12
// elements) on the DFMA pipe.  Three chained DMMAs (K = 4 each, C = 0 for the first) would take 13 of the 18
// FP64-pipe instructions per cell off that pipe — if the tensor core's internal order and rounding are those of the chain.
// The probe runs 8 x 8 dot products of length 12 on random data (and on data with heavy cancellation) both ways and counts
// mismatching bits.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/dmma_probe tools/dmma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

__global__ void probe(const double *A, const double *B, double *d_mma, double *d_chain, double *d_chain_rev)
{
    // A: 8 x 12 row-major, B: 12 x 8 (B[k][n]); one warp
    const int lane = threadIdx.x;
    const int row = lane >> 2, kq = lane & 3;       // A fragment: a = A[row][k0 + kq]; B fragment: b = B[k0 + kq][col = lane >> 2]
    double c0 = 0.0, c1 = 0.0;
    for (int k0 = 0; k0 < 12; k0 += 4) {
        const double a = A[row * 12 + k0 + kq];
        const double b = B[(k0 + kq) * 8 + (lane >> 2)];
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
    }
    // C fragment: c0 = D[row = lane >> 2][col = 2 * (lane & 3)], c1 = D[row][col + 1]
    d_mma[(lane >> 2) * 8 + 2 * (lane & 3)] = c0;
    d_mma[(lane >> 2) * 8 + 2 * (lane & 3) + 1] = c1;
    for (int e = lane; e < 64; e += 32) {
        const int m = e >> 3, n = e & 7;
        double s = A[m * 12] * B[n];
        for (int k = 1; k < 12; k++) s = fma(A[m * 12 + k], B[k * 8 + n], s);
        d_chain[e] = s;
        double r = A[m * 12 + 11] * B[11 * 8 + n];
        for (int k = 10; k >= 0; k--) r = fma(A[m * 12 + k], B[k * 8 + n], r);
        d_chain_rev[e] = r;
    }
}

int main()
{
    double *dA, *dB, *dM, *dC, *dR;
    cudaMalloc(&dA, 96 * 8); cudaMalloc(&dB, 96 * 8); cudaMalloc(&dM, 64 * 8); cudaMalloc(&dC, 64 * 8); cudaMalloc(&dR, 64 * 8);
    long long same = 0, same_rev = 0, total = 0;
    double worst = 0.0;
    srand(7);
    for (int trial = 0; trial < 2000; trial++) {
        std::vector<double> A(96), B(96);
        for (auto &v : A) v = (double)rand() / RAND_MAX;
        for (auto &v : B) v = (double)rand() / RAND_MAX;
        if (trial & 1) for (int i = 0; i < 96; i += 2) { A[i] = -A[i]; }        // cancellation
        cudaMemcpy(dA, A.data(), 96 * 8, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, B.data(), 96 * 8, cudaMemcpyHostToDevice);
        probe<<<1, 32>>>(dA, dB, dM, dC, dR);
        double M[64], C[64], R[64];
        cudaMemcpy(M, dM, 512, cudaMemcpyDeviceToHost);
        cudaMemcpy(C, dC, 512, cudaMemcpyDeviceToHost);
        cudaMemcpy(R, dR, 512, cudaMemcpyDeviceToHost);
        for (int e = 0; e < 64; e++) {
            total++;
            same += memcmp(&M[e], &C[e], 8) == 0;
            same_rev += memcmp(&M[e], &R[e], 8) == 0;
            double d = M[e] - C[e];
            if (d < 0) d = -d;
            if (d > worst) worst = d;
        }
    }
    printf("{\"dmma_m8n8k4_f64\": {\"dot_products\": %lld, \"bit_equal_to_forward_fma_chain\": %lld, \"bit_equal_to_reverse_chain\": %lld, "
           "\"max_abs_diff_vs_forward_chain\": %.3e, \"err\": \"%s\"}}\n", total, same, same_rev, worst, cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}

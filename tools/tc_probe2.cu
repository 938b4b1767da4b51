// Probe 2: (a) does the ~45-cycle floor of small-N tcgen05.mma come from the dependent accumulate chain — time the same
// MMAs round-robin over several independent accumulators; (b) tcgen05.ld throughput per SM for 4 / 8 / 16 warps.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I include -I real-time-audio-sync_b200/csrc
//             -o build/tc_probe2 tools/tc_probe2.cu real-time-audio-sync_b200/csrc/afs_common.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc05.cuh"

template <int N, int TS, int NACC, int M>
__global__ void __launch_bounds__(128) mma_time_kernel(long long *cycles, int reps)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    constexpr uint32_t a_tile = 128 * 128, b_tile = (uint32_t)N * 128;
    for (int i = tid; i < (int)((a_tile + b_tile) / 4); i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u + i;
    if (tid == 0) {
        afs::mbar_init(&s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc::fence_async_smem();
    if (warp == 0) tc::tmem_alloc(&s_tmem, 512);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tm = s_tmem;
    constexpr uint32_t idesc = tc::idesc_bf16_f32(M, N);
    if (tid == 0) {
        const uint64_t ad = tc::smem_desc_k_sw128(afs::smem_addr(smem));
        const uint64_t bd = tc::smem_desc_k_sw128(afs::smem_addr(smem + a_tile));
        const long long t0 = clock64();
        for (int rep = 0; rep < reps; rep++) {
#pragma unroll
            for (int ks = 0; ks < 4; ks++) {
#pragma unroll
                for (int acc = 0; acc < NACC; acc++) {
                    // accumulators at columns acc * N (A for TS mode at columns 448..479)
                    if (TS) tc::mma_ts(tm + acc * N, tm + 448 + ks * 8, bd + 2 * ks, idesc, 1u);
                    else tc::mma_ss(tm + acc * N, ad + 2 * ks, bd + 2 * ks, idesc, 1u);
                }
            }
        }
        tc::mma_commit(&s_bar);
        afs::mbar_wait(&s_bar, 0);
        cycles[blockIdx.x] = clock64() - t0;
    }
    __syncthreads();
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tm, 512);
}

template <int N, int TS, int NACC, int M>
static void run_mma(const char *name, int reps)
{
    const int blocks = 148;
    long long *dc;
    cudaMalloc(&dc, sizeof(long long) * blocks);
    const size_t smem = 128 * 128 + (size_t)N * 128 + 1024;
    cudaFuncSetAttribute(mma_time_kernel<N, TS, NACC, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int it = 0; it < 2; it++) mma_time_kernel<N, TS, NACC, M><<<blocks, 128, smem>>>(dc, reps);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<long long> cyc(blocks);
    cudaMemcpy(cyc.data(), dc, cyc.size() * 8, cudaMemcpyDeviceToHost);
    long long cmax = 0;
    for (long long c : cyc) cmax = c > cmax ? c : cmax;
    const int n_mma = reps * 4 * NACC;
    printf("{\"mma\": \"%s\", \"m\": %d, \"n\": %d, \"ts\": %d, \"accumulators\": %d, \"mmas\": %d, \"cycles_per_mma\": %.1f, \"floor\": %.1f, \"err\": \"%s\"}\n",
           name, M, N, TS, NACC, n_mma, (double)cmax / n_mma, (M > 64 ? 128.0 : 128.0) * N / 256.0, cudaGetErrorString(e));
    cudaFree(dc);
}

// ---- tcgen05.ld throughput ----
template <int X>
__device__ __forceinline__ uint32_t ld_sum(uint32_t taddr);
template <>
__device__ __forceinline__ uint32_t ld_sum<16>(uint32_t taddr)
{
    uint32_t v[16];
    tc::tmem_ld16(taddr, v);
    tc::tmem_wait_ld();
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) s ^= v[j];
    return s;
}
template <>
__device__ __forceinline__ uint32_t ld_sum<32>(uint32_t taddr)
{
    uint32_t a[16], b[16];
    tc::tmem_ld16(taddr, a);
    tc::tmem_ld16(taddr + 16, b);
    tc::tmem_wait_ld();
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) s ^= a[j] ^ b[j];
    return s;
}
template <>
__device__ __forceinline__ uint32_t ld_sum<64>(uint32_t taddr)
{
    uint32_t a[16], b[16], c[16], d[16];
    tc::tmem_ld16(taddr, a);
    tc::tmem_ld16(taddr + 16, b);
    tc::tmem_ld16(taddr + 32, c);
    tc::tmem_ld16(taddr + 48, d);
    tc::tmem_wait_ld();
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) s ^= a[j] ^ b[j] ^ c[j] ^ d[j];
    return s;
}

template <int X>
__global__ void __launch_bounds__(512) ld_time_kernel(long long *cycles, uint32_t *sink, int reps)
{
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tc::tmem_alloc(&s_tmem, 512);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tm = s_tmem;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t col0 = (uint32_t)(warp >> 2) * 64;       // up to 4 warps per lane quarter, each on its own 64 columns
    uint32_t s = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int rep = 0; rep < reps; rep++) {
#pragma unroll
        for (int c = 0; c < 64; c += X) s ^= ld_sum<X>(tm + lane_base + col0 + c);
    }
    __syncthreads();
    const long long t1 = clock64();
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    if (s == 0x12345678u) sink[tid] = s;
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tm, 512);
}

template <int X>
static void run_ld(int warps, int reps)
{
    const int blocks = 148;
    long long *dc;
    uint32_t *sink;
    cudaMalloc(&dc, sizeof(long long) * blocks);
    cudaMalloc(&sink, 4 * 512);
    for (int it = 0; it < 2; it++) ld_time_kernel<X><<<blocks, warps * 32, 0>>>(dc, sink, reps);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<long long> cyc(blocks);
    cudaMemcpy(cyc.data(), dc, cyc.size() * 8, cudaMemcpyDeviceToHost);
    long long cmax = 0;
    for (long long c : cyc) cmax = c > cmax ? c : cmax;
    const double bytes = (double)reps * 64 * 4 * 32 * warps;
    printf("{\"tmem_ld\": \"x%d\", \"warps\": %d, \"bytes_per_cycle_per_sm\": %.1f, \"cycles\": %lld, \"err\": \"%s\"}\n", X, warps, bytes / cmax,
           cmax, cudaGetErrorString(e));
    cudaFree(dc);
    cudaFree(sink);
}

int main()
{
    run_mma<64, 0, 1, 128>("ss_n64_acc1", 128);
    run_mma<64, 0, 2, 128>("ss_n64_acc2", 128);
    run_mma<64, 0, 4, 128>("ss_n64_acc4", 64);
    run_mma<32, 0, 1, 128>("ss_n32_acc1", 128);
    run_mma<32, 0, 2, 128>("ss_n32_acc2", 128);
    run_mma<32, 0, 4, 128>("ss_n32_acc4", 64);
    run_mma<32, 0, 8, 128>("ss_n32_acc8", 64);
    run_mma<16, 0, 4, 128>("ss_n16_acc4", 64);
    run_mma<16, 0, 8, 128>("ss_n16_acc8", 64);
    run_mma<8, 0, 8, 64>("ss_m64_n8_acc8", 64);
    run_mma<16, 0, 8, 64>("ss_m64_n16_acc8", 64);
    run_mma<32, 0, 4, 64>("ss_m64_n32_acc4", 64);
    run_mma<80, 1, 1, 128>("ts_n80_acc1", 128);
    run_mma<80, 1, 2, 128>("ts_n80_acc2", 128);
    run_mma<80, 1, 4, 128>("ts_n80_acc4", 64);
    run_mma<64, 1, 2, 128>("ts_n64_acc2", 128);
    run_mma<112, 1, 2, 128>("ts_n112_acc2", 64);
    for (int warps : {4, 8, 16}) {
        run_ld<16>(warps, 256);
        run_ld<32>(warps, 256);
        run_ld<64>(warps, 256);
    }
    return 0;
}

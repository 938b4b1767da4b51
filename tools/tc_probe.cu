// Probe for the tcgen05 building blocks K1 (chroma) is assembled from.  Each case runs one CTA of 128
// threads: operands are laid out in shared memory (K-major, 128-byte swizzle) or written into TMEM with
// tcgen05.st (TS mode), a chain of tcgen05.mma accumulates into TMEM, tcgen05.ld brings D back and the host
// compares with a float64 product.  A timing mode repeats the chain on every SM and reports cycles per MMA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I include -I real-time-audio-sync_b200/csrc
//             -o build/tc_probe tools/tc_probe.cu real-time-audio-sync_b200/csrc/afs_common.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc05.cuh"

struct ProbeArgs {
    const uint16_t *a_hi, *a_lo;   // [128][K] bf16 bits
    const uint16_t *b_hi, *b_lo;   // [N][K]
    float *d;                      // [128][N]
    long long *cycles;             // per block
    int n, kb;                     // N, number of 64-element K blocks
    int ts;                        // 1: A from TMEM
    int terms;                     // 1: hi*hi ; 3: hi*hi + lo*hi + hi*lo
    int reps;                      // timing: repeat the whole chain (first rep overwrites, the rest accumulate)
};

__global__ void __launch_bounds__(128) probe_kernel(ProbeArgs p)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_bar;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int K = p.kb * 64;
    const uint32_t a_tile = 128 * 128, b_tile = (uint32_t)p.n * 128;       // bytes per 64-element K block
    unsigned char *sa_hi = smem, *sa_lo = sa_hi + p.kb * a_tile;
    unsigned char *sb_hi = sa_lo + p.kb * a_tile, *sb_lo = sb_hi + p.kb * b_tile;
    // ---- operands -> shared memory in the canonical K-major SW128 layout ----
    for (int i = tid; i < 128 * K; i += 128) {
        const int r = i / K, k = i % K;
        const uint32_t off = (k >> 6) * a_tile + tc::sw128_offset(r, k & 63);
        *reinterpret_cast<uint16_t *>(sa_hi + off) = p.a_hi[i];
        *reinterpret_cast<uint16_t *>(sa_lo + off) = p.a_lo[i];
    }
    for (int i = tid; i < p.n * K; i += 128) {
        const int r = i / K, k = i % K;
        const uint32_t off = (k >> 6) * b_tile + tc::sw128_offset(r, k & 63);
        *reinterpret_cast<uint16_t *>(sb_hi + off) = p.b_hi[i];
        *reinterpret_cast<uint16_t *>(sb_lo + off) = p.b_lo[i];
    }
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
    const uint32_t d_tm = tm;                      // columns [0, n)
    const uint32_t ah_tm = tm + 256, al_tm = tm + 384;   // TS operands: K/2 columns each
    if (p.ts) {
        // thread = lane = row of A; 32-bit column c = {A[row][2c+1] : A[row][2c]}
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        for (int c0 = 0; c0 < K / 2; c0 += 16) {
            uint32_t vh[16], vl[16];
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const int k = 2 * (c0 + j);
                vh[j] = (uint32_t)p.a_hi[tid * K + k] | ((uint32_t)p.a_hi[tid * K + k + 1] << 16);
                vl[j] = (uint32_t)p.a_lo[tid * K + k] | ((uint32_t)p.a_lo[tid * K + k + 1] << 16);
            }
            tc::tmem_st16(ah_tm + lane_base + c0, vh);
            tc::tmem_st16(al_tm + lane_base + c0, vl);
        }
        tc::tmem_wait_st();
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
    }
    const uint32_t idesc = tc::idesc_bf16_f32(128, p.n);
    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        t0 = clock64();
        for (int rep = 0; rep < p.reps; rep++) {
            uint32_t acc = rep > 0;
            for (int term = 0; term < p.terms; term++) {
                // term 0: hi*hi, term 1: lo*hi, term 2: hi*lo
                const unsigned char *sa = term == 1 ? sa_lo : sa_hi;
                const unsigned char *sb = term == 2 ? sb_lo : sb_hi;
                const uint32_t a_tm = term == 1 ? al_tm : ah_tm;
                for (int ks = 0; ks < K / 16; ks++) {
                    const uint32_t boff = (ks >> 2) * b_tile + (ks & 3) * 32;
                    const uint64_t bd = tc::smem_desc_k_sw128(afs::smem_addr(sb) + boff);
                    if (p.ts) {
                        tc::mma_ts(d_tm, a_tm + ks * 8, bd, idesc, acc);
                    } else {
                        const uint32_t aoff = (ks >> 2) * a_tile + (ks & 3) * 32;
                        tc::mma_ss(d_tm, tc::smem_desc_k_sw128(afs::smem_addr(sa) + aoff), bd, idesc, acc);
                    }
                    acc = 1;
                }
            }
        }
        tc::mma_commit(&s_bar);
    }
    afs::mbar_wait(&s_bar, 0);
    if (tid == 0) {
        t1 = clock64();
        p.cycles[blockIdx.x] = t1 - t0;
    }
    tc::fence_after_sync();
    if (blockIdx.x == 0) {
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        for (int c0 = 0; c0 < p.n; c0 += 16) {
            uint32_t v[16];
            tc::tmem_ld16(d_tm + lane_base + c0, v);
            tc::tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; j++) p.d[(size_t)(warp * 32 + lane) * p.n + c0 + j] = __uint_as_float(v[j]);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tm, 512);
}


// Timing variant: everything the issue loop needs is a compile-time constant, so the single issuing thread spends
// one descriptor add per MMA and the measurement shows the tensor pipe (plus operand fetch), not the issue loop.
template <int N, int KB, int TS, int TERMS>
__global__ void __launch_bounds__(128) probe_time_kernel(ProbeArgs p)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    constexpr uint32_t a_tile = 128 * 128, b_tile = (uint32_t)N * 128;
    unsigned char *sa_hi = smem, *sa_lo = sa_hi + KB * a_tile;
    unsigned char *sb_hi = sa_lo + KB * a_tile, *sb_lo = sb_hi + KB * b_tile;
    for (int i = tid; i < (int)(KB * (2 * a_tile + 2 * b_tile) / 4); i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u + i;
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
    constexpr uint32_t idesc = tc::idesc_bf16_f32(128, N);
    if (tid == 0) {
        const uint64_t ah = tc::smem_desc_k_sw128(afs::smem_addr(sa_hi)), al = tc::smem_desc_k_sw128(afs::smem_addr(sa_lo));
        const uint64_t bh = tc::smem_desc_k_sw128(afs::smem_addr(sb_hi)), bl = tc::smem_desc_k_sw128(afs::smem_addr(sb_lo));
        const long long t0 = clock64();
        for (int rep = 0; rep < p.reps; rep++) {
#pragma unroll
            for (int term = 0; term < TERMS; term++) {
#pragma unroll
                for (int ks = 0; ks < KB * 4; ks++) {
                    const uint32_t boff = ((ks >> 2) * b_tile + (ks & 3) * 32) >> 4;
                    const uint32_t aoff = ((ks >> 2) * a_tile + (ks & 3) * 32) >> 4;
                    const uint64_t bd = (term == 2 ? bl : bh) + boff;
                    if (TS) tc::mma_ts(tm, tm + 256 + (term == 1 ? 128 : 0) + ks * 8, bd, idesc, 1u);
                    else tc::mma_ss(tm, (term == 1 ? al : ah) + aoff, bd, idesc, 1u);
                }
            }
        }
        tc::mma_commit(&s_bar);
        afs::mbar_wait(&s_bar, 0);
        p.cycles[blockIdx.x] = clock64() - t0;
    }
    __syncthreads();
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tm, 512);
}

template <int N, int KB, int TS, int TERMS>
static void run_time(const char *name, int reps, int blocks)
{
    ProbeArgs p{};
    long long *dc;
    cudaMalloc(&dc, sizeof(long long) * blocks);
    p.cycles = dc;
    p.reps = reps;
    const size_t smem = (size_t)KB * (2 * 128 * 128 + 2 * (size_t)N * 128) + 1024;
    cudaFuncSetAttribute(probe_time_kernel<N, KB, TS, TERMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int it = 0; it < 2; it++) probe_time_kernel<N, KB, TS, TERMS><<<blocks, 128, smem>>>(p);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<long long> cyc(blocks);
    cudaMemcpy(cyc.data(), dc, cyc.size() * 8, cudaMemcpyDeviceToHost);
    long long cmax = 0;
    for (long long c : cyc) cmax = c > cmax ? c : cmax;
    const int n_mma = reps * TERMS * KB * 4;
    printf("{\"time\": \"%s\", \"n\": %d, \"k\": %d, \"ts\": %d, \"terms\": %d, \"blocks\": %d, \"mmas\": %d, \"cycles_per_mma\": %.1f, \"floor\": %.1f, \"err\": \"%s\"}\n",
           name, N, KB * 64, TS, TERMS, blocks, n_mma, (double)cmax / n_mma, 128.0 * N / 256.0, cudaGetErrorString(e));
    cudaFree(dc);
}

static uint16_t trunc_bf16(float x, float *back)
{
    uint32_t u;
    memcpy(&u, &x, 4);
    u &= 0xFFFF0000u;
    memcpy(back, &u, 4);
    return (uint16_t)(u >> 16);
}

#define CK(x)                                                                      \
    do {                                                                           \
        cudaError_t e = (x);                                                       \
        if (e != cudaSuccess) {                                                    \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            exit(2);                                                               \
        }                                                                          \
    } while (0)

static int run_case(const char *name, int n, int kb, int ts, int terms, int reps, int blocks)
{
    const int K = kb * 64;
    std::vector<float> A(128 * K), B((size_t)n * K);
    std::vector<uint16_t> ah(A.size()), al(A.size()), bh(B.size()), bl(B.size());
    std::vector<float> Aeff(A.size()), Beff(B.size());
    srand(1234 + n + kb);
    for (size_t i = 0; i < A.size(); i++) {
        A[i] = (float)rand() / RAND_MAX * 2.f - 1.f;
        float h, l;
        ah[i] = trunc_bf16(A[i], &h);
        al[i] = trunc_bf16(A[i] - h, &l);
        Aeff[i] = h;
    }
    for (size_t i = 0; i < B.size(); i++) {
        B[i] = (float)rand() / RAND_MAX * 2.f - 1.f;
        float h, l;
        bh[i] = trunc_bf16(B[i], &h);
        bl[i] = trunc_bf16(B[i] - h, &l);
        Beff[i] = h;
    }
    ProbeArgs p{};
    uint16_t *dah, *dal, *dbh, *dbl;
    CK(cudaMalloc(&dah, ah.size() * 2)); CK(cudaMalloc(&dal, al.size() * 2));
    CK(cudaMalloc(&dbh, bh.size() * 2)); CK(cudaMalloc(&dbl, bl.size() * 2));
    CK(cudaMemcpy(dah, ah.data(), ah.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dal, al.data(), al.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dbh, bh.data(), bh.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dbl, bl.data(), bl.size() * 2, cudaMemcpyHostToDevice));
    float *dd;
    long long *dc;
    CK(cudaMalloc(&dd, sizeof(float) * 128 * n));
    CK(cudaMalloc(&dc, sizeof(long long) * blocks));
    CK(cudaMemset(dd, 0, sizeof(float) * 128 * n));
    p.a_hi = dah; p.a_lo = dal; p.b_hi = dbh; p.b_lo = dbl; p.d = dd; p.cycles = dc;
    p.n = n; p.kb = kb; p.ts = ts; p.terms = terms; p.reps = reps;
    const size_t smem = (size_t)kb * (2 * 128 * 128 + 2 * (size_t)n * 128) + 1024;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe_kernel<<<blocks, 128, smem>>>(p);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> D(128 * n);
    std::vector<long long> cyc(blocks);
    CK(cudaMemcpy(D.data(), dd, D.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(cyc.data(), dc, cyc.size() * 8, cudaMemcpyDeviceToHost));
    // reference: terms == 1 -> product of the truncated operands (exact in fp32 up to accumulation order);
    //            terms == 3 -> product of the original fp32 operands
    double max_err = 0.0, max_ref = 0.0;
    for (int m = 0; m < 128; m++)
        for (int j = 0; j < n; j++) {
            double s = 0.0;
            for (int k = 0; k < K; k++)
                s += terms == 1 ? (double)Aeff[m * K + k] * Beff[(size_t)j * K + k] : (double)A[m * K + k] * B[(size_t)j * K + k];
            s *= reps;
            max_err = fmax(max_err, fabs(s - D[(size_t)m * n + j]));
            max_ref = fmax(max_ref, fabs(s));
        }
    long long cmax = 0, cmin = 1LL << 60;
    for (long long c : cyc) { cmax = c > cmax ? c : cmax; cmin = c < cmin ? c : cmin; }
    const int n_mma = reps * terms * (K / 16);
    const double tol = terms == 1 ? 1e-4 * reps : 2e-3 * reps;      // 3-term: dropped lo*lo and truncation ~ 2^-15 per product
    const int ok = max_err <= tol * fmax(1.0, max_ref / 8);
    printf("{\"case\": \"%s\", \"n\": %d, \"k\": %d, \"ts\": %d, \"terms\": %d, \"reps\": %d, \"blocks\": %d, \"max_abs_err\": %.3e, "
           "\"max_ref\": %.3f, \"ok\": %s, \"cycles_min\": %lld, \"cycles_max\": %lld, \"cycles_per_mma\": %.1f}\n",
           name, n, K, ts, terms, reps, blocks, max_err, max_ref, ok ? "true" : "false", cmin, cmax, (double)cmax / n_mma);
    cudaFree(dah); cudaFree(dal); cudaFree(dbh); cudaFree(dbl); cudaFree(dd); cudaFree(dc);
    return ok;
}

int main()
{
    int ok = 1;
    ok &= run_case("ss_n128_k64", 128, 1, 0, 1, 1, 1);
    ok &= run_case("ss_n80_k64", 80, 1, 0, 1, 1, 1);
    ok &= run_case("ss_n144_k128", 144, 2, 0, 1, 1, 1);
    ok &= run_case("ts_n128_k64", 128, 1, 1, 1, 1, 1);
    ok &= run_case("ts_n144_k128", 144, 2, 1, 1, 1, 1);
    ok &= run_case("ss_3term_n80_k64", 80, 1, 0, 3, 1, 1);
    ok &= run_case("ts_3term_n144_k128", 144, 2, 1, 3, 1, 1);
    // timing on every SM at once (cycles per MMA of the chain; floor = 128 * N / 256)
    run_time<80, 1, 0, 3>("ss_n80", 64, 148);
    run_time<80, 1, 1, 3>("ts_n80", 64, 148);
    run_time<144, 2, 0, 3>("ss_n144", 32, 148);
    run_time<144, 2, 1, 3>("ts_n144", 32, 148);
    run_time<128, 2, 0, 3>("ss_n128", 32, 148);
    run_time<128, 2, 1, 3>("ts_n128", 32, 148);
    run_time<256, 1, 0, 1>("ss_n256", 128, 148);
    run_time<256, 1, 1, 1>("ts_n256", 128, 148);
    run_time<128, 2, 1, 3>("ts_n128_1sm", 32, 1);
    run_time<32, 1, 1, 1>("ts_n32", 256, 148);
    run_time<16, 1, 1, 1>("ts_n16", 256, 148);
    printf("{\"all_ok\": %s}\n", ok ? "true" : "false");
    return ok ? 0 : 1;
}

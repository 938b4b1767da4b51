// K2/K3 — full DTW for a batch of independent pairs on sm_100a.
//
// Replaces dtw.DTW (reference dtw.py:5-53).  The N x M cost and accumulated-cost
// matrices are never written to HBM: the cosine cost 1 - a_i.b_j (K = 12) is
// computed inline inside a skewed warp wavefront, and only a 2-bit direction code
// per cell is stored (0 = left (i,j-1), 1 = up (i-1,j), 2 or 3 = diag; dtw.py:30,38-40).
//
// Work decomposition
//   band   = 128 consecutive rows (i) of one pair, swept over all N columns by ONE warp;
//   lane l = rows band*128 + 4l .. +3, its 4 x 12 a-features live in registers;
//   step s : lane l processes column j = s - l  (anti-diagonal skew inside the warp);
//            the value under the lane's last row travels to lane l+1 by warp shuffle
//            (the `up`/`diag` inputs of the next lane's first row);
//   seq_b  : re-packed once per call to column-major [N][12 (+2 pad)] so that a chunk of
//            32 columns is ONE contiguous TMA bulk copy (cp.async.bulk -> mbarrier) into a
//            4-slot shared-memory ring per warp, and a lane fetches its column with
//            128-bit shared loads at a compile-time stride (bank-conflict free);
//   band b needs the last row of band b-1: it is streamed through an L2-resident
//   ring of two rows per pair in chunks of 32 columns.  Each hand-off record validates
//   itself ({value lo32, tag, value hi32, tag}, tag = launch epoch | band; one 128-bit
//   volatile store / load): no flags, no fences, no L1 invalidation on the hot path.
//   Bands are claimed from one global atomic ticket in an order where every dependency
//   has a smaller ticket (band-major, pairs interleaved), so a waiting warp always
//   waits on a warp that is already running.
//
// Direction map layout (per pair): 16-byte units; unit (cbp, g) covers row group
// g = i >> 2 (4 rows) and 16 consecutive SKEWED columns  jj = j + (g & 31),
// cbp = jj >> 4.  Inside a unit: byte (jj & 15), bits 2*(i & 3).  The skew makes
// all 32 lanes of a warp finish a unit on the same step, so a warp stores
// 32 x 16 B = 512 contiguous bytes every 16 steps (unit index = cbp * gpad + g).
#include <math_constants.h>

#include <atomic>
#include <cstdlib>
#include <type_traits>

#include <vector>

#include "afs_common.cuh"

namespace {

constexpr int kF = 12;           // chroma features (F = 12 everywhere)
constexpr int kRows = 4;         // rows per lane
constexpr int kBandRows = 32 * kRows;
constexpr int kWarpsPerBlock = 4;
constexpr int kChunkCols = 32;   // columns per staged chunk of seq_b
constexpr int kRingSlots = 4;    // chunks resident per warp
constexpr int kMirrorCols = 4;   // ring columns 0..3 repeated after column 127 (one unrolled group of steps looks 4 ahead)
constexpr int kSlackGroups = 0;  // default extra band-to-band distance, in groups of 32 columns (see the kernel)

// packed seq_b: elements per column.  fp64: 12 + 2 pad = 112 B (lane stride 28 words: the 8
// lanes of a 128-bit shared-load phase hit 8 distinct 4-word bank groups); fp32: 12 = 48 B
// (stride 12 words: also conflict free).
template <typename T> struct ColStride { static constexpr int value = sizeof(T) == 8 ? 14 : 12; };

struct DtwPair {
    int64_t a_off, b_off;   // element offsets of (12,M) / (12,N)
    int64_t dir_off;        // uint4 units into the direction area
    int64_t brow_off;       // elements into the hand-off area (2 rows of nsteps)
    int64_t bt_off;         // elements into the packed seq_b area
    int64_t path_off;       // pairs into the path area
    int32_t M, N;
    int32_t nbands, gpad;   // gpad = nbands * 32 row groups
    int32_t nsteps;         // roundup32(N + 31)
    int32_t band0;          // index of this pair's first band in the flat band numbering (diagnostics)
    int32_t path_cap;       // M + N
    int32_t nchunks;        // ceil(N / 32)
};

struct DtwItem { int32_t pair, band; };

template <typename T> struct Arith;
template <> struct Arith<double> {
    static __device__ __forceinline__ double inf() { return CUDART_INF; }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    // one packed column (12 values) from shared memory: 6 x ld.shared.v2.f64 at immediate offsets
    template <int BASE> static __device__ __forceinline__ void load_col(uint32_t saddr, double (&bk)[kF])
    {
        lds2<BASE + 0>(saddr, bk[0], bk[1]);
        lds2<BASE + 16>(saddr, bk[2], bk[3]);
        lds2<BASE + 32>(saddr, bk[4], bk[5]);
        lds2<BASE + 48>(saddr, bk[6], bk[7]);
        lds2<BASE + 64>(saddr, bk[8], bk[9]);
        lds2<BASE + 80>(saddr, bk[10], bk[11]);
    }
    template <int OFF> static __device__ __forceinline__ void lds2(uint32_t saddr, double &x, double &y)
    {
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(x), "=d"(y) : "r"(saddr), "n"(OFF));
    }
    static __device__ __forceinline__ double lds(uint32_t saddr)
    {
        double v;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(saddr));
        return v;
    }
    static __device__ __forceinline__ void sts(uint32_t saddr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(saddr), "d"(v) : "memory"); }
    // out = a < b ? a : b;  if (a < b) bits |= bit   (bit is a compile-time constant after unrolling)
    static __device__ __forceinline__ void min_first(double &out, uint32_t &bits, double a, double b, uint32_t bit)
    {
        asm("{\n\t.reg .pred p;\n\tsetp.lt.f64 p, %2, %3;\n\tselp.f64 %0, %2, %3, p;\n\t@p or.b32 %1, %1, %4;\n\t}"
            : "=d"(out), "+r"(bits) : "d"(a), "d"(b), "r"(bit));
    }
};
template <> struct Arith<float> {
    static __device__ __forceinline__ float inf() { return CUDART_INF_F; }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    template <int BASE> static __device__ __forceinline__ void load_col(uint32_t saddr, float (&bk)[kF])
    {
        lds4<BASE + 0>(saddr, bk[0], bk[1], bk[2], bk[3]);
        lds4<BASE + 16>(saddr, bk[4], bk[5], bk[6], bk[7]);
        lds4<BASE + 32>(saddr, bk[8], bk[9], bk[10], bk[11]);
    }
    template <int OFF> static __device__ __forceinline__ void lds4(uint32_t saddr, float &x, float &y, float &z, float &w)
    {
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(saddr), "n"(OFF));
    }
    static __device__ __forceinline__ float lds(uint32_t saddr)
    {
        float v;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
        return v;
    }
    static __device__ __forceinline__ void sts(uint32_t saddr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(saddr), "f"(v) : "memory"); }
    static __device__ __forceinline__ void min_first(float &out, uint32_t &bits, float a, float b, uint32_t bit)
    {
        asm("{\n\t.reg .pred p;\n\tsetp.lt.f32 p, %2, %3;\n\tselp.f32 %0, %2, %3, p;\n\t@p or.b32 %1, %1, %4;\n\t}"
            : "=f"(out), "+r"(bits) : "f"(a), "f"(b), "r"(bit));
    }
};

template <typename T>
struct DtwArgs {
    const T *a, *b;
    const DtwPair *pairs;
    const DtwItem *items;
    int n_items;
    uint4 *dir;
    uint4 *brow;      // band hand-off rows: {value lo32, tag, value hi32, tag} records (absolute fp64 costs)
    uint32_t epoch;   // 1..250, different for consecutive launches on the same workspace
    T *bt;
    int *ticket;
    double *acc_end;
    T *dense_cost, *dense_acc;
    // K4 (column stripe of one long pair; all null for ordinary batches)
    const double *leftb;   // [M] acc of the column left of this stripe (null: this is the first stripe)
    double *rightb;        // [M] out: acc of this stripe's last column (may be peer-mapped memory)
    const int *in_flag;    // [nbands] raised by the left stripe when leftb rows of a band are valid (null: valid already)
    int *out_flag;         // [nbands] raised for the right stripe (may be peer-mapped; null: none)
    unsigned long long wait_ns;   // bound on the in_flag wait
    int slack_groups;             // extra distance (groups of 32 columns) a band keeps from the band above
};

// Per-lane wavefront state.
template <typename T>
struct Lane {
    T ar[kRows][kF];
    T left[kRows];
    T up_prev;
    T bottom;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K4 only: longest a band waits for the stripe on its left before the launch is declared failed
// (default; the environment variable AFS_STRIPE_WAIT_MS overrides it per launch).
constexpr unsigned long long kStripeWaitNs = 20ull * 1000ull * 1000ull * 1000ull;
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
    } while (!done);
}

// Hand-off records are read/written with single 128-bit volatile accesses (never cached in L1).
// What the protocol relies on, and what it does not.  A record is {value lo, tag, value hi, tag}: each 64-bit half carries
// its own copy of the (launch epoch, band) tag, and a reader accepts a record only when BOTH tags match, so the protocol
// needs single-copy atomicity of aligned 64-bit halves only (which PTX guarantees for naturally aligned accesses of up to
// 64 bits), not of the whole 16-byte vector: a torn 128-bit access shows one stale half, whose tag is the previous band's or
// launch's and fails the comparison; the reader polls again.  What PTX does not promise is that the two halves of ONE
// st.v4 become visible in program order to a concurrent ld.v4 — irrelevant here because either order is detected.  The
// payload needs no fence: it IS the flagged word.  tests/test_dtw_gpu.py::test_band_handoff_stress_under_memory_pressure
// hammers the path (240 launches, alternating inputs on one workspace, memory hog on a second stream) in lieu of racecheck.
__device__ __forceinline__ uint4 ld_record(const uint4 *p)
{
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_record(uint4 *p, uint4 v)
{
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <typename T>
struct WarpSmem {
    // column (j & 127) at ring + (j & 127) * stride; the first kMirrorCols columns are mirrored behind the ring so
    // that a lane can address the next few columns from one base without wrapping (immediate offsets)
    T ring[(kRingSlots * kChunkCols + kMirrorCols) * ColStride<T>::value];
    T ubuf[32];
    T obuf[32];
    uint64_t mbar[kRingSlots];
};

// seq_b (12, N) feature-major -> [N_pad][stride] column-major (+ zero padding)
template <typename T>
__global__ void dtw_pack_b_kernel(const T *__restrict__ b, T *__restrict__ bt, const DtwPair *pairs, int n_pairs)
{
    constexpr int S = ColStride<T>::value;
    const int p = blockIdx.y;
    const DtwPair pm = pairs[p];
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= pm.nchunks * kChunkCols) return;
    const T *src = b + pm.b_off;
    T *dst = bt + pm.bt_off + (int64_t)j * S;
#pragma unroll
    for (int k = 0; k < S; k++) dst[k] = (k < kF && j < pm.N) ? __ldg(src + (int64_t)k * pm.N + j) : (T)0;
}

// cost of the lane's rows against column (s - lane): c[r] = 1 - a_r . b_j   (dtw.py:11; sequential fma chain over
// k per row = the rounding order of the reference's dgemm).  The four row chains are independent of each other and
// of the DP state, which is what lets dtw_step run them in the shadow of the previous column's DP chain.
// col_s: shared-space address of a ring column; COLS_AHEAD: compile-time column offset from it (mirror: no wrap).
template <typename T, int COLS_AHEAD>
__device__ __forceinline__ void dtw_cost(const Lane<T> &L, const uint32_t col_s, T (&c)[kRows])
{
    using A = Arith<T>;
    constexpr int S = ColStride<T>::value;
    T bk[kF];
    A::template load_col<COLS_AHEAD * S * (int)sizeof(T)>(col_s, bk);
#pragma unroll
    for (int r = 0; r < kRows; r++) c[r] = A::mul(L.ar[r][0], bk[0]);
#pragma unroll
    for (int k = 1; k < kF; k++)
#pragma unroll
        for (int r = 0; r < kRows; r++) c[r] = A::fma(L.ar[r][k], bk[k], c[r]);
#pragma unroll
    for (int r = 0; r < kRows; r++) c[r] = A::sub((T)1, c[r]);
}

// One wavefront step, software-pipelined: on entry c[] holds the costs of THIS step's column (computed during the
// previous step); the costs of the next step's column are computed first (48 independent-chain FMAs, throughput
// work) so that the scheduler can interleave them with this step's DP recurrence (a serial chain of
// add / compare / select through the four rows, latency work).
template <typename T, bool DENSE, int U, bool ALL>
__device__ __forceinline__ void dtw_step(Lane<T> &L, const int s, const int lane, const int band, const DtwPair &pm,
                                         const uint32_t ring_g, const uint32_t ubuf_s, T *obuf, const bool feeds_next,
                                         uint32_t &dw, const DtwArgs<T> &args, const double base, T (&c)[kRows])
{
    // ubuf_s / obuf point at the slots of the current group of FOUR steps (s & ~3): slot U is this step's
    using A = Arith<T>;
    const unsigned full = 0xffffffffu;
    const int N = pm.N;
    const int j = s - lane;
    // inputs for the lane's first row: value under the previous lane's last row, one step ago;
    // lane 0 takes it from the band above (staged in ubuf)
    T up = __shfl_up_sync(full, L.bottom, 1);
    if (lane == 0) up = A::lds(ubuf_s + U * (int)sizeof(T));
    // ALL: the caller guarantees 0 <= j < N for every lane (steady state): the commit selects fold away.
    // Otherwise lanes outside [0, N) (start-up / drain of the skew) compute on whatever the ring holds and
    // simply do not commit (branch-free: the unrolled steps stay one basic block).
    const bool act = ALL ? true : ((unsigned)j < (unsigned)N);
    T cn[kRows];
    dtw_cost<T, U + 1>(L, ring_g, cn);      // ring_g: this lane's column of the group's first step
    T diag = L.up_prev;
    T upv = up;
#pragma unroll
    for (int r = 0; r < kRows; r++) {
        T x = A::add(L.left[r], c[r]);                  // (i, j-1)   dtw.py:35
        T y = A::add(upv, c[r]);                        // (i-1, j)   dtw.py:36
        T z = A::fma((T)2, c[r], diag);                 // (i-1, j-1) dtw.py:37 (2c exact)
        // np.argmin: first minimum wins (strict <).  The direction code is the two raw predicates (bit 0: up
        // beat left, bit 1: diag beat both; 3 reads as diag in the backtrack), OR-ed into dw by ONE predicated
        // instruction each (inline PTX: the compiler otherwise builds a P2R / SEL / LOP3 chain, ~18 per step)
        T m, v;
        A::min_first(m, dw, y, x, 1u << (8 * U + 2 * r));      // m = y < x ? y : x
        A::min_first(v, dw, z, m, 2u << (8 * U + 2 * r));      // v = z < m ? z : m
        diag = L.left[r];
        L.left[r] = act ? v : L.left[r];
        upv = v;
        if (DENSE) {
            const int64_t i = (int64_t)band * kBandRows + lane * kRows + r;
            if (act && i < pm.M) {
                args.dense_cost[i * N + j] = c[r];
                args.dense_acc[i * N + j] = (T)((double)v + base);
            }
        }
    }
    L.up_prev = act ? up : L.up_prev;
    L.bottom = L.left[kRows - 1];
    if (feeds_next && lane == 31 && act) obuf[U] = L.bottom;
#pragma unroll
    for (int r = 0; r < kRows; r++) c[r] = cn[r];
}

template <typename T, bool DENSE>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 3) dtw_wavefront_kernel(const DtwArgs<T> args)
{
    using A = Arith<T>;
    constexpr int S = ColStride<T>::value;
    constexpr uint32_t kChunkBytes = kChunkCols * S * sizeof(T);
    constexpr uint32_t kMirrorBytes = kMirrorCols * S * sizeof(T);
    static_assert(kMirrorBytes % 16 == 0, "bulk copies move multiples of 16 bytes");
    extern __shared__ __align__(128) unsigned char s_dyn[];
    const int lane = threadIdx.x & 31;
    const int w = threadIdx.x >> 5;
    WarpSmem<T> &sm = reinterpret_cast<WarpSmem<T> *>(s_dyn)[w];
    // 32-bit shared-space addresses for the per-step accesses (keeps address arithmetic out of the hot loop)
    const uint32_t ring_s = smem_u32(sm.ring), ubuf_s = smem_u32(sm.ubuf);
    const unsigned full = 0xffffffffu;

    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < kRingSlots; q++) mbar_init(&sm.mbar[q], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    uint32_t phase_bits = 0;      // bit q: mbarrier phase parity of ring slot q's next completed fill

    for (;;) {
        int q = 0;
        if (lane == 0) q = atomicAdd(args.ticket, 1);
        q = __shfl_sync(full, q, 0);
        if (q >= args.n_items) break;
        const DtwItem it = args.items[q];
        const DtwPair pm = args.pairs[it.pair];
        const int band = it.band;
        const int N = pm.N;
        const bool feeds_next = band + 1 < pm.nbands;
        const T *ap = args.a + pm.a_off;
        const T *btp = args.bt + pm.bt_off;
        uint4 *brow_cur = args.brow + pm.brow_off + (int64_t)(band & 1) * pm.nsteps;
        const uint4 *brow_prev = args.brow + pm.brow_off + (int64_t)((band + 1) & 1) * pm.nsteps;
        // tag = launch epoch (1..250, changes every launch) | band: a stale record of the previous launch or of
        // band-2 in the same ring slot can never be mistaken for the one this band waits for
        const uint32_t tag_cur = (args.epoch << 24) | (uint32_t)band;
        const uint32_t tag_prev = (args.epoch << 24) | (uint32_t)(band - 1);
        uint4 *dirp = args.dir + pm.dir_off + (int64_t)band * 32 + lane;
        const int nchunks = pm.nchunks;

        // stage chunk c of packed seq_b into ring slot (c & 3): one TMA bulk copy
        auto stage = [&](int c) {
            if (c < nchunks && lane == 0) {
                const int slot = c & (kRingSlots - 1);
                uint64_t *bar = &sm.mbar[slot];
                mbar_expect_tx(bar, slot == 0 ? kChunkBytes + kMirrorBytes : kChunkBytes);
                bulk_g2s(sm.ring + slot * (kChunkCols * S), btp + (int64_t)c * (kChunkCols * S), kChunkBytes, bar);
                // slot 0 again behind the ring (first kMirrorCols columns), same barrier
                if (slot == 0)
                    bulk_g2s(sm.ring + kRingSlots * kChunkCols * S, btp + (int64_t)c * (kChunkCols * S), kMirrorBytes, bar);
            }
        };
        __syncwarp();          // previous band's readers are done with every slot
        stage(0);
        stage(1);

        Lane<T> L;
        {
            const int r0 = band * kBandRows + lane * kRows;
#pragma unroll
            for (int k = 0; k < kF; k++)
#pragma unroll
                for (int r = 0; r < kRows; r++)
                    L.ar[r][k] = (r0 + r < pm.M) ? __ldg(ap + (int64_t)k * pm.M + r0 + r) : (T)0;
#pragma unroll
            for (int r = 0; r < kRows; r++) L.left[r] = A::inf();
            L.up_prev = A::inf();
            L.bottom = A::inf();
            if (args.leftb != nullptr) {
                // column stripe of a longer pair (K4): column -1 of this stripe is the last column of the
                // stripe to the left.  Wait until that stripe (another GPU, or an earlier launch) has
                // finished this band, then take acc[i, c0-1] as `left` and acc[i-1, c0-1] as the first diag.
                if (args.in_flag != nullptr) {
                    // Bounded wait: a neighbour that never launches (crashed rank) must not hang this
                    // GPU.  After kStripeWaitNs the launch is marked failed (ticket[1]) and acc_end is
                    // poisoned with NaN; the host side turns that into an error on every rank.
                    volatile int *failed = args.ticket + 1;
                    const unsigned long long t0 = globaltimer_ns();
                    while (afs::ld_acquire_sys(args.in_flag + band) == 0) {
                        __nanosleep(200);
                        if (*failed != 0) break;
                        if (globaltimer_ns() - t0 > args.wait_ns) {
                            *failed = 1;
                            args.acc_end[0] = __longlong_as_double(0x7ff8000000000000ll);
                            break;
                        }
                    }
                }
#pragma unroll
                for (int r = 0; r < kRows; r++)
                    if (r0 + r < pm.M) L.left[r] = (T)__ldcg(args.leftb + r0 + r);
                if (r0 > 0 && r0 - 1 < pm.M) L.up_prev = (T)__ldcg(args.leftb + r0 - 1);
            }
            // dtw.py:20-21: acc[0,0] = cost[0,0], back = 2.  A virtual diagonal neighbour of
            // -cost[0,0] makes (0,0) an ordinary cell: fma(2, c, -c) == c exactly, code 2; the
            // slot (up_prev) is overwritten right after, so nothing else ever sees it.
            if (band == 0 && lane == 0 && args.leftb == nullptr) {
                T s = A::mul(L.ar[0][0], __ldg(btp));
#pragma unroll
                for (int k = 1; k < kF; k++) s = A::fma(L.ar[0][k], __ldg(btp + k), s);
                L.up_prev = -A::sub((T)1, s);
            }
        }
        sm.ubuf[lane] = A::inf();          // band 0: nothing above the first row
        __syncwarp();

        // fp32 mode: lane state holds OFFSETS against a per-band fp64 base that is moved every 32
        // steps (min-plus recurrences are shift invariant), so offsets stay O(band height) and the
        // accumulated cost keeps ~1e-7 relative accuracy over 4e4 additions.  fp64 mode: base == 0.
        // Slack between consecutive bands of a pair.  A band can follow the one above at a distance of two
        // groups (64 steps), but a band running AT that distance stalls on every hiccup of its producer, and
        // the stall propagates down the whole chain of in-flight bands.  Starting `slack` groups later costs
        // that many steps once per band of the first wave and lets the chain absorb jitter afterwards
        // (later waves inherit the spacing: slots free up in the rhythm the bands finish).
        if (band > 0 && args.slack_groups > 0) {
            const int wcol = min(N - 1, 31 + 32 * args.slack_groups);
            if (lane == 0) {
                for (;;) {
                    const uint4 r = ld_record(brow_prev + wcol);
                    if (r.y == tag_prev && r.w == tag_prev) break;
                    __nanosleep(200);
                }
            }
            __syncwarp();
        }
        double base = 0.0;
        uint4 pref = make_uint4(0u, 0u, 0u, 0u);
        // software pipeline prologue: costs of column 0 (chunk 0 must have landed)
        T cst[kRows];
        mbar_wait(&sm.mbar[0], phase_bits & 1u);
        phase_bits ^= 1u;
        dtw_cost<T, 0>(L, ring_s + ((0 - lane) & (kRingSlots * kChunkCols - 1)) * (S * (int)sizeof(T)), cst);
        uint32_t d0 = 0, d1 = 0, d2 = 0, d3 = 0;
        for (int s0 = 0; s0 < pm.nsteps; s0 += 32) {
            const int c0 = s0 >> 5;
            if (sizeof(T) == 4 && s0 > 0) {
                T lo = L.left[0];
#pragma unroll
                for (int r = 1; r < kRows; r++) lo = fminf((float)lo, (float)L.left[r]);
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) lo = fminf((float)lo, (float)__shfl_xor_sync(full, lo, off));
                if (lo < A::inf()) {
                    base += (double)lo;
#pragma unroll
                    for (int r = 0; r < kRows; r++) L.left[r] -= lo;      // inf - lo == inf
                    L.up_prev -= lo;
                    L.bottom -= lo;
                }
            }
            // ---- chunk c0+1 of seq_b must have landed (the last step of this group already computes the
            // costs of the next group's first column); then put chunk c0+2 in flight ----
            if (c0 + 1 < nchunks) {
                const int slot = (c0 + 1) & (kRingSlots - 1);
                mbar_wait(&sm.mbar[slot], (phase_bits >> slot) & 1u);
                phase_bits ^= 1u << slot;      // every fill of a slot is waited for exactly once, in order
            }
            __syncwarp();                      // all lanes are done with chunk c0-2 (same slot as c0+2)
            stage(c0 + 2);
            if (band > 0 && s0 < N) {
                // Columns [s0, s0+32) of the band above.  Every hand-off record validates itself: the two
                // halves of the fp64 value each travel with the producer band's tag, so one 128-bit load
                // either shows both tags (value complete) or is retried — no flag, no fence, no L1 flush.
                const int col = s0 + lane;
                const bool want = col < N;
                uint4 rec = pref;                 // issued one group ago (all zero before the first group)
                bool ok = !want || ((rec.y == tag_prev) && (rec.w == tag_prev));
                for (;;) {
                    if (!ok) {
                        rec = ld_record(brow_prev + col);
                        ok = (rec.y == tag_prev) && (rec.w == tag_prev);
                    }
                    if (__all_sync(full, ok)) break;
                    __nanosleep(20);
                }
                const double vabs = __hiloint2double((int)rec.z, (int)rec.x);
                if (sizeof(T) == 4 && s0 == 0) base = __shfl_sync(full, vabs, 0);   // start from the level of the row above
                T v = want ? (T)(vabs - base) : A::inf();
                __syncwarp();
                sm.ubuf[lane] = v;
                __syncwarp();
                // next group's record goes in flight now and is checked in 32 steps; if the band above is
                // not that far ahead yet the check fails and the poll loop above takes over
                pref = make_uint4(0u, 0u, 0u, 0u);
                if (col + 32 < N) pref = ld_record(brow_prev + col + 32);
            }
            auto run_group = [&](auto all_tag) {
                constexpr bool ALL = decltype(all_tag)::value;
#pragma unroll 1
                for (int g4 = 0; g4 < 8; g4++) {
                    const int s = s0 + g4 * 4;
                    uint32_t dw = 0;
                    const uint32_t ubuf_g = ubuf_s + g4 * 4 * (int)sizeof(T);     // s0 is a multiple of 32: slot = s & 31
                    const uint32_t ring_g = ring_s + ((s - lane) & (kRingSlots * kChunkCols - 1)) * (S * (int)sizeof(T));
                    T *obuf_g = sm.obuf + g4 * 4;
                    dtw_step<T, DENSE, 0, ALL>(L, s + 0, lane, band, pm, ring_g, ubuf_g, obuf_g, feeds_next, dw, args, base, cst);
                    dtw_step<T, DENSE, 1, ALL>(L, s + 1, lane, band, pm, ring_g, ubuf_g, obuf_g, feeds_next, dw, args, base, cst);
                    dtw_step<T, DENSE, 2, ALL>(L, s + 2, lane, band, pm, ring_g, ubuf_g, obuf_g, feeds_next, dw, args, base, cst);
                    dtw_step<T, DENSE, 3, ALL>(L, s + 3, lane, band, pm, ring_g, ubuf_g, obuf_g, feeds_next, dw, args, base, cst);
                    d0 = d1; d1 = d2; d2 = d3; d3 = dw;
                    if ((g4 & 3) == 3) {
                        const int cbp = s >> 4;
                        __stcs(dirp + (int64_t)cbp * pm.gpad, make_uint4(d0, d1, d2, d3));
                    }
                }
            };
            // steady state: every lane's column s - lane lies in [1, N) for all 32 steps of the group
            if (s0 >= 32 && s0 + 32 <= N) run_group(std::true_type{});
            else run_group(std::false_type{});
            if (feeds_next) {
                // slots 0..31 hold lane 31's columns s0-31 .. s0
                __syncwarp();
                const int col = s0 - 31 + lane;
                if (col >= 0 && col < N) {
                    const double vabs = (double)sm.obuf[lane] + base;
                    st_record(brow_cur + col, make_uint4((uint32_t)__double2loint(vabs), tag_cur, (uint32_t)__double2hiint(vabs), tag_cur));
                }
                __syncwarp();      // obuf is rewritten by lane 31 in the next group
            }
        }
        // after the sweep every lane's left[] holds column N-1: acc_cost[M-1, N-1] sits in one of them
        {
            const int rl = pm.M - 1 - (band * kBandRows + lane * kRows);
            if (rl >= 0 && rl < kRows) {
                T e = L.left[0];
#pragma unroll
                for (int r = 1; r < kRows; r++) if (rl == r) e = L.left[r];
                double ev = (double)e + base;
                if (args.in_flag != nullptr && *(volatile int *)(args.ticket + 1) != 0)
                    ev = __longlong_as_double(0x7ff8000000000000ll);     // a stripe wait timed out: poisoned
                args.acc_end[it.pair] = ev;
            }
        }
        if (args.rightb != nullptr) {
            // K4: hand the stripe's last column of this band to the stripe on the right (possibly
            // peer-mapped memory of the next GPU over NVLink), then raise its flag at system scope
            const int r0 = band * kBandRows + lane * kRows;
#pragma unroll
            for (int r = 0; r < kRows; r++)
                if (r0 + r < pm.M) args.rightb[r0 + r] = (double)L.left[r] + base;
            __threadfence_system();
            __syncwarp();
            if (lane == 0 && args.out_flag != nullptr) afs::st_release_sys(args.out_flag + band, 1);
        }
    }
}

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Backtrack helper.  The walk leaves the tile (band, cb8) to the left (band, cb8-1) or upwards; going up into the
// band above, the column skew (g & 31) jumps from 0 to 31, so the tile above is (band-1, cb8) or (band-1, cb8+1).
// Pull those three tiles (written with a streaming hint by K2, so they sit in HBM) into L2 while lane 0 walks the
// current one.  A tile is 8 rows of 512 contiguous bytes = 32 lines of 128 B: one line per lane.
__device__ __forceinline__ void prefetch_next_tiles(const uint4 *dunits, int gpad, int ncbp, int band, int cb8, int lane)
{
    const int q = lane >> 2, line8 = (lane & 3) * 8;
    const int cbl = (cb8 - 1) * 8 + q, cbc = cb8 * 8 + q, cbr = (cb8 + 1) * 8 + q;
    if (cb8 > 0) prefetch_l2(dunits + (int64_t)cbl * gpad + band * 32 + line8);
    if (band > 0) {
        if (cbc < ncbp) prefetch_l2(dunits + (int64_t)cbc * gpad + (band - 1) * 32 + line8);
        if (cbr < ncbp) prefetch_l2(dunits + (int64_t)cbr * gpad + (band - 1) * 32 + line8);
    }
}

// K3 for one stripe of a column-striped pair: walk from (i0, j0) (stripe-local column) until the
// path leaves the stripe through its left edge (or reaches (0,0) in the first stripe).  Points are
// written back-to-front with GLOBAL column indices (col0 + j).  exit_i = row at which the walk
// continues in the stripe to the left (its last column), or -1 when (0,0) was reached.
__global__ void __launch_bounds__(32) dtw_backtrack_stripe_kernel(const DtwPair *pairs, const uint4 *dir, int i0, int j0, int col0,
                                                                  int first, int32_t *path, int32_t *out3)
{
    // same tile-staged, register-decoded walk as dtw_backtrack_kernel (below), one warp
    __shared__ uint4 s_tile[8 * 32];
    const int lane = threadIdx.x;
    const DtwPair pm = pairs[0];
    const uint4 *dunits = dir + pm.dir_off;
    int2 *out = reinterpret_cast<int2 *>(path) + pm.path_off;
    const int ncbp = pm.nsteps >> 4;
    int i = i0, j = j0;
    int pos = pm.path_cap - 1;
    int exit_i = -1;
    int done = (first && i == 0 && j == 0) ? 1 : 0;
    if (lane == 0) out[pos] = make_int2(i, j + col0);
    while (!done) {
        const int band = i >> 7;
        const int cb8 = (j + ((i >> 2) & 31)) >> 7;
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int cbp = cb8 * 8 + q;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (cbp < ncbp) v = __ldcs(dunits + (int64_t)cbp * pm.gpad + band * 32 + lane);
            s_tile[q * 32 + lane] = v;
        }
        prefetch_next_tiles(dunits, pm.gpad, ncbp, band, cb8, lane);
        __syncwarp();
        if (lane == 0) {
            while (!done) {
                const int g = i >> 2;
                const int jj = j + (g & 31);
                if ((i >> 7) != band || (jj >> 7) != cb8) break;
                const uint4 u = s_tile[((jj >> 4) & 7) * 32 + (g & 31)];
                const uint64_t lo = (uint64_t)u.x | ((uint64_t)u.y << 32), hi = (uint64_t)u.z | ((uint64_t)u.w << 32);
                int cs = jj & 15, r = i & 3;
                for (;;) {
                    const uint64_t sel = (cs & 8) ? hi : lo;
                    const uint32_t code = (uint32_t)(sel >> (((cs & 7) << 3) + 2 * r)) & 3u;
                    const int di = code != 0u, dj = code != 1u;      // 0: left, 1: up, 2: diag
                    i -= di; r -= di;
                    j -= dj; cs -= dj;
                    if (j < 0) { exit_i = i; done = 1; break; }      // continues in the stripe to the left
                    pos -= 1;
                    out[pos] = make_int2(i, j + col0);
                    if (first && (i | j) == 0) { done = 1; break; }
                    if (r < 0 || cs < 0) break;
                }
            }
        }
        i = __shfl_sync(0xffffffffu, i, 0);
        j = __shfl_sync(0xffffffffu, j, 0);
        done = __shfl_sync(0xffffffffu, done, 0);
        __syncwarp();
    }
    if (lane == 0) {
        out3[0] = pos;
        out3[1] = pm.path_cap - pos;
        out3[2] = exit_i;
    }
}

// K3: backtrack over the direction map (dtw.py:43-52).  One warp per pair; lane 0
// walks, writing the path back-to-front so that no reversal pass is needed.
// The walk is a chain of dependent loads; to keep HBM latency off that chain the warp stages a
// whole tile of the direction map — one band (128 rows) x 128 skewed columns = 8 x 512 contiguous
// bytes — into shared memory with coalesced 16-byte loads, and lane 0 then walks inside the tile
// (>= ~100 steps per tile on a diagonal-ish path).
constexpr int kBtWarps = 2;
__global__ void __launch_bounds__(kBtWarps * 32) dtw_backtrack_kernel(const DtwPair *pairs, int n_pairs, const uint4 *dir,
                                                                      int32_t *path, int32_t *path_start, int32_t *path_len)
{
    __shared__ uint4 s_tile[kBtWarps][8 * 32];        // [cbp & 7][g & 31]
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = blockIdx.x * kBtWarps + w;
    if (p >= n_pairs) return;
    const DtwPair pm = pairs[p];
    const uint4 *dunits = dir + pm.dir_off;
    int2 *out = reinterpret_cast<int2 *>(path) + pm.path_off;
    const int ncbp = pm.nsteps >> 4;
    int i = pm.M - 1, j = pm.N - 1;
    int pos = pm.path_cap - 1;
    if (lane == 0) out[pos] = make_int2(i, j);
    while (i > 0 || j > 0) {                          // uniform: lane 0's (i, j) is broadcast below
        const int band = i >> 7;
        const int cb8 = (j + ((i >> 2) & 31)) >> 7;   // block of 8 skewed column units
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int cbp = cb8 * 8 + q;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (cbp < ncbp) v = __ldcs(dunits + (int64_t)cbp * pm.gpad + band * 32 + lane);
            s_tile[w][q * 32 + lane] = v;
        }
        prefetch_next_tiles(dunits, pm.gpad, ncbp, band, cb8, lane);
        __syncwarp();
        if (lane == 0) {
            while (i > 0 || j > 0) {
                const int g = i >> 2;
                const int jj = j + (g & 31);
                if ((i >> 7) != band || (jj >> 7) != cb8) break;
                // one 16-byte unit (4 rows x 16 skewed columns) into registers; walk inside it with bit ops only
                const uint4 u = s_tile[w][((jj >> 4) & 7) * 32 + (g & 31)];
                const uint64_t lo = (uint64_t)u.x | ((uint64_t)u.y << 32), hi = (uint64_t)u.z | ((uint64_t)u.w << 32);
                int cs = jj & 15, r = i & 3;
                for (;;) {
                    const uint64_t sel = (cs & 8) ? hi : lo;
                    const uint32_t code = (uint32_t)(sel >> (((cs & 7) << 3) + 2 * r)) & 3u;
                    const int di = code != 0u, dj = code != 1u;      // 0: left, 1: up, 2: diag
                    i -= di; r -= di;
                    j -= dj; cs -= dj;
                    pos -= 1;
                    out[pos] = make_int2(i, j);
                    if (r < 0 || cs < 0 || (i | j) == 0) break;
                }
            }
        }
        i = __shfl_sync(0xffffffffu, i, 0);
        j = __shfl_sync(0xffffffffu, j, 0);
        pos = __shfl_sync(0xffffffffu, pos, 0);
    }
    if (lane == 0) {
        path_start[p] = pos;
        path_len[p] = pm.path_cap - pos;
    }
}

}  // namespace

struct afs_dtw_plan {
    int n_pairs = 0;
    int dtype = AFS_F64;
    std::vector<DtwPair> pairs;
    std::vector<DtwItem> items;
    DtwPair *d_pairs = nullptr;
    DtwItem *d_items = nullptr;
    size_t dir_bytes = 0, brow_bytes = 0, bt_bytes = 0, ticket_bytes = 0;
    uint32_t epoch = 0;
    const void *last_ws = nullptr;      // workspace whose hand-off area has been cleared once
    int64_t total_path = 0;
    int total_bands = 0;
    int max_cols_pad = 0;
};

extern "C" {

int afs_dtw_plan_create(afs_dtw_plan **out, int n_pairs, const int64_t *h_len_a, const int64_t *h_len_b,
                        const int64_t *h_off_a, const int64_t *h_off_b, int n_features, int dtype)
{
    if (!out || n_pairs <= 0 || !h_len_a || !h_len_b || !h_off_a || !h_off_b)
        return afs::fail(AFS_ERR_INVALID, "afs_dtw_plan_create: null argument or n_pairs <= 0");
    if (n_features != kF) return afs::fail(AFS_ERR_UNSUPPORTED, "afs_dtw: n_features must be 12 (got %d)", n_features);
    if (dtype != AFS_F64 && dtype != AFS_F32) return afs::fail(AFS_ERR_INVALID, "afs_dtw: bad dtype %d", dtype);
    afs_dtw_plan *pl = new afs_dtw_plan();
    pl->n_pairs = n_pairs;
    pl->dtype = dtype;
    pl->pairs.resize(n_pairs);
    const int stride = dtype == AFS_F64 ? ColStride<double>::value : ColStride<float>::value;
    const size_t esz = dtype == AFS_F64 ? 8 : 4;
    int64_t dir_units = 0, brow_elems = 0, bt_elems = 0, path_pairs = 0;
    int prog = 0, max_bands = 0;
    for (int p = 0; p < n_pairs; p++) {
        const int64_t M = h_len_a[p], N = h_len_b[p];
        if (M <= 0 || N <= 0 || M > (1 << 30) || N > (1 << 30)) {
            delete pl;
            return afs::fail(AFS_ERR_INVALID, "afs_dtw: pair %d has invalid lengths %lld x %lld", p, (long long)M, (long long)N);
        }
        DtwPair &q = pl->pairs[p];
        q.a_off = h_off_a[p];
        q.b_off = h_off_b[p];
        q.M = (int32_t)M;
        q.N = (int32_t)N;
        q.nbands = (int32_t)((M + kBandRows - 1) / kBandRows);
        q.gpad = q.nbands * 32;
        q.nsteps = (int32_t)((N + 31 + 31) / 32 * 32);
        q.nchunks = (int32_t)((N + kChunkCols - 1) / kChunkCols);
        q.dir_off = dir_units;
        q.brow_off = brow_elems;
        q.bt_off = bt_elems;
        q.path_off = path_pairs;
        q.band0 = prog;
        q.path_cap = (int32_t)(M + N);
        dir_units += (int64_t)(q.nsteps / 16) * q.gpad;
        brow_elems += 2 * (int64_t)q.nsteps;
        bt_elems += (int64_t)q.nchunks * kChunkCols * stride;     // whole chunks: every bulk copy is full and 16 B aligned
        path_pairs += q.path_cap;
        prog += q.nbands;
        if (q.nbands > max_bands) max_bands = q.nbands;
        if (q.nchunks * kChunkCols > pl->max_cols_pad) pl->max_cols_pad = q.nchunks * kChunkCols;
    }
    // Ticket order.  Bands are interleaved band-major inside groups of `group` pairs, group after group; every dependency
    // (band b-1 of the same pair) has a smaller ticket.  Whole batch as one group (default): each pair has (warp slots /
    // pairs) bands in flight, and in steady state they are spread evenly over the columns — about a millisecond apart — so
    // every band re-reads its pair's packed seq_b from DRAM once the batch's seq_b (2.2 MB per 20k-column pair) exceeds L2:
    // 118 GB of DRAM traffic per launch at 256 pairs against 26.6 GB algorithmic.  Small groups keep a pair's bands close
    // together (64 steps, ~20 us, at group = 1) and the re-reads become L2 hits — measured at 256 pairs (GCUPS, DRAM GB):
    // batch 645 / 118, 64 pairs 625 / 75, 32: 618 / 51, 16: 612 / 34, 1 (pair-major): 507 / 27.1 = 1.02 x algorithmic.
    // The kernel is bound by the FP64 pipe, not by DRAM (0.76 TB/s), and every grouping costs ramp at the group edges, so
    // the default stays the whole batch; AFS_DTW_ORDER=g<N> | p selects the others.
    // group: pairs per interleaved group (1 = pair-major, n_pairs = band-major over the whole batch)
    int group = n_pairs;
    if (const char *o = getenv("AFS_DTW_ORDER")) {
        if (o[0] == 'p') group = 1;
        else if (o[0] == 'g') group = atoi(o + 1) > 0 ? atoi(o + 1) : n_pairs;
    }
    for (int p0 = 0; p0 < n_pairs; p0 += group) {
        const int p1 = p0 + group < n_pairs ? p0 + group : n_pairs;
        int gb = 0;
        for (int p = p0; p < p1; p++) gb = pl->pairs[p].nbands > gb ? pl->pairs[p].nbands : gb;
        for (int b = 0; b < gb; b++)
            for (int p = p0; p < p1; p++)
                if (b < pl->pairs[p].nbands) pl->items.push_back(DtwItem{p, b});
    }
    pl->total_bands = prog;
    pl->total_path = path_pairs;
    pl->dir_bytes = afs::align_up((size_t)dir_units * 16, 256);
    pl->brow_bytes = afs::align_up((size_t)brow_elems * sizeof(uint4), 256);
    pl->bt_bytes = afs::align_up((size_t)bt_elems * esz, 256);
    pl->ticket_bytes = 256;    // the band ticket (one int)
    cudaError_t e = cudaMalloc(&pl->d_pairs, sizeof(DtwPair) * n_pairs);
    if (e == cudaSuccess) e = cudaMalloc(&pl->d_items, sizeof(DtwItem) * pl->items.size());
    if (e == cudaSuccess) e = cudaMemcpy(pl->d_pairs, pl->pairs.data(), sizeof(DtwPair) * n_pairs, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(pl->d_items, pl->items.data(), sizeof(DtwItem) * pl->items.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cudaFree(pl->d_pairs);
        cudaFree(pl->d_items);
        delete pl;
        return afs::fail(AFS_ERR_CUDA, "afs_dtw_plan_create: %s", cudaGetErrorString(e));
    }
    *out = pl;
    return AFS_OK;
}

int afs_dtw_plan_destroy(afs_dtw_plan *pl)
{
    if (!pl) return AFS_OK;
    cudaFree(pl->d_pairs);
    cudaFree(pl->d_items);
    delete pl;
    return AFS_OK;
}

int afs_dtw_plan_workspace_bytes(const afs_dtw_plan *pl, size_t *bytes)
{
    if (!pl || !bytes) return afs::fail(AFS_ERR_INVALID, "afs_dtw_plan_workspace_bytes: null argument");
    *bytes = pl->dir_bytes + pl->brow_bytes + pl->bt_bytes + pl->ticket_bytes;
    return AFS_OK;
}

int afs_dtw_plan_path_layout(const afs_dtw_plan *pl, int pair, int64_t *offset, int64_t *capacity)
{
    if (!pl || pair < -1 || pair >= pl->n_pairs) return afs::fail(AFS_ERR_INVALID, "afs_dtw_plan_path_layout: bad pair");
    if (pair == -1) {   // total size of the path area
        if (offset) *offset = 0;
        if (capacity) *capacity = pl->total_path;
        return AFS_OK;
    }
    if (offset) *offset = pl->pairs[pair].path_off;
    if (capacity) *capacity = pl->pairs[pair].path_cap;
    return AFS_OK;
}

}  // extern "C"

static uint32_t next_launch_epoch()
{
    static std::atomic<uint32_t> g_epoch{0};
    return g_epoch.fetch_add(1u) % 250u + 1u;
}

template <typename T>
static int launch_accumulate(afs_dtw_plan *pl, const void *d_a, const void *d_b, void *ws, double *d_acc_end,
                             void *dense_cost, void *dense_acc, cudaStream_t st, const double *leftb = nullptr,
                             double *rightb = nullptr, const int *in_flag = nullptr, int *out_flag = nullptr)
{
    char *base = static_cast<char *>(ws);
    DtwArgs<T> args;
    args.a = static_cast<const T *>(d_a);
    args.b = static_cast<const T *>(d_b);
    args.pairs = pl->d_pairs;
    args.items = pl->d_items;
    args.n_items = (int)pl->items.size();
    args.dir = reinterpret_cast<uint4 *>(base);
    args.brow = reinterpret_cast<uint4 *>(base + pl->dir_bytes);
    args.bt = reinterpret_cast<T *>(base + pl->dir_bytes + pl->brow_bytes);
    args.ticket = reinterpret_cast<int *>(base + pl->dir_bytes + pl->brow_bytes + pl->bt_bytes);
    // Every launch starts from a hand-off area without valid records: the tags only carry an 8-bit launch epoch, so
    // a relaunch of this plan after exactly 250 k launches elsewhere in the process would otherwise meet records of
    // its own previous launch that carry the tags it waits for.  The clear costs microseconds (2 rows of 16-byte
    // records per pair) against a launch of milliseconds.
    AFS_CUDA(cudaMemsetAsync(args.brow, 0, pl->brow_bytes, st));
    pl->last_ws = ws;
    // the epoch still separates launches of different plans that were (against the header's advice) pointed at the
    // same workspace while both are in flight; one counter for the whole process (fp32 and fp64 plans alike)
    pl->epoch = next_launch_epoch();
    args.epoch = pl->epoch;
    args.acc_end = d_acc_end;
    args.dense_cost = static_cast<T *>(dense_cost);
    args.dense_acc = static_cast<T *>(dense_acc);
    args.leftb = leftb;
    args.rightb = rightb;
    args.in_flag = in_flag;
    args.out_flag = out_flag;
    args.wait_ns = kStripeWaitNs;
    args.slack_groups = kSlackGroups;
    if (const char *sg = getenv("AFS_DTW_SLACK")) args.slack_groups = atoi(sg);      // tuning knob
    if (in_flag != nullptr) {
        const char *ms = getenv("AFS_STRIPE_WAIT_MS");
        if (ms != nullptr && atoll(ms) > 0) args.wait_ns = (unsigned long long)atoll(ms) * 1000000ull;
    }
    AFS_CUDA(cudaMemsetAsync(args.ticket, 0, pl->ticket_bytes, st));
    {
        const int threads = 128;
        dim3 grid((pl->max_cols_pad + threads - 1) / threads, pl->n_pairs);
        dtw_pack_b_kernel<T><<<grid, threads, 0, st>>>(args.b, args.bt, pl->d_pairs, pl->n_pairs);
        afs::count_launch();
        AFS_CUDA(cudaGetLastError());
    }
    const bool dense = dense_cost != nullptr;
    auto kern = dense ? dtw_wavefront_kernel<T, true> : dtw_wavefront_kernel<T, false>;
    const size_t smem = sizeof(WarpSmem<T>) * kWarpsPerBlock;
    AFS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    AFS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kWarpsPerBlock * 32, smem));
    if (occ < 1) return afs::fail(AFS_ERR_CUDA, "dtw kernel does not fit on an SM");
    int blocks = afs::sm_count() * occ;
    const int need = (args.n_items + kWarpsPerBlock - 1) / kWarpsPerBlock;
    if (blocks > need) blocks = need;
    kern<<<blocks, kWarpsPerBlock * 32, smem, st>>>(args);
    afs::count_launch();
    AFS_CUDA(cudaGetLastError());
    return AFS_OK;
}

extern "C" {

int afs_dtw_accumulate(afs_dtw_plan *pl, const void *d_a, const void *d_b, void *d_workspace, double *d_acc_end,
                       void *d_dense_cost, void *d_dense_acc, void *stream)
{
    if (!pl || !d_a || !d_b || !d_workspace || !d_acc_end)
        return afs::fail(AFS_ERR_INVALID, "afs_dtw_accumulate: null argument");
    if ((d_dense_cost == nullptr) != (d_dense_acc == nullptr))
        return afs::fail(AFS_ERR_INVALID, "afs_dtw_accumulate: dense cost and acc must be given together");
    if (d_dense_cost && pl->n_pairs != 1)
        return afs::fail(AFS_ERR_UNSUPPORTED, "afs_dtw_accumulate: dense outputs need a single-pair plan");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (pl->dtype == AFS_F64)
        return launch_accumulate<double>(pl, d_a, d_b, d_workspace, d_acc_end, d_dense_cost, d_dense_acc, st);
    return launch_accumulate<float>(pl, d_a, d_b, d_workspace, d_acc_end, d_dense_cost, d_dense_acc, st);
}

int afs_dtw_accumulate_stripe(afs_dtw_plan *pl, const void *d_a, const void *d_b_stripe, void *d_workspace,
                              double *d_acc_end, const double *d_leftb, const int *d_in_flag, double *d_rightb,
                              int *d_out_flag, void *stream)
{
    if (!pl || !d_a || !d_b_stripe || !d_workspace || !d_acc_end)
        return afs::fail(AFS_ERR_INVALID, "afs_dtw_accumulate_stripe: null argument");
    if (pl->n_pairs != 1 || pl->dtype != AFS_F64)
        return afs::fail(AFS_ERR_UNSUPPORTED, "afs_dtw_accumulate_stripe: needs a single-pair fp64 plan (M x stripe columns)");
    return launch_accumulate<double>(pl, d_a, d_b_stripe, d_workspace, d_acc_end, nullptr, nullptr,
                                     static_cast<cudaStream_t>(stream), d_leftb, d_rightb, d_in_flag, d_out_flag);
}

int afs_dtw_backtrack_stripe(afs_dtw_plan *pl, const void *d_workspace, int start_i, int start_j, int col0,
                             int first_stripe, int32_t *d_path, int32_t *d_out3, void *stream)
{
    if (!pl || !d_workspace || !d_path || !d_out3) return afs::fail(AFS_ERR_INVALID, "afs_dtw_backtrack_stripe: null argument");
    if (pl->n_pairs != 1) return afs::fail(AFS_ERR_UNSUPPORTED, "afs_dtw_backtrack_stripe: single-pair plan only");
    if (start_i < 0 || start_i >= pl->pairs[0].M || start_j < 0 || start_j >= pl->pairs[0].N)
        return afs::fail(AFS_ERR_INVALID, "afs_dtw_backtrack_stripe: start cell outside the stripe");
    dtw_backtrack_stripe_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(
        pl->d_pairs, static_cast<const uint4 *>(d_workspace), start_i, start_j, col0, first_stripe, d_path, d_out3);
    afs::count_launch();
    AFS_CUDA(cudaGetLastError());
    return AFS_OK;
}

// ---- CUDA IPC helpers: an exchange block (boundary column + flags) one process allocates and its
// right-hand neighbour rank maps, so that the stripe kernel can store into it over NVLink ----
int afs_ipc_alloc(size_t bytes, void **d_ptr, void *handle64)
{
    if (!d_ptr || !handle64 || bytes == 0) return afs::fail(AFS_ERR_INVALID, "afs_ipc_alloc: bad argument");
    AFS_CUDA(cudaMalloc(d_ptr, bytes));
    AFS_CUDA(cudaMemset(*d_ptr, 0, bytes));
    cudaIpcMemHandle_t h;
    AFS_CUDA(cudaIpcGetMemHandle(&h, *d_ptr));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle64, &h, 64);
    return AFS_OK;
}

int afs_ipc_open(const void *handle64, void **d_ptr)
{
    if (!d_ptr || !handle64) return afs::fail(AFS_ERR_INVALID, "afs_ipc_open: bad argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    AFS_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return AFS_OK;
}

int afs_ipc_clear(void *d_ptr, size_t bytes, void *stream)
{
    if (!d_ptr) return afs::fail(AFS_ERR_INVALID, "afs_ipc_clear: null pointer");
    AFS_CUDA(cudaMemsetAsync(d_ptr, 0, bytes, static_cast<cudaStream_t>(stream)));
    return AFS_OK;
}

int afs_ipc_close(void *d_ptr)
{
    if (d_ptr) AFS_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return AFS_OK;
}

int afs_ipc_free(void *d_ptr)
{
    if (d_ptr) AFS_CUDA(cudaFree(d_ptr));
    return AFS_OK;
}

int afs_dtw_backtrack(afs_dtw_plan *pl, const void *d_workspace, int32_t *d_path, int32_t *d_path_start,
                      int32_t *d_path_len, void *stream)
{
    if (!pl || !d_workspace || !d_path || !d_path_start || !d_path_len)
        return afs::fail(AFS_ERR_INVALID, "afs_dtw_backtrack: null argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int wpb = 2;
    const int blocks = (pl->n_pairs + wpb - 1) / wpb;
    dtw_backtrack_kernel<<<blocks, wpb * 32, 0, st>>>(pl->d_pairs, pl->n_pairs, static_cast<const uint4 *>(d_workspace),
                                                      d_path, d_path_start, d_path_len);
    afs::count_launch();
    AFS_CUDA(cudaGetLastError());
    return AFS_OK;
}

}  // extern "C"

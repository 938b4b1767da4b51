// K1-TC — chroma on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// Replaces the same reference code as chroma.cu (chroma.py:44-75: framing, Hann window, rfft, |X|^2, 12 x 2049
// filterbank, L2 normalisation) with the real DFT of length 4096 written as two matrix products (4096 = 64 x 64):
//
//   n = 64 n1 + n2,  k = k1 + 64 k2
//   stage 1   Y[k1][n2]  = sum_n1 x[64 n1 + n2] W64^(n1 k1)            k1 = 0..32 (real input: the rest is conjugate)
//   twiddle   Y'[k1][n2] = W4096^(n2 k1) Y[k1][n2]
//   stage 2   X[k1 + 64 k2] = sum_n2 W64^(n2 k2) Y'[k1][n2]            k2 = 0..63 (bins above 2048 mirror the ones below)
//
// Both products run as tcgen05.mma (kind::f16, bf16 operands, fp32 accumulation in TMEM).  Every operand is split
// into two bf16 terms by truncation (x = hi + lo, 16 mantissa bits together) and a product is three MMAs
// (hi*hi + lo*hi + hi*lo): measured 1.2e-5 worst absolute error on normalised chroma of real audio against the
// float64 reference (tolerance 1e-4), independent of the signal's scale (bf16 keeps the float32 exponent).
//
// Kernel A (chroma_tc_spectrum_kernel), one CTA per SM, two independent 256-thread pipelines per CTA that run the
// same sequential program on alternate groups of 4 frames — while one waits for its MMAs the other converts,
// twiddles or squares, so the tensor pipe and the CUDA cores overlap without any cross-pipeline synchronisation:
//   convert   thread (frame, n2) loads its 64 samples x[64 n1 + n2] (coalesced across n2), applies the window, splits,
//             and stores them as the K-major A operand [128 rows = 2 frames x 64 n2][K = n1] (128-byte swizzle)
//   MMA 1     D1[(f, n2)][64] = A1 . F^T     (12 MMAs, N = 64: Re Y[0], Re Y[32], then Re/Im Y[k1], k1 = 1..31)
//   epilogue1 tcgen05.ld, twiddle in registers (the thread's n2 is fixed, so its twiddles are too), split, store as
//             the B operand of stage 2: row (frame, k1), K = (n2, re/im)
//   MMA 2     D2[(k2, re/im)][(frame, k1)] = G . Y'^T   (24 MMAs, N = 144; G lives in TMEM as the A operand)
//   epilogue2 squares; re^2 + im^2 meet through one shuffle between neighbouring lanes; split; the power spectrum
//             goes to a scratch buffer as two bf16 planes (L2 resident: the host runs the batch in chunks)
// Kernel B (chroma_tc_filterbank_kernel): [64 frames x (hi | lo) planes] x [bins x 12] as tcgen05.mma with the weight
// matrix resident in shared memory, then L2 normalisation and the feature-major store.
#include <cfloat>
#include <cmath>
#include <cstdlib>

#include "chroma_tc.cuh"
#include "tc05.cuh"

namespace {

constexpr int kNfft = 4096;
constexpr int kBins = 2049;
constexpr int kChroma = 12;

// ---------------- kernel A geometry ----------------
constexpr int kThreadsA = 512;                 // two pipelines of 256
constexpr int kRowsPerFrame = 36;              // 33 k1 rows + 3 pad rows (N of stage 2 must be a multiple of 16)
constexpr int kN1 = 64;                        // stage-1 N: 64 real outputs
constexpr int kN2 = 4 * kRowsPerFrame;         // stage-2 N: 144
constexpr int kPlaneWords = 9 * 128;           // power spectrum: 9 words per lane per frame and plane (2 bf16 each)
constexpr uint32_t kFPlane = 64 * 128;         // F^T image, one plane (hi or lo)
constexpr uint32_t kA1Plane = 128 * 128;       // stage-1 A operand, one plane
constexpr uint32_t kYBlk = kN2 * 128;          // stage-2 B operand: one 64-element K block of one plane
constexpr uint32_t kYPlane = 2 * kYBlk;        // K = 128
constexpr uint32_t kPipeBytes = 2 * kA1Plane + 2 * kYPlane;
constexpr uint32_t kSmemA = 2 * kFPlane + 2 * kPipeBytes;        // 229 376 B
// TMEM columns: G_hi [0,64) | G_lo [64,128) | pipeline 0 accumulators [128,272) | pipeline 1 [272,416)
constexpr uint32_t kColGhi = 0, kColGlo = 64, kColD0 = 128, kColDStride = kN2;

struct FrameMeta {
    long long base;     // element index of the frame's sample 0 in the audio array (may point before the track)
    int lo, hi;         // samples lo <= n < hi of the frame exist, the rest is the reference's zero padding
    int valid;          // frame index inside the chunk
};

struct SpectrumArgs {
    const void *audio;
    const int64_t *sample_off, *frame_off;
    int n_tracks;
    int64_t frame_begin, frame_end;            // this launch covers frames [begin, end)
    int hop, center_pad;
    const uint4 *f_img;                        // 16 KB: shared-memory images of F^T hi | lo
    const uint32_t *g_img;                     // [2][128][64] words: TMEM images of G hi | lo
    const float2 *tw;                          // [33][64]: W4096^(k1 n2) = (cos, -sin)
    const float *hann;                         // 4096
    uint32_t *p_hi, *p_lo;                     // [frame - frame_begin][kPlaneWords]
};

__device__ __forceinline__ void named_bar(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&v)[4])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
                 : "r"(taddr)
                 : "memory");
}

__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// where the frame `f` (global numbering) lives; `track` is a cursor that only moves forward
__device__ __forceinline__ FrameMeta locate_frame(const SpectrumArgs &a, int64_t f, int &track)
{
    FrameMeta m;
    m.base = 0; m.lo = 0; m.hi = 0; m.valid = 0;
    if (f >= a.frame_end) return m;
    while (track + 1 < a.n_tracks && a.frame_off[track + 1] <= f) track++;
    const int64_t s_begin = a.sample_off[track];
    const int64_t n_samp = a.sample_off[track + 1] - s_begin;
    const int64_t start = (f - a.frame_off[track]) * a.hop - (a.center_pad ? kNfft / 2 : 0);     // chroma.py:49 left zero pad
    m.base = s_begin + start;
    m.lo = start < 0 ? (int)(-start) : 0;
    const int64_t hi = n_samp - start;
    m.hi = hi > kNfft ? kNfft : (hi < 0 ? 0 : (int)hi);
    m.valid = 1;
    return m;
}

template <bool PCM16>
__global__ void __launch_bounds__(kThreadsA, 1) chroma_tc_spectrum_kernel(const SpectrumArgs args)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ FrameMeta s_meta[2][2][4];          // [pipeline][parity][frame of the group]

    const int tid = threadIdx.x;
    const int pipe = tid >> 8, t = tid & 255;
    const int wg = t >> 7, ln = t & 127;           // ln = TMEM lane = row of the 128-row tiles
    const int n2 = ln & 63, fp = ln >> 6;          // stage 1: row = (frame of the pair, n2)
    const uint32_t lane_base = (uint32_t)(ln & ~31) << 16;

    const uint32_t s_f = afs::smem_addr(smem);
    const uint32_t s_pipe = s_f + 2 * kFPlane + pipe * kPipeBytes;
    const uint32_t s_a1_hi = s_pipe, s_a1_lo = s_pipe + kA1Plane;
    const uint32_t s_y_hi = s_pipe + 2 * kA1Plane, s_y_lo = s_y_hi + kYPlane;

    // ---- one-time setup: F^T images, zeroed Y' tiles (pad rows stay zero), barriers, TMEM, G -> TMEM ----
    {
        uint4 *dst = reinterpret_cast<uint4 *>(smem);
        for (int i = tid; i < (int)(2 * kFPlane / 16); i += kThreadsA) dst[i] = args.f_img[i];
        uint4 *y0 = reinterpret_cast<uint4 *>(smem + 2 * kFPlane);
        for (int i = tid; i < (int)(2 * kPipeBytes / 16); i += kThreadsA) y0[i] = make_uint4(0, 0, 0, 0);
    }
    if (tid == 0) {
        afs::mbar_init(&s_bar[0], 1);
        afs::mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc::fence_async_smem();
    if (tid < 32) tc::tmem_alloc(&s_tmem, 512);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tm = s_tmem;
    if (tid < 128) {
        // thread = lane = row (2 k2 + part) of G; word c = {G[row][2c+1] : G[row][2c]}
#pragma unroll
        for (int plane = 0; plane < 2; plane++) {
            const uint32_t *src = args.g_img + ((size_t)plane * 128 + tid) * 64;
#pragma unroll
            for (int c0 = 0; c0 < 64; c0 += 16) {
                uint32_t v[16];
#pragma unroll
                for (int j = 0; j < 16; j++) v[j] = __ldg(src + c0 + j);
                tc::tmem_st16(tm + lane_base + (plane ? kColGlo : kColGhi) + c0, v);
            }
        }
        tc::tmem_wait_st();
    }
    // per-thread constants of the whole persistent loop
    float win[32];
#pragma unroll
    for (int j = 0; j < 32; j++) {
        win[j] = __ldg(args.hann + 64 * (32 * wg + j) + n2);
        if (PCM16) win[j] *= (1.0f / 32768.0f);      // librosa.load's scaling, a power of two
    }
    // twiddles W4096^(k1 n2): wg 0 handles k1 = 32 (slot 0) and 1..15, wg 1 handles 16..31
    float twc[16], tws[16];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const int k1 = wg ? 16 + i : (i == 0 ? 32 : i);
        const float2 w = __ldg(args.tw + k1 * 64 + n2);
        twc[i] = w.x;
        tws[i] = w.y;
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();

    const uint32_t d_tm = tm + kColD0 + pipe * kColDStride;
    const uint64_t a1_hi_desc = tc::smem_desc_k_sw128(s_a1_hi), a1_lo_desc = tc::smem_desc_k_sw128(s_a1_lo);
    const uint64_t f_hi_desc = tc::smem_desc_k_sw128(s_f), f_lo_desc = tc::smem_desc_k_sw128(s_f + kFPlane);
    const uint64_t y_hi_desc = tc::smem_desc_k_sw128(s_y_hi), y_lo_desc = tc::smem_desc_k_sw128(s_y_lo);
    constexpr uint32_t idesc1 = tc::idesc_bf16_f32(128, kN1), idesc2 = tc::idesc_bf16_f32(128, kN2);
    uint64_t *bar = &s_bar[pipe];
    uint32_t phase = 0;
    const int bar_id = 1 + pipe;

    const int64_t n_frames = args.frame_end - args.frame_begin;
    const int64_t n_groups = (n_frames + 3) >> 2;
    const int64_t g_stride = 2 * (int64_t)gridDim.x;
    int64_t g = 2 * (int64_t)blockIdx.x + pipe;
    int track = 0;                                      // cursor of the (at most four) threads that locate frames
    if (t < 4 && g < n_groups) s_meta[pipe][0][t] = locate_frame(args, args.frame_begin + 4 * g + t, track);
    int par = 0;

    for (; g < n_groups; g += g_stride, par ^= 1) {
        named_bar(bar_id, 256);                          // this group's frame table is visible
        if (t < 4 && g + g_stride < n_groups)
            s_meta[pipe][par ^ 1][t] = locate_frame(args, args.frame_begin + 4 * (g + g_stride) + t, track);

#pragma unroll 1
        for (int pr = 0; pr < 2; pr++) {
            const int fq = 2 * pr + fp;                  // frame of the group this row belongs to
            // ---- convert: x[64 n1 + n2] * window -> A1[row][n1], two bf16 planes ----
            {
                const FrameMeta m = s_meta[pipe][par][fq];
                const bool full = m.lo == 0 && m.hi == kNfft;
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    float x[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const int n = 64 * (32 * wg + 8 * c + j) + n2;
                        float v = 0.f;
                        if (full || (n >= m.lo && n < m.hi)) {
                            if (PCM16) v = (float)__ldg(static_cast<const short *>(args.audio) + m.base + n);
                            else v = __ldg(static_cast<const float *>(args.audio) + m.base + n);
                        }
                        x[j] = __fmul_rn(v, win[8 * c + j]);              // chroma.py:62 section * np.hanning
                    }
                    uint32_t h[4], l[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) tc::split_bf16x2(x[2 * j], x[2 * j + 1], h[j], l[j]);
                    const uint32_t off = (uint32_t)ln * 128 + ((uint32_t)((4 * wg + c) ^ (ln & 7)) << 4);
                    sts128(s_a1_hi + off, h[0], h[1], h[2], h[3]);
                    sts128(s_a1_lo + off, l[0], l[1], l[2], l[3]);
                }
            }
            tc::fence_async_smem();
            tc::fence_before_sync();
            named_bar(bar_id, 256);
            if (t == 0) {
                tc::fence_after_sync();
#pragma unroll
                for (int term = 0; term < 3; term++) {
                    const uint64_t ad = term == 1 ? a1_lo_desc : a1_hi_desc;     // hi*hi, lo*hi, hi*lo
                    const uint64_t bd = term == 2 ? f_lo_desc : f_hi_desc;
#pragma unroll
                    for (int ks = 0; ks < 4; ks++) tc::mma_ss(d_tm, ad + 2 * ks, bd + 2 * ks, idesc1, (term | ks) != 0);
                }
                tc::mma_commit(bar);
            }
            afs::mbar_wait(bar, phase);
            phase ^= 1u;
            tc::fence_after_sync();
            // ---- epilogue 1: twiddle, split, store as rows (frame, k1) of the stage-2 B operand ----
            {
                uint32_t v[32];
                {
                    uint32_t va[16], vb[16];
                    tc::tmem_ld16(d_tm + lane_base + 32 * wg, va);
                    tc::tmem_ld16(d_tm + lane_base + 32 * wg + 16, vb);
                    tc::tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 16; j++) { v[j] = va[j]; v[16 + j] = vb[j]; }
                }
                const uint32_t kb_off = (uint32_t)(n2 >> 5) * kYBlk + (uint32_t)(n2 & 3) * 4;
                const uint32_t cw = (uint32_t)(n2 & 31) >> 2;
                const int row0 = fq * kRowsPerFrame;
                auto emit = [&](int k1, float re, float im) {
                    uint32_t h, l;
                    tc::split_bf16x2(re, im, h, l);
                    const uint32_t row = (uint32_t)(row0 + k1);
                    const uint32_t off = kb_off + row * 128 + ((cw ^ (row & 7)) << 4);
                    sts32(s_y_hi + off, h);
                    sts32(s_y_lo + off, l);
                };
                if (wg == 0) {
                    emit(0, __uint_as_float(v[0]), 0.f);                                   // Y[0] is real and its twiddle is 1
                    const float y32 = __uint_as_float(v[1]);                               // Y[32] is real
                    emit(32, y32 * twc[0], y32 * tws[0]);
#pragma unroll
                    for (int i = 1; i < 16; i++) {
                        const float yr = __uint_as_float(v[2 * i]), yi = __uint_as_float(v[2 * i + 1]);
                        emit(i, yr * twc[i] - yi * tws[i], yr * tws[i] + yi * twc[i]);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        const float yr = __uint_as_float(v[2 * i]), yi = __uint_as_float(v[2 * i + 1]);
                        emit(16 + i, yr * twc[i] - yi * tws[i], yr * tws[i] + yi * twc[i]);
                    }
                }
            }
        }
        tc::fence_async_smem();
        tc::fence_before_sync();
        named_bar(bar_id, 256);
        if (t == 0) {
            tc::fence_after_sync();
#pragma unroll
            for (int term = 0; term < 3; term++) {
                const uint32_t a_tm = tm + (term == 1 ? kColGlo : kColGhi);           // hi*hi, lo*hi, hi*lo
                const uint64_t bd = term == 2 ? y_lo_desc : y_hi_desc;
#pragma unroll
                for (int ks = 0; ks < 8; ks++)
                    tc::mma_ts(d_tm, a_tm + 8 * ks, bd + (uint64_t)(((ks >> 2) * kYBlk + (ks & 3) * 32) >> 4), idesc2, (term | ks) != 0);
            }
            tc::mma_commit(bar);
        }
        afs::mbar_wait(bar, phase);
        phase ^= 1u;
        tc::fence_after_sync();
        // ---- epilogue 2: lane = (k2, re/im); |X|^2 pairs up through the neighbouring lane; wg takes two frames ----
#pragma unroll 1
        for (int q = 0; q < 2; q++) {
            const int fq = 2 * wg + q;
            uint32_t v[36];
            {
                uint32_t va[16], vb[16], vc[4];
                const uint32_t c0 = d_tm + lane_base + fq * kRowsPerFrame;
                tc::tmem_ld16(c0, va);
                tc::tmem_ld16(c0 + 16, vb);
                tmem_ld4(c0 + 32, vc);
                tc::tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 16; j++) { v[j] = va[j]; v[16 + j] = vb[j]; }
#pragma unroll
                for (int j = 0; j < 4; j++) v[32 + j] = vc[j];
            }
            const bool odd = ln & 1;
            float own[18];
#pragma unroll
            for (int j = 0; j < 17; j++) {
                // columns k1 = 2j (kept by even lanes) and 2j + 1 (kept by odd lanes); column 33 is padding
                const float a = __uint_as_float(v[2 * j]), b = __uint_as_float(v[2 * j + 1]);
                const float sa = a * a, sb = b * b;
                const float other = __shfl_xor_sync(0xffffffffu, odd ? sa : sb, 1);
                own[j] = (odd ? sb : sa) + other;
            }
            if (odd) own[16] = 0.f;
            own[17] = 0.f;
            const FrameMeta m = s_meta[pipe][par][fq];
            if (m.valid) {
                const int64_t fl = 4 * g + fq;                       // frame inside the chunk
                uint32_t *ph = args.p_hi + fl * kPlaneWords + ln, *pl = args.p_lo + fl * kPlaneWords + ln;
#pragma unroll
                for (int i = 0; i < 9; i++) {
                    uint32_t h, l;
                    tc::split_bf16x2(own[2 * i], own[2 * i + 1], h, l);
                    ph[i * 128] = h;
                    pl[i * 128] = l;
                }
            }
        }
        tc::fence_before_sync();      // the next group's MMAs overwrite the accumulator columns read above
    }
    tc::fence_before_sync();
    __syncthreads();
    if (tid < 32) tc::tmem_dealloc(tm, 512);
}

// ---------------- kernel B: filterbank + normalisation ----------------
constexpr int kFbThreads = 128;
constexpr int kFbTile = 64;                      // frames per tile: rows 0..63 = hi plane, 64..127 = lo plane
constexpr int kFbKBlocks = kPlaneWords * 2 / 64; // 36 blocks of 64 bf16
constexpr int kFbStages = 4;
constexpr uint32_t kFbWBlk = 32 * 128;           // weights: 32 rows (12 hi + 4 zero | 12 lo + 4 zero) x 64 elements
constexpr uint32_t kFbABlk = 128 * 128;
constexpr uint32_t kSmemFb = kFbKBlocks * kFbWBlk + kFbStages * kFbABlk;     // 147 456 + 65 536

struct FilterbankArgs {
    const uint32_t *p_hi, *p_lo;
    int64_t frame_begin, n_frames;               // chunk
    const uint4 *w_img;                          // [36][4 KB]
    const int64_t *frame_off, *out_off;
    int n_tracks, normalize, out_f64;
    void *out;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, int src_bytes)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

__global__ void __launch_bounds__(kFbThreads, 1) chroma_tc_filterbank_kernel(const FilterbankArgs args)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_free[kFbStages];
    __shared__ __align__(8) uint64_t s_done;
    __shared__ float s_part[kFbTile][kChroma];
    const int tid = threadIdx.x;
    const uint32_t lane_base = (uint32_t)(tid & ~31) << 16;
    const uint32_t s_w = afs::smem_addr(smem), s_a = s_w + kFbKBlocks * kFbWBlk;
    {
        uint4 *dst = reinterpret_cast<uint4 *>(smem);
        for (int i = tid; i < (int)(kFbKBlocks * kFbWBlk / 16); i += kFbThreads) dst[i] = args.w_img[i];
    }
    if (tid == 0) {
        for (int s = 0; s < kFbStages; s++) afs::mbar_init(&s_free[s], 1);
        afs::mbar_init(&s_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc::fence_async_smem();
    if (tid < 32) tc::tmem_alloc(&s_tmem, 32);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tm = s_tmem;
    constexpr uint32_t idesc = tc::idesc_bf16_f32(128, 32);
    const uint64_t w_desc = tc::smem_desc_k_sw128(s_w), a_desc = tc::smem_desc_k_sw128(s_a);

    const int64_t n_tiles = (args.n_frames + kFbTile - 1) / kFbTile;
    const int64_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t total_blocks = my_tiles * kFbKBlocks;

    // row `tid` of the A tile: frame (tid & 63) of the tile, plane hi (rows 0..63) or lo (rows 64..127)
    auto issue_load = [&](int64_t seq) {
        if (seq < total_blocks) {
            const int64_t tile = blockIdx.x + (seq / kFbKBlocks) * gridDim.x;
            const int kb = (int)(seq % kFbKBlocks);
            const int64_t fl = tile * kFbTile + (tid & 63);
            const bool ok = fl < args.n_frames;
            const uint32_t *src = (tid < 64 ? args.p_hi : args.p_lo) + (ok ? fl : 0) * kPlaneWords + kb * 32;
            const uint32_t dst = s_a + (uint32_t)(seq % kFbStages) * kFbABlk + (uint32_t)tid * 128;
#pragma unroll
            for (int c = 0; c < 8; c++) cp_async16(dst + ((uint32_t)(c ^ (tid & 7)) << 4), src + 4 * c, ok ? 16 : 0);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    for (int s = 0; s < kFbStages - 1; s++) issue_load(s);

    uint32_t done_phase = 0;
    for (int64_t seq = 0; seq < total_blocks; seq++) {
        const int kb = (int)(seq % kFbKBlocks);
        const int64_t tile = blockIdx.x + (seq / kFbKBlocks) * gridDim.x;
        // this thread's output frame (rows 0..63): located while the loads are in flight
        int track = 0;
        int64_t m_idx = 0, frames_k = 0, o_base = 0;
        bool ok_frame = false;
        if (kb == 0 && tid < kFbTile) {
            const int64_t fl = tile * kFbTile + tid;
            ok_frame = fl < args.n_frames;
            if (ok_frame) {
                const int64_t f = args.frame_begin + fl;
                int lo = 0, hi = args.n_tracks;
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (args.frame_off[mid] <= f) lo = mid; else hi = mid;
                }
                track = lo;
                m_idx = f - args.frame_off[track];
                frames_k = args.frame_off[track + 1] - args.frame_off[track];
                o_base = kChroma * args.out_off[track];
            }
            s_part[tid][0] = __int_as_float(track);            // parked in shared memory until the tile's epilogue
            s_part[tid][1] = __int_as_float(ok_frame ? 1 : 0);
        }
        asm volatile("cp.async.wait_group %0;" ::"n"(kFbStages - 2) : "memory");
        tc::fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            tc::fence_after_sync();
            const uint64_t ad = a_desc + (uint64_t)(((uint32_t)(seq % kFbStages) * kFbABlk) >> 4);
            const uint64_t wd = w_desc + (uint64_t)(((uint32_t)kb * kFbWBlk) >> 4);
#pragma unroll
            for (int ks = 0; ks < 4; ks++) tc::mma_ss(tm, ad + 2 * ks, wd + 2 * ks, idesc, (kb | ks) != 0);
            tc::mma_commit(&s_free[seq % kFbStages]);
            if (kb == kFbKBlocks - 1) tc::mma_commit(&s_done);
        }
        // the stage that block seq + S - 1 goes into was last read by the MMAs of block seq - 1
        if (seq >= 1) afs::mbar_wait(&s_free[(seq - 1) % kFbStages], (uint32_t)(((seq - 1) / kFbStages) & 1));
        issue_load(seq + kFbStages - 1);
        if (kb == kFbKBlocks - 1) {
            afs::mbar_wait(&s_done, done_phase);
            done_phase ^= 1u;
            tc::fence_after_sync();
            uint32_t va[16], vb[16];
            tc::tmem_ld16(tm + lane_base, va);
            tc::tmem_ld16(tm + lane_base + 16, vb);
            tc::tmem_wait_ld();
            int trk = 0, okf = 0;
            if (tid < kFbTile) { trk = __float_as_int(s_part[tid][0]); okf = __float_as_int(s_part[tid][1]); }
            __syncthreads();
            if (tid >= kFbTile) {
#pragma unroll
                for (int c = 0; c < kChroma; c++) s_part[tid - kFbTile][c] = __uint_as_float(va[c]);      // P_lo . w_hi
            }
            __syncthreads();
            if (tid < kFbTile && okf) {
                float raw[kChroma];
                float ss = 0.f;
#pragma unroll
                for (int c = 0; c < kChroma; c++) {
                    raw[c] = (__uint_as_float(va[c]) + __uint_as_float(vb[c])) + s_part[tid][c];      // chroma.py:70 np.dot(chromafb, spec)
                    ss = fmaf(raw[c], raw[c], ss);
                }
                float len = 1.f;
                if (args.normalize) {
                    // librosa.util.normalize(norm=2, axis=0), chroma.py:74: tiny lengths -> 1
                    len = sqrtf(ss);
                    if (len < FLT_MIN) len = 1.f;
                }
                const int64_t f = args.frame_begin + tile * kFbTile + tid;
                const int64_t m = f - args.frame_off[trk];
                const int64_t fk = args.frame_off[trk + 1] - args.frame_off[trk];
                const int64_t ob = kChroma * args.out_off[trk] + m;
#pragma unroll
                for (int c = 0; c < kChroma; c++) {
                    const float val = args.normalize ? raw[c] / len : raw[c];
                    if (args.out_f64) static_cast<double *>(args.out)[ob + (int64_t)c * fk] = (double)val;
                    else static_cast<float *>(args.out)[ob + (int64_t)c * fk] = val;
                }
            }
            tc::fence_before_sync();
            __syncthreads();          // accumulator and s_part are free for the next tile
        }
        (void)m_idx; (void)frames_k; (void)o_base;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    tc::fence_before_sync();
    __syncthreads();
    if (tid < 32) tc::tmem_dealloc(tm, 32);
}

uint16_t bf16_trunc(float x, float *back)
{
    uint32_t u;
    memcpy(&u, &x, 4);
    u &= 0xFFFF0000u;
    if (back) memcpy(back, &u, 4);
    return (uint16_t)(u >> 16);
}
void bf16_split(double v, uint16_t &hi, uint16_t &lo)
{
    const float x = (float)v;
    float h;
    hi = bf16_trunc(x, &h);
    lo = bf16_trunc(x - h, nullptr);
}

template <typename T>
int upload(T **dst, const std::vector<T> &src)
{
    AFS_CUDA(cudaMalloc(dst, sizeof(T) * src.size()));
    AFS_CUDA(cudaMemcpy(*dst, src.data(), sizeof(T) * src.size(), cudaMemcpyHostToDevice));
    return AFS_OK;
}

}  // namespace

struct afs_chroma_tc {
    uint4 *f_img = nullptr, *w_img = nullptr;
    uint32_t *g_img = nullptr;
    float2 *tw = nullptr;
    float *hann = nullptr;
    uint32_t *scratch = nullptr;      // [2 planes][chunk frames][kPlaneWords]
    int64_t scratch_frames = 0;
    int64_t chunk_frames = 16384;
};

void chroma_tc_destroy(afs_chroma_tc *tc)
{
    if (!tc) return;
    cudaFree(tc->f_img); cudaFree(tc->w_img); cudaFree(tc->g_img); cudaFree(tc->tw); cudaFree(tc->hann); cudaFree(tc->scratch);
    delete tc;
}

int chroma_tc_create(afs_chroma_tc **out, const std::vector<double> &fb, const std::vector<double> &hann)
{
    *out = nullptr;
    const double pi = 3.14159265358979323846;
    // ---- F^T: row o = output, K = n1.  o = 0: Re Y[0]; o = 1: Re Y[32]; o = 2 k1, 2 k1 + 1: Re, Im Y[k1] ----
    std::vector<uint16_t> f_img(2 * kFPlane / 2, 0);
    for (int o = 0; o < 64; o++)
        for (int n1 = 0; n1 < 64; n1++) {
            double v;
            if (o == 0) v = 1.0;
            else if (o == 1) v = (n1 & 1) ? -1.0 : 1.0;
            else {
                const int k1 = o >> 1;
                const double ang = 2.0 * pi * (double)((k1 * n1) & 63) / 64.0;
                v = (o & 1) ? -std::sin(ang) : std::cos(ang);
            }
            uint16_t hi, lo;
            bf16_split(v, hi, lo);
            const uint32_t off = tc::sw128_offset(o, n1) / 2;
            f_img[off] = hi;
            f_img[kFPlane / 2 + off] = lo;
        }
    // ---- G as the TMEM A operand: row 2 k2 + part, K = 2 n2 + (re | im of Y') ----
    std::vector<uint32_t> g_img((size_t)2 * 128 * 64);
    for (int k2 = 0; k2 < 64; k2++)
        for (int n2 = 0; n2 < 64; n2++) {
            const double ang = 2.0 * pi * (double)((k2 * n2) & 63) / 64.0;
            const double c = std::cos(ang), s = std::sin(ang);
            // X = sum (c - i s)(yr + i yi):  Re = c yr + s yi,  Im = -s yr + c yi
            const double rowv[2][2] = {{c, s}, {-s, c}};
            for (int part = 0; part < 2; part++) {
                uint16_t h0, l0, h1, l1;
                bf16_split(rowv[part][0], h0, l0);
                bf16_split(rowv[part][1], h1, l1);
                const size_t idx = (size_t)(2 * k2 + part) * 64 + n2;
                g_img[idx] = (uint32_t)h0 | ((uint32_t)h1 << 16);
                g_img[(size_t)128 * 64 + idx] = (uint32_t)l0 | ((uint32_t)l1 << 16);
            }
        }
    std::vector<float2> tw((size_t)33 * 64);
    for (int k1 = 0; k1 <= 32; k1++)
        for (int n2 = 0; n2 < 64; n2++) {
            const double ang = 2.0 * pi * (double)(k1 * n2) / 4096.0;
            tw[(size_t)k1 * 64 + n2] = make_float2((float)std::cos(ang), (float)-std::sin(ang));
        }
    // ---- filterbank as the B operand of kernel B.  K element e = 2 * (j * 128 + lane) + h is the power of
    //      k1 = 2 (2 j + h) + (lane & 1), k2 = lane >> 1, i.e. bin k1 + 64 k2 (mirrored above 2048); every bin of
    //      0..2048 must be hit exactly once, all other elements get weight zero ----
    std::vector<uint16_t> w_img((size_t)kFbKBlocks * kFbWBlk / 2, 0);
    std::vector<int> seen(kBins, 0);
    for (int e = 0; e < kPlaneWords * 2; e++) {
        const int word = e >> 1, h = e & 1, j = word >> 7, lane = word & 127;
        const int k1 = 2 * (2 * j + h) + (lane & 1), k2 = lane >> 1;
        if (k1 > 32) continue;
        int k = k1 + 64 * k2;
        if (k1 == 0 && k2 > 32) continue;              // X[64 k2] repeats conj X[4096 - 64 k2]
        if (k1 == 32 && k2 > 31) continue;
        if (k > 2048) k = kNfft - k;
        if (k < 0 || k >= kBins) return afs::fail(AFS_ERR_INVALID, "chroma_tc: bin map out of range");
        seen[k]++;
        const int kb = e >> 6, kk = e & 63;
        for (int c = 0; c < kChroma; c++) {
            uint16_t hi, lo;
            bf16_split(fb[(size_t)k * kChroma + c], hi, lo);
            w_img[(size_t)kb * kFbWBlk / 2 + tc::sw128_offset(c, kk) / 2] = hi;
            w_img[(size_t)kb * kFbWBlk / 2 + tc::sw128_offset(16 + c, kk) / 2] = lo;
        }
    }
    for (int k = 0; k < kBins; k++)
        if (seen[k] != 1) return afs::fail(AFS_ERR_INVALID, "chroma_tc: bin %d mapped %d times", k, seen[k]);
    std::vector<float> hannf(hann.begin(), hann.end());
    afs_chroma_tc *tcp = new afs_chroma_tc();
    int rc = AFS_OK;
    std::vector<uint4> f_u4(2 * kFPlane / 16), w_u4((size_t)kFbKBlocks * kFbWBlk / 16);
    memcpy(f_u4.data(), f_img.data(), f_img.size() * 2);
    memcpy(w_u4.data(), w_img.data(), w_img.size() * 2);
    if ((rc = upload(&tcp->f_img, f_u4)) || (rc = upload(&tcp->w_img, w_u4)) || (rc = upload(&tcp->g_img, g_img)) ||
        (rc = upload(&tcp->tw, tw)) || (rc = upload(&tcp->hann, hannf))) {
        chroma_tc_destroy(tcp);
        return rc;
    }
    cudaError_t e = cudaFuncSetAttribute(chroma_tc_spectrum_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemA);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(chroma_tc_spectrum_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemA);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(chroma_tc_filterbank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemFb);
    if (e != cudaSuccess) {
        chroma_tc_destroy(tcp);
        return afs::fail(AFS_ERR_CUDA, "chroma_tc_create: %s", cudaGetErrorString(e));
    }
    if (const char *c = getenv("AFS_CHROMA_TC_CHUNK")) {
        const long long v = atoll(c);
        if (v >= 64) tcp->chunk_frames = (v + 63) / 64 * 64;
    }
    *out = tcp;
    return AFS_OK;
}

int chroma_tc_run(afs_chroma_tc *tcp, const ChromaBatch &bt, cudaStream_t st)
{
    const int64_t chunk = tcp->chunk_frames < bt.total_frames ? tcp->chunk_frames : (bt.total_frames + 63) / 64 * 64;
    if (tcp->scratch_frames < chunk) {
        // grow-only scratch for the power spectrum of one chunk (stream-ordered use: one batch at a time per plan)
        AFS_CUDA(cudaStreamSynchronize(st));
        cudaFree(tcp->scratch);
        tcp->scratch = nullptr;
        tcp->scratch_frames = 0;
        AFS_CUDA(cudaMalloc(&tcp->scratch, sizeof(uint32_t) * 2 * (size_t)chunk * kPlaneWords));
        tcp->scratch_frames = chunk;
    }
    const int n_sm = afs::sm_count();
    for (int64_t f0 = 0; f0 < bt.total_frames; f0 += chunk) {
        const int64_t f1 = f0 + chunk < bt.total_frames ? f0 + chunk : bt.total_frames;
        SpectrumArgs sa;
        sa.audio = bt.audio;
        sa.sample_off = bt.sample_off;
        sa.frame_off = bt.frame_off;
        sa.n_tracks = bt.n_tracks;
        sa.frame_begin = f0;
        sa.frame_end = f1;
        sa.hop = bt.hop;
        sa.center_pad = bt.center_pad;
        sa.f_img = tcp->f_img;
        sa.g_img = tcp->g_img;
        sa.tw = tcp->tw;
        sa.hann = tcp->hann;
        sa.p_hi = tcp->scratch;
        sa.p_lo = tcp->scratch + (size_t)tcp->scratch_frames * kPlaneWords;
        const int64_t groups = (f1 - f0 + 3) / 4;
        int64_t blocks = (groups + 1) / 2;
        if (blocks > n_sm) blocks = n_sm;
        if (bt.pcm16) chroma_tc_spectrum_kernel<true><<<(unsigned)blocks, kThreadsA, kSmemA, st>>>(sa);
        else chroma_tc_spectrum_kernel<false><<<(unsigned)blocks, kThreadsA, kSmemA, st>>>(sa);
        afs::count_launch();
        AFS_CUDA(cudaGetLastError());
        FilterbankArgs fa;
        fa.p_hi = sa.p_hi;
        fa.p_lo = sa.p_lo;
        fa.frame_begin = f0;
        fa.n_frames = f1 - f0;
        fa.w_img = tcp->w_img;
        fa.frame_off = bt.frame_off;
        fa.out_off = bt.out_off;
        fa.n_tracks = bt.n_tracks;
        fa.normalize = bt.normalize;
        fa.out_f64 = bt.out_f64;
        fa.out = bt.out;
        int64_t tiles = (f1 - f0 + kFbTile - 1) / kFbTile;
        if (tiles > n_sm) tiles = n_sm;
        chroma_tc_filterbank_kernel<<<(unsigned)tiles, kFbThreads, kSmemFb, st>>>(fa);
        afs::count_launch();
        AFS_CUDA(cudaGetLastError());
    }
    return AFS_OK;
}

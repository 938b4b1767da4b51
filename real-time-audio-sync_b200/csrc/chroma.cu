// K1 — fused framing + Hann + real FFT-4096 + |X|^2 + 12 x 2049 chroma filterbank + L2
// normalisation, batched over tracks, on sm_100a.
//
// Replaces chroma.create_stft + chroma.create_chroma (reference chroma.py:44-75; the
// same code inlined in wtw.py:137-160 and :37-41) and the single-frame
// chroma.wav_to_chroma_col (chroma.py:35-42 == wtw.py:82-90).
//
// One CTA of 128 threads transforms one frame at a time (persistent grid-stride loop
// over all frames of all tracks; neighbouring CTAs work on neighbouring frames, so the
// 50 %-overlapping half of each frame is an L2 hit and every sample crosses HBM once).
// The real 4096-point FFT is a packed 2048-point complex FFT  z[n] = x[2n] + i x[2n+1]
// decomposed 2048 = 16 x 16 x 8 (three register-resident radix passes, two shared-memory
// exchanges), followed by the real-FFT untangle, |X|^2, the filterbank and the norm —
// nothing but the 12 chroma values per frame is written back.
#include <math_constants.h>

#include <cfloat>
#include <algorithm>
#include <cmath>
#include <vector>

#include <mutex>

#include "afs_common.cuh"
#include "chroma_tc.cuh"

namespace {

constexpr int kNfft = 4096;
constexpr int kNc = 2048;          // packed complex length
constexpr int kBins = 2049;
constexpr int kChroma = 12;
constexpr int kThreads = 128;
constexpr int kStrideA = 136;      // [k1][m] row stride (complex), padded: 2-wavefront 64-bit access
constexpr int kStrideB = 130;      // [k1][k2*8+m2] row stride (complex), padded for 128-bit reads

template <typename T> struct Cx { T x, y; };

template <typename T> __device__ __forceinline__ Cx<T> cadd(Cx<T> a, Cx<T> b) { return {a.x + b.x, a.y + b.y}; }
template <typename T> __device__ __forceinline__ Cx<T> csub(Cx<T> a, Cx<T> b) { return {a.x - b.x, a.y - b.y}; }
template <typename T> __device__ __forceinline__ Cx<T> cmul(Cx<T> a, Cx<T> b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
template <typename T> __device__ __forceinline__ Cx<T> mul_mi(Cx<T> a) { return {a.y, -a.x}; }   // a * (-i)

// natural-order 4-point DFT (forward, e^{-2 pi i nk/4})
template <typename T> __device__ __forceinline__ void fft4(Cx<T> &a, Cx<T> &b, Cx<T> &c, Cx<T> &d)
{
    const Cx<T> s0 = cadd(a, c), s1 = csub(a, c), s2 = cadd(b, d), s3 = mul_mi(csub(b, d));
    a = cadd(s0, s2);
    b = cadd(s1, s3);
    c = csub(s0, s2);
    d = csub(s1, s3);
}

// natural-order 16-point DFT in registers: 16 = 4 x 4
template <typename T> __device__ __forceinline__ void fft16(Cx<T> (&v)[16])
{
    constexpr T c1 = (T)0.92387953251128675613, s1 = (T)0.38268343236508977173;   // cos/sin(pi/8)
    constexpr T r2 = (T)0.70710678118654752440;
#pragma unroll
    for (int n2 = 0; n2 < 4; n2++) fft4(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);
    // v[4*k1 + n2] now holds t[n2][k1]; multiply by W16^(n2*k1)
    v[4 * 1 + 1] = cmul(v[4 * 1 + 1], Cx<T>{c1, -s1});     // W16^1
    v[4 * 1 + 2] = cmul(v[4 * 1 + 2], Cx<T>{r2, -r2});     // W16^2
    v[4 * 1 + 3] = cmul(v[4 * 1 + 3], Cx<T>{s1, -c1});     // W16^3
    v[4 * 2 + 1] = cmul(v[4 * 2 + 1], Cx<T>{r2, -r2});     // W16^2
    v[4 * 2 + 2] = mul_mi(v[4 * 2 + 2]);                   // W16^4
    v[4 * 2 + 3] = cmul(v[4 * 2 + 3], Cx<T>{-r2, -r2});    // W16^6
    v[4 * 3 + 1] = cmul(v[4 * 3 + 1], Cx<T>{s1, -c1});     // W16^3
    v[4 * 3 + 2] = cmul(v[4 * 3 + 2], Cx<T>{-r2, -r2});    // W16^6
    v[4 * 3 + 3] = cmul(v[4 * 3 + 3], Cx<T>{-c1, s1});     // W16^9
#pragma unroll
    for (int k1 = 0; k1 < 4; k1++) fft4(v[4 * k1 + 0], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
    // v[4*k1 + k2] = X[k1 + 4*k2]  -> reorder to natural
    Cx<T> o[16];
#pragma unroll
    for (int k1 = 0; k1 < 4; k1++)
#pragma unroll
        for (int k2 = 0; k2 < 4; k2++) o[k1 + 4 * k2] = v[4 * k1 + k2];
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = o[i];
}

// natural-order 8-point DFT: 8 = 4 x 2
template <typename T> __device__ __forceinline__ void fft8(Cx<T> (&v)[8])
{
    constexpr T r2 = (T)0.70710678118654752440;
    fft4(v[0], v[2], v[4], v[6]);      // t[0][k1] in v[0],v[2],v[4],v[6]
    fft4(v[1], v[3], v[5], v[7]);      // t[1][k1]
    v[3] = cmul(v[3], Cx<T>{r2, -r2}); // W8^1
    v[5] = mul_mi(v[5]);               // W8^2
    v[7] = cmul(v[7], Cx<T>{-r2, -r2});// W8^3
    Cx<T> o[8];
#pragma unroll
    for (int k1 = 0; k1 < 4; k1++) {
        o[k1] = cadd(v[2 * k1], v[2 * k1 + 1]);
        o[k1 + 4] = csub(v[2 * k1], v[2 * k1 + 1]);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = o[i];
}

template <typename T>
struct ChromaTables {
    const T *hann;          // 4096
    const Cx<T> *tw2048;    // exp(-2 pi i j / 2048), j < 2048
    const Cx<T> *tw4096;    // exp(-2 pi i k / 4096), k <= 1024
    const T *fb;            // [2049][12]
};


// sample s of a track that starts at element s_begin, as librosa.load would deliver it (float32)
__device__ __forceinline__ float load_sample(const ChromaBatch &bt, int64_t idx)
{
    if (bt.pcm16) return (float)__ldg(static_cast<const short *>(bt.audio) + idx) * (1.0f / 32768.0f);
    return __ldg(static_cast<const float *>(bt.audio) + idx);
}

template <typename T>
__global__ void __launch_bounds__(kThreads) chroma_kernel(const ChromaTables<T> tb, const ChromaBatch bt)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    Cx<T> *sA = reinterpret_cast<Cx<T> *>(s_raw);          // 16 x kStrideA   (later: Z natural, 2048)
    Cx<T> *sB = sA + 16 * kStrideA;                        // 16 x kStrideB
    T *sRed = reinterpret_cast<T *>(sB + 16 * kStrideB);   // 4 warps x 12
    const int t = threadIdx.x;
    const int lane = t & 31, warp = t >> 5;

    for (int64_t f = blockIdx.x; f < bt.total_frames; f += gridDim.x) {
        // ---- locate (track, frame) ----
        int lo = 0, hi = bt.n_tracks;            // frame_off[lo] <= f < frame_off[hi]
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (bt.frame_off[mid] <= f) lo = mid; else hi = mid;
        }
        const int track = lo;
        const int64_t m_idx = f - bt.frame_off[track];
        const int64_t s_begin = bt.sample_off[track];
        const int64_t n_samp = bt.sample_off[track + 1] - s_begin;
        const int64_t frames_k = bt.frame_off[track + 1] - bt.frame_off[track];
        const int64_t start = m_idx * bt.hop - (bt.center_pad ? kNfft / 2 : 0);     // chroma.py:49 left zero pad

        // ---- pass 1: thread m = t, n1 = 0..15: z[128 n1 + m] = x[256 n1 + 2m] + i x[256 n1 + 2m + 1], windowed ----
        Cx<T> v[16];
#pragma unroll
        for (int n1 = 0; n1 < 16; n1++) {
            const int n = 256 * n1 + 2 * t;
            const int64_t s = start + n;
            float2 xv = make_float2(0.f, 0.f);
            if (s >= 0 && s < n_samp) xv.x = load_sample(bt, s_begin + s);
            if (s + 1 >= 0 && s + 1 < n_samp) xv.y = load_sample(bt, s_begin + s + 1);
            v[n1].x = (T)xv.x * __ldg(tb.hann + n);          // chroma.py:62 section * np.hanning
            v[n1].y = (T)xv.y * __ldg(tb.hann + n + 1);
        }
        fft16(v);
#pragma unroll
        for (int k1 = 0; k1 < 16; k1++) {
            Cx<T> y = v[k1];
            if (k1 > 0) y = cmul(y, tb.tw2048[t * k1]);                // W_2048^(m k1)
            sA[k1 * kStrideA + t] = y;
        }
        __syncthreads();
        // ---- pass 2: thread (k1 = t / 8, m2 = t % 8): u[m1] = A[k1][8 m1 + m2] ----
        {
            const int k1 = t >> 3, m2 = t & 7;
#pragma unroll
            for (int m1 = 0; m1 < 16; m1++) v[m1] = sA[k1 * kStrideA + 8 * m1 + m2];
            fft16(v);
#pragma unroll
            for (int k2 = 0; k2 < 16; k2++) {
                Cx<T> y = v[k2];
                if (k2 > 0) y = cmul(y, tb.tw2048[16 * m2 * k2]);      // W_128^(m2 k2)
                sB[k1 * kStrideB + k2 * 8 + m2] = y;
            }
        }
        __syncthreads();
        // ---- pass 3: two (k1, k2) groups per thread; 8-point DFT over m2 -> Z[k1 + 16 k2 + 256 k3] ----
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int p = t + h * kThreads;        // p = k2 * 16 + k1  (lanes run along k1)
            const int k1 = p & 15, k2 = p >> 4;
            Cx<T> u[8];
#pragma unroll
            for (int m2 = 0; m2 < 8; m2++) u[m2] = sB[k1 * kStrideB + k2 * 8 + m2];
            fft8(u);
#pragma unroll
            for (int k3 = 0; k3 < 8; k3++) sA[k1 + 16 * k2 + 256 * k3] = u[k3];
        }
        __syncthreads();
        // ---- untangle + power + filterbank ----
        Cx<T> *spec = bt.stft_out ? static_cast<Cx<T> *>(bt.stft_out) + (bt.out_off[track] + m_idx) * kBins : nullptr;
        T acc[kChroma];
#pragma unroll
        for (int c = 0; c < kChroma; c++) acc[c] = (T)0;
        auto add_bin = [&](int k, T p) {
            const T *w = tb.fb + (size_t)k * kChroma;
#pragma unroll
            for (int c = 0; c < kChroma; c++) acc[c] = fma(__ldg(w + c), p, acc[c]);
        };
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int k = t + kThreads * i;            // 0 .. 1023
            if (k == 0) {
                const Cx<T> z0 = sA[0];
                const T x0 = z0.x + z0.y, xn = z0.x - z0.y;       // X[0], X[2048] (real)
                add_bin(0, x0 * x0);
                add_bin(2048, xn * xn);
                const Cx<T> zq = sA[1024];                        // X[1024] = conj(Z[1024])
                add_bin(1024, zq.x * zq.x + zq.y * zq.y);
                if (spec) { spec[0] = Cx<T>{x0, (T)0}; spec[2048] = Cx<T>{xn, (T)0}; spec[1024] = Cx<T>{zq.x, -zq.y}; }
            } else {
                const Cx<T> zk = sA[k], zn = sA[kNc - k];
                const Cx<T> e = {(T)0.5 * (zk.x + zn.x), (T)0.5 * (zk.y - zn.y)};     // (Zk + conj Zn)/2
                const Cx<T> d = {(T)0.5 * (zk.x - zn.x), (T)0.5 * (zk.y + zn.y)};     // (Zk - conj Zn)/2
                const Cx<T> wk = tb.tw4096[k];
                const Cx<T> o = mul_mi(cmul(wk, d));                                  // -i w_k d
                const Cx<T> xa = cadd(e, o), xb = csub(e, o);                         // X[k], conj X[2048-k]
                add_bin(k, xa.x * xa.x + xa.y * xa.y);                                // chroma.py:68 abs(ft)**2
                add_bin(kNc - k, xb.x * xb.x + xb.y * xb.y);
                if (spec) { spec[k] = xa; spec[kNc - k] = Cx<T>{xb.x, -xb.y}; }       // chroma.py:63 np.fft.rfft
            }
        }
        // ---- block reduction of the 12 partial sums (chroma.py:70 np.dot(chromafb, spec)) ----
#pragma unroll
        for (int c = 0; c < kChroma; c++) {
            T a = acc[c];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
            acc[c] = a;
        }
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < kChroma; c++) sRed[warp * kChroma + c] = acc[c];
        }
        __syncthreads();
        if (warp == 0) {
            const int c = lane < kChroma ? lane : 0;
            T raw = sRed[c] + sRed[kChroma + c] + sRed[2 * kChroma + c] + sRed[3 * kChroma + c];
            if (lane >= kChroma) raw = (T)0;
            T val = raw;
            if (bt.normalize) {
                // librosa.util.normalize(norm=2, axis=0): chroma.py:74; tiny lengths -> 1
                T ss = raw * raw;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
                T len = sqrt(ss);
                const T tiny = sizeof(T) == 4 ? (T)FLT_MIN : (T)DBL_MIN;
                if (len < tiny) len = (T)1;
                val = raw / len;
            }
            if (lane < kChroma && bt.out) {
                const int64_t o = kChroma * bt.out_off[track] + (int64_t)lane * frames_k + m_idx;
                if (bt.out_f64) static_cast<double *>(bt.out)[o] = (double)val;
                else static_cast<float *>(bt.out)[o] = (float)val;
            }
        }
        __syncthreads();      // sRed / sA reused by the next frame
    }
}

// ------------------------------------------------------------------------------------------
// Fast float32 variant (the throughput path).  Same transform as chroma_kernel<float>, but
//   * the thread's 32 window values, its 15 pass-1 twiddles and its untangle twiddle base live
//     in registers for the whole persistent loop; the 128 pass-2 twiddles sit in shared memory;
//   * the 12 x 2049 filterbank is applied as class-sorted sparse windows: for every bin above the
//     lowest few, only 6 cyclically adjacent chroma classes carry weight (the rest are < 1e-7 of
//     the column maximum: Gaussian of 1 semitone width, librosa.filters.chroma).  Bins are sorted
//     by the first class c0 of their window and dealt to threads so that all bins of a thread share
//     c0: the thread accumulates 6 statically indexed sums from a [slot][6][thread] weight table
//     (coalesced, L1 resident) and power values gathered from shared memory.  The handful of
//     wide low-frequency bins keep all 12 weights ("dense list").
// The plan builder verifies the structure on the actual matrix and falls back to the generic
// kernel if it does not hold.
constexpr int kWin = 6;
constexpr int kMaxBpt = 24;
constexpr int kMaxDense = 32;
constexpr int kZeroSlot = kBins;      // sP[2049] == 0 for unused (thread, slot) entries

struct ChromaFastTables {
    const float *hann;          // 4096
    const float2 *tw2048;       // exp(-2 pi i j / 2048)
    const float2 *tw4096;       // exp(-2 pi i k / 4096), k <= 1024
    const float2 *tw1;          // [15][128]: W_2048^(t k1) at [(k1-1)][t]
    const float *wsp;           // [bpt][6][128]
    const uint16_t *paddr;      // [bpt][128]
    const int *cls_start;       // [13] threads [cls_start[c], cls_start[c+1]) have window start c
    const float *wdense;        // [nd][12]
    const uint16_t *kdense;     // [nd]
    int bpt, nd;
};

constexpr int kFastStrideA = 129;   // odd row strides: 64-bit accesses whose lanes run along k1 are conflict free
constexpr int kFastStrideB = 129;

// BPT = bins per thread of the sparse filterbank when known at compile time (17 for the standard
// librosa filterbank at sr 22050 / n_fft 4096), 0 = read it from the tables
#ifndef AFS_CHROMA_MINB
#define AFS_CHROMA_MINB 4      // resident CTAs per SM the register budget is tuned for
#endif
template <int BPT, bool PCM16>
__global__ void __launch_bounds__(kThreads, AFS_CHROMA_MINB) chroma_fast_kernel(const ChromaFastTables tb, const ChromaBatch bt)
{
    constexpr int kSampleBytes = PCM16 ? 2 : 4;
    using C = Cx<float>;
    __shared__ __align__(128) C sA[16 * kFastStrideA];     // staged audio of the NEXT frame (16 KB, TMA) | pass-1 output | Z natural
    __shared__ __align__(16) C sB[16 * kFastStrideB];      // pass-2 output; later power spectrum float[2052] + reduction scratch
    __shared__ C sTw2[16 * 8];                              // W_128^(m2 k2) at [k2][m2]
    __shared__ int sCls[13];
    __shared__ __align__(8) uint64_t sBar;                  // completion of the staged-audio bulk copy
    const int t = threadIdx.x;
    const int lane = t & 31;

    // per-thread constants for the whole persistent loop: 32 window values; pass-1 twiddles W^(t k1) are
    // kept for k1 = 1, 2, 4, 8 only (the other eleven are products of those: 44 extra flops, 22 fewer registers)
    float2 win[16];
#pragma unroll
    for (int n1 = 0; n1 < 16; n1++) {
        win[n1] = reinterpret_cast<const float2 *>(tb.hann)[128 * n1 + t];
        // PCM16: fold librosa.load's 1/32768 into the window (a power of two: the products round identically)
        if (PCM16) { win[n1].x *= (1.0f / 32768.0f); win[n1].y *= (1.0f / 32768.0f); }
    }
    C w1, w2, w4, w8;
    { const float2 a = tb.tw1[0 * kThreads + t], b = tb.tw1[1 * kThreads + t], c = tb.tw1[3 * kThreads + t], d = tb.tw1[7 * kThreads + t];
      w1 = C{a.x, a.y}; w2 = C{b.x, b.y}; w4 = C{c.x, c.y}; w8 = C{d.x, d.y}; }
    const float2 twu0 = tb.tw4096[t];                       // W_4096^t ; W_4096^(t + 128 i) = twu0 * W_32^i
    { const float2 w = tb.tw2048[16 * (t & 7) * (t >> 3)]; sTw2[t] = C{w.x, w.y}; }
    if (t < 13) sCls[t] = tb.cls_start[t];
    if (t == 0) {
        afs::mbar_init(&sBar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    float *sP = reinterpret_cast<float *>(sB);
    float *sRed = sP + 2112;                                // [6][128] window sums + dense[12] + (class, w)[72]; sP ends at 2050
    __syncthreads();

    // Stage the 4096 samples of frame f (16 KB, contiguous) into sA with ONE TMA bulk copy when the frame lies
    // inside the track and is 16-byte aligned; otherwise pass 1 reads it with guarded global loads.
    int ld_track = 0;
    auto stage_frame = [&](int64_t f) -> bool {
        while (bt.frame_off[ld_track + 1] <= f) ld_track++;          // frames are visited in increasing order
        const int64_t s_begin = bt.sample_off[ld_track];
        const int64_t n_samp = bt.sample_off[ld_track + 1] - s_begin;
        const int64_t start = (f - bt.frame_off[ld_track]) * bt.hop - (bt.center_pad ? kNfft / 2 : 0);
        const char *src = static_cast<const char *>(bt.audio) + (s_begin + start) * kSampleBytes;
        const bool ok = (start >= 0) && (start + kNfft <= n_samp) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
        if (ok && t == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            afs::mbar_expect_tx(&sBar, kNfft * kSampleBytes);
            afs::bulk_g2s(sA, src, kNfft * kSampleBytes, &sBar);
        }
        return ok;
    };
    bool staged = ((int64_t)blockIdx.x < bt.total_frames) ? stage_frame(blockIdx.x) : false;
    uint32_t bar_phase = 0;

    int track = 0;
    for (int64_t f = blockIdx.x; f < bt.total_frames; f += gridDim.x) {
        while (bt.frame_off[track + 1] <= f) track++;
        const int64_t m_idx = f - bt.frame_off[track];
        const int64_t frames_k = bt.frame_off[track + 1] - bt.frame_off[track];

        // ---- pass 1: thread m = t ----
        C v[16];
        if (staged) {
            afs::mbar_wait(&sBar, bar_phase);
            bar_phase ^= 1u;
            if (PCM16) {
                const short2 *xs = reinterpret_cast<const short2 *>(sA);
#pragma unroll
                for (int n1 = 0; n1 < 16; n1++) {
                    const short2 xv = xs[128 * n1 + t];
                    v[n1] = C{__fmul_rn((float)xv.x, win[n1].x), __fmul_rn((float)xv.y, win[n1].y)};
                }
            } else {
                const float2 *xs = reinterpret_cast<const float2 *>(sA);
#pragma unroll
                for (int n1 = 0; n1 < 16; n1++) {
                    const float2 xv = xs[128 * n1 + t];
                    v[n1] = C{__fmul_rn(xv.x, win[n1].x), __fmul_rn(xv.y, win[n1].y)};      // no fma contraction: float and PCM16 inputs must round alike
                }
            }
        } else {
            const int64_t s_begin = bt.sample_off[track];
            const int64_t n_samp = bt.sample_off[track + 1] - s_begin;
            const int64_t start = m_idx * bt.hop - (bt.center_pad ? kNfft / 2 : 0);
#pragma unroll
            for (int n1 = 0; n1 < 16; n1++) {
                const int64_t s = start + 256 * n1 + 2 * t;
                float2 xv = make_float2(0.f, 0.f);
                if (PCM16) {
                    const short *x = static_cast<const short *>(bt.audio) + s_begin;
                    if (s >= 0 && s < n_samp) xv.x = (float)__ldg(x + s);
                    if (s + 1 >= 0 && s + 1 < n_samp) xv.y = (float)__ldg(x + s + 1);
                } else {
                    const float *x = static_cast<const float *>(bt.audio) + s_begin;
                    if (s >= 0 && s < n_samp) xv.x = __ldg(x + s);
                    if (s + 1 >= 0 && s + 1 < n_samp) xv.y = __ldg(x + s + 1);
                }
                v[n1] = C{__fmul_rn(xv.x, win[n1].x), __fmul_rn(xv.y, win[n1].y)};      // no fma contraction: float and PCM16 inputs must round alike
            }
        }
        fft16(v);
        __syncthreads();                 // every thread has taken its samples out of sA
        {
            const C w3 = cmul(w2, w1), w5 = cmul(w4, w1), w6 = cmul(w4, w2), w7 = cmul(w4, w3);
            sA[t] = v[0];
            sA[1 * kFastStrideA + t] = cmul(v[1], w1);
            sA[2 * kFastStrideA + t] = cmul(v[2], w2);
            sA[3 * kFastStrideA + t] = cmul(v[3], w3);
            sA[4 * kFastStrideA + t] = cmul(v[4], w4);
            sA[5 * kFastStrideA + t] = cmul(v[5], w5);
            sA[6 * kFastStrideA + t] = cmul(v[6], w6);
            sA[7 * kFastStrideA + t] = cmul(v[7], w7);
            sA[8 * kFastStrideA + t] = cmul(v[8], w8);
            sA[9 * kFastStrideA + t] = cmul(v[9], cmul(w8, w1));
            sA[10 * kFastStrideA + t] = cmul(v[10], cmul(w8, w2));
            sA[11 * kFastStrideA + t] = cmul(v[11], cmul(w8, w3));
            sA[12 * kFastStrideA + t] = cmul(v[12], cmul(w8, w4));
            sA[13 * kFastStrideA + t] = cmul(v[13], cmul(w8, w5));
            sA[14 * kFastStrideA + t] = cmul(v[14], cmul(w8, w6));
            sA[15 * kFastStrideA + t] = cmul(v[15], cmul(w8, w7));
        }
        __syncthreads();
        // ---- pass 2: thread (k1 = t % 16, m2 = t / 16) ----
        {
            const int k1 = t & 15, m2 = t >> 4;
#pragma unroll
            for (int m1 = 0; m1 < 16; m1++) v[m1] = sA[k1 * kFastStrideA + 8 * m1 + m2];
            fft16(v);
            sB[k1 * kFastStrideB + m2] = v[0];
#pragma unroll
            for (int k2 = 1; k2 < 16; k2++) sB[k1 * kFastStrideB + k2 * 8 + m2] = cmul(v[k2], sTw2[k2 * 8 + m2]);
        }
        __syncthreads();
        // ---- pass 3: (k1 = p % 16, k2 = p / 16), 8-point DFT over m2 ----
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int p = t + h * kThreads;
            const int k1 = p & 15, k2 = p >> 4;
            C u[8];
#pragma unroll
            for (int m2 = 0; m2 < 8; m2++) u[m2] = sB[k1 * kFastStrideB + k2 * 8 + m2];
            fft8(u);
#pragma unroll
            for (int k3 = 0; k3 < 8; k3++) sA[k1 + 16 * k2 + 256 * k3] = u[k3];
            if (p == 0) sA[kNc] = u[0];        // Z[2048] == Z[0]: makes k = 0 an ordinary pair in the untangle
        }
        __syncthreads();
        // ---- untangle + power -> sP ----
        {
            constexpr float kc[8] = {1.f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
                                     0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f, 0.19509032201612826785f};
            constexpr float ks[8] = {0.f, 0.19509032201612826785f, 0.38268343236508977173f, 0.55557023301960222474f,
                                     0.70710678118654752440f, 0.83146961230254523708f, 0.92387953251128675613f, 0.98078528040323044913f};
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int k = t + kThreads * i;            // 0 .. 1023, partner 2048 - k
                const C zk = sA[k], zn = sA[kNc - k];
                const C e = {0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y)};
                const C d = {0.5f * (zk.x - zn.x), 0.5f * (zk.y + zn.y)};
                const C wk = cmul(C{twu0.x, twu0.y}, C{kc[i], -ks[i]});       // W_4096^(t + 128 i)
                const C o = mul_mi(cmul(wk, d));
                const C xa = cadd(e, o), xb = csub(e, o);
                sP[k] = xa.x * xa.x + xa.y * xa.y;
                sP[kNc - k] = xb.x * xb.x + xb.y * xb.y;
            }
            if (t == 0) {
                const C zq = sA[1024];                     // X[1024] = conj(Z[1024])
                sP[1024] = zq.x * zq.x + zq.y * zq.y;
                sP[kZeroSlot] = 0.f;
            }
        }
        __syncthreads();
        // sA is dead now: the next frame's samples are bulk-copied into it while the filterbank runs
        staged = (f + gridDim.x < bt.total_frames) ? stage_frame(f + gridDim.x) : false;
        // ---- sparse filterbank: 6 statically indexed partial sums per thread ----
        float acc[kWin];
#pragma unroll
        for (int w = 0; w < kWin; w++) acc[w] = 0.f;
        if (BPT > 0) {
            // slot count known at compile time: every table access is base + immediate offset
            const uint16_t *pa = tb.paddr + t;
            const float *wp = tb.wsp + t;
#pragma unroll
            for (int i = 0; i < BPT; i++) {
                const float p = sP[__ldg(pa + i * kThreads)];
#pragma unroll
                for (int w = 0; w < kWin; w++) acc[w] = fmaf(__ldg(wp + (i * kWin + w) * kThreads), p, acc[w]);
            }
        } else {
#pragma unroll 4
            for (int i = 0; i < tb.bpt; i++) {
                const float p = sP[__ldg(tb.paddr + i * kThreads + t)];
                const float *wp = tb.wsp + (size_t)i * kWin * kThreads + t;
#pragma unroll
                for (int w = 0; w < kWin; w++) acc[w] = fmaf(__ldg(wp + w * kThreads), p, acc[w]);
            }
        }
        float dacc = 0.f;
        if (t >= kThreads - kChroma) {          // 12 threads: the wide low bins with all 12 weights
            const int c = t - (kThreads - kChroma);
            for (int j = 0; j < tb.nd; j++) dacc = fmaf(__ldg(tb.wdense + j * kChroma + c), sP[__ldg(tb.kdense + j)], dacc);
        }
#pragma unroll
        for (int w = 0; w < kWin; w++) sRed[w * kThreads + t] = acc[w];
        if (t >= kThreads - kChroma) sRed[kWin * kThreads + t - (kThreads - kChroma)] = dacc;
        __syncthreads();
        // ---- combine: (class c, window position w) sums over the threads whose window starts at c - w ----
        if (t < kChroma * kWin) {
            const int c = t / kWin, w = t - c * kWin;
            int c0 = c - w;
            if (c0 < 0) c0 += kChroma;
            float sacc = 0.f;
            for (int q = sCls[c0]; q < sCls[c0 + 1]; q++) sacc += sRed[w * kThreads + q];
            sRed[kWin * kThreads + kChroma + t] = sacc;
        }
        __syncthreads();
        if (t < 32) {
            float raw = 0.f;
            if (lane < kChroma) {
                raw = sRed[kWin * kThreads + lane];
#pragma unroll
                for (int w = 0; w < kWin; w++) raw += sRed[kWin * kThreads + kChroma + lane * kWin + w];
            }
            float val = raw;
            if (bt.normalize) {
                float ss = raw * raw;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
                float len = sqrtf(ss);
                if (len < FLT_MIN) len = 1.f;
                val = raw / len;
            }
            if (lane < kChroma) {
                const int64_t o = kChroma * bt.out_off[track] + (int64_t)lane * frames_k + m_idx;
                if (bt.out_f64) static_cast<double *>(bt.out)[o] = (double)val;
                else static_cast<float *>(bt.out)[o] = val;
            }
        }
        // no barrier here: the reduction scratch lives in sB, which is next written in pass 2 — two
        // barriers away; pass 1 of the next frame only touches sA
    }
}

}  // namespace

struct afs_chroma_plan {
    int n_fft = kNfft, hop = 2048;
    // device tables (float and double variants)
    float *f_hann = nullptr, *f_fb = nullptr;
    Cx<float> *f_tw2048 = nullptr, *f_tw4096 = nullptr;
    double *d_hann = nullptr, *d_fb = nullptr;
    Cx<double> *d_tw2048 = nullptr, *d_tw4096 = nullptr;
    // tensor-core pipeline (csrc/chroma_tc.cu): tables + the chunk scratch for the power spectrum
    afs_chroma_tc *tc = nullptr;
    // launches of one plan are serialised across streams: the tensor path owns one scratch buffer per plan
    std::mutex mu;
    cudaEvent_t last_done = nullptr;
    cudaStream_t last_stream = nullptr;
    bool has_last = false;
    // fast-path tables (class-sorted sparse filterbank); fast_ok == false -> generic kernel only
    bool fast_ok = false;
    int bpt = 0, nd = 0;
    float *f_wsp = nullptr, *f_wdense = nullptr;
    Cx<float> *f_tw1 = nullptr;
    uint16_t *u_paddr = nullptr, *u_kdense = nullptr;
    int *i_cls = nullptr;
};

template <typename T> static size_t chroma_smem_bytes()
{
    return sizeof(Cx<T>) * (16 * kStrideA + 16 * kStrideB) + sizeof(T) * 4 * kChroma;
}

template <typename T>
static int upload(T **dst, const std::vector<T> &src)
{
    AFS_CUDA(cudaMalloc(dst, sizeof(T) * src.size()));
    AFS_CUDA(cudaMemcpy(*dst, src.data(), sizeof(T) * src.size(), cudaMemcpyHostToDevice));
    return AFS_OK;
}


// Derive the class-sorted sparse form of the filterbank (see chroma_fast_kernel).  fb is [k][12].
static int build_fast_tables(afs_chroma_plan *pl, const std::vector<double> &fb)
{
    const double tol = 1e-6;
    std::vector<int> c0(kBins, 0);
    std::vector<int> dense;
    std::vector<std::vector<int>> by_class(kChroma);
    for (int k = 0; k < kBins; k++) {
        const double *col = &fb[(size_t)k * kChroma];
        double cmax = 0.0;
        for (int c = 0; c < kChroma; c++) cmax = std::max(cmax, std::fabs(col[c]));
        int best = 0;
        double best_kept = -1.0;
        for (int s = 0; s < kChroma; s++) {
            double kept = 0.0;
            for (int w = 0; w < kWin; w++) kept += col[(s + w) % kChroma] * col[(s + w) % kChroma];
            if (kept > best_kept) { best_kept = kept; best = s; }
        }
        double dropped = 0.0;
        for (int c = 0; c < kChroma; c++) {
            const int rel = (c - best + kChroma) % kChroma;
            if (rel >= kWin) dropped = std::max(dropped, std::fabs(col[c]));
        }
        c0[k] = best;
        if (cmax > 0.0 && dropped > tol * cmax) dense.push_back(k);
        else by_class[best].push_back(k);
    }
    if ((int)dense.size() > kMaxDense) return AFS_OK;        // structure absent: generic kernel only
    int bpt = 0;
    for (int cand = 8; cand <= kMaxBpt; cand++) {
        int thr = 0;
        for (int c = 0; c < kChroma; c++) thr += ((int)by_class[c].size() + cand - 1) / cand;
        if (thr <= kThreads) { bpt = cand; break; }
    }
    if (!bpt) return AFS_OK;
    std::vector<float> wsp((size_t)bpt * kWin * kThreads, 0.f);
    std::vector<uint16_t> paddr((size_t)bpt * kThreads, (uint16_t)kZeroSlot);
    std::vector<int> cls(kChroma + 1, 0);
    int thr = 0;
    for (int c = 0; c < kChroma; c++) {
        cls[c] = thr;
        const std::vector<int> &bins = by_class[c];
        const int nthr = ((int)bins.size() + bpt - 1) / bpt;
        for (size_t q = 0; q < bins.size(); q++) {
            // deal bins round-robin over this class's threads: neighbouring bins -> neighbouring lanes
            const int tt = thr + (int)(q % nthr), slot = (int)(q / nthr);
            paddr[(size_t)slot * kThreads + tt] = (uint16_t)bins[q];
            for (int w = 0; w < kWin; w++)
                wsp[((size_t)slot * kWin + w) * kThreads + tt] = (float)fb[(size_t)bins[q] * kChroma + (c + w) % kChroma];
        }
        thr += nthr;
    }
    cls[kChroma] = thr;
    std::vector<float> wdense(std::max<size_t>(1, dense.size() * kChroma), 0.f);
    std::vector<uint16_t> kdense(std::max<size_t>(1, dense.size()), 0);
    for (size_t j = 0; j < dense.size(); j++) {
        kdense[j] = (uint16_t)dense[j];
        for (int c = 0; c < kChroma; c++) wdense[j * kChroma + c] = (float)fb[(size_t)dense[j] * kChroma + c];
    }
    std::vector<Cx<float>> tw1((size_t)15 * kThreads);
    const double pi = 3.14159265358979323846;
    for (int k1 = 1; k1 < 16; k1++)
        for (int t = 0; t < kThreads; t++) {
            const double ang = 2.0 * pi * (double)(t * k1) / kNc;
            tw1[(size_t)(k1 - 1) * kThreads + t] = {(float)std::cos(ang), (float)-std::sin(ang)};
        }
    int rc = AFS_OK;
    if ((rc = upload(&pl->f_tw1, tw1))) return rc;
    if ((rc = upload(&pl->f_wsp, wsp)) || (rc = upload(&pl->u_paddr, paddr)) || (rc = upload(&pl->i_cls, cls)) ||
        (rc = upload(&pl->f_wdense, wdense)) || (rc = upload(&pl->u_kdense, kdense)))
        return rc;
    pl->bpt = bpt;
    pl->nd = (int)dense.size();
    pl->fast_ok = true;
    return AFS_OK;
}

extern "C" {

int afs_chroma_plan_create(afs_chroma_plan **out, const double *h_filterbank, int n_fft, int hop, int n_chroma)
{
    if (!out || !h_filterbank) return afs::fail(AFS_ERR_INVALID, "afs_chroma_plan_create: null argument");
    if (n_fft != kNfft || n_chroma != kChroma)
        return afs::fail(AFS_ERR_UNSUPPORTED, "afs_chroma: only n_fft = 4096 and 12 chroma bins are implemented (got %d, %d)", n_fft, n_chroma);
    if (hop <= 0 || hop > n_fft || (hop & 1)) return afs::fail(AFS_ERR_INVALID, "afs_chroma: hop must be even and in (0, n_fft]");
    afs_chroma_plan *pl = new afs_chroma_plan();
    pl->hop = hop;
    const double pi = 3.14159265358979323846;
    std::vector<double> hann(kNfft), fb((size_t)kBins * kChroma);
    std::vector<Cx<double>> t2(kNc), t4(1025);
    for (int n = 0; n < kNfft; n++) hann[n] = 0.5 - 0.5 * std::cos(2.0 * pi * n / (kNfft - 1));   // np.hanning (symmetric)
    for (int j = 0; j < kNc; j++) t2[j] = {std::cos(2.0 * pi * j / kNc), -std::sin(2.0 * pi * j / kNc)};
    for (int k = 0; k <= 1024; k++) t4[k] = {std::cos(2.0 * pi * k / kNfft), -std::sin(2.0 * pi * k / kNfft)};
    for (int c = 0; c < kChroma; c++)
        for (int k = 0; k < kBins; k++) fb[(size_t)k * kChroma + c] = h_filterbank[(size_t)c * kBins + k];
    std::vector<float> hannf(hann.begin(), hann.end()), fbf(fb.begin(), fb.end());
    std::vector<Cx<float>> t2f(kNc), t4f(1025);
    for (int j = 0; j < kNc; j++) t2f[j] = {(float)t2[j].x, (float)t2[j].y};
    for (int k = 0; k <= 1024; k++) t4f[k] = {(float)t4[k].x, (float)t4[k].y};
    int rc = AFS_OK;
    if ((rc = upload(&pl->d_hann, hann)) || (rc = upload(&pl->d_fb, fb)) || (rc = upload(&pl->d_tw2048, t2)) ||
        (rc = upload(&pl->d_tw4096, t4)) || (rc = upload(&pl->f_hann, hannf)) || (rc = upload(&pl->f_fb, fbf)) ||
        (rc = upload(&pl->f_tw2048, t2f)) || (rc = upload(&pl->f_tw4096, t4f))) {
        afs_chroma_plan_destroy(pl);
        return rc;
    }
    cudaError_t e = cudaFuncSetAttribute(chroma_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chroma_smem_bytes<double>());
    if (e != cudaSuccess) {
        afs_chroma_plan_destroy(pl);
        return afs::fail(AFS_ERR_CUDA, "afs_chroma_plan_create: %s", cudaGetErrorString(e));
    }
    if (int rc2 = build_fast_tables(pl, fb)) {
        afs_chroma_plan_destroy(pl);
        return rc2;
    }
    if (int rc3 = chroma_tc_create(&pl->tc, fb, hann)) {
        afs_chroma_plan_destroy(pl);
        return rc3;
    }
    *out = pl;
    return AFS_OK;
}

int afs_chroma_plan_destroy(afs_chroma_plan *pl)
{
    if (!pl) return AFS_OK;
    cudaFree(pl->f_hann); cudaFree(pl->f_fb); cudaFree(pl->f_tw2048); cudaFree(pl->f_tw4096);
    cudaFree(pl->d_hann); cudaFree(pl->d_fb); cudaFree(pl->d_tw2048); cudaFree(pl->d_tw4096);
    chroma_tc_destroy(pl->tc);
    if (pl->last_done) cudaEventDestroy(pl->last_done);
    cudaFree(pl->f_wsp); cudaFree(pl->f_wdense); cudaFree(pl->f_tw1); cudaFree(pl->u_paddr); cudaFree(pl->u_kdense); cudaFree(pl->i_cls);
    delete pl;
    return AFS_OK;
}

int64_t afs_chroma_num_frames(const afs_chroma_plan *pl, int64_t n_samples, int center_pad)
{
    if (!pl || n_samples < 0) return 0;
    // chroma.py:49-54: x = [zeros(L/2), wav]; num_hops = floor((len(x) - L) / H) + 1
    const int64_t n = n_samples + (center_pad ? pl->n_fft / 2 : 0);
    if (n < pl->n_fft) return 0;
    return (n - pl->n_fft) / pl->hop + 1;
}

}  // extern "C"

template <typename T>
static int launch_chroma(const ChromaTables<T> &tb, const ChromaBatch &bt, cudaStream_t st)
{
    const size_t smem = chroma_smem_bytes<T>();
    int occ = 0;
    AFS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, chroma_kernel<T>, kThreads, smem));
    if (occ < 1) return afs::fail(AFS_ERR_CUDA, "chroma kernel does not fit on an SM");
    int64_t blocks = (int64_t)afs::sm_count() * occ;
    if (blocks > bt.total_frames) blocks = bt.total_frames;
    chroma_kernel<T><<<(unsigned)blocks, kThreads, smem, st>>>(tb, bt);
    afs::count_launch();
    AFS_CUDA(cudaGetLastError());
    return AFS_OK;
}

static int chroma_launch(afs_chroma_plan *pl, ChromaBatch &bt, int compute_dtype, cudaStream_t st)
{
    if (compute_dtype == AFS_BF16X3) {
        if (!pl->tc) return afs::fail(AFS_ERR_UNSUPPORTED, "afs_chroma_batch: the tensor-core path is not available for this plan");
        return chroma_tc_run(pl->tc, bt, st);
    }
    if (compute_dtype == AFS_F32 && pl->fast_ok && !bt.stft_out) {
        ChromaFastTables ft{pl->f_hann, reinterpret_cast<const float2 *>(pl->f_tw2048), reinterpret_cast<const float2 *>(pl->f_tw4096),
                            reinterpret_cast<const float2 *>(pl->f_tw1), pl->f_wsp, pl->u_paddr, pl->i_cls, pl->f_wdense, pl->u_kdense, pl->bpt, pl->nd};
        auto kern = bt.pcm16 ? (pl->bpt == 17 ? chroma_fast_kernel<17, true> : chroma_fast_kernel<0, true>)
                             : (pl->bpt == 17 ? chroma_fast_kernel<17, false> : chroma_fast_kernel<0, false>);
        int occ = 0;
        AFS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, 0));
        if (occ < 1) return afs::fail(AFS_ERR_CUDA, "chroma fast kernel does not fit on an SM");
        int64_t blocks = (int64_t)afs::sm_count() * occ;
        if (blocks > bt.total_frames) blocks = bt.total_frames;
        kern<<<(unsigned)blocks, kThreads, 0, st>>>(ft, bt);
        afs::count_launch();
        AFS_CUDA(cudaGetLastError());
        return AFS_OK;
    }
    if (compute_dtype == AFS_F32) {
        ChromaTables<float> tb{pl->f_hann, pl->f_tw2048, pl->f_tw4096, pl->f_fb};
        return launch_chroma<float>(tb, bt, st);
    }
    ChromaTables<double> tb{pl->d_hann, pl->d_tw2048, pl->d_tw4096, pl->d_fb};
    return launch_chroma<double>(tb, bt, st);
}

static int chroma_batch_impl(afs_chroma_plan *pl, const void *d_audio, int pcm16, const int64_t *h_offsets, int n_tracks,
                             int center_pad, int normalize, void *d_out, const int64_t *h_out_offsets, int out_dtype,
                             int compute_dtype, void *stream, void *d_stft = nullptr)
{
    if (!pl || !d_audio || !h_offsets || (!d_out && !d_stft) || n_tracks <= 0)
        return afs::fail(AFS_ERR_INVALID, "afs_chroma_batch: null argument or n_tracks <= 0");
    if ((out_dtype != AFS_F32 && out_dtype != AFS_F64) ||
        (compute_dtype != AFS_F32 && compute_dtype != AFS_F64 && compute_dtype != AFS_BF16X3))
        return afs::fail(AFS_ERR_INVALID, "afs_chroma_batch: bad dtype");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // per-call offsets: sample_off | frame_off | out_off.  The device copy is allocated and freed in stream order,
    // so concurrent calls (other streams, other threads) never share it.
    std::vector<int64_t> meta((size_t)3 * n_tracks + 2, 0);
    int64_t *s_off = meta.data(), *f_off = s_off + n_tracks + 1, *o_off = f_off + n_tracks + 1;
    f_off[0] = 0;
    for (int k = 0; k < n_tracks; k++) {
        if (h_offsets[k + 1] < h_offsets[k]) return afs::fail(AFS_ERR_INVALID, "afs_chroma_batch: offsets must be non-decreasing");
        if (h_offsets[k] & 1) return afs::fail(AFS_ERR_INVALID, "afs_chroma_batch: track offsets must be even (8-byte aligned samples)");
        s_off[k] = h_offsets[k];
        f_off[k + 1] = f_off[k] + afs_chroma_num_frames(pl, h_offsets[k + 1] - h_offsets[k], center_pad);
        o_off[k] = h_out_offsets ? h_out_offsets[k] : f_off[k];
    }
    s_off[n_tracks] = h_offsets[n_tracks];
    const int64_t total = f_off[n_tracks];
    if (total == 0) return AFS_OK;
    std::lock_guard<std::mutex> lock(pl->mu);
    if (pl->has_last && pl->last_stream != st) AFS_CUDA(cudaStreamWaitEvent(st, pl->last_done, 0));
    int64_t *d_meta = nullptr;
    AFS_CUDA(cudaMallocAsync(&d_meta, sizeof(int64_t) * meta.size(), st));
    // pageable source: the runtime stages it before cudaMemcpyAsync returns, so `meta` may go out of scope
    cudaError_t ce = cudaMemcpyAsync(d_meta, meta.data(), sizeof(int64_t) * meta.size(), cudaMemcpyHostToDevice, st);
    if (ce != cudaSuccess) {
        cudaFreeAsync(d_meta, st);
        return afs::fail(AFS_ERR_CUDA, "afs_chroma_batch: %s", cudaGetErrorString(ce));
    }
    ChromaBatch bt;
    bt.audio = d_audio;
    bt.pcm16 = pcm16;
    bt.sample_off = d_meta;
    bt.frame_off = d_meta + n_tracks + 1;
    bt.out_off = d_meta + 2 * (n_tracks + 1);
    bt.n_tracks = n_tracks;
    bt.total_frames = total;
    bt.hop = pl->hop;
    bt.center_pad = center_pad;
    bt.normalize = normalize;
    bt.out_f64 = out_dtype == AFS_F64;
    bt.out = d_out;
    bt.stft_out = d_stft;
    const int rc = chroma_launch(pl, bt, compute_dtype, st);
    cudaFreeAsync(d_meta, st);
    if (rc == AFS_OK) {
        if (!pl->last_done) AFS_CUDA(cudaEventCreateWithFlags(&pl->last_done, cudaEventDisableTiming));
        AFS_CUDA(cudaEventRecord(pl->last_done, st));
        pl->last_stream = st;
        pl->has_last = true;
    }
    return rc;
}

extern "C" int afs_chroma_batch(afs_chroma_plan *pl, const float *d_audio, const int64_t *h_offsets, int n_tracks,
                                int center_pad, int normalize, void *d_out, const int64_t *h_out_offsets, int out_dtype,
                                int compute_dtype, void *stream)
{
    return chroma_batch_impl(pl, d_audio, 0, h_offsets, n_tracks, center_pad, normalize, d_out, h_out_offsets, out_dtype,
                             compute_dtype, stream);
}

extern "C" int afs_chroma_batch_pcm16(afs_chroma_plan *pl, const int16_t *d_pcm, const int64_t *h_offsets, int n_tracks,
                                      int center_pad, int normalize, void *d_out, const int64_t *h_out_offsets, int out_dtype,
                                      int compute_dtype, void *stream)
{
    return chroma_batch_impl(pl, d_pcm, 1, h_offsets, n_tracks, center_pad, normalize, d_out, h_out_offsets, out_dtype,
                             compute_dtype, stream);
}

// create_stft (chroma.py:44-65): the complex spectrum itself, [frame][2049] interleaved (re, im) of the compute type
// (complex64 for AFS_F32, complex128 for AFS_F64); frame k of track t sits at row h_out_offsets[t] + k.
extern "C" int afs_stft_batch(afs_chroma_plan *pl, const float *d_audio, const int64_t *h_offsets, int n_tracks, int center_pad,
                              void *d_spec, const int64_t *h_out_offsets, int compute_dtype, void *stream)
{
    if (!d_spec) return afs::fail(AFS_ERR_INVALID, "afs_stft_batch: null output");
    if (compute_dtype != AFS_F32 && compute_dtype != AFS_F64) return afs::fail(AFS_ERR_INVALID, "afs_stft_batch: compute_dtype must be AFS_F32 or AFS_F64");
    return chroma_batch_impl(pl, d_audio, 0, h_offsets, n_tracks, center_pad, 0, nullptr, h_out_offsets, AFS_F32, compute_dtype, stream,
                             d_spec);
}

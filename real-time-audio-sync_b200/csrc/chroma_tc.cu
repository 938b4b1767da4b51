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
// One persistent launch (chroma_tc_spectrum_kernel), one CTA per SM, two kinds of CTA:
//  * spectrum CTAs (all but 16): groups of 4 frames flow through warp-specialised roles that only meet through mbarriers
//      L   TMA bulk copies of the audio into a shared-memory ring (the memory system answers after ~1.6 us, so ~50 KB per
//          SM have to be in flight; overlapping frames share the hop they have in common)
//      C   samples x window -> bf16 hi/lo -> A operand of stage 1, written straight into TMEM (tcgen05.st)
//      M   one elected lane issues every tcgen05.mma:  D1[(f, n2)][64] = A1 . F^T   (12 MMAs per 2 frames, N = 64),
//          D2[(k2, re/im)][(frame, k1)] = G . Y'^T   (24 MMAs per 4 frames, N = 144; G lives in TMEM as the A operand)
//      E1  tcgen05.ld of D1, twiddle in registers (the thread's n2 is fixed, so its twiddles are too), split, store as the
//          B operand of stage 2 in shared memory: row (frame, k1), K = (n2, re/im)
//      E2  tcgen05.ld of D2, squares; re^2 + im^2 meet through one shuffle between neighbouring lanes; split; the power
//          spectrum goes to a ring in global memory (28 MB, L2 resident) already laid out as the next MMA's A operand
//      P   publishes stored groups to the filterbank CTAs (gpu-scope release) and checks that ring slots are free again
//  * filterbank CTAs (16): [64 frames x (hi | lo) planes] x [bins x 12] as tcgen05.mma with the weight matrix resident
//    in shared memory and the power spectrum streamed in by TMA bulk copies, then L2 normalisation and the
//    feature-major store.
// What bounds it (profiles/README.md): the CUDA-core work around the MMAs — 3 500 warp instructions per frame, a quarter
// of them the bf16 hi/lo splits — not the tensor pipe (720 cycles per frame) and not HBM.
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <mutex>

#include "chroma_tc.cuh"
#include "tc05.cuh"

namespace {

constexpr int kNfft = 4096;
constexpr int kBins = 2049;
constexpr int kChroma = 12;

// ---------------- kernel A geometry ----------------
// Warp-specialised, one CTA per SM, persistent over groups of 4 frames (two stage-1 tiles of 2 frames each):
//   warps  0..7   C   convert:   staged audio -> x window -> bf16 hi/lo -> A operand of stage 1, written straight into TMEM
//   warps  8..15  E1  epilogue 1: D1 (TMEM) -> twiddle -> bf16 hi/lo -> Y' (shared memory, B operand of stage 2)
//   warps 16..23  E2  epilogue 2: D2 (TMEM) -> |X|^2 -> bf16 hi/lo planes of the power spectrum (global scratch, L2)
//   warp  24      M   one thread issues every tcgen05.mma; completion is tracked with tcgen05.commit -> mbarrier
//   warps 25, 26  L   loaders: TMA bulk copies (cp.async.bulk) of the audio into a shared-memory ring; issuing a copy or an
//                     mbarrier transaction costs a warp 100-200 cycles, so two warps take alternate slots
//   warp  27      P   publisher: the gpu-scope release of a stored group (and the check that its ring slot is free) take
//                     a round trip to L2, so they run here and not in the epilogue warps
// The roles only meet through mbarriers, so while the tensor pipe works on one tile the converters are two tiles
// ahead and the two epilogues drain the previous results.
constexpr int kWarpsC = 8, kWarpsE1 = 8, kWarpsE2 = 8;
constexpr int kWarpsL = 2;
constexpr int kThreadsA = 32 * (kWarpsC + kWarpsE1 + kWarpsE2 + 1 + kWarpsL + 1);      // 896: + M + L + P
constexpr int kRowsPerFrame = 36;              // 33 k1 rows + 3 pad rows (N of stage 2 must be a multiple of 16)
constexpr int kN1 = 64;                        // stage-1 N: 64 real outputs
constexpr int kN2 = 4 * kRowsPerFrame;         // stage-2 N: 144
constexpr int kPlaneWords = 9 * 128;           // power spectrum: 9 words per lane per frame and plane (2 bf16 each)
constexpr uint32_t kFPlane = 64 * 128;         // F^T image, one plane (hi or lo)
constexpr uint32_t kYBlk = kN2 * 128;          // stage-2 B operand: one 64-element K block of one plane
constexpr uint32_t kYPlane = 2 * kYBlk;        // K = 128
constexpr uint32_t kYBuf = 2 * kYPlane;        // hi | lo
constexpr uint32_t kOffF = 0, kOffY = kOffF + 2 * kFPlane, kOffWin = kOffY + 2 * kYBuf;
// Audio ring.  A slot holds the samples one half (16 of 32 n1-rows per half frame) of a tile needs, as three pieces of
// 1024 samples: [first half of frame 0 | second half of frame 0 = first half of frame 1 | second half of frame 1] —
// consecutive frames of a track overlap by one hop, so the middle piece serves both.  The memory system answers a
// bulk copy after ~1.6 us, so the ring has to hold ~50 KB per SM in flight to stream at the rate the tensor pipe
// consumes; that is why G lives in TMEM and the stage-1 A operand is single-buffered.  The window table holds the
// first half of the (symmetric) Hann window.
constexpr int kRingSlots = 4;                  // power of two
constexpr uint32_t kPieceBytes = 1024 * 4, kSlotBytes = 3 * kPieceBytes;
constexpr uint32_t kOffRing = kOffWin + kNfft * 4;      // window: w[n] for n < 2048, then w[2047 - m] = w[2048 + m] for m < 2048
constexpr uint32_t kSmemA = kOffRing + kRingSlots * kSlotBytes;          // 229 376 B
// TMEM columns: G_hi [0,64) | G_lo [64,128) | A1 hi [128,160) lo [160,192) | D1[s] [192 + 64 s, +64) | D2 [320, 464)
constexpr uint32_t kColGhi = 0, kColGlo = 64, kColA1 = 128, kColD1 = 192, kColD2 = 320;
// mbarriers
enum { kBarA1Full = 0, kBarM1Done = 1, kBarD1Free = 3, kBarYFull = 5, kBarM2Done = 7, kBarD2Free = 9, kBarStored = 10, kBarSlotOk = 11,
       kBarRingFull = 13, kBarRingEmpty = 13 + kRingSlots, kNumBars = 13 + 2 * kRingSlots };

// what the loader tells the converters about a ring slot
struct SlotInfo {
    long long base1;    // frame 1 of the tile: element index of its sample 0, and the samples that exist
    int lo1, hi1;
    int shared;         // 1: frame 1 starts one hop after frame 0 in the same track (the middle piece serves both)
    int pad;
};

struct FrameMeta {
    long long base;     // element index of the frame's sample 0 in the audio array (may point before the track)
    int lo, hi;         // samples lo <= n < hi of the frame exist, the rest is the reference's zero padding
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
    // power spectrum ring (global memory, sized to stay in L2): tile slot = 64 frames, planes hi | lo
    uint32_t *p_hi, *p_lo;                     // [ring_tiles * 64][kPlaneWords]
    int ring_tiles;
    int *prod, *cons;                          // per slot: groups stored so far / times consumed so far (monotonic)
    int n_fb, n_spec;                          // CTAs [0, n_fb) run the filterbank role, the rest the spectrum role
    // filterbank role
    const uint4 *w_img;                        // [36][4 KB]
    const int64_t *out_off;
    int normalize, out_f64;
    void *out;
    long long *trace;                          // optional (AFS_CHROMA_TC_TRACE): [block][warp][kTraceLen] clock64 stamps
};
constexpr int kTraceLen = 256;

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&v)[4])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
                 : "r"(taddr)
                 : "memory");
}

__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
// mbarrier wait that backs off with nanosleep between probes, for the roles that idle most of the time (epilogues,
// loaders, publisher).  A bare try_wait loop re-issues every few hundred cycles (the suspend-time hint does not change
// that on this part): those probes were 53 % of all instructions the kernel executed (profiles/ncu_lines_r2_chroma_tc.txt).
// Measured effect on throughput: none either way (169 vs 166 Mframes/s: the probes only take issue slots nobody else
// wanted), while sleeping on the converter <-> issuer chain costs 15 % (wake-up latency) — so those two spin.
#ifndef AFS_TC_WAIT_NS
#define AFS_TC_WAIT_NS 200
#endif
// The only waits that span CTAs are the two ring counters.  They cannot deadlock (see the publisher), but a wait on another
// CTA is where a bug would turn into a hung GPU, so they are bounded: after kRingWaitNs the kernel traps (the launch fails
// with an error instead of never returning).
constexpr unsigned long long kRingWaitNs = 4000000000ull;
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ bool mbar_try(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done)
                 : "r"(afs::smem_addr(bar)), "r"(parity)
                 : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_wait_sleep(uint64_t *bar, uint32_t parity)
{
    if (mbar_try(bar, parity)) return;
    do {
        __nanosleep(AFS_TC_WAIT_NS);
    } while (!mbar_try(bar, parity));
}
// the waits on the critical chain (converter <-> MMA issuer) keep probing
__device__ __forceinline__ void mbar_wait_spin(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try(bar, parity)) {
    }
}
// keep the computation of v before whatever volatile operation follows (the scheduler otherwise sinks it behind the wait)
__device__ __forceinline__ void pin16(uint32_t (&v)[16])
{
    asm volatile("" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]),
                 "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(afs::smem_addr(bar)) : "memory");
}

// Where the frame `f` (global numbering) lives.  The cursor keeps the current track's offsets in registers, so the
// common case (next frame, same track) touches no memory; frames only move forward.
struct TrackCursor {
    int track = -1;
    int64_t f_begin = 0, f_end = 0, s_begin = 0, n_samp = 0;
};
__device__ __forceinline__ FrameMeta locate_frame(const SpectrumArgs &a, int64_t f, TrackCursor &c)
{
    FrameMeta m;
    m.base = 0; m.lo = 0; m.hi = 0;
    if (f >= a.frame_end) return m;
    if (c.track < 0 || f >= c.f_end) {
        int t = c.track < 0 ? 0 : c.track;
        while (t + 1 < a.n_tracks && __ldg(a.frame_off + t + 1) <= f) t++;
        c.track = t;
        c.f_begin = __ldg(a.frame_off + t);
        c.f_end = __ldg(a.frame_off + t + 1);
        c.s_begin = __ldg(a.sample_off + t);
        c.n_samp = __ldg(a.sample_off + t + 1) - c.s_begin;
    }
    const int64_t start = (f - c.f_begin) * a.hop - (a.center_pad ? kNfft / 2 : 0);     // chroma.py:49 left zero pad
    m.base = c.s_begin + start;
    m.lo = start < 0 ? (int)(-start) : 0;
    const int64_t hi = c.n_samp - start;
    m.hi = hi > kNfft ? kNfft : (hi < 0 ? 0 : (int)hi);
    return m;
}

// ---------------- filterbank role: [64 frames x (hi | lo) planes] x [bins x 12] + normalisation ----------------
// Runs on a few CTAs of the same launch: it consumes tiles of 64 frames from the power-spectrum ring as soon as the 16
// spectrum CTAs that produce a tile have stored their groups (prod counter, release/acquire through L2) and hands the
// slot back (cons counter).  The ring is a few tens of MB, so the power spectrum never leaves L2.  A tile slot is laid
// out as the shared-memory image of the MMA's A operand — [36 K blocks][128 rows][128 B, 128-byte swizzle], row
// 2 r + plane for frame r — so one cp.async.bulk of 16 KB per K block brings it in.
//   warp 0      producer: waits for the tile, streams its 36 K blocks through a 3-stage ring, two K blocks (32 KB) per bulk
//               copy — an mbarrier transaction or a bulk copy costs the issuing warp 100-200 cycles, so few and large ones
//   warp 1      MMA issuer (elected lane): 8 MMAs (N = 32) per stage into one of two TMEM accumulators, which start from
//               zero (cleared by the epilogue).  (Two issuer warps on alternate K blocks were tried: the kernel hangs.)
//   warps 2..5  epilogue: TMEM -> add the hi/lo partial products (neighbouring lanes) -> normalise -> store
constexpr int kFbTile = 64;                      // frames per tile
constexpr int kFbGroups = kFbTile / 4;           // spectrum groups per tile
constexpr int kFbKBlocks = kPlaneWords * 2 / 64; // 36 blocks of 64 bf16
constexpr int kFbStages = 3;                     // of kFbStageBlocks K blocks each
constexpr int kFbStageBlocks = 2;
#ifndef AFS_FB_MMA_WARPS
#define AFS_FB_MMA_WARPS 1
#endif
constexpr int kFbMmaWarps = AFS_FB_MMA_WARPS;
constexpr int kFbWarps = 5 + kFbMmaWarps;
// weights per K block: 24 rows (12 hi | 12 lo) x 64 elements.  The MMA reads N = 32 rows: rows 24..31 alias the first
// eight rows of the next K block (behind the last block: the first A stage) and only fill accumulator columns 24..31,
// which nobody reads.
constexpr uint32_t kFbWBlk = 24 * 128;
constexpr uint32_t kFbABlk = 128 * 128;
constexpr uint32_t kFbTileBytes = kFbKBlocks * kFbABlk;                      // 589 824 B per ring slot
constexpr uint32_t kFbStageBytes = kFbStageBlocks * kFbABlk;
constexpr uint32_t kSmemFb = kFbKBlocks * kFbWBlk + kFbStages * kFbStageBytes;     // 110 592 + 98 304
static_assert(kSmemFb <= kSmemA, "the filterbank role lives in the spectrum kernel's shared memory");
enum { kFbBarFull = 0, kFbBarFree = kFbStages, kFbBarAccFull = 2 * kFbStages, kFbBarAccFree = 2 * kFbStages + 2, kFbNumBars = 2 * kFbStages + 4 };

__device__ __forceinline__ void named_bar(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

template <bool TRACE>
__device__ __noinline__ void filterbank_role(const SpectrumArgs &args, unsigned char *smem)
{
    __shared__ uint32_t s_tmem_fb;
    __shared__ __align__(8) uint64_t s_fbar[kFbNumBars];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp >= kFbWarps) return;
    constexpr int kThreads = 32 * kFbWarps;
    const uint32_t s_w = afs::smem_addr(smem), s_a = s_w + kFbKBlocks * kFbWBlk;
    {
        uint4 *dst = reinterpret_cast<uint4 *>(smem);
        for (int i = tid; i < (int)(kFbKBlocks * kFbWBlk / 16); i += kThreads) dst[i] = __ldg(args.w_img + i);
    }
    if (tid == 0) {
        for (int s = 0; s < kFbStages; s++) {
            afs::mbar_init(&s_fbar[kFbBarFull + s], 1);
            afs::mbar_init(&s_fbar[kFbBarFree + s], 1);
        }
        for (int a = 0; a < 2; a++) {
            afs::mbar_init(&s_fbar[kFbBarAccFull + a], kFbMmaWarps);
            afs::mbar_init(&s_fbar[kFbBarAccFree + a], 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc::fence_async_smem();
    if (warp == 0) tc::tmem_alloc(&s_tmem_fb, 64);
    tc::fence_before_sync();
    named_bar(1, kThreads);
    tc::fence_after_sync();
    const uint32_t tm = s_tmem_fb;
    if (warp >= 1 + kFbMmaWarps) {
        // both accumulators start from zero
        uint32_t z[16];
#pragma unroll
        for (int j = 0; j < 16; j++) z[j] = 0u;
#pragma unroll
        for (int c = 0; c < 64; c += 16) tc::tmem_st16(tm + ((uint32_t)((warp & 3) * 32) << 16) + c, z);
        tc::tmem_wait_st();
    }
    tc::fence_before_sync();
    named_bar(1, kThreads);
    tc::fence_after_sync();

    const int64_t n_frames = args.frame_end - args.frame_begin;
    const int64_t n_groups = (n_frames + 3) >> 2;
    const int64_t n_tiles = (n_frames + kFbTile - 1) / kFbTile;
    const unsigned char *ring = reinterpret_cast<const unsigned char *>(args.p_hi);
    auto trace = [&](int ev, int idx) {
        if (TRACE && lane == 0 && 4 * idx + ev < kTraceLen)
            args.trace[((size_t)(args.n_spec + blockIdx.x) * (kThreadsA / 32) + warp) * kTraceLen + 4 * idx + ev] = clock64();
    };

    if (warp == 0) {
        // ================= producer =================
        uint32_t seq = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += args.n_fb) {
            const int slot = (int)(tile % args.ring_tiles), use = (int)(tile / args.ring_tiles);
            if (lane == 0) {
                const int64_t left = n_groups - tile * kFbGroups;
                const int want = use * kFbGroups + (int)(left < kFbGroups ? left : kFbGroups);
                const unsigned long long t0 = global_ns();
                while (afs::ld_acquire(args.prod + slot) < want) {
                    __nanosleep(100);
                    if (global_ns() - t0 > kRingWaitNs) __trap();
                }
                // the tile was written with ordinary stores by other SMs; the bulk copies below read it through the async proxy
                asm volatile("fence.proxy.async.global;" ::: "memory");
            }
            __syncwarp();
            const unsigned char *src = ring + (size_t)slot * kFbTileBytes;
            for (int kb = 0; kb < kFbKBlocks; kb += kFbStageBlocks, seq++) {
                const uint32_t st = seq % kFbStages, u = seq / kFbStages;
                if (u > 0) afs::mbar_wait(&s_fbar[kFbBarFree + st], (u - 1) & 1);
                trace(0, seq);
                if (lane == 0) {
                    afs::mbar_expect_tx(&s_fbar[kFbBarFull + st], kFbStageBytes);
                    afs::bulk_g2s(smem + kFbKBlocks * kFbWBlk + st * kFbStageBytes, src + (size_t)kb * kFbABlk, kFbStageBytes, &s_fbar[kFbBarFull + st]);
                }
                __syncwarp();
                trace(1, seq);
            }
        }
    } else if (warp <= kFbMmaWarps) {
        // ================= MMA issuers: warp 1 takes the even K blocks, warp 2 the odd ones =================
        constexpr uint32_t idesc = tc::idesc_bf16_f32(128, 32);
        const uint64_t w_desc = tc::smem_desc_k_sw128(s_w), a_desc = tc::smem_desc_k_sw128(s_a);
        const int mw = warp - 1;
        uint32_t seq = 0, it = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += args.n_fb, it++) {
            const uint32_t acc = it & 1, au = it >> 1;
            if (au > 0) afs::mbar_wait(&s_fbar[kFbBarAccFree + acc], (au - 1) & 1);
            tc::fence_after_sync();
            for (int kb = 0; kb < kFbKBlocks; kb += kFbStageBlocks, seq++) {
                if ((int)(seq % kFbMmaWarps) != mw) continue;
                const uint32_t st = seq % kFbStages, u = seq / kFbStages;
                trace(2, seq);
                afs::mbar_wait(&s_fbar[kFbBarFull + st], u & 1);
                tc::fence_after_sync();
                trace(0, seq);
#pragma unroll
                for (int b2 = 0; b2 < kFbStageBlocks; b2++) {
                    const uint64_t ad = a_desc + (uint64_t)((st * kFbStageBytes + b2 * kFbABlk) >> 4);
                    const uint64_t wd = w_desc + (uint64_t)(((uint32_t)(kb + b2) * kFbWBlk) >> 4);
#pragma unroll
                    for (int ks = 0; ks < 4; ks++) tc::mma_ss_elect(tm + 32 * acc, ad + 2 * ks, wd + 2 * ks, idesc, 1u);
                }
                tc::mma_commit_elect(&s_fbar[kFbBarFree + st]);
                trace(1, seq);
            }
            tc::mma_commit_elect(&s_fbar[kFbBarAccFull + acc]);
        }
    } else {
        // ================= epilogue: warp w reads TMEM lanes 32 (w % 4) ..; lane 2 r + plane belongs to frame r =================
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const int r = ((warp & 3) * 32 + lane) >> 1;             // frame of the tile
        const bool lo_plane = lane & 1;
        uint32_t it = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += args.n_fb, it++) {
            const uint32_t acc = it & 1, au = it >> 1;
            // this lane's output frame, located while the tile is being accumulated
            const int64_t f = args.frame_begin + tile * kFbTile + r;
            const bool okf = !lo_plane && f < args.frame_end;
            int trk = 0;
            if (okf) {
                int lo = 0, hi = args.n_tracks;
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (__ldg(args.frame_off + mid) <= f) lo = mid; else hi = mid;
                }
                trk = lo;
            }
            afs::mbar_wait(&s_fbar[kFbBarAccFull + acc], au & 1);
            tc::fence_after_sync();
            // all MMAs of the tile are done, so every K block has left the ring slot: hand it back to the spectrum CTAs
            if (warp == 1 + kFbMmaWarps && lane == 0)
                afs::st_release(args.cons + (int)(tile % args.ring_tiles), (int)(tile / args.ring_tiles) + 1);
            uint32_t va[16], vb[16];
            tc::tmem_ld16(tm + lane_base + 32 * acc, va);
            tc::tmem_ld16(tm + lane_base + 32 * acc + 16, vb);
            tc::tmem_wait_ld();
            {
                uint32_t z[16];
#pragma unroll
                for (int j = 0; j < 16; j++) z[j] = 0u;
                tc::tmem_st16(tm + lane_base + 32 * acc, z);
                tc::tmem_st16(tm + lane_base + 32 * acc + 16, z);
                tc::tmem_wait_st();
            }
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_fbar[kFbBarAccFree + acc]);
            float raw[kChroma];
            float ss = 0.f;
#pragma unroll
            for (int c = 0; c < kChroma; c++) {
                // columns c: w_hi, 12 + c: w_lo; the lo-plane lane only contributes its w_hi product (lo * lo is noise)
                const float w_lo_part = __uint_as_float(12 + c < 16 ? va[(12 + c) & 15] : vb[(12 + c) & 15]);
                const float mine = lo_plane ? __uint_as_float(va[c]) : __uint_as_float(va[c]) + w_lo_part;
                raw[c] = mine + __shfl_down_sync(0xffffffffu, mine, 1);                 // chroma.py:70 np.dot(chromafb, spec)
                ss = fmaf(raw[c], raw[c], ss);
            }
            if (okf) {
                float len = 1.f;
                if (args.normalize) {
                    // librosa.util.normalize(norm=2, axis=0), chroma.py:74: tiny lengths -> 1
                    len = sqrtf(ss);
                    if (len < FLT_MIN) len = 1.f;
                }
                const int64_t m = f - __ldg(args.frame_off + trk);
                const int64_t fk = __ldg(args.frame_off + trk + 1) - __ldg(args.frame_off + trk);
                const int64_t ob = kChroma * __ldg(args.out_off + trk) + m;
#pragma unroll
                for (int c = 0; c < kChroma; c++) {
                    const float val = args.normalize ? raw[c] / len : raw[c];
                    if (args.out_f64) static_cast<double *>(args.out)[ob + (int64_t)c * fk] = (double)val;
                    else static_cast<float *>(args.out)[ob + (int64_t)c * fk] = val;
                }
            }
        }
    }
    tc::fence_before_sync();
    named_bar(1, kThreads);
    if (warp == 0) tc::tmem_dealloc(tm, 64);
}

template <bool PCM16, bool TRACE>
__global__ void __launch_bounds__(kThreadsA, 1) chroma_tc_spectrum_kernel(const SpectrumArgs args)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_bar[kNumBars];
    __shared__ __align__(16) SlotInfo s_slot[kRingSlots];

    if ((int)blockIdx.x < args.n_fb) {
        filterbank_role<TRACE>(args, smem);
        return;
    }
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t s_base = afs::smem_addr(smem);
    const int sb = (int)blockIdx.x - args.n_fb;      // index among the spectrum CTAs

    // ---- one-time setup: operand images, zeroed Y' tiles (pad rows stay zero), window, barriers, TMEM, G_hi -> TMEM ----
    {
        uint4 *dst = reinterpret_cast<uint4 *>(smem + kOffF);
        for (int i = tid; i < (int)(2 * kFPlane / 16); i += kThreadsA) dst[i] = __ldg(args.f_img + i);
        uint4 *y0 = reinterpret_cast<uint4 *>(smem + kOffY);
        for (int i = tid; i < (int)(2 * kYBuf / 16); i += kThreadsA) y0[i] = make_uint4(0, 0, 0, 0);
        float *w = reinterpret_cast<float *>(smem + kOffWin);
        // PCM16: librosa.load's 1/32768 scaling (a power of two) rides on the window
        for (int i = tid; i < kNfft; i += kThreadsA) {
            // np.hanning is symmetric, w[n] = w[4095 - n]: the second half is stored as the mirror of the first
            const int n = i < kNfft / 2 ? i : kNfft - 1 - i;
            w[i] = __ldg(args.hann + n) * (PCM16 ? (1.0f / 32768.0f) : 1.0f);
        }
    }
    if (tid == 0) {
        afs::mbar_init(&s_bar[kBarA1Full], kWarpsC);
        for (int s = 0; s < 2; s++) {
            afs::mbar_init(&s_bar[kBarM1Done + s], 1);
            afs::mbar_init(&s_bar[kBarD1Free + s], kWarpsE1);
            afs::mbar_init(&s_bar[kBarYFull + s], 2 * kWarpsE1);
            afs::mbar_init(&s_bar[kBarM2Done + s], 1);
        }
        afs::mbar_init(&s_bar[kBarD2Free], kWarpsE2);
        afs::mbar_init(&s_bar[kBarStored], kWarpsE2);
        afs::mbar_init(&s_bar[kBarSlotOk], 1);
        afs::mbar_init(&s_bar[kBarSlotOk + 1], 1);
        for (int r = 0; r < kRingSlots; r++) {
            afs::mbar_init(&s_bar[kBarRingFull + r], 1);
            afs::mbar_init(&s_bar[kBarRingEmpty + r], kWarpsC);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc::fence_async_smem();
    if (warp == 0) tc::tmem_alloc(&s_tmem, 512);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    // the CTA owns the whole tensor memory (512 columns, one CTA per SM), so the allocation starts at lane 0, column 0;
    // with a literal base every TMEM address of the issue loop is a compile-time constant
    if (s_tmem != 0) __trap();
    constexpr uint32_t tm = 0;
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
                tc::tmem_st16(tm + ((uint32_t)(tid & ~31) << 16) + (plane ? kColGlo : kColGhi) + c0, v);
            }
        }
        tc::tmem_wait_st();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();

    const int64_t n_frames = args.frame_end - args.frame_begin;
    const int64_t n_groups = (n_frames + 3) >> 2;
    const int n_local = (int)((n_groups - sb + args.n_spec - 1) / args.n_spec);      // groups of this CTA (>= 1: grid <= groups)
    const int n_tiles = 2 * n_local;
    // event trace of this warp (debugging aid, off unless the host passes a buffer): slot = 4 * index + event
    auto trace = [&](int ev, int idx) {
        if (TRACE && lane == 0 && 4 * idx + ev < kTraceLen)
            args.trace[((size_t)sb * (kThreadsA / 32) + warp) * kTraceLen + 4 * idx + ev] = clock64();
    };

    if (warp < kWarpsC) {
        // ================= C: convert =================
        const int ln = tid & 127, wg = tid >> 7;          // ln = TMEM lane = row (frame of the pair, n2); wg = half of n1
        const int n2 = ln & 63, fp = ln >> 6;
        const uint32_t lane_base = (uint32_t)(ln & ~31) << 16;
        constexpr int kSampleBytes = PCM16 ? 2 : 4;
        // this thread's samples of a tile: n = 64 (32 wg + r) + n2, r = 0..31; rows 0..15 come from the tile's first ring
        // slot, rows 16..31 from the second; piece of the slot: frame fp, half wg -> piece fp + wg
        const float *wptr = reinterpret_cast<const float *>(smem + kOffWin) + (kNfft / 2) * wg + n2;
        const unsigned char *xptr = smem + kOffRing + (uint32_t)(fp + wg) * kPieceBytes + n2 * kSampleBytes;
        const bool straddle_warp = fp == 1 && wg == 0;      // the half frame that is missing from the ring when the tile straddles tracks
        const uint32_t a1_col = tm + lane_base + kColA1 + 16 * wg;
        uint32_t stage = 0, ring_phase = 0;
#pragma unroll 1
        for (int T = 0; T < n_tiles; T++) {
            float x[32];
#pragma unroll
            for (int half = 0; half < 2; half++) {
                mbar_wait_spin(&s_bar[kBarRingFull + stage], ring_phase);
                if (straddle_warp && !s_slot[stage].shared) {
                    // the tile straddles two tracks (or ends the chunk): guarded loads from global memory instead
                    const SlotInfo si = s_slot[stage];
#pragma unroll
                    for (int r = 0; r < 16; r++) {
                        const int n = 64 * (16 * half + r) + n2;
                        float v = 0.f;
                        if (n >= si.lo1 && n < si.hi1) {
                            if (PCM16) v = (float)__ldg(static_cast<const short *>(args.audio) + si.base1 + n);
                            else v = __ldg(static_cast<const float *>(args.audio) + si.base1 + n);
                        }
                        x[16 * half + r] = v;
                    }
                } else {
                    const unsigned char *src = xptr + stage * kSlotBytes;
#pragma unroll
                    for (int r = 0; r < 16; r++) {
                        if (PCM16) x[16 * half + r] = (float)*reinterpret_cast<const short *>(src + 64 * r * kSampleBytes);
                        else x[16 * half + r] = *reinterpret_cast<const float *>(src + 64 * r * kSampleBytes);
                    }
                }
#pragma unroll
                for (int r = 0; r < 16; r++) x[16 * half + r] = __fmul_rn(x[16 * half + r], wptr[64 * (16 * half + r)]);   // chroma.py:62 section * np.hanning
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_bar[kBarRingEmpty + stage]);                          // the slot's samples are in registers
                stage = (stage + 1) & (kRingSlots - 1);
                if (stage == 0) ring_phase ^= 1u;
            }
            uint32_t h[16], l[16];
#pragma unroll
            for (int j = 0; j < 16; j++) tc::split_bf16x2(x[2 * j], x[2 * j + 1], h[j], l[j]);
            pin16(h);
            pin16(l);
            trace(1, T);
            // A1 is single-buffered: the MMAs of the previous tile must have read it
            if (T > 0) mbar_wait_spin(&s_bar[kBarM1Done + ((T - 1) & 1)], (uint32_t)(((T - 1) >> 1) & 1));
            trace(2, T);
            tc::tmem_st16(a1_col, h);
            tc::tmem_st16(a1_col + 32, l);
            tc::tmem_wait_st();
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_bar[kBarA1Full]);
            trace(0, T);
        }
    } else if (warp < kWarpsC + kWarpsE1) {
        // ================= E1: twiddle, split, store as rows (frame, k1) of the stage-2 B operand =================
        const int t = tid - 32 * kWarpsC;
        const int ln = t & 127, wg = t >> 7;
        const int n2 = ln & 63, fp = ln >> 6;
        const uint32_t lane_base = (uint32_t)(ln & ~31) << 16;
        // twiddles W4096^(k1 n2): wg 0 handles k1 = 32 (slot 0) and 1..15, wg 1 handles 16..31
        float twc[16], tws[16];
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const int k1 = wg ? 16 + i : (i == 0 ? 32 : i);
            const float2 w = __ldg(args.tw + k1 * 64 + n2);
            twc[i] = w.x;
            tws[i] = w.y;
        }
        const uint32_t kb_off = (uint32_t)(n2 >> 5) * kYBlk + (uint32_t)(n2 & 3) * 4;
        const uint32_t cw16 = ((((uint32_t)(n2 & 31) >> 2) ^ (uint32_t)(4 * fp)) << 4);
#pragma unroll 1
        for (int T = 0; T < n_tiles; T++) {
            const int s = T & 1, use = T >> 1, p = T & 1, gi = T >> 1, b = gi & 1, u = gi >> 1;
            const uint32_t s_y_hi = s_base + kOffY + (uint32_t)b * kYBuf;
            const int fq = 2 * p + fp;                   // frame of the group this row belongs to
            // row = 36 fq + k1: (row & 7) = (k1 & 7) ^ 4 fp, so the swizzled 16-byte chunk is (cw16 ^ (k1 & 7) << 4) with a
            // per-thread constant cw16; the rest of the address has bits 4..6 clear
            const uint32_t y_row0 = s_y_hi + kb_off + (uint32_t)(fq * kRowsPerFrame) * 128;
            auto emit = [&](int k1, float re, float im) {
                uint32_t h, l;
                tc::split_bf16x2(re, im, h, l);
                const uint32_t addr = (y_row0 | (cw16 ^ (uint32_t)((k1 & 7) << 4))) + (uint32_t)k1 * 128;
                sts32(addr, h);
                sts32(addr + kYPlane, l);
            };
            mbar_wait_sleep(&s_bar[kBarM1Done + s], (uint32_t)(use & 1));
            tc::fence_after_sync();
            trace(0, T);
            // Y'[b] was the B operand of stage 2 two groups ago
            if (p == 0 && u > 0) mbar_wait_sleep(&s_bar[kBarM2Done + b], (uint32_t)((u - 1) & 1));
            const uint32_t d1 = tm + lane_base + kColD1 + 64 * s + 32 * wg;
            {
                uint32_t v[16];
                tc::tmem_ld16(d1, v);
                tc::tmem_wait_ld();
                if (wg == 0) {
                    emit(0, __uint_as_float(v[0]), 0.f);                                   // Y[0] is real and its twiddle is 1
                    const float y32 = __uint_as_float(v[1]);                               // Y[32] is real
                    emit(32, y32 * twc[0], y32 * tws[0]);
                }
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    if (i == 0 && wg == 0) continue;
                    const float yr = __uint_as_float(v[2 * i]), yi = __uint_as_float(v[2 * i + 1]);
                    emit(16 * wg + i, yr * twc[i] - yi * tws[i], yr * tws[i] + yi * twc[i]);
                }
            }
            {
                uint32_t v[16];
                tc::tmem_ld16(d1 + 16, v);
                tc::tmem_wait_ld();
                tc::fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_bar[kBarD1Free + s]);
                trace(1, T);
#pragma unroll
                for (int i = 8; i < 16; i++) {
                    const float yr = __uint_as_float(v[2 * i - 16]), yi = __uint_as_float(v[2 * i - 15]);
                    emit(16 * wg + i, yr * twc[i] - yi * tws[i], yr * tws[i] + yi * twc[i]);
                }
            }
            tc::fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_bar[kBarYFull + b]);
            trace(2, T);
        }
    } else if (warp < kWarpsC + kWarpsE1 + kWarpsE2) {
        // ================= E2: lane = (k2, re/im); |X|^2 pairs up through the neighbouring lane; wg takes two frames =================
        const int t = tid - 32 * (kWarpsC + kWarpsE1);
        const int ln = t & 127, wg = t >> 7;
        const uint32_t lane_base = (uint32_t)(ln & ~31) << 16;
        const bool odd = ln & 1;
        for (int gi = 0; gi < n_local; gi++) {
            const int b = gi & 1, u = gi >> 1;
            const int64_t g = sb + (int64_t)gi * args.n_spec;
            const int64_t ring_tile = g / kFbGroups;
            const int ring_slot = (int)(ring_tile % args.ring_tiles);
            mbar_wait_sleep(&s_bar[kBarM2Done + b], (uint32_t)(u & 1));
            tc::fence_after_sync();
            trace(0, gi);
#pragma unroll 1
            for (int q = 0; q < 2; q++) {
                const int fq = 2 * wg + q;
                uint32_t v[36];
                {
                    uint32_t va[16], vb[16], vc[4];
                    const uint32_t c0 = tm + lane_base + kColD2 + fq * kRowsPerFrame;
                    tc::tmem_ld16(c0, va);
                    tc::tmem_ld16(c0 + 16, vb);
                    tmem_ld4(c0 + 32, vc);
                    tc::tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 16; j++) { v[j] = va[j]; v[16 + j] = vb[j]; }
#pragma unroll
                    for (int j = 0; j < 4; j++) v[32 + j] = vc[j];
                }
                if (q == 1) {
                    // everything this warp reads from D2 is in registers: the next group's stage 2 may overwrite it
                    tc::fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&s_bar[kBarD2Free]);
                    trace(1, gi);
                }
                float own[18];
#pragma unroll
                for (int j = 0; j < 17; j++) {
                    // columns k1 = 2j (kept by even lanes) and 2j + 1 (kept by odd lanes); column 33 is padding
                    const float a = __uint_as_float(v[2 * j]), bb = __uint_as_float(v[2 * j + 1]);
                    const float sa = a * a, sb = bb * bb;
                    const float other = __shfl_xor_sync(0xffffffffu, odd ? sa : sb, 1);
                    own[j] = (odd ? sb : sa) + other;
                }
                if (odd) own[16] = 0.f;
                own[17] = 0.f;
                const int64_t fl = 4 * g + fq;                           // frame of the launch
                if (q == 0) mbar_wait_sleep(&s_bar[kBarSlotOk + (gi & 1)], (uint32_t)((gi >> 1) & 1));     // the ring slot is free (publisher warp)
                if (fl < n_frames) {
                    // A-operand image of the filterbank tile: word i * 128 + ln of the frame's plane -> K block 4 i + ln / 32,
                    // row 2 r + plane, 16-byte chunk ((ln % 32) / 4) ^ (row & 7)
                    const uint32_t r = (uint32_t)(fl & (kFbTile - 1));
                    unsigned char *base = reinterpret_cast<unsigned char *>(args.p_hi) + (size_t)ring_slot * kFbTileBytes +
                                          (uint32_t)(ln >> 5) * kFbABlk + (uint32_t)(ln & 3) * 4;
                    const uint32_t row_hi = 2 * r, row_lo = 2 * r + 1, ch = (uint32_t)(ln & 31) >> 2;
                    uint32_t *ph = reinterpret_cast<uint32_t *>(base + row_hi * 128 + ((ch ^ (row_hi & 7)) << 4));
                    uint32_t *pl = reinterpret_cast<uint32_t *>(base + row_lo * 128 + ((ch ^ (row_lo & 7)) << 4));
#pragma unroll
                    for (int i = 0; i < 9; i++) {
                        uint32_t h, l;
                        tc::split_bf16x2(own[2 * i], own[2 * i + 1], h, l);
                        ph[i * (4 * kFbABlk / 4)] = h;
                        pl[i * (4 * kFbABlk / 4)] = l;
                    }
                }
            }
            // the group is stored: the publisher warp releases it at gpu scope
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_bar[kBarStored]);
            trace(2, gi);
        }
    } else if (warp == kWarpsC + kWarpsE1 + kWarpsE2 + 1 + kWarpsL) {
        // ================= P: ring-slot hand-shake with the filterbank CTAs =================
        auto slot_of = [&](int gi, int &slot, int &use) {
            const int64_t tile = (sb + (int64_t)gi * args.n_spec) / kFbGroups;
            slot = (int)(tile % args.ring_tiles);
            use = (int)(tile / args.ring_tiles);
        };
        // Two duties, polled without ever blocking one on the other (a publisher that waits for a slot while it holds back
        // a stored group deadlocks a small ring: the slot may only free up once that very group has been published):
        //   q: next group whose ring slot must be confirmed free (at most two groups ahead of the epilogue: two barriers)
        //   p: next group to publish once the epilogue warps have stored it
        int pg = 0, qg = 0;
        unsigned long long t_idle = global_ns();
        while (pg < n_local) {
            bool progress = false;
            if (qg < n_local && qg < pg + 2) {
                int slot, use;
                slot_of(qg, slot, use);
                int ok = 0;
                if (lane == 0) ok = afs::ld_acquire(args.cons + slot) >= use;       // the slot's previous tile has been consumed
                ok = __shfl_sync(0xffffffffu, ok, 0);
                if (ok) {
                    if (lane == 0) mbar_arrive(&s_bar[kBarSlotOk + (qg & 1)]);
                    qg++;
                    progress = true;
                }
            }
            if (mbar_try(&s_bar[kBarStored], (uint32_t)(pg & 1))) {
                int slot, use;
                slot_of(pg, slot, use);
                if (lane == 0) {
                    __threadfence();
                    atomicAdd(args.prod + slot, 1);
                }
                __syncwarp();
                trace(1, pg);
                pg++;
                progress = true;
            }
            if (progress) {
                t_idle = global_ns();
            } else {
                __nanosleep(100);
                if (global_ns() - t_idle > kRingWaitNs) __trap();
            }
        }
    } else if (warp > kWarpsC + kWarpsE1 + kWarpsE2) {
        // ================= L: audio -> ring.  A piece (1024 consecutive samples) that lies inside its track and
        // starts 16-byte aligned is one cp.async.bulk; anything else (zero padding at the ends of a track, frames past
        // the end of the chunk, odd track offsets) is copied by the warp with guarded loads.  Loader warp hl takes the
        // slot of half hl of every tile; its lanes 0..2 own the three pieces =================
        constexpr int kSampleBytes = PCM16 ? 2 : 4;
        const int hl = warp - (kWarpsC + kWarpsE1 + kWarpsE2 + 1);
        TrackCursor track;
        const int pc_l = lane < 3 ? lane : 0;
        const int n0_l = (pc_l == 0 ? 0 : 2048) + 1024 * hl;
#pragma unroll 1
        for (int T = 0; T < n_tiles; T++) {
            // frames only move forward through the cursor: both frames of the tile are located once per tile
            const int64_t g = sb + (int64_t)(T >> 1) * args.n_spec;
            const int64_t f0 = args.frame_begin + 4 * g + 2 * (T & 1);
            const FrameMeta m0 = locate_frame(args, f0, track);
            const FrameMeta m1 = locate_frame(args, f0 + 1, track);
            // frame 1's first half is the same run of real samples as frame 0's second half
            const bool shared = m1.hi > 0 && m1.lo == 0 && m0.hi == kNfft && m1.base == m0.base + kNfft / 2;
            const long long base_l = pc_l == 2 ? m1.base : m0.base;
            const int lo_l = pc_l == 2 ? m1.lo : m0.lo, hi_l = pc_l == 2 ? m1.hi : m0.hi;
            const char *src_l = static_cast<const char *>(args.audio) + (base_l + n0_l) * kSampleBytes;
            const bool bulk_l = lane < 3 && n0_l >= lo_l && n0_l + 1024 <= hi_l && (reinterpret_cast<uintptr_t>(src_l) & 15) == 0;
            const unsigned qmask = __ballot_sync(0xffffffffu, bulk_l) & 7u;
            const int k = 2 * T + hl, stage = k % kRingSlots, ruse = k / kRingSlots;
            if (ruse > 0) mbar_wait_sleep(&s_bar[kBarRingEmpty + stage], (uint32_t)((ruse - 1) & 1));
            unsigned char *slot = smem + kOffRing + stage * kSlotBytes;
            if (qmask != 7u) {
                // guarded pieces first, so that the one arrival below (release) also publishes them
#pragma unroll
                for (int pc = 0; pc < 3; pc++) {
                    if (qmask & (1u << pc)) continue;
                    const FrameMeta &fm = pc == 2 ? m1 : m0;
                    const int n0 = (pc == 0 ? 0 : 2048) + 1024 * hl;
                    unsigned char *dst = slot + pc * kPieceBytes;
                    for (int i = lane; i < 1024; i += 32) {
                        const int n = n0 + i;
                        const bool in = n >= fm.lo && n < fm.hi;
                        if (PCM16) reinterpret_cast<short *>(dst)[i] = in ? __ldg(static_cast<const short *>(args.audio) + fm.base + n) : (short)0;
                        else reinterpret_cast<float *>(dst)[i] = in ? __ldg(static_cast<const float *>(args.audio) + fm.base + n) : 0.f;
                    }
                }
            }
            if (lane == 0) {
                SlotInfo si;
                si.base1 = m1.base; si.lo1 = m1.lo; si.hi1 = m1.hi; si.shared = shared ? 1 : 0; si.pad = 0;
                s_slot[stage] = si;
            }
            __syncwarp();
            if (lane == 0) {
                if (qmask) afs::mbar_expect_tx(&s_bar[kBarRingFull + stage], (uint32_t)__popc(qmask) * 1024 * kSampleBytes);
                else mbar_arrive(&s_bar[kBarRingFull + stage]);
            }
            __syncwarp();
            if (bulk_l) afs::bulk_g2s(slot + pc_l * kPieceBytes, src_l, 1024 * kSampleBytes, &s_bar[kBarRingFull + stage]);
            trace(0, T);
        }
    } else {
        // ================= M: the warp that issues tensor-core work (converged; one elected lane issues) =================
        const uint64_t f_hi_desc = tc::smem_desc_k_sw128(s_base + kOffF), f_lo_desc = tc::smem_desc_k_sw128(s_base + kOffF + kFPlane);
        constexpr uint32_t idesc1 = tc::idesc_bf16_f32(128, kN1), idesc2 = tc::idesc_bf16_f32(128, kN2);
        auto stage1 = [&](int T) {
            const int s = T & 1, use = T >> 1;
            mbar_wait_spin(&s_bar[kBarA1Full], (uint32_t)(T & 1));
            if (use > 0) mbar_wait_spin(&s_bar[kBarD1Free + s], (uint32_t)((use - 1) & 1));
            tc::fence_after_sync();
            trace(0, T);
            const uint32_t d1 = tm + kColD1 + 64 * s, a_hi = tm + kColA1, a_lo = a_hi + 32;
#pragma unroll
            for (int ks = 0; ks < 4; ks++) tc::mma_ts_elect(d1, a_hi + 8 * ks, f_hi_desc + 2 * ks, idesc1, ks != 0);     // x_hi F_hi
#pragma unroll
            for (int ks = 0; ks < 4; ks++) tc::mma_ts_elect(d1, a_lo + 8 * ks, f_hi_desc + 2 * ks, idesc1, 1u);          // x_lo F_hi
#pragma unroll
            for (int ks = 0; ks < 4; ks++) tc::mma_ts_elect(d1, a_hi + 8 * ks, f_lo_desc + 2 * ks, idesc1, 1u);          // x_hi F_lo
            tc::mma_commit_elect(&s_bar[kBarM1Done + s]);
            trace(1, T);
        };
        auto stage2 = [&](int gi) {
            const int b = gi & 1, u = gi >> 1;
            mbar_wait_spin(&s_bar[kBarYFull + b], (uint32_t)(u & 1));
            if (gi > 0) mbar_wait_spin(&s_bar[kBarD2Free], (uint32_t)((gi - 1) & 1));
            tc::fence_after_sync();
            trace(2, gi);
            const uint64_t y_hi_desc = tc::smem_desc_k_sw128(s_base + kOffY + (uint32_t)b * kYBuf);
            const uint64_t y_lo_desc = tc::smem_desc_k_sw128(s_base + kOffY + (uint32_t)b * kYBuf + kYPlane);
#pragma unroll
            for (int term = 0; term < 3; term++) {
                const uint32_t a_tm = tm + (term == 1 ? kColGlo : kColGhi);           // hi*hi, lo*hi, hi*lo
                const uint64_t bd = term == 2 ? y_lo_desc : y_hi_desc;
#pragma unroll
                for (int ks = 0; ks < 8; ks++)
                    tc::mma_ts_elect(tm + kColD2, a_tm + 8 * ks, bd + (uint64_t)(((ks >> 2) * kYBlk + (ks & 3) * 32) >> 4), idesc2, (term | ks) != 0);
            }
            tc::mma_commit_elect(&s_bar[kBarM2Done + b]);
            trace(3, gi);
        };
        // issue order: the single-buffered A1 is refilled while the previous group's stage 2 keeps the tensor pipe busy
        for (int gi = 0; gi < n_local; gi++) {
            stage1(2 * gi);
            if (gi > 0) stage2(gi - 1);
            stage1(2 * gi + 1);
        }
        stage2(n_local - 1);
        // drain: the last commit covers every MMA issued before it
        mbar_wait_spin(&s_bar[kBarM2Done + ((n_local - 1) & 1)], (uint32_t)(((n_local - 1) >> 1) & 1));
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tm, 512);
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
    uint32_t *ring = nullptr;         // power spectrum ring: [2 planes][ring_tiles * 64 frames][kPlaneWords]
    int *flags = nullptr;             // [2][ring_tiles]: prod | cons
    int ring_tiles = 48;              // 48 x 590 KB = 28 MB: stays in L2
    int n_fb = 16;                    // CTAs that run the filterbank role
    // the ring and its counters belong to the plan: launches of one plan from different streams or threads are put in
    // order on the device (each waits for the previous one's completion event) instead of sharing them concurrently
    std::mutex mu;
    cudaEvent_t done = nullptr;
};

void chroma_tc_destroy(afs_chroma_tc *tc)
{
    if (!tc) return;
    cudaFree(tc->f_img); cudaFree(tc->w_img); cudaFree(tc->g_img); cudaFree(tc->tw); cudaFree(tc->hann); cudaFree(tc->ring); cudaFree(tc->flags);
    if (tc->done) cudaEventDestroy(tc->done);
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
    // ---- G as the TMEM A operand: row 2 k2 + part, K = 2 n2 + (re | im of Y'); lane = row, word c = K elements 2c, 2c + 1 ----
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
            w_img[(size_t)kb * kFbWBlk / 2 + tc::sw128_offset(12 + c, kk) / 2] = lo;
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
    cudaError_t e = cudaFuncSetAttribute(chroma_tc_spectrum_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemA);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(chroma_tc_spectrum_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemA);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(chroma_tc_spectrum_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemA);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(chroma_tc_spectrum_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemA);
    if (e != cudaSuccess) {
        chroma_tc_destroy(tcp);
        return afs::fail(AFS_ERR_CUDA, "chroma_tc_create: %s", cudaGetErrorString(e));
    }
    if (const char *c = getenv("AFS_CHROMA_TC_RING")) {
        const int v = atoi(c);
        if (v >= 2) tcp->ring_tiles = v;
    }
    if (const char *c = getenv("AFS_CHROMA_TC_FB")) {
        const int v = atoi(c);
        if (v >= 1 && v < afs::sm_count()) tcp->n_fb = v;
    }
    cudaError_t e2 = cudaMalloc(&tcp->ring, sizeof(uint32_t) * 2 * (size_t)tcp->ring_tiles * kFbTile * kPlaneWords);
    if (e2 == cudaSuccess) e2 = cudaMalloc(&tcp->flags, sizeof(int) * 2 * tcp->ring_tiles);
    if (e2 == cudaSuccess) e2 = cudaEventCreateWithFlags(&tcp->done, cudaEventDisableTiming);
    if (e2 != cudaSuccess) {
        chroma_tc_destroy(tcp);
        return afs::fail(AFS_ERR_CUDA, "chroma_tc_create: %s", cudaGetErrorString(e2));
    }
    *out = tcp;
    return AFS_OK;
}

int chroma_tc_run(afs_chroma_tc *tcp, const ChromaBatch &bt, cudaStream_t st)
{
    // One launch for the whole batch: CTAs [0, n_fb) run the filterbank role, the others the spectrum role; both are
    // persistent and every CTA is resident (grid <= number of SMs, one CTA per SM), which the hand-off through the
    // power-spectrum ring relies on.
    if (bt.total_frames <= 0) return AFS_OK;
    std::lock_guard<std::mutex> lock(tcp->mu);
    AFS_CUDA(cudaStreamWaitEvent(st, tcp->done, 0));      // the previous launch of this plan has released the ring
    const int n_sm = afs::sm_count();
    const int64_t groups = (bt.total_frames + 3) / 4, tiles = (bt.total_frames + kFbTile - 1) / kFbTile;
    SpectrumArgs sa;
    sa.audio = bt.audio;
    sa.sample_off = bt.sample_off;
    sa.frame_off = bt.frame_off;
    sa.n_tracks = bt.n_tracks;
    sa.frame_begin = 0;
    sa.frame_end = bt.total_frames;
    sa.hop = bt.hop;
    sa.center_pad = bt.center_pad;
    sa.f_img = tcp->f_img;
    sa.g_img = tcp->g_img;
    sa.tw = tcp->tw;
    sa.hann = tcp->hann;
    sa.ring_tiles = tcp->ring_tiles;
    sa.p_hi = tcp->ring;
    sa.p_lo = tcp->ring + (size_t)tcp->ring_tiles * kFbTile * kPlaneWords;
    sa.prod = tcp->flags;
    sa.cons = tcp->flags + tcp->ring_tiles;
    sa.n_fb = (int)(tiles < tcp->n_fb ? tiles : tcp->n_fb);
    sa.n_spec = (int)(groups < n_sm - sa.n_fb ? groups : n_sm - sa.n_fb);
    sa.w_img = tcp->w_img;
    sa.out_off = bt.out_off;
    sa.normalize = bt.normalize;
    sa.out_f64 = bt.out_f64;
    sa.out = bt.out;
    sa.trace = nullptr;
    AFS_CUDA(cudaMemsetAsync(tcp->flags, 0, sizeof(int) * 2 * tcp->ring_tiles, st));
    // AFS_CHROMA_TC_TRACE=<file>: clock64 stamps of every warp's pipeline events (debugging aid)
    const char *trace_path = getenv("AFS_CHROMA_TC_TRACE");
    const size_t trace_n = (size_t)n_sm * (kThreadsA / 32) * kTraceLen;
    if (trace_path) {
        AFS_CUDA(cudaMalloc(&sa.trace, trace_n * sizeof(long long)));
        AFS_CUDA(cudaMemsetAsync(sa.trace, 0, trace_n * sizeof(long long), st));
    }
    const unsigned blocks = (unsigned)(sa.n_fb + sa.n_spec);
    if (sa.trace) {
        if (bt.pcm16) chroma_tc_spectrum_kernel<true, true><<<blocks, kThreadsA, kSmemA, st>>>(sa);
        else chroma_tc_spectrum_kernel<false, true><<<blocks, kThreadsA, kSmemA, st>>>(sa);
    } else if (bt.pcm16) chroma_tc_spectrum_kernel<true, false><<<blocks, kThreadsA, kSmemA, st>>>(sa);
    else chroma_tc_spectrum_kernel<false, false><<<blocks, kThreadsA, kSmemA, st>>>(sa);
    afs::count_launch();
    AFS_CUDA(cudaGetLastError());
    AFS_CUDA(cudaEventRecord(tcp->done, st));
    if (trace_path) {
        std::vector<long long> host(trace_n);
        AFS_CUDA(cudaMemcpyAsync(host.data(), sa.trace, trace_n * sizeof(long long), cudaMemcpyDeviceToHost, st));
        AFS_CUDA(cudaStreamSynchronize(st));
        cudaFree(sa.trace);
        if (FILE *fp = fopen(trace_path, "wb")) {
            const int hdr[4] = {n_sm, kThreadsA / 32, kTraceLen, sa.n_spec};
            fwrite(hdr, sizeof(int), 4, fp);
            fwrite(host.data(), sizeof(long long), trace_n, fp);
            fclose(fp);
        }
    }
    return AFS_OK;
}

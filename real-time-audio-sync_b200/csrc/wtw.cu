// K6 — windowed time warping (Macrae & Dixon 2010) for many concurrent streams.
//
// Replaces the window loop of wtw.WTW.insert (reference wtw.py:94-128) with its
// helpers get_cost_matrix (:162-171, cosine with explicit norms), run_dtw (:173-217,
// weights 1,1,1; candidates (i-1,j), (i,j-1), (i-1,j-1) compared with strict <) and
// find_path (:219-240), fed one live chroma column per stream per step.
//
// One CTA per stream.  The W x W cost matrix is built in shared memory by all threads,
// accumulated in place along anti-diagonals (one barrier per diagonal), and thread 0
// backtracks and stitches the sub-path (wtw.py:107-128).  Only the last W live columns
// are kept (ring), since a window is consumed as soon as W columns are pending.
#include <math_constants.h>

#include <vector>

#include "afs_common.cuh"

namespace {

constexpr int kF = 12;
constexpr int kWtwThreads = 128;

struct WtwStream {
    int64_t ref_off;    // elements into d_ref, (12, M) feature-major
    int64_t path_off;   // pairs
    int32_t M;
    int32_t path_cap;
};

struct WtwScalars { int32_t chroma_ptr, live_ptr, ref_ptr, path_len; };

struct WtwArgs {
    const double *ref;
    const WtwStream *streams;
    WtwScalars *scal;
    double *ring;         // n x W x 12
    int32_t *path;
    int32_t *ptrs;        // (n,3) mirror
    int32_t *path_len;    // (n) mirror
    int n_streams, W, h;
    const double *cols;
    int n_frames;
    const uint8_t *active;
    int32_t *out_status;
    int latch_stop;       // 1: frames that follow a STOP inside this launch are not consumed (one insert() call of many frames)
};

// np.dot on two strided column views: OpenBLAS generic ddot
__device__ __forceinline__ double dot_strided12(const double *x, const double *y)
{
    double t1 = 0.0, t2 = 0.0;
#pragma unroll
    for (int i = 0; i < kF; i += 4) {
        const double m3 = __dmul_rn(y[i + 2], x[i + 2]);
        const double m4 = __dmul_rn(y[i + 3], x[i + 3]);
        t1 = __dadd_rn(t1, __fma_rn(y[i], x[i], m3));
        t2 = __dadd_rn(t2, __fma_rn(y[i + 1], x[i + 1], m4));
    }
    return __dadd_rn(t1, t2);
}

// np.linalg.norm of a column: ravel() copy, contiguous ddot (n < 32: fused tail loop), sqrt
__device__ __forceinline__ double norm12(const double *x)
{
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < kF; i++) s = __fma_rn(x[i], x[i], s);
    return __dsqrt_rn(s);
}

// K1's output (stream, feature, frame) -> K6's input (frame, stream, feature)
__global__ void wtw_transpose_cols(const double *__restrict__ src, double *__restrict__ dst, int n_streams, int n_frames)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)n_streams * n_frames * kF;
    if (i >= total) return;
    const int k = (int)(i % kF);
    const int64_t fs = i / kF;
    const int s = (int)(fs % n_streams), f = (int)(fs / n_streams);
    dst[i] = src[((int64_t)s * kF + k) * n_frames + f];
}

__global__ void __launch_bounds__(kWtwThreads) wtw_push_kernel(const WtwArgs a)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int W = a.W;
    double *D = reinterpret_cast<double *>(s_raw);        // W x W (cost, then accumulated cost)
    double *xs = D + (size_t)W * W;                       // W x 12 live window
    double *ys = xs + (size_t)W * kF;                     // W x 12 ref window
    double *nx = ys + (size_t)W * kF;                     // W
    double *ny = nx + W;                                  // W
    uint8_t *B = reinterpret_cast<uint8_t *>(ny + W);     // W x W back-pointers
    __shared__ WtwScalars S;
    __shared__ int s_do_window;

    const int s = blockIdx.x;
    const int t = threadIdx.x;
    const WtwStream sd = a.streams[s];
    const int M = sd.M, Ncap = 2 * sd.M;
    const double *ref = a.ref + sd.ref_off;
    double *ring = a.ring + (int64_t)s * W * kF;
    int2 *path = reinterpret_cast<int2 *>(a.path) + sd.path_off;
    if (t == 0) S = a.scal[s];
    __syncthreads();

    // wtw.py:96-97 returns from insert() at the first stopping frame and leaves the rest of the buffer unread: the frames
    // that follow a STOP inside this launch are not consumed (the caller pushes them again, as the reference re-reads them)
    bool stopped_in_launch = false;          // thread 0 only
    for (int f = 0; f < a.n_frames; f++) {
        const int64_t oidx = (int64_t)f * a.n_streams + s;
        const bool on = (a.active == nullptr) || (a.active[s] != 0);
        if (t == 0) {
            int status = AFS_STEP_NONE;
            s_do_window = 0;
            if (on && stopped_in_launch && a.latch_stop) {
                status = AFS_STEP_STOP;
            } else if (on) {
                if (S.chroma_ptr >= Ncap) {
                    status = AFS_STEP_FULL;            // the reference would raise IndexError at wtw.py:92
                } else {
                    // wtw.py:92-93
                    const double *col = a.cols + oidx * kF;
                    double *dst = ring + (int64_t)(S.chroma_ptr % W) * kF;
                    for (int k = 0; k < kF; k++) dst[k] = col[k];
                    S.chroma_ptr += 1;
                    if (S.ref_ptr >= (M - 1 - W) || S.live_ptr >= (Ncap - 1 - W)) { status = AFS_STEP_STOP; stopped_in_launch = true; }   // wtw.py:96-97
                    else if (S.chroma_ptr - S.live_ptr >= W) s_do_window = 1;                             // wtw.py:100
                }
            }
            if (a.out_status) a.out_status[oidx] = status;
        }
        __syncthreads();
        const int do_window = s_do_window;
        __syncthreads();                  // everyone has read the flag before thread 0 may rewrite it
        if (!do_window) continue;         // uniform

        const int lp = S.live_ptr, rp = S.ref_ptr;
        // ---- load windows: live columns lp .. lp+W-1 (ring), ref columns rp .. rp+W-1 ----
        for (int q = t; q < W * kF; q += kWtwThreads) {
            const int i = q / kF, k = q - i * kF;
            xs[q] = ring[(int64_t)((lp + i) % W) * kF + k];
            ys[q] = __ldg(ref + (int64_t)k * M + rp + i);
        }
        __syncthreads();
        for (int i = t; i < 2 * W; i += kWtwThreads) {
            if (i < W) nx[i] = norm12(xs + i * kF);
            else ny[i - W] = norm12(ys + (i - W) * kF);
        }
        __syncthreads();
        // ---- get_cost_matrix, wtw.py:162-171 ----
        for (int q = t; q < W * W; q += kWtwThreads) {
            const int i = q / W, j = q - i * W;
            const double num = dot_strided12(xs + i * kF, ys + j * kF);
            D[q] = __dsub_rn(1.0, __ddiv_rn(num, __dmul_rn(nx[i], ny[j])));
        }
        __syncthreads();
        // ---- run_dtw, wtw.py:173-217: in place along anti-diagonals ----
        for (int d = 0; d <= 2 * W - 2; d++) {
            const int i_lo = max(0, d - W + 1), i_hi = min(d, W - 1);
            for (int i = i_lo + t; i <= i_hi; i += kWtwThreads) {
                const int j = d - i;
                const double c = D[i * W + j];
                if (i == 0 && j == 0) { B[0] = 0; }
                else if (j == 0) { D[i * W] = __dadd_rn(D[(i - 1) * W], c); B[i * W] = 3; }      // wtw.py:189-193
                else if (i == 0) { D[j] = __dadd_rn(D[j - 1], c); B[j] = 1; }                    // wtw.py:196-200
                else {
                    double m = D[(i - 1) * W + j];        // (i-1, j): code 3
                    uint8_t code = 3;
                    double v = D[i * W + j - 1];          // (i, j-1): code 1
                    if (v < m) { m = v; code = 1; }
                    v = D[(i - 1) * W + j - 1];           // (i-1, j-1): code 2
                    if (v < m) { m = v; code = 2; }
                    D[i * W + j] = __dadd_rn(m, c);
                    B[i * W + j] = code;
                }
            }
            __syncthreads();
        }
        // ---- find_path (wtw.py:219-240) + stitch (wtw.py:107-128) ----
        if (t == 0) {
            // walk back from (W-1, W-1); remember, for every live index l, nothing more than
            // what the forward scan needs: the forward scan appends points while l <= h and
            // stops at the first l > h, whose predecessor becomes the new origin.
            // Backward walk stores the sub-path reversed in the (now free) xs/ys area.
            int32_t *sub = reinterpret_cast<int32_t *>(xs);      // 2W*12*8 B >= 2*(2W)*4 B
            int n = 0, ci = W - 1, cj = W - 1;
            sub[0] = ci; sub[1] = cj; n = 1;
            while (ci != 0 || cj != 0) {
                const uint8_t p = B[ci * W + cj];
                if (p == 1) cj -= 1;
                else if (p == 2) { ci -= 1; cj -= 1; }
                else ci -= 1;
                sub[2 * n] = ci; sub[2 * n + 1] = cj; n++;
            }
            bool change = false;
            int pl = 0, pr = 0;
            for (int q = n - 1; q >= 0; q--) {
                const int l = sub[2 * q], r = sub[2 * q + 1];
                if (l <= a.h) {
                    if (S.path_len < sd.path_cap) path[S.path_len] = make_int2(l + lp, r + rp);
                    S.path_len += 1;
                } else {
                    change = true;
                    pl = sub[2 * (q + 1)];
                    pr = sub[2 * (q + 1) + 1];
                    break;
                }
            }
            if (change) { S.live_ptr = lp + pl; S.ref_ptr = rp + pr; }
            else { S.live_ptr = lp + a.h; S.ref_ptr = rp + a.h; }      // wtw.py:126-128
        }
        __syncthreads();
    }
    if (t == 0) {
        a.scal[s] = S;
        a.ptrs[3 * s] = S.chroma_ptr;
        a.ptrs[3 * s + 1] = S.live_ptr;
        a.ptrs[3 * s + 2] = S.ref_ptr;
        a.path_len[s] = S.path_len;
    }
}

__global__ void wtw_reset_kernel(const WtwArgs a)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nring = (int64_t)a.n_streams * a.W * kF;
    if (i < nring) a.ring[i] = 0.0;                       // wtw.py:56 chroma_live = zeros
    if (i < a.n_streams) {
        a.scal[i] = WtwScalars{0, 0, 0, 0};
        a.ptrs[3 * i] = a.ptrs[3 * i + 1] = a.ptrs[3 * i + 2] = 0;
        a.path_len[i] = 0;
    }
}

}  // namespace

struct afs_wtw {
    WtwArgs args;
    std::vector<WtwStream> streams;
    WtwStream *d_streams = nullptr;
    size_t state_bytes = 0, off_scal = 0, off_ring = 0, off_path = 0, off_ptrs = 0, off_plen = 0;
    int64_t total_path = 0;
    size_t smem_bytes = 0;
    bool bound = false;
};

extern "C" {

int afs_wtw_create(afs_wtw **out, int n_streams, const double *d_ref, const int64_t *h_ref_len, const int64_t *h_ref_off,
                   int n_features, int W, int h)
{
    if (!out || n_streams <= 0 || !d_ref || !h_ref_len || !h_ref_off)
        return afs::fail(AFS_ERR_INVALID, "afs_wtw_create: null argument or n_streams <= 0");
    if (n_features != kF) return afs::fail(AFS_ERR_UNSUPPORTED, "afs_wtw: n_features must be 12 (got %d)", n_features);
    if (W < 2 || h < 1) return afs::fail(AFS_ERR_INVALID, "afs_wtw: need W >= 2 and h >= 1 (got %d, %d)", W, h);
    const size_t smem = sizeof(double) * ((size_t)W * W + 2 * (size_t)W * kF + 2 * (size_t)W) + (size_t)W * W;
    if (smem > 220 * 1024) return afs::fail(AFS_ERR_UNSUPPORTED, "afs_wtw: window of %d frames needs %zu B of shared memory (max 220 KiB)", W, smem);
    afs_wtw *hd = new afs_wtw();
    memset(&hd->args, 0, sizeof(hd->args));
    hd->streams.resize(n_streams);
    int64_t path_pairs = 0;
    for (int s = 0; s < n_streams; s++) {
        const int64_t M = h_ref_len[s];
        if (M <= 0 || M > (1 << 29)) { delete hd; return afs::fail(AFS_ERR_INVALID, "afs_wtw: stream %d has invalid reference length", s); }
        hd->streams[s].ref_off = h_ref_off[s];
        hd->streams[s].M = (int32_t)M;
        const int64_t windows = (2 * M) / h + 2;
        hd->streams[s].path_cap = (int32_t)(windows * (h + W + 1));
        hd->streams[s].path_off = path_pairs;
        path_pairs += hd->streams[s].path_cap;
    }
    hd->total_path = path_pairs;
    WtwArgs &a = hd->args;
    a.ref = d_ref;
    a.n_streams = n_streams;
    a.W = W;
    a.h = h;
    hd->smem_bytes = smem;
    size_t off = 0;
    hd->off_scal = off; off = afs::align_up(off + sizeof(WtwScalars) * n_streams, 256);
    hd->off_ring = off; off = afs::align_up(off + sizeof(double) * (size_t)n_streams * W * kF, 256);
    hd->off_path = off; off = afs::align_up(off + sizeof(int32_t) * 2 * (size_t)path_pairs, 256);
    hd->off_ptrs = off; off = afs::align_up(off + sizeof(int32_t) * 3 * n_streams, 256);
    hd->off_plen = off; off = afs::align_up(off + sizeof(int32_t) * n_streams, 256);
    hd->state_bytes = off;
    cudaError_t e = cudaMalloc(&hd->d_streams, sizeof(WtwStream) * n_streams);
    if (e == cudaSuccess) e = cudaMemcpy(hd->d_streams, hd->streams.data(), sizeof(WtwStream) * n_streams, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && smem > 48 * 1024)
        e = cudaFuncSetAttribute(wtw_push_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        cudaFree(hd->d_streams);
        delete hd;
        return afs::fail(AFS_ERR_CUDA, "afs_wtw_create: %s", cudaGetErrorString(e));
    }
    a.streams = hd->d_streams;
    *out = hd;
    return AFS_OK;
}

int afs_wtw_destroy(afs_wtw *h)
{
    if (!h) return AFS_OK;
    cudaFree(h->d_streams);
    delete h;
    return AFS_OK;
}

int afs_wtw_state_bytes(const afs_wtw *h, size_t *bytes)
{
    if (!h || !bytes) return afs::fail(AFS_ERR_INVALID, "afs_wtw_state_bytes: null argument");
    *bytes = h->state_bytes;
    return AFS_OK;
}

int afs_wtw_reset(afs_wtw *h, void *d_state, void *stream)
{
    if (!h || !d_state) return afs::fail(AFS_ERR_INVALID, "afs_wtw_reset: null argument");
    char *base = static_cast<char *>(d_state);
    WtwArgs &a = h->args;
    a.scal = reinterpret_cast<WtwScalars *>(base + h->off_scal);
    a.ring = reinterpret_cast<double *>(base + h->off_ring);
    a.path = reinterpret_cast<int32_t *>(base + h->off_path);
    a.ptrs = reinterpret_cast<int32_t *>(base + h->off_ptrs);
    a.path_len = reinterpret_cast<int32_t *>(base + h->off_plen);
    h->bound = true;
    const int64_t n = (int64_t)a.n_streams * a.W * kF;
    const int threads = 256;
    wtw_reset_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, static_cast<cudaStream_t>(stream)>>>(a);
    afs::count_launch();
    AFS_CUDA(cudaGetLastError());
    return AFS_OK;
}

static int wtw_push_impl(afs_wtw *h, const double *d_cols, int n_frames, const uint8_t *d_active, int32_t *d_status, int latch_stop,
                         void *stream)
{
    if (!h || !d_cols || n_frames <= 0) return afs::fail(AFS_ERR_INVALID, "afs_wtw_push: null argument or n_frames <= 0");
    if (!h->bound) return afs::fail(AFS_ERR_INVALID, "afs_wtw_push: call afs_wtw_reset first");
    WtwArgs a = h->args;
    a.cols = d_cols;
    a.n_frames = n_frames;
    a.active = d_active;
    a.out_status = d_status;
    a.latch_stop = latch_stop;
    wtw_push_kernel<<<a.n_streams, kWtwThreads, h->smem_bytes, static_cast<cudaStream_t>(stream)>>>(a);
    afs::count_launch();
    AFS_CUDA(cudaGetLastError());
    return AFS_OK;
}

int afs_wtw_push(afs_wtw *h, const double *d_cols, int n_frames, const uint8_t *d_active, int32_t *d_status, void *stream)
{
    // every column is an insert of its own (a STOP does not swallow the columns after it)
    return wtw_push_impl(h, d_cols, n_frames, d_active, d_status, 0, stream);
}

int afs_wtw_push_audio(afs_wtw *h, afs_chroma_plan *plan, const float *d_audio, const int64_t *h_sample_off, int n_frames,
                       double *d_scratch, const uint8_t *d_active, int32_t *d_status, int compute, void *stream)
{
    if (!h || !plan || !d_audio || !h_sample_off || !d_scratch || n_frames <= 0)
        return afs::fail(AFS_ERR_INVALID, "afs_wtw_push_audio: null argument or n_frames <= 0");
    if (!h->bound) return afs::fail(AFS_ERR_INVALID, "afs_wtw_push_audio: call afs_wtw_reset first");
    const int n = h->args.n_streams;
    std::vector<int64_t> out_off(n);
    for (int s = 0; s < n; s++) {
        const int64_t got = afs_chroma_num_frames(plan, h_sample_off[s + 1] - h_sample_off[s], 0);
        if (got != n_frames)
            return afs::fail(AFS_ERR_INVALID, "afs_wtw_push_audio: stream %d holds %lld frames, expected %d", s, (long long)got, n_frames);
        out_off[s] = (int64_t)s * n_frames;
    }
    // K1: frames without padding (wtw.py:81-90), float64 columns, track-major (12, n_frames) per stream
    double *d_chroma = d_scratch, *d_cols = d_scratch + (size_t)n * n_frames * kF;
    const int rc = afs_chroma_batch(plan, d_audio, h_sample_off, n, 0, 1, d_chroma, out_off.data(), AFS_F64, compute, stream);
    if (rc != AFS_OK) return rc;
    // (stream, feature, frame) -> (frame, stream, feature), the order K6 reads
    const int64_t total = (int64_t)n * n_frames * kF;
    wtw_transpose_cols<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_chroma, d_cols, n, n_frames);
    afs::count_launch();
    AFS_CUDA(cudaGetLastError());
    return wtw_push_impl(h, d_cols, n_frames, d_active, d_status, 1, stream);
}

int afs_wtw_path_layout(const afs_wtw *h, int stream_idx, int64_t *offset, int64_t *capacity)
{
    if (!h || stream_idx < -1 || stream_idx >= h->args.n_streams) return afs::fail(AFS_ERR_INVALID, "afs_wtw_path_layout: bad stream");
    if (stream_idx == -1) {
        if (offset) *offset = 0;
        if (capacity) *capacity = h->total_path;
        return AFS_OK;
    }
    if (offset) *offset = h->streams[stream_idx].path_off;
    if (capacity) *capacity = h->streams[stream_idx].path_cap;
    return AFS_OK;
}

int afs_wtw_path_ptr(const afs_wtw *h, const int32_t **d_path, const int32_t **d_path_len)
{
    if (!h || !h->bound) return afs::fail(AFS_ERR_INVALID, "afs_wtw_path_ptr: state not bound");
    if (d_path) *d_path = h->args.path;
    if (d_path_len) *d_path_len = h->args.path_len;
    return AFS_OK;
}

int afs_wtw_positions_ptr(const afs_wtw *h, const int32_t **d_ptrs)
{
    if (!h || !h->bound || !d_ptrs) return afs::fail(AFS_ERR_INVALID, "afs_wtw_positions_ptr: state not bound");
    *d_ptrs = h->args.ptrs;
    return AFS_OK;
}

}  // extern "C"

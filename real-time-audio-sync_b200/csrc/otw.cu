// K5 — online time warping (Dixon 2005) for thousands of concurrent streams.
//
// Replaces, per stream and per inserted live frame:
//   OnlineTimeWarping.insert      otw_eran.py:38-85   (eval_path_cost :215-239,
//                                 set_direction :153-188, best_point :192-211)
//   LiveNoteV2.insert             livenote_v2.py:43-104 (:165-236)
//   LiveNote.insert (v1)          livenote.py
//
// The reference keeps two dense (2N x N) float64 matrices per stream.  Every read
// the algorithm makes while the frontier is at (t, j) falls in row t, cols [j-c, j]
// or column j, rows [t-c, t], so the device state per stream is
//   Rw : ring of c+1 doubles = acc_cost[t, j-c .. j]   (slot = col mod (c+1))
//   Cl : ring of c+1 doubles = acc_cost[t-c .. t, j]   (slot = row mod (c+1))
//   LH : ring of c+1 live frames, feature-major [12][c+1] (slot = row mod (c+1))
// plus the scalars t, j, previous, run_count, direction.  Cells the reference never
// evaluated read as its fill value (1e10 for OTW, +inf for LiveNote); options that
// fall outside the matrix (x == 0 or y == 0) are excluded, i.e. +inf here.
//
// One warp advances one stream; a CTA advances a group of up to 7 streams in lock-step
// rounds (see otw_step_kernel).  A row (or column) sweep of <= c cells is a serial chain
// through acc[x, y-1] (resp. acc[x-1, y]); to stay bit-exact with the float64 reference
// the chain is NOT re-associated: the 32 lanes of the stream's warp compute each cell's
// cost and the two chain-independent candidates in parallel into shared memory, then ONE
// warp runs  v = min(v + cost_k, m_k)  for all streams of the group at once (one stream
// per lane), then every warp writes its line back.
#include <math_constants.h>

#include <limits>
#include <vector>

#include "afs_common.cuh"

namespace {

constexpr int kF = 12;
enum { DIR_BOTH = 0, DIR_ROW = 1, DIR_COL = 2 };

struct OtwStream {
    int64_t ref_off;     // elements into d_ref; reference is (12, N) feature-major
    int64_t path_off;    // pairs into the path area
    int32_t N;           // reference frames
    int32_t path_cap;
};

struct OtwScalars {      // 16 ints per stream
    int32_t t, j, previous, run_count, direction, first, status, path_len, last_x, last_y;
    int32_t pad[6];
};

struct OtwArgs {
    const double *ref;
    const OtwStream *streams;
    OtwScalars *scal;
    double *rw, *cl, *lh;
    int32_t *path;
    int32_t *tj;          // (n,2) mirror of (t, j) for the host
    int32_t *path_len;    // (n) mirror
    int n_streams, kind, c, max_run, metric, rs, cpad, pts;
    int group, cstride;   // streams per CTA; doubles of chain scratch per stream (2*cpad + 2: bank-skewed)
    double fill;
    // per launch
    const double *frames;
    int n_frames;
    const uint8_t *active;
    int32_t *out_status, *out_npoints, *out_points;
};

// otw_eran.py:220 / livenote_v2.py:170 — np.dot of two strided column views executes
// OpenBLAS' generic ddot: blocks of four, two running sums.
__device__ __forceinline__ double cost_cosine(const double *x, const double *y)
{
    double t1 = 0.0, t2 = 0.0;
#pragma unroll
    for (int i = 0; i < kF; i += 4) {
        const double m3 = __dmul_rn(y[i + 2], x[i + 2]);
        const double m4 = __dmul_rn(y[i + 3], x[i + 3]);
        t1 = __dadd_rn(t1, __fma_rn(y[i], x[i], m3));
        t2 = __dadd_rn(t2, __fma_rn(y[i + 1], x[i + 1], m4));
    }
    return __dsub_rn(1.0, __dadd_rn(t1, t2));
}

// livenote_v2.py:168 — sqrt(np.sum((live - ref)**2)): numpy's 8-lane pairwise sum.
__device__ __forceinline__ double cost_euclid(const double *x, const double *y)
{
    double d[kF];
#pragma unroll
    for (int i = 0; i < kF; i++) {
        const double e = __dsub_rn(x[i], y[i]);
        d[i] = __dmul_rn(e, e);
    }
    double r = __dadd_rn(__dadd_rn(__dadd_rn(d[0], d[1]), __dadd_rn(d[2], d[3])),
                         __dadd_rn(__dadd_rn(d[4], d[5]), __dadd_rn(d[6], d[7])));
#pragma unroll
    for (int i = 8; i < kF; i++) r = __dadd_rn(r, d[i]);
    return __dsqrt_rn(r);
}

__device__ __forceinline__ double cell_cost(int metric, const double *x, const double *y)
{
    return metric == 0 ? cost_cosine(x, y) : cost_euclid(x, y);
}

// ---------------------------------------------------------------------------------------------------
// A CTA advances a GROUP of up to 7 streams (one warp each) in lock-step rounds.  Round 0 is the row
// sweep of the frame, every further round one trip of the reference's column loop (otw_eran.py:64-85).
// In each round every warp does the PARALLEL part of its stream's sweep (costs + the two
// chain-independent candidates into shared memory), then warp 0 runs ALL the group's serial chains
//   v = min(v + cost_k, m_k)
// at once, one stream per lane, then every warp writes its line back and takes its own decisions.
// (With one warp per stream doing its own chain, 54 % of all warp-instructions ran with one active lane.)
constexpr int kOtwMaxGroup = 7;        // 7 warps x 4 CTAs/SM = 28 streams per SM: 4096 streams fit one wave on 148 SMs

// Parallel part of one sweep over cells k = k1 .. hi of the new line.
//   ROW sweep: new row t over ref columns k  (`fix` = live frame, sequence = ref columns)
//   COL sweep: new column j over live rows k (`fix` = ref column,  sequence = live history)
// `fix` points to shared memory (12 doubles).  `ring` holds the previous line in the same index space.
__device__ __forceinline__ void sweep_parallel(const OtwArgs &a, const bool is_row, const int k1, const int hi, const double *fix,
                                               const double *ref, const int N, const double *lh, const double *ring, double *sc,
                                               double *sm, const int lane)
{
    const int n = hi - k1 + 1;
    const int rs = a.rs;
    const int base_slot = k1 % rs;
    // Software pipelining through L2 (registers are too scarce to keep several iterations' loads in
    // flight): lanes 0..23 prefetch, for each of the 12 feature rows, the two 128-byte lines that the
    // iteration kPf steps ahead will read.  pf_row = feature, pf_half = which line.
    constexpr int kPf = 4;
    const int pf_row = lane % kF, pf_half = lane / kF;
    auto prefetch_iter = [&](int first_cell) {         // first_cell: index (k) of the iteration's lane 0
        if (lane < 2 * kF) {
            const int kk = first_cell + 16 * pf_half;
            if (kk <= hi) {
                const double *p = is_row ? ref + (int64_t)pf_row * N + kk : lh + (int64_t)pf_row * rs + (kk % rs);
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
            }
        }
    };
#pragma unroll
    for (int q = 0; q < kPf; q++) prefetch_iter(k1 + 32 * q);
    for (int idx = lane; idx < n; idx += 32) {
        const int k = k1 + idx;
        prefetch_iter(k - lane + 32 * kPf);
        double seq[kF];
        if (is_row) {
#pragma unroll
            for (int f = 0; f < kF; f++) seq[f] = __ldg(ref + (int64_t)f * N + k);
        } else {
            // live history is feature-major [12][rs] like the reference: consecutive lanes (consecutive rows k)
            // read consecutive doubles -> every sector is used in full
            const double *p = lh + (k % rs);
#pragma unroll
            for (int f = 0; f < kF; f++) seq[f] = p[(int64_t)f * rs];
        }
        // live is always the first np.dot operand (otw_eran.py:220)
        const double cst = is_row ? cell_cost(a.metric, fix, seq) : cell_cost(a.metric, seq, fix);
        int slot = base_slot + idx;
        if (slot >= rs) slot -= rs;
        const int slot1 = (slot == 0) ? rs - 1 : slot - 1;
        const double pk = ring[slot];                                    // same index, previous line
        const double pk1 = (k > 0) ? ring[slot1] : CUDART_INF;           // index-1, previous line (diag)
        const double o1 = __dadd_rn(pk, cst);
        const double o2 = __fma_rn(2.0, cst, pk1);
        sc[idx] = cst;
        sm[idx] = (o2 < o1) ? o2 : o1;
    }
    // neutral cells up to a multiple of four (the chain is unrolled by four): v + 0 == v, min(v, +inf) == v
    const int n4 = (n + 3) & ~3;
    if (lane < n4 - n) {
        sc[n + lane] = 0.0;
        sm[n + lane] = CUDART_INF;
    }
}

// Write the finished line back (sc now holds the accumulated costs) and return the last cell's value.
__device__ __forceinline__ double sweep_finish(const OtwArgs &a, const int k1, const int hi, double *ring, const double *sc, const int lane)
{
    const int n = hi - k1 + 1;
    const int rs = a.rs;
    const int base_slot = k1 % rs;
    for (int idx = lane; idx < n; idx += 32) {
        int slot = base_slot + idx;
        if (slot >= rs) slot -= rs;
        ring[slot] = sc[idx];
    }
    if (lane == 0 && hi - a.c >= 0) ring[(hi - a.c) % rs] = a.fill;   // index hi-c was not evaluated in this line
    const double last = sc[n - 1];
    __syncwarp();
    return last;
}

// The serial chains of the whole group, one stream per lane (executed by warp 0).
// To stay bit-exact with the float64 reference the chain is NOT re-associated.
__device__ __forceinline__ void chain_exec(double *scratch, const int cstride, const int cpad, const int *s_n, const double *s_v0,
                                           const int G, const int lane)
{
    const int n = (lane < G) ? s_n[lane] : 0;
    const int n4 = (n + 3) & ~3;
    int nmax = n4;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, off));
    double v = (lane < G) ? s_v0[lane] : 0.0;
    double *c = scratch + (size_t)(lane < G ? lane : 0) * cstride;
    const double *m = c + cpad;
    for (int idx = 0; idx < nmax; idx += 4) {
        if (idx < n4) {
            const double2 c01 = *reinterpret_cast<const double2 *>(c + idx), c23 = *reinterpret_cast<const double2 *>(c + idx + 2);
            const double2 m01 = *reinterpret_cast<const double2 *>(m + idx), m23 = *reinterpret_cast<const double2 *>(m + idx + 2);
            double2 r01, r23;
            double x;
            x = __dadd_rn(v, c01.x); v = (x < m01.x) ? x : m01.x; r01.x = v;
            x = __dadd_rn(v, c01.y); v = (x < m01.y) ? x : m01.y; r01.y = v;
            x = __dadd_rn(v, c23.x); v = (x < m23.x) ? x : m23.x; r23.x = v;
            x = __dadd_rn(v, c23.y); v = (x < m23.y) ? x : m23.y; r23.y = v;
            *reinterpret_cast<double2 *>(c + idx) = r01;
            *reinterpret_cast<double2 *>(c + idx + 2) = r23;
        }
    }
}

// first-minimum argmin of ring over indices [k1, hi] (np.argmin, otw_eran.py:197,204)
__device__ __forceinline__ void otw_argmin(const double *ring, const int rs, const int k1, const int hi, const int lane, double &bv,
                           int &bk)
{
    double v = CUDART_INF;
    int kk = 0x7fffffff;
    bool have = false;
    // four independent loads in flight per trip (the rings are read from L2), compared in ascending k
    for (int k = k1 + lane; k <= hi; k += 128) {
        double x[4];
#pragma unroll
        for (int q = 0; q < 4; q++) x[q] = (k + 32 * q <= hi) ? ring[(k + 32 * q) % rs] : CUDART_INF;
#pragma unroll
        for (int q = 0; q < 4; q++)
            if (k + 32 * q <= hi && (!have || x[q] < v)) { v = x[q]; kk = k + 32 * q; have = true; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, v, off);
        const int ok = __shfl_xor_sync(0xffffffffu, kk, off);
        const bool ohave = __shfl_xor_sync(0xffffffffu, (int)have, off) != 0;
        const bool take = ohave && (!have || ov < v || (ov == v && ok < kk));
        if (take) { v = ov; kk = ok; have = true; }
    }
    bv = v;
    bk = kk;
}

__global__ void __launch_bounds__(kOtwMaxGroup * 32, 4) otw_step_kernel(const OtwArgs a)
{
    extern __shared__ __align__(16) double s_scratch[];      // [G][cstride] chain scratch, then [G][12] fixed vectors
    __shared__ int s_n[8];
    __shared__ double s_v0[8];
    const int G = blockDim.x >> 5;
    const int lane = threadIdx.x & 31;
    const int w = threadIdx.x >> 5;
    const bool valid = blockIdx.x * G + w < a.n_streams;
    const int s = valid ? blockIdx.x * G + w : 0;
    double *sc = s_scratch + (size_t)w * a.cstride;
    double *sm = sc + a.cpad;
    double *fix = s_scratch + (size_t)G * a.cstride + w * kF;

    const OtwStream sd = a.streams[s];
    const int N = sd.N;
    const int Lcap = 2 * N;
    const double *ref = a.ref + sd.ref_off;
    double *rw = a.rw + (int64_t)s * a.rs;
    double *cl = a.cl + (int64_t)s * a.rs;
    double *lh = a.lh + (int64_t)s * a.rs * kF;
    int2 *path = reinterpret_cast<int2 *>(a.path) + sd.path_off;
    OtwScalars S = a.scal[s];
    const int c = a.c, rs = a.rs;

    for (int f = 0; f < a.n_frames; f++) {
        const int64_t oidx = (int64_t)f * a.n_streams + s;
        int status = AFS_STEP_NONE, npts = 0;
        const bool on = valid && ((a.active == nullptr) || (a.active[s] != 0));
        // ---------------- round 0: the row sweep ----------------
        bool do_row = false;
        if (on && S.status == AFS_STEP_STOP) {
            status = AFS_STEP_STOP;            // the reference object is finished; stay finished
        } else if (on) {
            if (lane < kF) fix[lane] = __ldg(a.frames + oidx * kF + lane);
            __syncwarp();
            if (S.first) {
                // otw_eran.py:41-45: store column 0 and evaluate (0,0); no path point
                S.first = 0;
                if (lane < kF) lh[(int64_t)lane * rs] = fix[lane];
                double r0[kF];
#pragma unroll
                for (int k = 0; k < kF; k++) r0[k] = __ldg(ref + (int64_t)k * N);
                const double cst = cell_cost(a.metric, fix, r0);
                if (lane == 0) { rw[0] = cst; cl[0] = cst; }
                __syncwarp();
            } else {
                S.t += 1;
                if (S.t >= Lcap) {
                    status = AFS_STEP_FULL;    // otw_eran.py:53-55: live buffer exhausted, returns None forever
                                               // (t keeps counting, exactly like the reference's self.t)
                } else {
                    if (lane < kF) lh[(int64_t)lane * rs + (S.t % rs)] = fix[lane];
                    __syncwarp();
                    do_row = true;
                }
            }
        }
        int k1 = 0;
        if (do_row) {
            // ROW: eval(t, k) for k in [max(0, j-c+1), j]   otw_eran.py:58-62
            k1 = max(0, S.j - c + 1);
            sweep_parallel(a, true, k1, S.j, fix, ref, N, lh, rw, sc, sm, lane);
        }
        if (lane == 0) {
            s_n[w] = do_row ? S.j - k1 + 1 : 0;
            // neighbour before the first evaluated cell: outside the matrix (excluded) when k1 == 0, otherwise a
            // never-evaluated cell of the new line (reads the fill value)
            s_v0[w] = (k1 == 0) ? CUDART_INF : a.fill;
        }
        __syncthreads();
        if (w == 0) chain_exec(s_scratch, a.cstride, a.cpad, s_n, s_v0, G, lane);
        __syncthreads();
        bool looping = false;
        if (do_row) {
            const double last = sweep_finish(a, k1, S.j, rw, sc, lane);
            if (lane == 0) cl[S.t % rs] = last;        // acc[t, j] joins column j
            __syncwarp();
            looping = true;
        }
        // ---------------- further rounds: the column loop, otw_eran.py:64-85 ----------------
        while (__syncthreads_or(looping ? 1 : 0)) {
            bool do_col = false;
            if (looping && S.direction != DIR_ROW) {
                S.j += 1;
                if (S.j >= N) {                                  // otw_eran.py:69-71
                    status = AFS_STEP_STOP;
                    S.status = AFS_STEP_STOP;
                    looping = false;
                } else {
                    if (lane < kF) fix[lane] = __ldg(ref + (int64_t)lane * N + S.j);
                    __syncwarp();
                    // COLUMN: eval(k, j) for k in [max(0, t-c+1), t]   otw_eran.py:73-77
                    k1 = max(0, S.t - c + 1);
                    sweep_parallel(a, false, k1, S.t, fix, ref, N, lh, cl, sc, sm, lane);
                    do_col = true;
                }
            }
            if (lane == 0) {
                s_n[w] = do_col ? S.t - k1 + 1 : 0;
                s_v0[w] = (k1 == 0) ? CUDART_INF : a.fill;
            }
            __syncthreads();
            if (w == 0) chain_exec(s_scratch, a.cstride, a.cpad, s_n, s_v0, G, lane);
            __syncthreads();
            if (looping) {
                if (do_col) {
                    const double last = sweep_finish(a, k1, S.t, cl, sc, lane);
                    if (lane == 0) rw[S.j % rs] = last;    // acc[t, j] joins row t
                    __syncwarp();
                }
                // best_point: otw_eran.py:192-211
                double cj, ct;
                int bj, bt;
                otw_argmin(rw, rs, max(0, S.j - c + 1), S.j, lane, cj, bj);
                otw_argmin(cl, rs, max(0, S.t - c + 1), S.t, lane, ct, bt);
                int x, y;
                if (cj < ct) { x = S.t; y = bj; } else { x = bt; y = S.j; }
                // path append: always (otw_eran.py:160), forward-only for LiveNoteV2 (livenote_v2.py:198-199)
                bool app = true;
                if (a.kind == AFS_LIVENOTE_V2 && S.path_len > 0) app = (x > S.last_x) && (y >= S.last_y);
                if (app) {
                    if (lane == 0) {
                        if (S.path_len < sd.path_cap) path[S.path_len] = make_int2(x, y);
                        if (a.out_points && npts < a.pts) {
                            a.out_points[(oidx * a.pts + npts) * 2 + 0] = x;
                            a.out_points[(oidx * a.pts + npts) * 2 + 1] = y;
                        }
                    }
                    S.path_len += 1;
                    S.last_x = x;
                    S.last_y = y;
                    npts += 1;
                }
                // set_direction: otw_eran.py:162-188
                int nd;
                if (S.t < c) nd = DIR_BOTH;
                else if (S.run_count >= a.max_run) nd = (S.previous == DIR_ROW) ? DIR_COL : DIR_ROW;
                else if (x < S.t) nd = DIR_COL;
                else if (y < S.j) nd = DIR_ROW;
                else nd = DIR_BOTH;
                if (nd != DIR_BOTH && nd == S.previous) S.run_count += 1; else S.run_count = 1;
                if (nd != DIR_BOTH) S.previous = nd;
                S.direction = nd;
                if (nd != DIR_COL) looping = false;
            }
        }
        if (valid && lane == 0) {
            if (a.out_status) a.out_status[oidx] = status;
            if (a.out_npoints) a.out_npoints[oidx] = npts;
        }
    }
    if (valid && lane == 0) {
        a.scal[s] = S;
        a.tj[2 * s] = S.t;
        a.tj[2 * s + 1] = S.j;
        a.path_len[s] = S.path_len;
    }
}

__global__ void otw_reset_kernel(const OtwArgs a)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nring = (int64_t)a.n_streams * a.rs;
    if (i < nring) {
        a.rw[i] = a.fill;
        a.cl[i] = a.fill;
    }
    if (i < a.n_streams) {
        OtwScalars S;
        memset(&S, 0, sizeof(S));
        S.run_count = (a.kind == AFS_OTW) ? 1 : 0;     // otw_eran.py:33 / livenote_v2.py:35
        S.first = 1;
        a.scal[i] = S;
        a.tj[2 * i] = 0;
        a.tj[2 * i + 1] = 0;
        a.path_len[i] = 0;
    }
}

// set_live (otw_eran.py:91-142, livenote_v2.py:108-155) calls set_direction once at
// (0,0) before the first step: best point (0,0) is appended, t < c gives Both, and
// run_count becomes 1.  Applied to freshly reset streams.
__global__ void otw_seed_set_live_kernel(const OtwArgs a)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.n_streams) return;
    OtwScalars S = a.scal[s];
    if (S.path_len == 0 && S.first) {
        reinterpret_cast<int2 *>(a.path)[a.streams[s].path_off] = make_int2(0, 0);
        S.path_len = 1;
        S.last_x = 0;
        S.last_y = 0;
        S.run_count = 1;
        a.scal[s] = S;
        a.path_len[s] = 1;
    }
}

}  // namespace

struct afs_otw {
    OtwArgs args;
    std::vector<OtwStream> streams;
    OtwStream *d_streams = nullptr;
    size_t state_bytes = 0;
    size_t off_scal = 0, off_rw = 0, off_cl = 0, off_lh = 0, off_path = 0, off_tj = 0, off_plen = 0;
    int64_t total_path = 0;
    bool bound = false;
    size_t smem_bytes = 0;
};

extern "C" {

int afs_otw_create(afs_otw **out, int kind, int n_streams, const double *d_ref, const int64_t *h_ref_len,
                   const int64_t *h_ref_off, int n_features, int c, int max_run, int cost_kind)
{
    if (!out || n_streams <= 0 || !d_ref || !h_ref_len || !h_ref_off)
        return afs::fail(AFS_ERR_INVALID, "afs_otw_create: null argument or n_streams <= 0");
    if (n_features != kF) return afs::fail(AFS_ERR_UNSUPPORTED, "afs_otw: n_features must be 12 (got %d)", n_features);
    if (kind < AFS_OTW || kind > AFS_LIVENOTE_V1) return afs::fail(AFS_ERR_INVALID, "afs_otw: bad kind %d", kind);
    if (cost_kind != AFS_COST_COSINE && cost_kind != AFS_COST_EUCLID) return afs::fail(AFS_ERR_INVALID, "afs_otw: bad cost kind");
    if (c < 1 || max_run < 1) return afs::fail(AFS_ERR_INVALID, "afs_otw: c and max_run_count must be >= 1");
    const int cpad = (c + 3) / 4 * 4;
    const int cstride = 2 * cpad + 2;                                   // +2 doubles: consecutive streams start 4 banks apart
    const size_t per_stream = (size_t)(cstride + kF) * sizeof(double);
    // four CTAs per SM (232448 B of shared memory, 1 KB reserved per CTA) when c allows it
    int group = (int)((232448 / 4 - 1024 - 256) / per_stream);
    if (group > kOtwMaxGroup) group = kOtwMaxGroup;
    if (group < 1) group = (int)((227 * 1024 - 256) / per_stream) >= 1 ? 1 : 0;
    if (group < 1) return afs::fail(AFS_ERR_UNSUPPORTED, "afs_otw: c = %d needs %zu B of shared memory per stream (max 227 KiB)", c, per_stream);
    if (group > n_streams) group = n_streams;
    const size_t smem = (size_t)group * per_stream;
    afs_otw *h = new afs_otw();
    memset(&h->args, 0, sizeof(h->args));
    h->streams.resize(n_streams);
    int64_t path_pairs = 0;
    for (int s = 0; s < n_streams; s++) {
        const int64_t N = h_ref_len[s];
        if (N <= 0 || N > (1 << 29)) { delete h; return afs::fail(AFS_ERR_INVALID, "afs_otw: stream %d has invalid reference length", s); }
        h->streams[s].ref_off = h_ref_off[s];
        h->streams[s].N = (int32_t)N;
        h->streams[s].path_cap = (int32_t)(3 * N + 8);   // one point per set_direction call, <= 2N rows + N columns
        h->streams[s].path_off = path_pairs;
        path_pairs += h->streams[s].path_cap;
    }
    h->total_path = path_pairs;
    OtwArgs &a = h->args;
    a.ref = d_ref;
    a.n_streams = n_streams;
    a.kind = kind;
    a.c = c;
    a.max_run = max_run;
    a.metric = cost_kind;
    a.rs = c + 1;
    a.cpad = cpad;
    a.group = group;
    a.cstride = cstride;
    a.pts = max_run + 2;
    a.fill = (kind == AFS_OTW) ? 1e10 : (double)INFINITY;   // otw_eran.py:27 / livenote_v2.py:22-23
    h->smem_bytes = smem;
    size_t off = 0;
    h->off_scal = off; off = afs::align_up(off + sizeof(OtwScalars) * n_streams, 256);
    h->off_rw = off;   off = afs::align_up(off + sizeof(double) * (size_t)n_streams * a.rs, 256);
    h->off_cl = off;   off = afs::align_up(off + sizeof(double) * (size_t)n_streams * a.rs, 256);
    h->off_lh = off;   off = afs::align_up(off + sizeof(double) * (size_t)n_streams * a.rs * kF, 256);
    h->off_path = off; off = afs::align_up(off + sizeof(int32_t) * 2 * (size_t)path_pairs, 256);
    h->off_tj = off;   off = afs::align_up(off + sizeof(int32_t) * 2 * n_streams, 256);
    h->off_plen = off; off = afs::align_up(off + sizeof(int32_t) * n_streams, 256);
    h->state_bytes = off;
    cudaError_t e = cudaMalloc(&h->d_streams, sizeof(OtwStream) * n_streams);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_streams, h->streams.data(), sizeof(OtwStream) * n_streams, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && smem > 48 * 1024)
        e = cudaFuncSetAttribute(otw_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        cudaFree(h->d_streams);
        delete h;
        return afs::fail(AFS_ERR_CUDA, "afs_otw_create: %s", cudaGetErrorString(e));
    }
    a.streams = h->d_streams;
    *out = h;
    return AFS_OK;
}

int afs_otw_destroy(afs_otw *h)
{
    if (!h) return AFS_OK;
    cudaFree(h->d_streams);
    delete h;
    return AFS_OK;
}

int afs_otw_points_per_step(const afs_otw *h)
{
    return h ? h->args.pts : 0;
}

int afs_otw_state_bytes(const afs_otw *h, size_t *bytes)
{
    if (!h || !bytes) return afs::fail(AFS_ERR_INVALID, "afs_otw_state_bytes: null argument");
    *bytes = h->state_bytes;
    return AFS_OK;
}

int afs_otw_reset(afs_otw *h, void *d_state, void *stream)
{
    if (!h || !d_state) return afs::fail(AFS_ERR_INVALID, "afs_otw_reset: null argument");
    char *base = static_cast<char *>(d_state);
    OtwArgs &a = h->args;
    a.scal = reinterpret_cast<OtwScalars *>(base + h->off_scal);
    a.rw = reinterpret_cast<double *>(base + h->off_rw);
    a.cl = reinterpret_cast<double *>(base + h->off_cl);
    a.lh = reinterpret_cast<double *>(base + h->off_lh);
    a.path = reinterpret_cast<int32_t *>(base + h->off_path);
    a.tj = reinterpret_cast<int32_t *>(base + h->off_tj);
    a.path_len = reinterpret_cast<int32_t *>(base + h->off_plen);
    h->bound = true;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t n = (int64_t)a.n_streams * a.rs;
    const int threads = 256;
    otw_reset_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, st>>>(a);
    afs::count_launch();
    AFS_CUDA(cudaGetLastError());
    return AFS_OK;
}

int afs_otw_seed_set_live(afs_otw *h, void *stream)
{
    if (!h || !h->bound) return afs::fail(AFS_ERR_INVALID, "afs_otw_seed_set_live: call afs_otw_reset first");
    const int threads = 128;
    otw_seed_set_live_kernel<<<(h->args.n_streams + threads - 1) / threads, threads, 0, static_cast<cudaStream_t>(stream)>>>(h->args);
    afs::count_launch();
    AFS_CUDA(cudaGetLastError());
    return AFS_OK;
}

int afs_otw_step(afs_otw *h, const double *d_frames, int n_frames, const uint8_t *d_active, int32_t *d_status,
                 int32_t *d_npoints, int32_t *d_points, void *stream)
{
    if (!h || !d_frames || n_frames <= 0) return afs::fail(AFS_ERR_INVALID, "afs_otw_step: null argument or n_frames <= 0");
    if (!h->bound) return afs::fail(AFS_ERR_INVALID, "afs_otw_step: call afs_otw_reset first");
    OtwArgs a = h->args;
    a.frames = d_frames;
    a.n_frames = n_frames;
    a.active = d_active;
    a.out_status = d_status;
    a.out_npoints = d_npoints;
    a.out_points = d_points;
    const int blocks = (a.n_streams + a.group - 1) / a.group;
    otw_step_kernel<<<blocks, a.group * 32, h->smem_bytes, static_cast<cudaStream_t>(stream)>>>(a);
    afs::count_launch();
    AFS_CUDA(cudaGetLastError());
    return AFS_OK;
}

int afs_otw_path_layout(const afs_otw *h, int stream_idx, int64_t *offset, int64_t *capacity)
{
    if (!h || stream_idx < -1 || stream_idx >= h->args.n_streams) return afs::fail(AFS_ERR_INVALID, "afs_otw_path_layout: bad stream");
    if (stream_idx == -1) {
        if (offset) *offset = 0;
        if (capacity) *capacity = h->total_path;
        return AFS_OK;
    }
    if (offset) *offset = h->streams[stream_idx].path_off;
    if (capacity) *capacity = h->streams[stream_idx].path_cap;
    return AFS_OK;
}

int afs_otw_path_ptr(const afs_otw *h, const int32_t **d_path, const int32_t **d_path_len)
{
    if (!h || !h->bound) return afs::fail(AFS_ERR_INVALID, "afs_otw_path_ptr: state not bound");
    if (d_path) *d_path = h->args.path;
    if (d_path_len) *d_path_len = h->args.path_len;
    return AFS_OK;
}

int afs_otw_positions_ptr(const afs_otw *h, const int32_t **d_tj)
{
    if (!h || !h->bound || !d_tj) return afs::fail(AFS_ERR_INVALID, "afs_otw_positions_ptr: state not bound");
    *d_tj = h->args.tj;
    return AFS_OK;
}

int afs_otw_read_window(afs_otw *h, int stream_idx, int32_t *h_scalars, double *h_row, double *h_col, double *h_live, void *stream)
{
    if (!h || !h->bound) return afs::fail(AFS_ERR_INVALID, "afs_otw_read_window: state not bound");
    if (stream_idx < 0 || stream_idx >= h->args.n_streams) return afs::fail(AFS_ERR_INVALID, "afs_otw_read_window: bad stream");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const OtwArgs &a = h->args;
    const int rs = a.rs, c = a.c;
    OtwScalars S;
    std::vector<double> rw(rs), cl(rs), lh((size_t)rs * kF);
    AFS_CUDA(cudaMemcpyAsync(&S, a.scal + stream_idx, sizeof(S), cudaMemcpyDeviceToHost, st));
    AFS_CUDA(cudaMemcpyAsync(rw.data(), a.rw + (int64_t)stream_idx * rs, sizeof(double) * rs, cudaMemcpyDeviceToHost, st));
    AFS_CUDA(cudaMemcpyAsync(cl.data(), a.cl + (int64_t)stream_idx * rs, sizeof(double) * rs, cudaMemcpyDeviceToHost, st));
    AFS_CUDA(cudaMemcpyAsync(lh.data(), a.lh + (int64_t)stream_idx * rs * kF, sizeof(double) * rs * kF, cudaMemcpyDeviceToHost, st));
    AFS_CUDA(cudaStreamSynchronize(st));
    if (h_scalars) {
        h_scalars[0] = S.t; h_scalars[1] = S.j; h_scalars[2] = S.previous; h_scalars[3] = S.run_count; h_scalars[4] = S.direction;
        h_scalars[5] = S.first; h_scalars[6] = S.status; h_scalars[7] = S.path_len;
    }
    const double nan = std::numeric_limits<double>::quiet_NaN();
    for (int k = 0; k <= c; k++) {
        const int col = S.j - c + k, row = S.t - c + k;      // slot = index mod (c + 1)
        if (h_row) h_row[k] = col >= 0 ? rw[col % rs] : nan;
        if (h_col) h_col[k] = row >= 0 ? cl[row % rs] : nan;
        if (h_live)
            for (int f = 0; f < kF; f++) h_live[(size_t)f * (c + 1) + k] = row >= 0 ? lh[(size_t)f * rs + row % rs] : nan;
    }
    return AFS_OK;
}

}  // extern "C"

/*
 * TEST INFRASTRUCTURE — scalar C restatement of the reference's alignment
 * algorithms (smritip/real-time-audio-sync).  This is the CPU oracle the CUDA
 * path is checked against; it is never linked into, imported by, or called
 * from the product package.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may use it.
 *
 * Parity status: PINNED.  tests/test_oracle_cpu.py runs this file
 * against the unmodified reference sources (oracle/ref_shim.py) in the build
 * container — bit-equal acc_cost matrices and identical paths — and
 * tests/golden/ holds vectors generated from the reference itself, including
 * the reference's own golden file Songs/chopin/tests/wtw_test_20b.txt.
 *
 * Every floating-point operation is written out explicitly (compile with
 * -ffp-contract=off): the rounding order below is the one the reference's
 * numpy/OpenBLAS calls execute:
 *   - dgemm  K=12  (dtw.py:11)                : s=a0*b0; s=fma(ak,bk,s), k=1..11
 *   - strided ddot (otw_eran.py:220, livenote_v2.py:170, wtw.py:169 numerator)
 *                                             : 4-blocked two-accumulator form
 *   - contiguous ddot n<32 (np.linalg.norm, wtw.py:169): sequential fma chain
 *   - np.sum of 12 contiguous doubles (livenote_v2.py:168): 8-lane pairwise
 *
 * Layout convention everywhere: feature-major C-ordered arrays (F, frames),
 * exactly what the reference passes around.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ dots */

/* dgemm micro-kernel order for K small: one fma chain starting from the
 * rounded first product (fma(a,b,+0) == round(a*b)).  dtw.py:11 */
static double dot_gemm(const double *a, long sa, const double *b, long sb, int n)
{
    double s = a[0] * b[0];
    for (int k = 1; k < n; k++)
        s = fma(a[k * sa], b[k * sb], s);
    return s;
}

/* OpenBLAS generic strided ddot: blocks of 4, two running sums, products of
 * elements 2,3 rounded first then fused with elements 0,1.
 * otw_eran.py:220 / livenote_v2.py:170 / wtw.py:169 (np.dot of two column views) */
static double dot_strided(const double *x, long sx, const double *y, long sy, int n)
{
    double t1 = 0.0, t2 = 0.0;
    int i = 0, n4 = n & ~3;
    for (; i < n4; i += 4) {
        double m3 = y[(i + 2) * sy] * x[(i + 2) * sx];
        double m4 = y[(i + 3) * sy] * x[(i + 3) * sx];
        t1 = t1 + fma(y[i * sy], x[i * sx], m3);
        t2 = t2 + fma(y[(i + 1) * sy], x[(i + 1) * sx], m4);
    }
    for (; i < n; i++)
        t1 = fma(y[i * sy], x[i * sx], t1);
    return t1 + t2;
}

/* contiguous ddot for n < 32: plain fused tail loop.  (np.linalg.norm ravel()s
 * the column view into a contiguous copy first, wtw.py:169) */
static double dot_contig_small(const double *x, long sx, const double *y, long sy, int n)
{
    double s = 0.0;
    for (int i = 0; i < n; i++)
        s = fma(y[i * sy], x[i * sx], s);
    return s;
}

/* numpy pairwise sum of a contiguous double vector, n < 128 branch. */
static double np_sum_small(const double *v, int n)
{
    if (n < 8) {
        double r = 0.0;
        for (int i = 0; i < n; i++) r = r + v[i];
        return r;
    }
    double r[8];
    for (int k = 0; k < 8; k++) r[k] = v[k];
    int i = 8;
    for (; i < n - (n % 8); i += 8)
        for (int k = 0; k < 8; k++) r[k] = r[k] + v[i + k];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; i++) res = res + v[i];
    return res;
}

/* ------------------------------------------------------------------ DTW
 * dtw.py:5-53.  seq_a (F,M), seq_b (F,N).  Steps (0,-1),(-1,0),(-1,-1) with
 * weights 1,1,2; first minimum wins (np.argmin) in the order left, up, diag.
 * cost / acc may be NULL (then only two acc rows are kept); back is M*N bytes
 * supplied by the caller (or NULL -> allocated internally).
 * Returns path length; path is written start-to-end as (i,j) int64 pairs.
 */
int64_t orc_dtw(const double *a, const double *b, int F, int64_t M, int64_t N,
                double *cost, double *acc, int64_t *path, double *acc_end)
{
    uint8_t *back = (uint8_t *)malloc((size_t)M * (size_t)N);
    double *rows = NULL;
    if (!acc) rows = (double *)malloc(sizeof(double) * 2 * (size_t)N);
    if (!back || (!acc && !rows)) { free(back); free(rows); return -1; }

    for (int64_t i = 0; i < M; i++) {
        double *cur = acc ? acc + i * N : rows + (i & 1) * N;
        const double *prv = acc ? acc + (i - 1) * N : rows + ((i - 1) & 1) * N;
        uint8_t *bk = back + i * N;
        for (int64_t j = 0; j < N; j++) {
            double c = 1.0 - dot_gemm(a + i, M, b + j, N, F);           /* dtw.py:11 */
            if (cost) cost[i * N + j] = c;
            if (i == 0 && j == 0) { cur[0] = c; bk[0] = 2; continue; }  /* dtw.py:20-21 */
            if (j == 0) { cur[0] = c + prv[0]; bk[0] = 1; continue; }   /* dtw.py:23-25 */
            if (i == 0) { cur[j] = c + cur[j - 1]; bk[j] = 0; continue; } /* dtw.py:26-28 */
            double o0 = cur[j - 1] + c;                                 /* dtw.py:35-37 */
            double o1 = prv[j] + c;
            double o2 = prv[j - 1] + 2.0 * c;
            int best = 0; double v = o0;                                /* np.argmin: first min */
            if (o1 < v) { v = o1; best = 1; }
            if (o2 < v) { v = o2; best = 2; }
            cur[j] = v; bk[j] = (uint8_t)best;
        }
        if (i == M - 1 && acc_end) *acc_end = cur[N - 1];
    }
    /* dtw.py:43-52 backtrack, then reverse */
    int64_t i = M - 1, j = N - 1, n = 0;
    path[0] = i; path[1] = j; n = 1;
    while (i > 0 || j > 0) {
        int s = back[i * N + j];
        if (s == 0) j -= 1; else if (s == 1) i -= 1; else { i -= 1; j -= 1; }
        path[2 * n] = i; path[2 * n + 1] = j; n++;
    }
    for (int64_t lo = 0, hi = n - 1; lo < hi; lo++, hi--) {
        int64_t t0 = path[2 * lo], t1 = path[2 * lo + 1];
        path[2 * lo] = path[2 * hi]; path[2 * lo + 1] = path[2 * hi + 1];
        path[2 * hi] = t0; path[2 * hi + 1] = t1;
    }
    free(back); free(rows);
    return n;
}

/* Throughput-only variant for the CPU baseline: same arithmetic as orc_dtw
 * (1-byte backpointers, two acc rows) over several pairs.  Thread-safe and
 * GIL-free under ctypes, so the bench runs one call per host thread.
 * Returns the sum of path lengths (so the work cannot be elided). */
int64_t orc_dtw_many(const double *a, const double *b, int F, int64_t M, int64_t N,
                     int n_pairs, double *acc_end_out)
{
    int64_t total = 0;
    for (int p = 0; p < n_pairs; p++) {
        int64_t *path = (int64_t *)malloc(sizeof(int64_t) * 2 * (size_t)(M + N));
        double e = 0.0;
        int64_t n = orc_dtw(a + (size_t)p * F * M, b + (size_t)p * F * N, F, M, N, NULL, NULL, path, &e);
        if (acc_end_out) acc_end_out[p] = e;
        total += n;
        free(path);
    }
    return total;
}

/* ------------------------------------------------------------------ OTW family
 * One dense state object per stream, exactly like the reference keeps it:
 *   OnlineTimeWarping  otw_eran.py:6-36     (kind 0)
 *   LiveNote (v1)      livenote.py          (kind 2)
 *   LiveNoteV2         livenote_v2.py:8-40  (kind 1)
 * rows = live (capacity 2N), cols = ref (N).
 */
enum { DIR_BOTH = 0, DIR_ROW = 1, DIR_COL = 2 };
enum { KIND_OTW = 0, KIND_LN2 = 1, KIND_LN1 = 2 };

typedef struct {
    int kind, F, metric;          /* metric 0: 1-dot, 1: euclid (livenote_v2.py:167-170) */
    int64_t N, L;                 /* ref frames, live capacity = 2N */
    int64_t c; int max_run;
    double *ref, *live, *acc;     /* (F,N) (F,L) (L,N) */
    int64_t t, j;
    int previous, run_count, direction, first;
    int64_t *path; int64_t path_len, path_cap;
    int64_t n_evals;
} orc_otw;

orc_otw *orc_otw_create(int kind, const double *ref, int F, int64_t N, int64_t c, int max_run, int metric)
{
    orc_otw *s = (orc_otw *)calloc(1, sizeof(orc_otw));
    s->kind = kind; s->F = F; s->N = N; s->L = 2 * N; s->c = c; s->max_run = max_run; s->metric = metric;
    s->ref = (double *)malloc(sizeof(double) * F * N);
    memcpy(s->ref, ref, sizeof(double) * F * N);
    s->live = (double *)malloc(sizeof(double) * F * s->L);
    s->acc = (double *)malloc(sizeof(double) * (size_t)s->L * (size_t)N);
    /* otw_eran.py:27 finite 1e10 sentinel; livenote_v2.py:22-23 +inf */
    double fill = (kind == KIND_OTW) ? 1e10 : INFINITY;
    for (size_t q = 0; q < (size_t)s->L * (size_t)N; q++) s->acc[q] = fill;
    s->t = s->j = 0; s->previous = 0;
    s->run_count = (kind == KIND_OTW) ? 1 : 0;        /* otw_eran.py:33 / livenote_v2.py:35 */
    s->direction = DIR_BOTH; s->first = 1;
    s->path_cap = 4 * (s->L + N) + 16;
    s->path = (int64_t *)malloc(sizeof(int64_t) * 2 * s->path_cap);
    s->path_len = 0;
    return s;
}

void orc_otw_destroy(orc_otw *s)
{
    if (!s) return;
    free(s->ref); free(s->live); free(s->acc); free(s->path); free(s);
}

/* otw_eran.py:215-239 / livenote_v2.py:165-189 */
static void otw_eval(orc_otw *s, int64_t x, int64_t y)
{
    s->n_evals++;
    double cst;
    if (s->metric == 1) {
        double d[64];
        for (int k = 0; k < s->F; k++) {
            double e = s->live[k * s->L + x] - s->ref[k * s->N + y];
            d[k] = e * e;
        }
        cst = sqrt(np_sum_small(d, s->F));
    } else {
        cst = 1.0 - dot_strided(s->live + x, s->L, s->ref + y, s->N, s->F);
    }
    double *A = s->acc;
    int64_t N = s->N;
    if (x == 0 && y == 0) { A[0] = cst; return; }
    double best = INFINITY; int have = 0;
    if (y > 0) { double v = A[x * N + (y - 1)] + cst; if (!have || v < best) best = v; have = 1; }
    if (x > 0) { double v = A[(x - 1) * N + y] + cst; if (!have || v < best) best = v; have = 1; }
    if (x > 0 && y > 0) { double v = A[(x - 1) * N + (y - 1)] + 2.0 * cst; if (!have || v < best) best = v; have = 1; }
    A[x * N + y] = best;
}

/* otw_eran.py:192-211 / livenote_v2.py:219-236: first-min argmin along the
 * current row and the current column; row candidate wins only if strictly less. */
static void otw_best_point(const orc_otw *s, int64_t *bx, int64_t *by)
{
    int64_t N = s->N;
    int64_t j1 = s->j - s->c + 1; if (j1 < 0) j1 = 0;
    int64_t bj = j1; double cj = s->acc[s->t * N + j1];
    for (int64_t k = j1 + 1; k <= s->j; k++) { double v = s->acc[s->t * N + k]; if (v < cj) { cj = v; bj = k; } }
    int64_t t1 = s->t - s->c + 1; if (t1 < 0) t1 = 0;
    int64_t bt = t1; double ct = s->acc[t1 * N + s->j];
    for (int64_t k = t1 + 1; k <= s->t; k++) { double v = s->acc[k * N + s->j]; if (v < ct) { ct = v; bt = k; } }
    if (cj < ct) { *bx = s->t; *by = bj; } else { *bx = bt; *by = s->j; }
}

static void otw_append(orc_otw *s, int64_t x, int64_t y)
{
    if (s->kind == KIND_LN2 && s->path_len > 0) {             /* livenote_v2.py:198-199 */
        int64_t px = s->path[2 * (s->path_len - 1)], py = s->path[2 * (s->path_len - 1) + 1];
        if (!(x > px && y >= py)) return;
    }
    if (s->path_len >= s->path_cap) {
        s->path_cap *= 2;
        s->path = (int64_t *)realloc(s->path, sizeof(int64_t) * 2 * s->path_cap);
    }
    s->path[2 * s->path_len] = x; s->path[2 * s->path_len + 1] = y; s->path_len++;
}

/* otw_eran.py:153-188 == livenote_v2.py:193-217 + :91-100 */
static void otw_set_direction(orc_otw *s)
{
    int64_t x, y;
    otw_best_point(s, &x, &y);
    otw_append(s, x, y);
    int nd;
    if (s->t < s->c) nd = DIR_BOTH;
    else if (s->run_count >= s->max_run) nd = (s->previous == DIR_ROW) ? DIR_COL : DIR_ROW;
    else if (x < s->t) nd = DIR_COL;
    else if (y < s->j) nd = DIR_ROW;
    else nd = DIR_BOTH;
    /* `direction == previous` compares against None/'Row'/'Column': BOTH never equals previous */
    if (nd != DIR_BOTH && nd == s->previous) s->run_count += 1; else s->run_count = 1;
    if (nd != DIR_BOTH) s->previous = nd;
    s->direction = nd;
}

/* otw_eran.py:38-85 / livenote_v2.py:43-104.  Returns 0 = None, 1 = "stop",
 * 2 = None because the pre-allocated live buffer is full. */
int orc_otw_insert(orc_otw *s, const double *frame)
{
    if (s->first) {
        s->first = 0;
        for (int k = 0; k < s->F; k++) s->live[k * s->L + s->t] = frame[k];
        otw_eval(s, s->t, s->j);
        return 0;
    }
    s->t += 1;
    if (s->t >= s->L) return 2;
    for (int k = 0; k < s->F; k++) s->live[k * s->L + s->t] = frame[k];
    int64_t k1 = s->j - s->c + 1; if (k1 < 0) k1 = 0;
    for (int64_t k = k1; k <= s->j; k++) otw_eval(s, s->t, k);
    for (;;) {
        if (s->direction != DIR_ROW) {
            s->j += 1;
            if (s->j >= s->N) return 1;
            int64_t r1 = s->t - s->c + 1; if (r1 < 0) r1 = 0;
            for (int64_t k = r1; k <= s->t; k++) otw_eval(s, k, s->j);
        }
        otw_set_direction(s);
        if (s->direction != DIR_COL) break;
    }
    return 0;
}

int64_t orc_otw_path_len(const orc_otw *s) { return s->path_len; }
const int64_t *orc_otw_path(const orc_otw *s) { return s->path; }
int64_t orc_otw_t(const orc_otw *s) { return s->t; }
int64_t orc_otw_j(const orc_otw *s) { return s->j; }
int64_t orc_otw_evals(const orc_otw *s) { return s->n_evals; }
double orc_otw_acc(const orc_otw *s, int64_t x, int64_t y) { return s->acc[x * s->N + y]; }

/* ------------------------------------------------------------------ WTW
 * wtw.py:71-128 on chroma columns (the audio->chroma part is restated in
 * oracle/chroma_oracle.py with numpy's own rfft).  W = dtw_win_size/hop_size,
 * h = dtw_hop_size/hop_size.
 */
typedef struct {
    int F; int64_t M, Ncap;       /* ref frames M, live capacity 2M */
    int W, h;
    double *ref, *live;           /* (F,M) (F,Ncap) */
    int64_t chroma_ptr, live_ptr, ref_ptr;
    int64_t *path; int64_t path_len, path_cap;
    double *C, *D; uint8_t *B; int64_t *sub;
} orc_wtw;

orc_wtw *orc_wtw_create(const double *ref, int F, int64_t M, int W, int h)
{
    orc_wtw *s = (orc_wtw *)calloc(1, sizeof(orc_wtw));
    s->F = F; s->M = M; s->Ncap = 2 * M; s->W = W; s->h = h;
    s->ref = (double *)malloc(sizeof(double) * F * M);
    memcpy(s->ref, ref, sizeof(double) * F * M);
    s->live = (double *)calloc((size_t)F * s->Ncap, sizeof(double));   /* wtw.py:56 zeros */
    s->path_cap = 8 * (s->Ncap + M) + 64;
    s->path = (int64_t *)malloc(sizeof(int64_t) * 2 * s->path_cap);
    s->C = (double *)malloc(sizeof(double) * W * W);
    s->D = (double *)malloc(sizeof(double) * W * W);
    s->B = (uint8_t *)malloc((size_t)W * W);
    s->sub = (int64_t *)malloc(sizeof(int64_t) * 2 * (2 * W + 2));
    return s;
}

void orc_wtw_destroy(orc_wtw *s)
{
    if (!s) return;
    free(s->ref); free(s->live); free(s->path); free(s->C); free(s->D); free(s->B); free(s->sub); free(s);
}

static void wtw_push(orc_wtw *s, int64_t x, int64_t y)
{
    if (s->path_len >= s->path_cap) {
        s->path_cap *= 2;
        s->path = (int64_t *)realloc(s->path, sizeof(int64_t) * 2 * s->path_cap);
    }
    s->path[2 * s->path_len] = x; s->path[2 * s->path_len + 1] = y; s->path_len++;
}

/* one W x W window at the current pointers: wtw.py:101-128 */
static void wtw_window(orc_wtw *s)
{
    int W = s->W, F = s->F;
    const double *X = s->live + s->live_ptr;   /* column i at X + i, stride Ncap */
    const double *Y = s->ref + s->ref_ptr;     /* column j at Y + j, stride M    */
    double nx[1024], ny[1024];
    /* the reference recomputes both norms for every cell (wtw.py:169); values are identical */
    for (int i = 0; i < W; i++) nx[i] = sqrt(dot_contig_small(X + i, s->Ncap, X + i, s->Ncap, F));
    for (int j = 0; j < W; j++) ny[j] = sqrt(dot_contig_small(Y + j, s->M, Y + j, s->M, F));
    for (int i = 0; i < W; i++)
        for (int j = 0; j < W; j++)
            s->C[i * W + j] = 1.0 - dot_strided(X + i, s->Ncap, Y + j, s->M, F) / (nx[i] * ny[j]);
    /* run_dtw wtw.py:173-217: weights 1,1,1; candidates (i-1,j) then (i,j-1) then (i-1,j-1), strict < */
    double *C = s->C, *D = s->D; uint8_t *B = s->B;
    D[0] = C[0]; B[0] = 0;
    double run = C[0];
    for (int i = 1; i < W; i++) { run = run + C[i * W]; D[i * W] = run; B[i * W] = 3; }
    run = C[0];
    for (int j = 1; j < W; j++) { run = run + C[j]; D[j] = run; B[j] = 1; }
    for (int i = 1; i < W; i++)
        for (int j = 1; j < W; j++) {
            double m = D[(i - 1) * W + j]; uint8_t code = 3;
            double v = D[i * W + j - 1];
            if (v < m) { m = v; code = 1; }
            v = D[(i - 1) * W + j - 1];
            if (v < m) { m = v; code = 2; }
            D[i * W + j] = m + C[i * W + j];
            B[i * W + j] = code;
        }
    /* find_path wtw.py:219-240 */
    int n = 0; int ci = W - 1, cj = W - 1;
    s->sub[0] = ci; s->sub[1] = cj; n = 1;
    while (ci != 0 || cj != 0) {
        uint8_t p = B[ci * W + cj];
        if (p == 1) cj -= 1; else if (p == 2) { ci -= 1; cj -= 1; } else ci -= 1;
        s->sub[2 * n] = ci; s->sub[2 * n + 1] = cj; n++;
    }
    /* stitch wtw.py:107-128 (sub is end-to-start here; walk it reversed) */
    int change = 0; int64_t pl = 0, pr = 0;
    for (int q = n - 1, idx = 0; q >= 0; q--, idx++) {
        int64_t l = s->sub[2 * q], r = s->sub[2 * q + 1];
        if (l <= s->h) {
            wtw_push(s, l + s->live_ptr, r + s->ref_ptr);
        } else {
            change = 1;
            /* index = i-1: the previous sub-path point (python's [-1] wrap cannot occur: point 0 is (0,0)) */
            pl = s->sub[2 * (q + 1)]; pr = s->sub[2 * (q + 1) + 1];
            break;
        }
    }
    if (change) { s->live_ptr += pl; s->ref_ptr += pr; }
    else { s->live_ptr += s->h; s->ref_ptr += s->h; }
}

/* Feed ONE new live chroma column (what each trip of the wtw.py:81-93 loop
 * produces).  `pre_stop` mirrors the entry check wtw.py:76-77, evaluated by the
 * caller once per insert() call via orc_wtw_should_stop().  Returns 1 = "stop". */
int orc_wtw_should_stop(const orc_wtw *s)
{
    return (s->ref_ptr >= s->M - 1 || s->live_ptr >= s->Ncap - 1) ? 1 : 0;
}

int orc_wtw_push_chroma(orc_wtw *s, const double *col)
{
    if (s->chroma_ptr >= s->Ncap) return 3;      /* the reference would raise IndexError here */
    for (int k = 0; k < s->F; k++) s->live[k * s->Ncap + s->chroma_ptr] = col[k];
    s->chroma_ptr += 1;
    if (s->ref_ptr >= (s->M - 1 - s->W) || s->live_ptr >= (s->Ncap - 1 - s->W)) return 1;  /* wtw.py:96-97 */
    while (s->chroma_ptr - s->live_ptr >= s->W) wtw_window(s);                             /* wtw.py:100 */
    return 0;
}

int64_t orc_wtw_path_len(const orc_wtw *s) { return s->path_len; }
const int64_t *orc_wtw_path(const orc_wtw *s) { return s->path; }
int64_t orc_wtw_live_ptr(const orc_wtw *s) { return s->live_ptr; }
int64_t orc_wtw_ref_ptr(const orc_wtw *s) { return s->ref_ptr; }

/* exported probes so tests can pin the dot-product rounding orders against numpy */
double orc_dot_gemm(const double *a, const double *b, int n) { return dot_gemm(a, 1, b, 1, n); }
double orc_dot_strided(const double *a, const double *b, int n) { return dot_strided(a, 1, b, 1, n); }
double orc_dot_contig(const double *a, const double *b, int n) { return dot_contig_small(a, 1, b, 1, n); }
double orc_np_sum(const double *a, int n) { return np_sum_small(a, n); }

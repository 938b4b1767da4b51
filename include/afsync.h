/*
 * afsync.h — C ABI of libafsync.so, the B200 (sm_100a) implementation of the
 * alignment hot path of smritip/real-time-audio-sync.
 *
 * The reference has no FFI seam of its own (it is pure Python); the seam is the
 * set of Python callables in chroma.py, dtw.py, otw_eran.py, livenote_v2.py and
 * wtw.py.  Each entry point below names the reference callable whose arithmetic
 * it replaces (file:line in the reference tree).  The Python modules of the same
 * names in real-time-audio-sync_b200/ bind these symbols with ctypes and keep
 * the reference signatures (see INTEGRATION.md for the stub).
 *
 * Conventions
 *   - every pointer named d_* is a DEVICE pointer on the current CUDA device;
 *     h_* pointers are host memory; sizes are element counts unless named *_bytes;
 *   - sequences are feature-major, C-ordered (F, frames) exactly as the reference
 *     passes them; F must be 12 on the CUDA path;
 *   - `stream` is a cudaStream_t passed as void* (0 = default stream); calls are
 *     asynchronous with respect to the host unless stated otherwise;
 *   - return value: 0 = AFS_OK, negative = error (see afs_status); the message of
 *     the last error on the calling thread is returned by afs_last_error();
 *   - no exceptions, no ownership transfer: the caller owns every buffer it
 *     passes; opaque handles own only their small metadata.
 */
#ifndef AFSYNC_H_
#define AFSYNC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum afs_status {
    AFS_OK = 0,
    AFS_ERR_INVALID = -1,     /* bad argument (NULL, size <= 0, F != 12, ...) */
    AFS_ERR_CUDA = -2,        /* a CUDA runtime call failed */
    AFS_ERR_NOMEM = -3,       /* workspace too small / allocation failed */
    AFS_ERR_UNSUPPORTED = -4  /* parameter outside what the kernels implement */
} afs_status;

/* AFS_BF16X3: only as afs_chroma_batch's compute_dtype — the tcgen05 path (DFT as two matrix products, every
 * operand split into two bf16 terms, fp32 accumulation in TMEM); <= 2e-5 abs on normalised chroma. */
typedef enum afs_dtype { AFS_F64 = 0, AFS_F32 = 1, AFS_BF16X3 = 2 } afs_dtype;

const char *afs_last_error(void);
/* "libafsync <version> sm_100a" */
const char *afs_version(void);
/* number of kernels this library has launched in this process (bench: gpu_launches) */
int64_t afs_launch_count(void);

/* ===================================================================== DTW
 * Replaces dtw.DTW(seq_a, seq_b)  (dtw.py:5-53): cost = 1 - a_i.b_j computed
 * inline, steps left/up/diag with weights 1,1,2, first minimum wins in the order
 * left, up, diag (np.argmin, dtw.py:38), backtrack from (M-1,N-1) (dtw.py:43-52).
 * A plan describes a batch of independent pairs; the accumulated-cost matrix is
 * never materialised (2-bit direction map only) unless `dense` is requested.
 */
typedef struct afs_dtw_plan afs_dtw_plan;

/* h_len_a[p] = M_p (frames of seq_a), h_len_b[p] = N_p.  Pair p's sequences live
 * at d_a + h_off_a[p] and d_b + h_off_b[p] (element offsets), each (12, len).
 * dtype selects fp64 (bit-exact mode) or fp32 arithmetic. */
int afs_dtw_plan_create(afs_dtw_plan **out, int n_pairs, const int64_t *h_len_a, const int64_t *h_len_b,
                        const int64_t *h_off_a, const int64_t *h_off_b, int n_features, int dtype);
int afs_dtw_plan_destroy(afs_dtw_plan *plan);
/* bytes of caller-provided device workspace (direction maps + band hand-off rows).  A workspace belongs to ONE plan
 * and to one launch at a time (accumulate, then backtrack, on the same stream or properly ordered streams). */
int afs_dtw_plan_workspace_bytes(const afs_dtw_plan *plan, size_t *bytes);
/* capacity (in (i,j) pairs) of pair p's slot in d_path and its element offset */
int afs_dtw_plan_path_layout(const afs_dtw_plan *plan, int pair, int64_t *offset, int64_t *capacity);

/* Wavefront accumulate (kernel K2).  d_acc_end[p] receives acc_cost[M-1,N-1] as
 * double.  d_dense_cost / d_dense_acc (may be NULL) receive pair 0's full (M,N)
 * matrices in the plan dtype; allowed only for single-pair plans. */
int afs_dtw_accumulate(afs_dtw_plan *plan, const void *d_a, const void *d_b, void *d_workspace,
                       double *d_acc_end, void *d_dense_cost, void *d_dense_acc, void *stream);
/* Backtrack (kernel K3) over the direction map left in the workspace.  Pair p's
 * path occupies d_path[2*(off_p + d_path_start[p]) ...] for d_path_len[p] pairs,
 * ordered from (0,0) to (M-1,N-1), int32 (i,j). */
int afs_dtw_backtrack(afs_dtw_plan *plan, const void *d_workspace, int32_t *d_path,
                      int32_t *d_path_start, int32_t *d_path_len, void *stream);

/* ---- K4: one very long pair split into column stripes, one stripe per GPU (BASELINE config[4]).
 * No reference equivalent beyond the recurrence itself (dtw.py:32-40).  Each stripe is a single-pair
 * fp64 plan of (M x N_g) over the stripe's own (12, N_g) slice of seq_b.  d_leftb[i] = acc_cost of the
 * column left of the stripe (NULL for the first stripe); d_in_flag[band] != 0 once rows of that band
 * are valid (NULL: all valid).  d_rightb / d_out_flag: where this stripe publishes its last column and
 * raises the per-band flag for its right neighbour — normally memory of the NEXT GPU mapped with
 * afs_ipc_open (stores travel over NVLink, flags are released at system scope); NULL on the last stripe. */
int afs_dtw_accumulate_stripe(afs_dtw_plan *plan, const void *d_a, const void *d_b_stripe, void *d_workspace,
                              double *d_acc_end, const double *d_leftb, const int *d_in_flag, double *d_rightb,
                              int *d_out_flag, void *stream);
/* Backtrack inside one stripe from (start_i, start_j) (stripe-local column).  Points are written
 * back-to-front into d_path with global column = col0 + j.  d_out3 = {first valid index, count,
 * exit row}: exit row is where the path continues in the left neighbour's last column, -1 at (0,0). */
int afs_dtw_backtrack_stripe(afs_dtw_plan *plan, const void *d_workspace, int start_i, int start_j, int col0,
                             int first_stripe, int32_t *d_path, int32_t *d_out3, void *stream);
/* CUDA-IPC exchange blocks (zero-initialised): allocate + export a 64-byte handle / map a peer's block. */
int afs_ipc_alloc(size_t bytes, void **d_ptr, void *handle64);
int afs_ipc_open(const void *handle64, void **d_ptr);
int afs_ipc_clear(void *d_ptr, size_t bytes, void *stream);
int afs_ipc_close(void *d_ptr);
int afs_ipc_free(void *d_ptr);

/* ===================================================================== OTW family
 * Replaces OnlineTimeWarping.insert (otw_eran.py:38-85, eval_path_cost :215-239,
 * set_direction :153-188, best_point :192-211), LiveNoteV2.insert
 * (livenote_v2.py:43-104, :165-236) and livenote.LiveNote.insert, for a batch of
 * independent streams advanced together (kernel K5).
 */
typedef struct afs_otw afs_otw;
enum { AFS_OTW = 0, AFS_LIVENOTE_V2 = 1, AFS_LIVENOTE_V1 = 2 };
enum { AFS_COST_COSINE = 0, AFS_COST_EUCLID = 1 };   /* livenote_v2.py:167-170 */
enum { AFS_STEP_NONE = 0, AFS_STEP_STOP = 1, AFS_STEP_FULL = 2 };

/* d_ref holds stream s's reference (12, h_ref_len[s]) at element offset h_ref_off[s].
 * c = params['c'] / params['search_band_width'], max_run = params['max_run_count']. */
int afs_otw_create(afs_otw **out, int kind, int n_streams, const double *d_ref, const int64_t *h_ref_len,
                   const int64_t *h_ref_off, int n_features, int c, int max_run, int cost_kind);
int afs_otw_destroy(afs_otw *h);
/* bytes of caller-provided device state (ring windows, live history, paths) */
int afs_otw_state_bytes(const afs_otw *h, size_t *bytes);
/* bind + initialise the state block (must be called once before stepping; calling
 * it again resets every stream to the freshly constructed object) */
int afs_otw_reset(afs_otw *h, void *d_state, void *stream);
/* Put freshly reset streams into the state the reference's set_live() drivers
 * (otw_eran.py:91-142, livenote_v2.py:108-155) have after their first loop-top
 * best-point call: path = [(0,0)], run_count = 1.  Then afs_otw_step over the whole
 * live sequence reproduces set_live's path. */
int afs_otw_seed_set_live(afs_otw *h, void *stream);
/* slots per (frame, stream) in d_points: max_run + 2 (an insert appends at most
 * max_run + 1 points: the run-count limit bounds consecutive column steps) */
int afs_otw_points_per_step(const afs_otw *h);
/* Advance every stream by `n_frames` live frames.  d_frames is (n_frames, n_streams, 12)
 * double.  d_active (may be NULL) is n_streams bytes: 0 = skip this stream.
 * Per (frame, stream): d_status = AFS_STEP_*; d_npoints = points appended by that
 * insert; d_points (.., PTS, 2) int32 = the appended (live, ref) pairs, PTS =
 * afs_otw_points_per_step().  Any of the three outputs may be NULL. */
int afs_otw_step(afs_otw *h, const double *d_frames, int n_frames, const uint8_t *d_active,
                 int32_t *d_status, int32_t *d_npoints, int32_t *d_points, void *stream);
/* Device-resident full paths: per stream, capacity and offset (in pairs) inside the
 * path area, whose device address is returned by afs_otw_path_ptr. */
int afs_otw_path_layout(const afs_otw *h, int stream_idx, int64_t *offset, int64_t *capacity);
int afs_otw_path_ptr(const afs_otw *h, const int32_t **d_path, const int32_t **d_path_len);
/* per-stream scalars (t, j) as the reference exposes .t/.j (.live_ptr/.ref_ptr) */
int afs_otw_positions_ptr(const afs_otw *h, const int32_t **d_tj);
/* On-demand view of the state the reference keeps in dense matrices (otw_eran.py:23-35: .acc_cost, .t, .j, .previous,
 * .run_count, .direction; livenote_v2.py:22-38), for ONE stream, copied to the host (synchronises `stream`):
 *   h_scalars[8] = t, j, previous, run_count, direction (0 "Both", 1 "Row", 2 "Column"; previous 0 = None), first_insert,
 *                  status, path length
 *   h_row[c + 1]  acc_cost[t, j - c + k]      h_col[c + 1]  acc_cost[t - c + k, j]        (k = 0 .. c)
 *   h_live[12][c + 1]  the live frames t - c .. t (self.live[:, t - c + k])
 * Indices before the start of the matrix read NaN; cells of the two lines the algorithm has not evaluated hold the
 * reference's fill value (1e10 for OnlineTimeWarping, +inf for LiveNote).  Any output may be NULL. */
int afs_otw_read_window(afs_otw *h, int stream_idx, int32_t *h_scalars, double *h_row, double *h_col, double *h_live, void *stream);

/* ===================================================================== chroma
 * Replaces chroma.create_stft + create_chroma (chroma.py:44-75) == wtw.WTW.stft +
 * wtw.py:37-41, and the single-frame chroma.wav_to_chroma_col (chroma.py:35-42) ==
 * wtw.py:82-90 (kernel K1): Hann(4096, symmetric) * frame -> rfft -> |X|^2 ->
 * (12 x 2049) filterbank -> per-frame L2 normalise (zero frames stay zero).
 * d_filterbank is the (12, 2049) float64 matrix (host side builds the librosa
 * formula once); it is converted/packed by afs_chroma_plan_create.
 */
typedef struct afs_chroma_plan afs_chroma_plan;
int afs_chroma_plan_create(afs_chroma_plan **out, const double *h_filterbank, int n_fft, int hop, int n_chroma);
int afs_chroma_plan_destroy(afs_chroma_plan *plan);
/* frames produced for a track of n samples: center_pad=1 -> chroma.py:49-54
 * (left zero pad n_fft/2, tail dropped); center_pad=0 -> frames start at 0. */
int64_t afs_chroma_num_frames(const afs_chroma_plan *plan, int64_t n_samples, int center_pad);
/* Batched tracks.  Track k's samples are d_audio[h_offsets[k] .. h_offsets[k+1]);
 * its chroma goes to d_out + 12 * h_out_offsets[k] as (12, frames_k) feature-major
 * (h_out_offsets in frames; may be NULL = packed in track order).
 * out_dtype: AFS_F32 or AFS_F64 storage; compute_dtype: AFS_F32 (throughput) or
 * AFS_F64 (parity mode). normalize=0 returns raw chroma (create_chroma(normalize=False)). */
int afs_chroma_batch(afs_chroma_plan *plan, const float *d_audio, const int64_t *h_offsets, int n_tracks,
                     int center_pad, int normalize, void *d_out, const int64_t *h_out_offsets,
                     int out_dtype, int compute_dtype, void *stream);
/* Same, for 16-bit PCM samples as they sit in the reference's WAV files (Songs/ * / *.wav):
 * sample = value / 32768, exactly librosa.load's int16 scaling (chroma.py:26, test_simple.py:98-99), applied
 * on the device (folded into the window, a power of two: results are bit-identical to afs_chroma_batch on
 * the converted samples).  Halves the host->device bytes of a track. */
int afs_chroma_batch_pcm16(afs_chroma_plan *plan, const int16_t *d_pcm, const int64_t *h_offsets, int n_tracks,
                           int center_pad, int normalize, void *d_out, const int64_t *h_out_offsets,
                           int out_dtype, int compute_dtype, void *stream);
/* chroma.create_stft (chroma.py:44-65 == wtw.WTW.stft, wtw.py:137-160) on its own: the complex spectrum
 * np.fft.rfft(frame * np.hanning(4096)) of every frame, [frame][2049] interleaved (re, im) of the compute type
 * (AFS_F32: complex64, AFS_F64: complex128); frame k of track t is row h_out_offsets[t] + k (NULL = packed). */
int afs_stft_batch(afs_chroma_plan *plan, const float *d_audio, const int64_t *h_offsets, int n_tracks, int center_pad,
                   void *d_spec, const int64_t *h_out_offsets, int compute_dtype, void *stream);

/* ===================================================================== WTW
 * Replaces wtw.WTW.insert's window loop (wtw.py:100-128) with get_cost_matrix
 * (:162-171), run_dtw (:173-217, weights 1,1,1, order down/left/diag with strict <)
 * and find_path (:219-240), for a batch of independent streams fed chroma columns
 * (kernel K6).  W = dtw_win_size/hop_size, h = dtw_hop_size/hop_size.
 */
typedef struct afs_wtw afs_wtw;
int afs_wtw_create(afs_wtw **out, int n_streams, const double *d_ref, const int64_t *h_ref_len,
                   const int64_t *h_ref_off, int n_features, int W, int h);
int afs_wtw_destroy(afs_wtw *h);
int afs_wtw_state_bytes(const afs_wtw *h, size_t *bytes);
int afs_wtw_reset(afs_wtw *h, void *d_state, void *stream);
/* Push `n_frames` live chroma columns per stream: d_cols (n_frames, n_streams, 12).
 * d_status (n_frames, n_streams): AFS_STEP_NONE / AFS_STEP_STOP (wtw.py:96-97). */
int afs_wtw_push(afs_wtw *h, const double *d_cols, int n_frames, const uint8_t *d_active,
                 int32_t *d_status, void *stream);
/* Audio in, alignment out: wtw.py:71-93 (WTW.insert) for a batch of streams without leaving the device.  Stream s
 * contributes the samples d_audio[h_sample_off[s] .. h_sample_off[s+1]) (float32, offsets even), which must hold exactly
 * n_frames un-padded frames (frame q = samples [q hop, q hop + 4096), wtw.py:81-83).  K1 computes their chroma columns
 * (compute = AFS_F64 / AFS_F32 / AFS_BF16X3) into d_scratch (2 * n_frames * n_streams * 12 doubles, caller-provided),
 * a small kernel reorders them and K6 consumes them, all on `stream`.  A stream that reports AFS_STEP_STOP for a frame
 * does not consume the frames after it (the reference returns from insert() there; push them again). */
int afs_wtw_push_audio(afs_wtw *h, afs_chroma_plan *plan, const float *d_audio, const int64_t *h_sample_off, int n_frames,
                       double *d_scratch, const uint8_t *d_active, int32_t *d_status, int compute, void *stream);
int afs_wtw_path_layout(const afs_wtw *h, int stream_idx, int64_t *offset, int64_t *capacity);
int afs_wtw_path_ptr(const afs_wtw *h, const int32_t **d_path, const int32_t **d_path_len);
/* per-stream (chroma_ptr, live_ptr, ref_ptr) */
int afs_wtw_positions_ptr(const afs_wtw *h, const int32_t **d_ptrs);

#ifdef __cplusplus
}
#endif
#endif /* AFSYNC_H_ */

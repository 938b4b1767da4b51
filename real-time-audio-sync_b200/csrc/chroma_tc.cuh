// Internal interface between the chroma front end (chroma.cu) and the tensor-core path (chroma_tc.cu).
#pragma once
#include <cstdint>
#include <vector>

#include "afs_common.cuh"

// One batch of tracks as every chroma kernel sees it (offsets live in device memory for the launch).
struct ChromaBatch {
    const void *audio;           // float32 samples, or int16 PCM (pcm16 = 1: sample = value / 32768, librosa.load's scaling)
    int pcm16;
    const int64_t *sample_off;   // n_tracks + 1
    const int64_t *frame_off;    // n_tracks + 1 (prefix of frames per track)
    const int64_t *out_off;      // n_tracks (frames), output placement
    int n_tracks;
    int64_t total_frames;
    int hop, center_pad, normalize, out_f64;
    void *out;
    // optional: materialise the spectrum (create_stft, chroma.py:44-65) as interleaved (re, im) pairs of the
    // compute type, [frame][2049], frames numbered like the chroma output; generic kernel only
    void *stft_out;
};

struct afs_chroma_tc;   // tables + scratch of the tcgen05 pipeline

// fb is the filterbank as [bin][12] (double); returns AFS_OK and *out == nullptr when the tensor path does not apply
int chroma_tc_create(afs_chroma_tc **out, const std::vector<double> &fb, const std::vector<double> &hann);
void chroma_tc_destroy(afs_chroma_tc *tc);
int chroma_tc_run(afs_chroma_tc *tc, const ChromaBatch &bt, cudaStream_t st);

#!/usr/bin/env python
"""Align a live recording to a reference recording with the B200 drop-in modules — the flow of the reference's
``test_simple.py`` (load two WAVs, chroma, run an aligner, write the path log, optionally score against beat CSVs).

    python examples/align_pair.py ref.wav live.wav --method dtw|otw|livenote_v2|livenote|wtw [--c 50] [--max-run 3]
                                  [--log path.txt] [--ref-csv ref.csv --live-csv live.csv]

WAVs: 22 050 Hz PCM16 (mono or stereo), like ``Songs/*/*.wav`` in the reference.  Needs a B200 and the built library."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "real-time-audio-sync_b200"))      # the reference's module names, GPU inside

from chroma import wav_to_chroma, load_wav, hop_size, fft_len      # noqa: E402  (chroma.py:25-33)
from dtw import DTW                                                 # noqa: E402  (dtw.py:5)
from otw_eran import OnlineTimeWarping                              # noqa: E402  (otw_eran.py:5)
from livenote_v2 import LiveNoteV2                                  # noqa: E402  (livenote_v2.py:3)
from livenote import LiveNote                                       # noqa: E402
from wtw import WTW                                                 # noqa: E402  (wtw.py:19)
from evalutil import write_path_log, BeatScorer                     # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("ref_wav")
    ap.add_argument("live_wav")
    ap.add_argument("--method", default="dtw", choices=["dtw", "otw", "livenote_v2", "livenote", "wtw"])
    ap.add_argument("--c", type=int, default=50, help="search band width (frames)")
    ap.add_argument("--max-run", type=int, default=3)
    ap.add_argument("--log", default=None, help="write the path in the reference's log format")
    ap.add_argument("--ref-csv", default=None)
    ap.add_argument("--live-csv", default=None)
    args = ap.parse_args()

    t0 = time.perf_counter()
    if args.method == "wtw":
        # WTW takes audio (wtw.py:19-41): reference file name in, live samples fed in chunks as the live app does
        params = {"fft_len": fft_len, "hop_size": hop_size, "dtw_win_size": hop_size * 20, "dtw_hop_size": hop_size * 10}
        aligner = WTW(args.ref_wav, params)
        live, fs = load_wav(args.live_wav)
        assert fs == 22050
        chunk = 4096
        for s in range(0, len(live), chunk):
            if aligner.insert(list(live[s : s + chunk])) == "stop":
                break
        path = [tuple(p) for p in aligner.path]
        log_params = {"dtw_win_size": params["dtw_win_size"], "dtw_hop_size": params["dtw_hop_size"]}
    else:
        ref_seq, live_seq = wav_to_chroma(args.ref_wav), wav_to_chroma(args.live_wav)     # test_simple.py:98-99
        if args.method == "dtw":
            _, _, p = DTW(live_seq, ref_seq)                                              # test_simple.py:195
            path = [tuple(x) for x in np.asarray(p).tolist()]
            log_params = {"c": 0, "max_run_count": 0}
        else:
            if args.method == "otw":
                aligner = OnlineTimeWarping(ref_seq, {"c": args.c, "max_run_count": args.max_run})          # :137
                log_params = {"c": args.c, "max_run_count": args.max_run}
            else:
                cls = LiveNoteV2 if args.method == "livenote_v2" else LiveNote
                aligner = cls(ref_seq, {"search_band_width": args.c, "max_run_count": args.max_run})
                log_params = {"search_band_width": args.c, "max_run_count": args.max_run}
            for i in range(live_seq.shape[1]):
                if aligner.insert(live_seq[:, i]) == "stop":                               # None / "stop" protocol
                    break
            path = [tuple(p) for p in aligner.path]
    dt = time.perf_counter() - t0
    print("method %s: %d path points, first %s, last %s, %.2f s" % (args.method, len(path), path[0], path[-1], dt))
    if args.log:
        write_path_log(args.log, os.path.basename(args.ref_wav), fft_len, hop_size, log_params, path)
        print("path log written to", args.log)
    if args.ref_csv and args.live_csv:
        print("beat errors:", BeatScorer(args.ref_csv, args.live_csv).score(path))


if __name__ == "__main__":
    main()

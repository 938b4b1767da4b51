"""GPU: streaming front door (audio chunks of many streams -> K1 chroma columns -> K5 steps) vs the
reference-style per-stream loop run on the oracle (livenote_live.py:161-209 semantics)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_front_door_matches_per_stream_oracle_loop(entry, orc):
    fd_mod = entry.submodule("frontdoor")
    aud = np.load(os.path.join(GOLD, "audio_15s.npz"))
    ch = np.load(os.path.join(GOLD, "chopin_chroma.npz"))
    mono = {t: (aud[t + "_i16"].astype(np.float32) / np.float32(32768.0)).mean(axis=1, dtype=np.float32) for t in ("ref", "live")}
    ref = ch["ref"]
    # three "microphones": the live excerpt, the ref excerpt played against itself, and a delayed start
    feeds = [mono["live"], mono["ref"], np.concatenate((np.zeros(5000, np.float32), mono["live"][:200000]))]
    fd = fd_mod.StreamFrontDoor([ref, ref, ref], 50, 3, kind="otw", compute="fp64")
    # oracle: per stream, frames of 4096 with hop 2048, wav_to_chroma_col + insert
    oracles = [orc.OnlineTimeWarping(ref, {"c": 50, "max_run_count": 3}) for _ in feeds]
    for s, x in enumerate(feeds):
        data = x.copy()
        while len(data) >= 4096:
            if oracles[s].insert(orc.wav_to_chroma_col(data[:4096])) == "stop":
                break
            data = data[2048:]
    # feed the front door in uneven chunks (different chunk size per stream)
    pos = [0, 0, 0]
    sizes = [3000, 7777, 1024]
    while any(pos[s] < len(feeds[s]) for s in range(3)):
        chunks = []
        for s in range(3):
            chunks.append(feeds[s][pos[s] : pos[s] + sizes[s]])
            pos[s] += sizes[s]
        fd.feed(chunks)
    got = fd.paths()
    for s in range(3):
        want = oracles[s].path_array()
        assert np.array_equal(got[s], want), s
        assert fd.position[s] == tuple(int(v) for v in want[-1])
    fd.close()

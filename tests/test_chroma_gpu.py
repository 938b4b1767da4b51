"""GPU parity: fused STFT->chroma kernel (K1) vs the reference's own chroma (golden
vectors generated from chroma.py) and the numpy oracle.  Tolerance from
BASELINE.json north_star: 1e-4 absolute (float32 compute); float64 compute mode 1e-9."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL_F32 = 1e-4
TOL_F64 = 1e-9


@pytest.fixture(scope="module")
def chroma(entry):
    return entry.submodule("chroma")


@pytest.fixture(scope="module")
def audio():
    aud = np.load(os.path.join(GOLD, "audio_15s.npz"))
    out = {}
    for tag in ("ref", "live"):
        st = aud[tag + "_i16"]
        out[tag] = (st.astype(np.float32) / np.float32(32768.0)).mean(axis=1, dtype=np.float32)
        out[tag + "_chroma"] = aud[tag + "_chroma"]
        out[tag + "_raw"] = aud[tag + "_raw_chroma"]
        out[tag + "_cols"] = aud[tag + "_cols"]
    return out


@pytest.mark.parametrize("tag", ["ref", "live"])
def test_real_audio_vs_reference_golden(chroma, audio, tag):
    got = chroma.wav_samples_to_chroma(audio[tag])
    want = audio[tag + "_chroma"]
    assert got.shape == want.shape and got.dtype == np.float64
    err = np.abs(got - want).max()
    assert err < TOL_F32, err
    got64 = chroma.wav_samples_to_chroma(audio[tag], compute="fp64")
    assert np.abs(got64 - want).max() < TOL_F64
    # unit columns
    assert np.allclose(np.linalg.norm(got, axis=0), 1.0, atol=1e-5)


def test_raw_chroma_unnormalised(chroma, audio):
    got = chroma.wav_samples_to_chroma(audio["ref"], normalize=False, compute="fp64")
    want = audio["ref_raw"]
    assert np.abs(got - want).max() <= 1e-9 * np.abs(want).max()
    got32 = chroma.create_chroma(chroma.create_stft(audio["ref"]), normalize=False)
    assert np.abs(got32 - want).max() <= 1e-4 * np.abs(want).max()


def test_single_column(chroma, audio):
    for k, s in enumerate((0, 2048, 50000, 200000)):
        col = chroma.wav_to_chroma_col(audio["live"][s : s + 4096])
        assert col.shape == (12,)
        assert np.abs(col - audio["live_cols"][k]).max() < TOL_F32
        col64 = chroma.wav_to_chroma_col(audio["live"][s : s + 4096], compute="fp64")
        assert np.abs(col64 - audio["live_cols"][k]).max() < TOL_F64
    with pytest.raises(AssertionError):
        chroma.wav_to_chroma_col(np.zeros(100))


def test_batch_ragged_silence_and_short_tracks(chroma, orc):
    rng = np.random.default_rng(2)
    sr = 22050
    t = np.arange(3 * sr) / sr
    tone = (0.5 * np.sin(2 * np.pi * 440 * t) + 0.25 * np.sin(2 * np.pi * 660 * t)).astype(np.float32)
    tracks = [
        tone,
        (0.1 * rng.standard_normal(20001)).astype(np.float32),   # odd length
        np.zeros(10000, dtype=np.float32),                       # silence: zeros stay zeros, no NaN (librosa normalize)
        np.zeros(100, dtype=np.float32),                         # too short for one frame even with the pad
        (0.3 * rng.standard_normal(2048)).astype(np.float32),    # exactly one (half-padded) frame
    ]
    got = chroma.chroma_batch(tracks)
    for k, x in enumerate(tracks):
        want = orc.wav_samples_to_chroma(x)
        assert got[k].shape == want.shape, k
        if want.size:
            assert np.abs(got[k] - want).max() < TOL_F32, k
            assert np.isfinite(got[k]).all()
    assert got[2].shape == (12, 4) and not got[2].any()
    assert got[3].shape == (12, 0)
    # A4 dominates the tone: chroma class 9 (C-based) is the largest
    assert int(np.argmax(got[0][:, 10])) == 9


def test_chroma_diff(chroma, audio):
    c = chroma.wav_samples_to_chroma(audio["ref"], compute="fp64")
    d = chroma.chroma_to_diff(c)
    assert d.shape == (12, c.shape[1] - 1) and (d >= 0).all()


def test_wav_to_chroma_from_a_wav_file(chroma, audio, tmp_path):
    """chroma.wav_to_chroma(path) / wav_to_chroma_diff(path) on a PCM16 stereo WAV (the excerpt of the
    reference's own recording): same numbers as the reference's chroma of that audio."""
    import wave
    aud = np.load(os.path.join(GOLD, "audio_15s.npz"))
    path = os.path.join(str(tmp_path), "excerpt.wav")
    with wave.open(path, "wb") as w:
        w.setnchannels(2)
        w.setsampwidth(2)
        w.setframerate(22050)
        w.writeframes(aud["ref_i16"].astype("<i2").tobytes())
    got = chroma.wav_to_chroma(path)
    assert got.shape == audio["ref_chroma"].shape and np.abs(got - audio["ref_chroma"]).max() < TOL_F32
    d = chroma.wav_to_chroma_diff(path)
    want = np.clip(np.diff(audio["ref_chroma"]), 0, np.inf)
    assert d.shape == want.shape and np.abs(d - want).max() < 2 * TOL_F32


def test_pcm16_input_is_bit_identical_to_float_input(chroma, audio, tmp_path):
    """afs_chroma_batch_pcm16: int16 samples scaled by 1/32768 inside the kernel (librosa.load's int16 rule) give exactly
    the chroma of the float32 samples — fast fp32 kernel, fp64 parity kernel, raw (un-normalised) output, ragged batch
    with tracks shorter than a frame, and the mono-WAV loader path."""
    import wave
    rng = np.random.default_rng(11)
    aud = np.load(os.path.join(GOLD, "audio_15s.npz"))
    mono = aud["ref_i16"][:, 0].astype(np.int16)
    tracks16 = [mono, mono[1000:1000 + 4096 * 5 + 37], rng.integers(-32768, 32767, size=9001).astype(np.int16),
                np.zeros(6000, np.int16), mono[:100], mono[3:3 + 4096]]
    tracksf = [t.astype(np.float32) / np.float32(32768.0) for t in tracks16]
    for kw in (dict(), dict(compute="fp64"), dict(normalize=False), dict(center=False)):
        a = chroma.chroma_batch(tracks16, **kw)
        b = chroma.chroma_batch(tracksf, **kw)
        for x, y in zip(a, b):
            assert x.shape == y.shape and np.array_equal(x, y), kw
    # device-level call with track offsets that are NOT 16-byte aligned (guarded-load path instead of TMA staging)
    import torch
    plan = chroma.default_plan()
    offs = np.array([0, 30002, 30002 + 25000], dtype=np.int64)
    flat16 = rng.integers(-20000, 20000, size=int(offs[-1])).astype(np.int16)
    o16, f16 = plan.run(torch.from_numpy(flat16).cuda(), offs)
    o32, f32 = plan.run(torch.from_numpy(flat16.astype(np.float32) / np.float32(32768.0)).cuda(), offs)
    assert np.array_equal(f16, f32) and torch.equal(o16, o32)
    # mono PCM16 WAV: the loader hands int16 to the kernel
    path = os.path.join(str(tmp_path), "mono.wav")
    with wave.open(path, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(22050)
        w.writeframes(mono.astype("<i2").tobytes())
    pcm, rate = chroma.load_wav_pcm16(path)
    assert rate == 22050 and pcm.dtype == np.int16 and np.array_equal(pcm, mono)
    assert np.array_equal(chroma.wav_to_chroma(path), chroma.wav_samples_to_chroma(tracksf[0]))


def test_full_config_1024_tracks_sampled_parity(chroma, orc):
    """BASELINE config[1] at full size — 1024 five-minute tracks (27 GB of float32 audio, the bench's own generator) in ONE
    launch; the chromagrams of 5 sampled tracks are within 1e-4 of the numpy oracle's, and no frame of any track is
    left unwritten or un-normalised (every column has unit length)."""
    import sys
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    argv, sys.argv = sys.argv, sys.argv[:1]
    try:
        import bench
    finally:
        sys.argv = argv
    T, n = 1024, int(300.0 * 22050)
    audio = bench.synth_audio_tracks(torch, T, n, 4321, "cuda")
    plan = chroma.default_plan()
    offs = np.arange(T + 1, dtype=np.int64) * n
    out, foffs = plan.run(audio.reshape(-1), offs)
    torch.cuda.synchronize()
    frames = int(foffs[1] - foffs[0])
    assert frames == (n + 2048 - 4096) // 2048 + 1          # chroma.py:49-54: left pad of L/2, tail dropped
    cols = out.view(T, 12, frames)
    lens = torch.linalg.vector_norm(cols.double(), dim=1)
    assert bool(((lens - 1.0).abs() < 1e-5).all())          # synthetic tracks have no silent frames
    for k in (0, 1, 511, 1022, 1023):
        want = orc.wav_samples_to_chroma(audio[k].cpu().numpy())
        got = cols[k].cpu().numpy().astype(np.float64)
        assert got.shape == want.shape
        assert np.abs(got - want).max() < TOL_F32, k


def test_create_stft_is_the_reference_spectrum(chroma, audio, orc):
    """chroma.create_stft (chroma.py:44-65) returns the complex (2049, M) spectrum itself (float64 kernel, materialised by
    afs_stft_batch): equal to numpy's rfft of the windowed, left-padded frames to 1e-9 of the spectrum's peak; and
    create_chroma of an arbitrary spectrum array (not produced by create_stft) goes through filterbank + normalisation."""
    x = audio["ref"][: 22050 * 4]
    got = chroma.create_stft(x)
    want = orc.create_stft(x)
    assert got.shape == want.shape and got.shape[0] == 2049 and np.iscomplexobj(got)
    assert np.abs(np.asarray(got) - want).max() <= 1e-9 * np.abs(want).max()
    c1 = chroma.create_chroma(np.array(want))            # a plain ndarray: the spectrum path
    c2 = orc.create_chroma(want)
    assert np.abs(c1 - c2).max() < 1e-9
    c3 = chroma.create_chroma(got)                      # the object create_stft returned: fused path from the samples
    assert np.abs(c3 - c2).max() < 1e-4


@pytest.mark.parametrize("pcm", [False, True])
def test_tensor_core_path_many_short_ragged_tracks(chroma, orc, pcm, monkeypatch):
    """The fused tcgen05 launch on the shapes its fast paths do not cover: hundreds of short tracks of random length, so that
    almost every 2-frame tile straddles two tracks (the converter's guarded loads), every track has zero-padded first and
    last frames (the loader's guarded copies), and — with a 4-tile ring and 3 filterbank CTAs on a fresh plan — the
    power-spectrum ring wraps dozens of times (slot hand-back between the spectrum and the filterbank CTAs)."""
    import torch
    monkeypatch.setenv("AFS_CHROMA_TC_RING", "4")
    monkeypatch.setenv("AFS_CHROMA_TC_FB", "3")
    plan = chroma.ChromaPlan(chroma.fft_len, chroma.hop_size)
    rng = np.random.default_rng(21 + pcm)
    tracks = []
    for k in range(420):
        n = int(rng.integers(2048, 14000))
        t = np.arange(n) / 22050.0
        x = 0.3 * np.sin(2 * np.pi * float(rng.uniform(80, 4000)) * t) + 0.05 * rng.standard_normal(n)
        tracks.append(np.round(x * 20000).astype(np.int16) if pcm else x.astype(np.float32))
    q = 8 if pcm else 4
    lens = [(len(t) + q - 1) // q * q for t in tracks]
    offs = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
    flat = np.zeros(int(offs[-1]), dtype=np.int16 if pcm else np.float32)
    for k, t in enumerate(tracks):
        flat[offs[k] : offs[k] + len(t)] = t
    d_out, foffs = plan.run(torch.from_numpy(flat).cuda(), offs, out_dtype=torch.float64, compute="tc")
    host = d_out.cpu().numpy()
    assert int(foffs[-1]) > 64 * 4 * 4            # the 4-slot ring wrapped several times
    worst = 0.0
    for k, t in enumerate(tracks):
        x = np.zeros(lens[k], dtype=np.float32)
        x[: len(t)] = (t.astype(np.float32) / np.float32(32768.0)) if pcm else t
        want = orc.wav_samples_to_chroma(x)
        got = host[12 * foffs[k] : 12 * foffs[k + 1]].reshape(12, -1)
        assert got.shape == want.shape, k
        worst = max(worst, float(np.abs(got - want).max()))
    assert worst < TOL_F32, worst
    plan.close() if hasattr(plan, "close") else None


def test_one_plan_from_two_streams(chroma, orc):
    """Two launches of the same plan on different streams share the plan's power-spectrum ring: the library orders them on
    the device instead of letting them run into each other (plan state rewritten under a running kernel)."""
    import torch
    rng = np.random.default_rng(5)
    plan = chroma.default_plan()
    xs = [(0.2 * rng.standard_normal(22050 * 20)).astype(np.float32) for _ in range(2)]
    want = [orc.wav_samples_to_chroma(x) for x in xs]
    d = [torch.from_numpy(x).cuda() for x in xs]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    outs = []
    torch.cuda.synchronize()
    for rep in range(6):
        for k in range(2):
            with torch.cuda.stream(streams[k]):
                o, foffs = plan.run(d[k], [0, len(xs[k])], out_dtype=torch.float64)
                outs.append((k, o))
    torch.cuda.synchronize()
    for k, o in outs:
        got = o.cpu().numpy().reshape(12, -1)
        assert np.abs(got - want[k]).max() < TOL_F32

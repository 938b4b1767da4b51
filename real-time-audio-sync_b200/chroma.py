"""Drop-in for the reference's ``chroma.py``.

Reference: chroma.py:20-90 — ``wav_to_chroma(path)``, ``wav_to_chroma_col(buf)``,
``create_stft`` + ``create_chroma`` and ``wav_to_chroma_diff(path)``; globals ``fft_len``,
``hop_size``, ``fs``.  Framing, Hann window, real FFT, |X|^2, the 12 x 2049 filterbank
and the L2 normalisation run fused in kernel K1 (csrc/chroma_tc.cu on the tensor cores, csrc/chroma.cu on the
CUDA cores) through ``afs_chroma_batch``; WAV decoding stays on the CPU (it only feeds samples).

Differences from the reference, all explicit:
* three arithmetic modes: ``compute="tc"`` (default) runs the DFT as two matrix products on the
  tensor cores with bf16x3 split operands and fp32 accumulation (<= 2e-5 abs from the float64
  reference on normalised chroma); ``"fp32"`` is the CUDA-core FFT kernel in float32 (<= 1e-6);
  ``"fp64"`` the same kernel in float64 (<= 1e-9);
* ``create_stft`` returns the complex (2049, M) spectrum like the reference (float64 kernel) wrapped in an
  array subclass that remembers the samples, so ``create_chroma(create_stft(wav))`` runs the fused path;
  ``create_chroma`` on any other (2049, M) array applies filterbank + normalisation to that spectrum;
* ``librosa.load`` is replaced by a WAV reader for 22 050 Hz PCM16 files (no resampler).
"""
import ctypes as C
import wave

import numpy as np
import torch

try:
    from . import _native as nat
except ImportError:
    import _native as nat

# globals (chroma.py:20-22)
fft_len = 4096
hop_size = 2048
fs = 22050


def chroma_filterbank(sr=fs, n_fft=fft_len, n_chroma=12, A440=440.0, ctroct=5.0, octwidth=2.0):
    """librosa.filters.chroma(sr, n_fft) with its defaults (chroma.py:69): (12, 1 + n_fft/2) float64.
    Host-side, computed once; the kernel consumes it through afs_chroma_plan_create."""
    freqs = np.linspace(0, sr, n_fft, endpoint=False)[1:]
    bins = n_chroma * np.log2(freqs / (float(A440) / 16.0))
    bins = np.concatenate(([bins[0] - 1.5 * n_chroma], bins))
    width = np.concatenate((np.maximum(bins[1:] - bins[:-1], 1.0), [1.0]))
    D = np.subtract.outer(bins, np.arange(0, n_chroma, dtype="d")).T
    half = np.round(float(n_chroma) / 2)
    D = np.remainder(D + half + 10 * n_chroma, n_chroma) - half
    wts = np.exp(-0.5 * (2 * D / np.tile(width, (n_chroma, 1))) ** 2)
    wts = wts / np.sqrt(np.sum(wts ** 2, axis=0, keepdims=True))
    wts *= np.tile(np.exp(-0.5 * (((bins / n_chroma - ctroct) / octwidth) ** 2)), (n_chroma, 1))
    wts = np.roll(wts, -3 * (n_chroma // 12), axis=0)
    return np.ascontiguousarray(wts[:, : int(1 + n_fft / 2)])


_COMPUTE = {"tc": nat.AFS_BF16X3, "bf16x3": nat.AFS_BF16X3, "fp32": nat.AFS_F32, "f32": nat.AFS_F32,
            "fp64": nat.AFS_F64, "f64": nat.AFS_F64}


class ChromaPlan(object):
    """Device tables (window, twiddles, filterbank) for one (n_fft, hop) configuration."""

    def __init__(self, n_fft=fft_len, hop=hop_size, filterbank=None, device=None):
        nat.require_cuda()
        self.device = nat.device() if device is None else torch.device(device)
        self.n_fft, self.hop = int(n_fft), int(hop)
        fb = chroma_filterbank(fs, n_fft) if filterbank is None else np.ascontiguousarray(filterbank, dtype=np.float64)
        assert fb.shape == (12, 1 + n_fft // 2)
        self.filterbank = fb
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            nat.check(nat.lib().afs_chroma_plan_create(C.byref(h), C.c_void_p(fb.ctypes.data), self.n_fft, self.hop, 12))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            nat.lib().afs_chroma_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def num_frames(self, n_samples, center=True):
        return int(nat.lib().afs_chroma_num_frames(self._h, int(n_samples), 1 if center else 0))

    def run(self, d_audio, offsets, d_out=None, center=True, normalize=True, out_dtype=torch.float32, compute="tc",
            out_offsets=None):
        """K1 on the current stream.  d_audio: float32 device tensor holding all tracks, or int16 PCM
        (sample = value / 32768 as librosa.load scales WAV data; converted inside the kernel);
        offsets: (n_tracks+1) sample offsets.  Returns (d_out, frame_offsets): track k's chroma
        is d_out[12*frame_offsets[k] : 12*frame_offsets[k+1]].view(12, frames_k)."""
        assert d_audio.is_cuda and d_audio.dtype in (torch.float32, torch.int16)
        offs = np.ascontiguousarray(offsets, dtype=np.int64)
        n = offs.shape[0] - 1
        frames = np.array([self.num_frames(offs[k + 1] - offs[k], center) for k in range(n)], dtype=np.int64)
        foffs = np.concatenate(([0], np.cumsum(frames))).astype(np.int64)
        if d_out is None:
            d_out = torch.empty(int(12 * foffs[-1]), dtype=out_dtype, device=d_audio.device)
        oo = None if out_offsets is None else np.ascontiguousarray(out_offsets, dtype=np.int64)
        entry = nat.lib().afs_chroma_batch_pcm16 if d_audio.dtype == torch.int16 else nat.lib().afs_chroma_batch
        with torch.cuda.device(d_audio.device):
            nat.check(entry(
                self._h, nat.ptr(d_audio), offs.ctypes.data_as(nat._i64p), n, 1 if center else 0, 1 if normalize else 0,
                nat.ptr(d_out), None if oo is None else oo.ctypes.data_as(nat._i64p),
                nat.AFS_F64 if d_out.dtype == torch.float64 else nat.AFS_F32, _COMPUTE[compute], nat.stream_ptr()))
        return d_out, foffs

    def stft(self, d_audio, offsets, center=True, compute="fp64"):
        """create_stft (chroma.py:44-65) for a batch: returns (d_spec, frame_offsets) with d_spec a complex
        (total_frames, 2049) device tensor (complex128 for compute="fp64", complex64 for "fp32")."""
        assert d_audio.is_cuda and d_audio.dtype == torch.float32
        offs = np.ascontiguousarray(offsets, dtype=np.int64)
        n = offs.shape[0] - 1
        frames = np.array([self.num_frames(offs[k + 1] - offs[k], center) for k in range(n)], dtype=np.int64)
        foffs = np.concatenate(([0], np.cumsum(frames))).astype(np.int64)
        f64 = compute in ("fp64", "f64")
        d_spec = torch.empty((int(foffs[-1]), 1 + self.n_fft // 2), dtype=torch.complex128 if f64 else torch.complex64,
                             device=d_audio.device)
        with torch.cuda.device(d_audio.device):
            nat.check(nat.lib().afs_stft_batch(self._h, nat.ptr(d_audio), offs.ctypes.data_as(nat._i64p), n, 1 if center else 0,
                                               nat.ptr(d_spec), None, nat.AFS_F64 if f64 else nat.AFS_F32, nat.stream_ptr()))
        return d_spec, foffs


_default_plan = None


def default_plan():
    global _default_plan
    if _default_plan is None:
        _default_plan = ChromaPlan(fft_len, hop_size)
    return _default_plan


def chroma_batch(tracks, center=True, normalize=True, compute="tc"):
    """Many tracks in one launch: list of 1-D sample arrays -> list of (12, frames) float64 arrays."""
    plan = default_plan()
    # int16 tracks (raw PCM, see load_wav_pcm16) travel and are staged as int16; anything else as float32
    pcm = len(tracks) > 0 and all(np.asarray(t).dtype == np.int16 for t in tracks)
    sdt, quantum = (np.int16, 8) if pcm else (np.float32, 4)
    # every track starts on a 16-byte boundary (TMA bulk staging of whole frames); the zero padding at a
    # track's end can at most add one frame, which is dropped below (frames are counted on the true length)
    lens = [(len(t) + quantum - 1) // quantum * quantum for t in tracks]
    offs = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
    flat = np.zeros(int(offs[-1]), dtype=sdt)
    for k, t in enumerate(tracks):
        flat[offs[k] : offs[k] + len(t)] = np.asarray(t, dtype=sdt)
    d_audio = torch.from_numpy(flat).to(plan.device)
    outs = []
    d_out, foffs = plan.run(d_audio, offs, center=center, normalize=normalize, out_dtype=torch.float64, compute=compute)
    host = d_out.cpu().numpy()
    for k, t in enumerate(tracks):
        nfr = plan.num_frames(len(t), center)
        blk = host[12 * foffs[k] : 12 * foffs[k + 1]].reshape(12, -1)
        outs.append(np.ascontiguousarray(blk[:, :nfr]))
    return outs


def wav_samples_to_chroma(wav, normalize=True, compute="tc"):
    """create_chroma(create_stft(wav)) (chroma.py:31-33) -> (12, M) float64."""
    return chroma_batch([np.asarray(wav)], center=True, normalize=normalize, compute=compute)[0]


def load_wav(path_to_wav):
    """Stand-in for librosa.load(path) defaults (mono float32 at 22 050 Hz) for PCM16 WAVs; CPU only."""
    with wave.open(path_to_wav, "rb") as w:
        rate, nch, width = w.getframerate(), w.getnchannels(), w.getsampwidth()
        raw = w.readframes(w.getnframes())
    if width != 2:
        raise ValueError("only PCM16 WAV files are supported")
    x = np.frombuffer(raw, dtype="<i2").astype(np.float32) / np.float32(32768.0)
    if nch > 1:
        x = x.reshape(-1, nch).mean(axis=1, dtype=np.float32)
    return np.ascontiguousarray(x), rate


def load_wav_pcm16(path_to_wav):
    """The raw int16 samples of a mono PCM16 WAV (None for other layouts) and its rate: what K1 takes directly."""
    with wave.open(path_to_wav, "rb") as w:
        rate, nch, width = w.getframerate(), w.getnchannels(), w.getsampwidth()
        if width != 2 or nch != 1:
            return None, rate
        raw = w.readframes(w.getnframes())
    return np.frombuffer(raw, dtype="<i2").astype(np.int16), rate


def wav_to_chroma(path_to_wav):
    """chroma.py:25-33.  Mono PCM16 files go to the GPU as int16 (the 1/32768 scaling of librosa.load
    happens in the kernel, bit-identically); other layouts are mixed down on the CPU first."""
    pcm, wav_fs = load_wav_pcm16(path_to_wav)
    if pcm is None:
        pcm, wav_fs = load_wav(path_to_wav)
    assert(wav_fs == 22050)
    return wav_samples_to_chroma(pcm)


def wav_to_chroma_col(wav_buf, compute="tc"):
    """chroma.py:35-42: one un-padded 4096-sample frame -> (12,)."""
    assert(len(wav_buf) == fft_len)
    return chroma_batch([np.asarray(wav_buf)], center=False, compute=compute)[0][:, 0]


class Stft(np.ndarray):
    """What create_stft returns: the reference's complex (2049, M) spectrum, plus the samples it came from so that
    create_chroma can run the fused path instead of re-reading the spectrum."""

    def __new__(cls, spectrum, wav):
        obj = np.asarray(spectrum).view(cls)
        obj.wav = wav
        return obj

    def __array_finalize__(self, obj):
        self.wav = getattr(obj, "wav", None) if obj is not None and getattr(obj, "shape", None) == self.shape else None


def create_stft(wav, compute="fp64"):
    """chroma.py:44-65: zero-pad fft_len/2 on the left, frames of fft_len at hop_size, Hann window, rfft.
    Returns the complex (1 + fft_len/2, num_hops) array (complex128), computed by the float64 kernel."""
    plan = default_plan()
    x = np.ascontiguousarray(np.asarray(wav), dtype=np.float32)
    pad = (len(x) + 3) // 4 * 4
    flat = np.zeros(pad, dtype=np.float32)
    flat[: len(x)] = x
    d_audio = torch.from_numpy(flat).to(plan.device)
    d_spec, foffs = plan.stft(d_audio, np.array([0, len(x)], dtype=np.int64), center=True, compute=compute)
    spec = d_spec.cpu().numpy().astype(np.complex128, copy=False).T
    return Stft(np.ascontiguousarray(spec), x)


def create_chroma(ft, normalize=True):
    """chroma.py:67-75: |ft|^2 through the (12, 2049) filterbank, L2-normalised per frame.  A spectrum that
    came from create_stft is recomputed from its samples by the fused kernel; any other (2049, M) complex
    array goes through the filterbank + normalisation on the device."""
    if isinstance(ft, Stft) and ft.wav is not None:
        return wav_samples_to_chroma(ft.wav, normalize=normalize)
    plan = default_plan()
    spec = torch.from_numpy(np.ascontiguousarray(np.asarray(ft), dtype=np.complex128)).to(plan.device)
    assert spec.dim() == 2 and spec.shape[0] == 1 + fft_len // 2, "expected a (2049, M) spectrum"
    power = spec.real ** 2 + spec.imag ** 2                                    # chroma.py:68
    fbank = torch.from_numpy(plan.filterbank).to(plan.device)
    raw = fbank @ power                                                        # chroma.py:70
    if normalize:
        length = torch.sqrt((raw * raw).sum(dim=0, keepdim=True))             # chroma.py:74, librosa.util.normalize(norm=2)
        length = torch.where(length < np.finfo(np.float64).tiny, torch.ones_like(length), length)
        raw = raw / length
    return raw.cpu().numpy()


def chroma_to_diff(chroma):
    """chroma.py:88-90 on an existing chromagram: half-wave-rectified time difference."""
    return np.clip(np.diff(chroma), 0, float('inf'))


def wav_to_chroma_diff(path_to_wav):
    """chroma.py:77-90."""
    return chroma_to_diff(wav_to_chroma(path_to_wav))

"""CPU emulation of the tensor-core chroma pipeline's operand roundings (numpy), to choose the MMA operand formats
before writing kernels: 4096-point real DFT as 64 x 64 two-stage product, operands rounded to fp16 / bf16 / split terms,
products accumulated exactly (float64 stands in for the fp32 TMEM accumulator), compared with the float64 oracle.
python tools/emul_chroma_split.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import afs_oracle as orc
from oracle import librosa_restated as lr

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def bf16_trunc(x):
    u = np.asarray(x, np.float32).view(np.uint32) & np.uint32(0xFFFF0000)
    return u.view(np.float32)


def bf16_rn(x):
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    u = ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    return u.view(np.float32)


def fp16(x):
    return np.asarray(x, np.float32).astype(np.float16).astype(np.float32)


def terms(x, fmt, n):
    """x as a sum of n terms in the format (list of float32 arrays)"""
    out = []
    r = np.asarray(x, np.float32)
    for _ in range(n):
        t = fmt(r)
        out.append(t)
        r = (r - t).astype(np.float32)
    return out


def frames_of(x, hop=2048, n_fft=4096):
    x = np.asarray(x, np.float32)
    m = 1 + len(x) // hop if len(x) else 0                     # chroma.py:47-52 framing (see oracle)
    xp = np.concatenate([np.zeros(n_fft // 2, np.float32), x, np.zeros(n_fft + hop, np.float32)])
    return np.stack([xp[i * hop: i * hop + n_fft] for i in range(m)]) if m else np.zeros((0, n_fft), np.float32)


def pipeline(x, mode):
    fr = frames_of(x)
    want = orc.wav_samples_to_chroma(x)
    m = want.shape[1]
    fr = fr[:m]
    win = np.hanning(4096).astype(np.float32)
    xw = (fr * win).astype(np.float32)                           # [m][4096], n = 64 n1 + n2
    A = xw.reshape(m, 64, 64)                                    # [m][n1][n2]
    n1 = np.arange(64)
    F = np.exp(-2j * np.pi * np.outer(n1, np.arange(33)) / 64)   # [n1][k1]
    G = np.exp(-2j * np.pi * np.outer(np.arange(64), np.arange(64)) / 64)   # [n2][k2]
    tw = np.exp(-2j * np.pi * np.outer(np.arange(33), np.arange(64)) / 4096).astype(np.complex64)  # [k1][n2]
    fmt, nx, nf = mode["fmt"], mode["nx"], mode["nf"]
    nx2, nf2 = mode.get("nx2", nx), mode.get("nf2", nf)
    scale = mode.get("scale", False)
    if scale:
        mx = np.abs(xw).max(axis=1)
        e = np.where(mx > 0, np.floor(np.log2(np.maximum(mx, 1e-300))) + 1, 0)
        s1 = (2.0 ** -e).astype(np.float32)[:, None, None]
    else:
        s1 = np.float32(1.0)
    As = (A * s1).astype(np.float32)
    Fr, Fi = F.real.astype(np.float32), F.imag.astype(np.float32)
    At = terms(As, fmt, nx)
    Frt, Fit = terms(Fr, fmt, nf), terms(Fi, fmt, nf)
    # stage 1: Y[m][k1][n2] = sum_n1 A[m][n1][n2] F[n1][k1]; cross terms of order i + j < max(nx, nf)
    Yr = np.zeros((m, 33, 64)); Yi = np.zeros((m, 33, 64))
    for i, a in enumerate(At):
        for j in range(nf):
            if i + j >= max(nx, nf):
                continue
            Yr += np.einsum("mab,ak->mkb", a.astype(np.float64), Frt[j].astype(np.float64))
            Yi += np.einsum("mab,ak->mkb", a.astype(np.float64), Fit[j].astype(np.float64))
    Y = (Yr.astype(np.float32) + 1j * Yi.astype(np.float32)).astype(np.complex64)
    # twiddle in fp32, scale for stage 2
    Yp = Y * tw[None]
    s2 = np.float32(1.0 / 64) if scale else np.float32(1.0)
    Ypr, Ypi = (Yp.real * s2).astype(np.float32), (Yp.imag * s2).astype(np.float32)
    Ypr_t, Ypi_t = terms(Ypr, fmt, nx2), terms(Ypi, fmt, nx2)
    Gr_t, Gi_t = terms(G.real.astype(np.float32), fmt, nf2), terms(G.imag.astype(np.float32), fmt, nf2)
    Xr = np.zeros((m, 33, 64)); Xi = np.zeros((m, 33, 64))
    for i in range(nx2):
        for j in range(nf2):
            if i + j >= max(nx2, nf2):
                continue
            yr, yi = Ypr_t[i].astype(np.float64), Ypi_t[i].astype(np.float64)
            gr, gi = Gr_t[j].astype(np.float64), Gi_t[j].astype(np.float64)
            Xr += np.einsum("mkn,nq->mkq", yr, gr) - np.einsum("mkn,nq->mkq", yi, gi)
            Xi += np.einsum("mkn,nq->mkq", yr, gi) + np.einsum("mkn,nq->mkq", yi, gr)
    Xr = Xr.astype(np.float32); Xi = Xi.astype(np.float32)
    P = (Xr * Xr + Xi * Xi).astype(np.float32)                    # [m][k1][k2] -> bin k1 + 64 k2
    unscale = (1.0 / (s1.reshape(-1) * s2)) ** 2 if scale else np.ones(m)
    power = np.zeros((m, 2049), np.float32)
    for k1 in range(33):
        for k2 in range(64):
            k = k1 + 64 * k2
            power[:, k if k <= 2048 else 4096 - k] = P[:, k1, k2]
    late = mode.get("pfmt", bf16_trunc) is fp16           # fp16 power planes keep the frame scale until after the filterbank
    if not late:
        power = (power * unscale[:, None].astype(np.float32)).astype(np.float32)
    fb = lr.filters_chroma(22050, 4096).astype(np.float64)[:, :2049]
    pfmt, pn, wn = mode.get("pfmt", bf16_trunc), mode.get("pn", 2), mode.get("wn", 2)
    Pt = sum(t.astype(np.float64) for t in terms(power, pfmt, pn))
    Wt = sum(t.astype(np.float64) for t in terms(fb.astype(np.float32), pfmt, wn))
    raw = (Wt @ Pt.T).astype(np.float32).astype(np.float64)
    if late:
        raw = raw * unscale[None, :]
    nrm = np.sqrt((raw * raw).sum(axis=0, keepdims=True))
    nrm[nrm < np.finfo(np.float32).tiny] = 1.0
    got = raw / nrm
    return float(np.abs(got - want).max())


def signals():
    aud = np.load(os.path.join(GOLD, "audio_15s.npz"))
    out = {}
    for tag in ("ref", "live"):
        out["golden_" + tag] = (aud[tag + "_i16"].astype(np.float32) / np.float32(32768.0)).mean(axis=1, dtype=np.float32)
    rng = np.random.default_rng(3)
    t = np.arange(3 * 22050) / 22050
    out["sine440"] = (0.5 * np.sin(2 * np.pi * 440 * t)).astype(np.float32)
    out["noise.1"] = (0.1 * rng.standard_normal(20001)).astype(np.float32)
    out["noise3e4"] = (3e4 * rng.standard_normal(30000)).astype(np.float32)
    out["noise1e-6"] = (1e-6 * rng.standard_normal(30000)).astype(np.float32)
    out["dc+tone"] = (0.3 + 1e-3 * np.sin(2 * np.pi * 440 * np.arange(40000) / 22050)).astype(np.float32)
    out["two_tones"] = (0.5 * np.sin(2 * np.pi * 453.08 * t) + 0.5 * np.sin(2 * np.pi * 1244.5 * t)).astype(np.float32)
    imp = np.zeros(30000, np.float32); imp[5000] = 1.0; imp[17000] = -0.7
    out["impulses"] = imp
    out["chirp"] = (0.4 * np.sin(2 * np.pi * (100 * t + 1500 * t * t))).astype(np.float32)
    out["loud_edge"] = np.concatenate([0.9 * rng.standard_normal(300), 1e-3 * np.sin(2 * np.pi * 880 * t[:20000])]).astype(np.float32)
    return out


MODES = {
    "fp16 x1 F2": dict(fmt=fp16, nx=1, nf=2, scale=True),
    "fp16 x2 F2 (3 MMA)": dict(fmt=fp16, nx=2, nf=2, scale=True),
    "fp16 s1: x1F1, s2: x2F2": dict(fmt=fp16, nx=1, nf=1, nx2=2, nf2=2, scale=True),
    "fp16 s1: x2F2, s2: x1F1": dict(fmt=fp16, nx=2, nf=2, nx2=1, nf2=1, scale=True),
    "fp16 s1: x1F2, s2: x1F2": dict(fmt=fp16, nx=1, nf=2, nx2=1, nf2=2, scale=True),
    "fp16 s1: x2F1, s2: x1F2": dict(fmt=fp16, nx=2, nf=1, nx2=1, nf2=2, scale=True),
    "fp16 s1: x1F2, s2: x2F1": dict(fmt=fp16, nx=1, nf=2, nx2=2, nf2=1, scale=True),
}

if __name__ == "__main__":
    sig = signals()
    for name, mode in MODES.items():
        errs = {k: pipeline(x, mode) for k, x in sig.items()}
        print("%-34s worst %.2e | " % (name, max(errs.values())) + " ".join("%s %.1e" % (k, v) for k, v in errs.items()), flush=True)

"""Seeded synthetic frame pairs for the direct tracker.

A procedural texture (sum of 24 sinusoids) is evaluated analytically: frame 1 samples
it at the pixel grid, frame 2 samples it through the plane-induced homography of a known
SE3 motion (all scene points on Z = 1, the reference's mono depth assumption,
/root/reference/src/Tracker.cpp:1317,1354).  No dataset, no network.

`render_pair` runs in numpy (float64); `render_batch_torch` does the same arithmetic with
torch on any device (used to fill large batches on the GPU box).
"""
import math

import numpy as np

# level-0 pinhole calibrations of the BASELINE configs (fx, fy, cx, cy)
CALIB = {
    # /root/reference/calibration/calibrationTUM.xml:20 (TUM-RGBD, 640x480)
    "tum": (640, 480, 525.0, 525.0, 319.5, 239.5),
    # /root/reference/calibration/calibrationEUROC.xml:20 used as a pinhole at 752x480
    "euroc": (752, 480, 458.654, 457.296, 367.215, 248.375),
    # TUM-mono normalised intrinsics scaled to 1280x1024
    "tum_mono": (1280, 1024, 685.72, 685.64, 630.86, 511.92),
    # config 4: arbitrary fixed pinhole for a 3840x2160 frame
    "uhd": (3840, 2160, 2057.0, 2057.0, 1919.5, 1079.5),
    # small frames for fast tests (all dims divisible by 16)
    "tiny": (64, 48, 52.5, 52.5, 31.5, 23.5),
    "small": (160, 128, 131.25, 140.0, 79.5, 63.5),
}

N_WAVES = 24


def texture_params(seed):
    rng = np.random.default_rng(seed)
    ang = rng.uniform(0.0, 2.0 * math.pi, N_WAVES)
    k = rng.uniform(0.02, 0.25, N_WAVES)
    amp = rng.uniform(4.0, 22.0, N_WAVES)
    ph = rng.uniform(0.0, 2.0 * math.pi, N_WAVES)
    kx, ky = k * np.cos(ang), k * np.sin(ang)
    sigma = math.sqrt(float(np.sum(amp * amp)) / 2.0)
    scale = 127.5 / (3.2 * sigma)
    return kx, ky, amp * scale, ph


def motion(seed, rot=5e-3, trans=5e-3):
    """Seeded small motion: omega ~ U(-rot,rot)^3 rad, t ~ U(-trans,trans)^3."""
    rng = np.random.default_rng(1_000_003 + seed)
    return rng.uniform(-rot, rot, 3), rng.uniform(-trans, trans, 3)


def rotation_matrix(omega):
    th = float(np.linalg.norm(omega))
    O = np.array([[0, -omega[2], omega[1]], [omega[2], 0, -omega[0]], [-omega[1], omega[0], 0]])
    if th < 1e-12:
        return np.eye(3) + O
    return np.eye(3) + math.sin(th) / th * O + (1 - math.cos(th)) / (th * th) * (O @ O)


def homography_inv(calib, omega, t):
    """Maps frame-2 pixels to frame-1 pixels: K (R + t n^T)^-1 K^-1, n = e_z."""
    _, _, fx, fy, cx, cy = calib
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]])
    M = rotation_matrix(np.asarray(omega, float)).copy()
    M[:, 2] += np.asarray(t, float)
    return K @ np.linalg.inv(M) @ np.linalg.inv(K)


def _eval_texture_np(px, py, tp):
    kx, ky, amp, ph = tp
    v = np.zeros(px.shape, np.float64)
    for i in range(N_WAVES):
        v += amp[i] * np.sin(kx[i] * px + ky[i] * py + ph[i])
    return np.clip(np.rint(127.5 + v), 0, 255).astype(np.uint8)


def render_frame(calib, tex_seed, Hinv=None):
    w, h = calib[0], calib[1]
    tp = texture_params(tex_seed)
    py, px = np.mgrid[0:h, 0:w].astype(np.float64)
    if Hinv is not None:
        d = Hinv[2, 0] * px + Hinv[2, 1] * py + Hinv[2, 2]
        qx = (Hinv[0, 0] * px + Hinv[0, 1] * py + Hinv[0, 2]) / d
        qy = (Hinv[1, 0] * px + Hinv[1, 1] * py + Hinv[1, 2]) / d
        px, py = qx, qy
    return _eval_texture_np(px, py, tp)


def render_pair(calib_name, seed, rot=5e-3, trans=5e-3):
    """Returns (prev u8 HxW, cur u8 HxW, omega, t)."""
    calib = CALIB[calib_name] if isinstance(calib_name, str) else calib_name
    omega, t = motion(seed, rot, trans)
    prev = render_frame(calib, seed)
    cur = render_frame(calib, seed, homography_inv(calib, omega, t))
    return prev, cur, omega, t


def render_sequence(calib_name, seed, n_frames, rot=1e-3, trans=1e-3):
    """Frames along a smooth SE3 path (config 1): frame i is the texture seen after i small
    steps of one seeded motion, so consecutive frames differ by (omega, t)."""
    calib = CALIB[calib_name] if isinstance(calib_name, str) else calib_name
    omega, t = motion(seed, rot, trans)
    frames = []
    for i in range(n_frames):
        frames.append(render_frame(calib, seed, None if i == 0 else
                                   homography_inv(calib, omega * i, t * i)))
    return frames, omega, t


def render_batch_torch(calib_name, seeds, device, rot=5e-3, trans=5e-3, dtype=None):
    """torch version of render_pair for many seeds.  Returns two u8 tensors [B,H,W] on
    `device` (prev, cur).  Same formula as the numpy path (values may differ from numpy in
    the last rounding of sin(); callers feed the same bytes to every implementation)."""
    import torch

    calib = CALIB[calib_name] if isinstance(calib_name, str) else calib_name
    w, h = calib[0], calib[1]
    dtype = dtype or torch.float32
    ys, xs = torch.meshgrid(torch.arange(h, device=device, dtype=dtype),
                            torch.arange(w, device=device, dtype=dtype), indexing="ij")
    prev = torch.empty((len(seeds), h, w), dtype=torch.uint8, device=device)
    cur = torch.empty_like(prev)

    def ev(px, py, tp):
        kx, ky, amp, ph = [torch.as_tensor(a, device=device, dtype=dtype) for a in tp]
        v = torch.zeros_like(px)
        for i in range(N_WAVES):
            v += amp[i] * torch.sin(kx[i] * px + ky[i] * py + ph[i])
        return torch.clamp(torch.round(127.5 + v), 0, 255).to(torch.uint8)

    for b, seed in enumerate(seeds):
        tp = texture_params(seed)
        omega, t = motion(seed, rot, trans)
        Hi = homography_inv(calib, omega, t)
        prev[b] = ev(xs, ys, tp)
        d = Hi[2, 0] * xs + Hi[2, 1] * ys + Hi[2, 2]
        qx = (Hi[0, 0] * xs + Hi[0, 1] * ys + Hi[0, 2]) / d
        qy = (Hi[1, 0] * xs + Hi[1, 1] * ys + Hi[1, 2]) / d
        cur[b] = ev(qx, qy, tp)
    return prev, cur


def render_pairs_torch(calib_name, seeds, device, rot=5e-3, trans=5e-3, chunk=32):
    """render_batch_torch vectorised over chunks of seeds (float32): fills thousands of pairs
    in a second on a GPU.  Returns two u8 tensors [B,H,W] on `device` (prev, cur)."""
    import torch

    calib = CALIB[calib_name] if isinstance(calib_name, str) else calib_name
    w, h = calib[0], calib[1]
    B = len(seeds)
    prev = torch.empty((B, h, w), dtype=torch.uint8, device=device)
    cur = torch.empty_like(prev)
    ys, xs = torch.meshgrid(torch.arange(h, device=device, dtype=torch.float32),
                            torch.arange(w, device=device, dtype=torch.float32), indexing="ij")

    def ev(px, py, kx, ky, amp, ph):
        v = torch.zeros_like(px)
        for k in range(N_WAVES):
            v += amp[:, k] * torch.sin(kx[:, k] * px + ky[:, k] * py + ph[:, k])
        return torch.clamp(torch.round(127.5 + v), 0, 255).to(torch.uint8)

    for b0 in range(0, B, chunk):
        bs = seeds[b0:b0 + chunk]
        tps = [texture_params(s) for s in bs]
        kx, ky, amp, ph = [torch.tensor(np.stack([tp[j] for tp in tps]), device=device,
                                        dtype=torch.float32)[:, :, None, None] for j in range(4)]
        Hi = torch.tensor(np.stack([homography_inv(calib, *motion(s, rot, trans)) for s in bs]),
                          device=device, dtype=torch.float32)
        one = torch.ones((len(bs), 1, 1), device=device)
        prev[b0:b0 + len(bs)] = ev(xs * one, ys * one, kx, ky, amp, ph)
        d = Hi[:, 2, 0, None, None] * xs + Hi[:, 2, 1, None, None] * ys + Hi[:, 2, 2, None, None]
        qx = (Hi[:, 0, 0, None, None] * xs + Hi[:, 0, 1, None, None] * ys +
              Hi[:, 0, 2, None, None]) / d
        qy = (Hi[:, 1, 0, None, None] * xs + Hi[:, 1, 1, None, None] * ys +
              Hi[:, 1, 2, None, None]) / d
        cur[b0:b0 + len(bs)] = ev(qx, qy, kx, ky, amp, ph)
    return prev, cur

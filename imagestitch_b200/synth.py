"""Deterministic synthetic panorama inputs (``synth_v1``).

A smooth random scene on the viewing cylinder is sampled by N rotated pinhole cameras; every source
pixel looks up the scene at its own forward projection, so neighbouring images agree on the overlap up
to per-image gain and integer noise -- the situation the seam finder and the blender are built for.

Inputs are inputs: the generator is shared by the CUDA path, the oracle and the CPU baseline of one
run, it is not part of any parity claim.  numpy is used for small cases, torch (any device) for the
bench-size images.
"""
from __future__ import annotations

import math

import numpy as np

N_TERMS = 24


def rot_y(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], np.float64)


def rot_x(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]], np.float64)


def strip_cameras(n, w, h, f_over_w=1.2, overlap=0.25, pitch_amp=0.002, grid_rows=1):
    """K (n,3,3) f32, R (n,3,3) f32, scale.  Strip: yaw_i = (i-(n-1)/2)*delta, pitch_i = amp*sin(1.7 i).
    grid_rows > 1 lays the cameras out as a grid_rows x (n/grid_rows) mosaic (config 5)."""
    f = f_over_w * w
    K = np.array([[f, 0, w / 2.0], [0, f, h / 2.0], [0, 0, 1]], np.float64)
    delta = (1.0 - overlap) * 2.0 * math.atan(w / (2.0 * f))
    vfov = 2.0 * math.atan(h / (2.0 * f))
    per_row = n // grid_rows
    Ks, Rs = [], []
    for i in range(n):
        r, c = divmod(i, per_row)
        yaw = (c - (per_row - 1) / 2.0) * delta
        pitch = pitch_amp * math.sin(1.7 * i)
        if grid_rows > 1:
            pitch += (r - (grid_rows - 1) / 2.0) * (1.0 - overlap) * vfov
        Rs.append((rot_y(yaw) @ rot_x(pitch)).astype(np.float32))
        Ks.append(K.astype(np.float32))
    return np.stack(Ks), np.stack(Rs), float(f)


def scene_coefficients(seed=12345):
    rng = np.random.default_rng(seed)
    omega = rng.uniform(2.0, 400.0, N_TERMS)
    nu = rng.uniform(2.0, 400.0, N_TERMS)
    amp = rng.uniform(0.2, 1.0, (N_TERMS, 3))
    amp = amp / amp.sum(axis=0, keepdims=True) * 200.0      # random phases rarely align: ~[30, 225], clipped
    phi = rng.uniform(0, 2 * math.pi, (N_TERMS, 3))
    psi = rng.uniform(0, 2 * math.pi, (N_TERMS, 3))
    return omega, nu, amp, phi, psi


def make_image(i, w, h, K, R, seed=12345, device=None, rows_per_chunk=512):
    """uint8 (h, w, 3) BGR source image of camera i.  device=None -> numpy; else a torch device string."""
    omega, nu, amp, phi, psi = scene_coefficients(seed)
    gain = 1.0 + 0.03 * math.sin(i)
    M = (np.asarray(R, np.float64) @ np.linalg.inv(np.asarray(K, np.float64)))
    if device is None:
        out = np.empty((h, w, 3), np.uint8)
        rng = np.random.default_rng(777 + i)
        xs = np.arange(w, dtype=np.float64)[None, :]
        for y0 in range(0, h, rows_per_chunk):
            y1 = min(h, y0 + rows_per_chunk)
            ys = np.arange(y0, y1, dtype=np.float64)[:, None]
            x_ = M[0, 0] * xs + M[0, 1] * ys + M[0, 2]
            y_ = M[1, 0] * xs + M[1, 1] * ys + M[1, 2]
            z_ = M[2, 0] * xs + M[2, 1] * ys + M[2, 2]
            th = np.arctan2(x_, z_)
            t = y_ / np.hypot(x_, z_)
            val = np.full((y1 - y0, w, 3), 128.0)
            for k in range(N_TERMS):
                for c in range(3):
                    val[:, :, c] += amp[k, c] * np.sin(omega[k] * th + phi[k, c]) * np.sin(nu[k] * t + psi[k, c])
            val = val * gain + rng.integers(-2, 3, val.shape)
            out[y0:y1] = np.clip(np.rint(val), 0, 255).astype(np.uint8)
        return out
    import torch

    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(777 + i)
    out = torch.empty((h, w, 3), dtype=torch.uint8, device=dev)
    Mt = torch.tensor(M, dtype=torch.float64)
    xs = torch.arange(w, dtype=torch.float32, device=dev)[None, :]
    om = torch.tensor(omega, dtype=torch.float32, device=dev)
    nv = torch.tensor(nu, dtype=torch.float32, device=dev)
    am = torch.tensor(amp, dtype=torch.float32, device=dev)
    ph = torch.tensor(phi, dtype=torch.float32, device=dev)
    ps = torch.tensor(psi, dtype=torch.float32, device=dev)
    for y0 in range(0, h, rows_per_chunk):
        y1 = min(h, y0 + rows_per_chunk)
        ys = torch.arange(y0, y1, dtype=torch.float32, device=dev)[:, None]
        x_ = float(Mt[0, 0]) * xs + float(Mt[0, 1]) * ys + float(Mt[0, 2])
        y_ = float(Mt[1, 0]) * xs + float(Mt[1, 1]) * ys + float(Mt[1, 2])
        z_ = float(Mt[2, 0]) * xs + float(Mt[2, 1]) * ys + float(Mt[2, 2])
        th = torch.atan2(x_, z_)
        t = y_ / torch.hypot(x_, z_)
        val = torch.full((y1 - y0, w, 3), 128.0, dtype=torch.float32, device=dev)
        for k in range(N_TERMS):
            a = torch.sin(om[k] * th[..., None] + ph[k])
            b = torch.sin(nv[k] * t[..., None] + ps[k])
            val += am[k] * a * b
        noise = torch.randint(-2, 3, val.shape, generator=g, device=dev, dtype=torch.int32).to(torch.float32)
        val = val * gain + noise
        out[y0:y1] = torch.clamp(torch.round(val), 0, 255).to(torch.uint8)
    return out


def make_panorama_inputs(n, w, h, f_over_w=1.2, overlap=0.25, grid_rows=1, seed=12345, device=None):
    Ks, Rs, scale = strip_cameras(n, w, h, f_over_w, overlap, grid_rows=grid_rows)
    imgs = [make_image(i, w, h, Ks[i], Rs[i], seed, device) for i in range(n)]
    return imgs, Ks, Rs, scale

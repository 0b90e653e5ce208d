"""Synthetic DAVIS-shaped videos: T x 480 x 854 RGB frames with K moving textured
ellipses over a smooth random background, plus per-frame object-id labels.  Pixel/label ranges match
what the reference's datasets hand to the model (vos_dataset.py:233,276-279: RGB fp32 in [0,1],
labels as float ids).  Deterministic in `seed`."""
import numpy as np


def _smooth_noise(rs, h, w, c, cell):
    gh, gw = h // cell + 2, w // cell + 2
    g = rs.rand(gh, gw, c).astype(np.float32)
    ys = np.linspace(0, gh - 1.001, h, dtype=np.float32)
    xs = np.linspace(0, gw - 1.001, w, dtype=np.float32)
    y0, x0 = ys.astype(int), xs.astype(int)
    fy, fx = (ys - y0)[:, None, None], (xs - x0)[None, :, None]
    a = g[y0][:, x0] * (1 - fx) + g[y0][:, x0 + 1] * fx
    b = g[y0 + 1][:, x0] * (1 - fx) + g[y0 + 1][:, x0 + 1] * fx
    return a * (1 - fy) + b * fy


def make_video(seed, num_frames=10, height=480, width=854, num_objects=1):
    """-> frames uint8 [T,H,W,3], labels uint8 [T,H,W] (0 background, 1..K objects)."""
    rs = np.random.RandomState(seed)
    bg = _smooth_noise(rs, height, width, 3, 48) * 0.6 + _smooth_noise(rs, height, width, 3, 8) * 0.2 + 0.1
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float32)
    objs = []
    for k in range(num_objects):
        objs.append(dict(cx=rs.uniform(0.25, 0.75) * width, cy=rs.uniform(0.3, 0.7) * height,
                         rx=rs.uniform(0.08, 0.16) * width, ry=rs.uniform(0.12, 0.22) * height,
                         vx=rs.uniform(-6, 6), vy=rs.uniform(-3, 3), color=rs.uniform(0.2, 1.0, 3).astype(np.float32),
                         tex=_smooth_noise(rs, height, width, 3, 6)))
    frames = np.zeros((num_frames, height, width, 3), np.uint8)
    labels = np.zeros((num_frames, height, width), np.uint8)
    for t in range(num_frames):
        img = bg + 0.03 * np.sin(0.05 * (xx + 3 * t))[..., None]
        lab = np.zeros((height, width), np.uint8)
        for k, o in enumerate(objs):
            cx = np.clip(o["cx"] + o["vx"] * t, 0.15 * width, 0.85 * width)
            cy = np.clip(o["cy"] + o["vy"] * t, 0.2 * height, 0.8 * height)
            inside = ((xx - cx) / o["rx"]) ** 2 + ((yy - cy) / o["ry"]) ** 2 <= 1.0
            img = np.where(inside[..., None], 0.55 * o["color"] + 0.45 * o["tex"], img)
            lab[inside] = k + 1
        frames[t] = np.clip(img * 255.0, 0, 255).astype(np.uint8)
        labels[t] = lab
    return frames, labels


def davis_val_shaped_set(num_videos=30, seed=1):
    """(num_frames, num_objects) per video with DAVIS-2017-val-like statistics (SURVEY.md §8d cfg 2/3:
    30 videos, T ~ U[34,104], K in 1..5, sum K ~ 61)."""
    rs = np.random.RandomState(seed)
    spec = []
    for v in range(num_videos):
        spec.append((int(rs.randint(34, 105)), int(rs.choice([1, 1, 2, 2, 2, 3, 3, 4, 5]))))
    return spec

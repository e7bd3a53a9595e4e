"""Replay export: one environment's trajectory out of the rollout buffers, drawn like the reference's ``render``.

The reference draws the CURRENT state of its single env with matplotlib (v0: red AI players, blue opponents, green
ball on a 105 x 68 pitch, gym_futbol/envs/futbol_env.py:253-277; v1: the pymunk debug view, envs_v1/futbol_env.py:
236-243).  A batched simulator produces ``obs[K, n, D]`` instead; this module cuts env ``i`` out of such a buffer
(``trajectory``) and draws the frames with PIL, which needs no display (``frames`` / ``save_gif``).  Host-side
convenience for debugging and demos: nothing here is on the step path, and nothing here touches the GPU beyond one
device-to-host copy of the selected env's rows.
"""
from __future__ import annotations

import numpy as np

PITCH = (105.0, 68.0)
_V1_AVG = (52.5, 34.0)
_V1_RANGE_BALL = (52.5, 34.0)
_V1_RANGE_PLAYER = (55.5, 34.0)


def trajectory(obs, env=0, variant="v0"):
    """``obs``: [K, n, D] (torch tensor on any device, or numpy).  Returns a dict of float64 numpy arrays:
    ``team_a`` [K, N, 2], ``team_b`` [K, N, 2], ``ball`` [K, 2] in pitch coordinates, and for v0 ``owner`` [K]
    (0..3 = ai_1, ai_2, opp_1, opp_2; 4 = nobody; the row of 10s in observation row 5, futbol_env.py:720-736)."""
    rows = obs[:, env]
    rows = rows.detach().cpu().numpy() if hasattr(rows, "detach") else np.asarray(rows)
    rows = rows.astype(np.float64)
    K, D = rows.shape
    if variant == "v0":
        if D != 30:
            raise ValueError("v0 observations have 30 values, got %d" % D)
        r = rows[:, :25].reshape(K, 5, 5)
        onehot = rows[:, 25:30]
        owner = np.where(onehot.max(axis=1) > 0, onehot.argmax(axis=1), 4)
        return {"team_a": r[:, 0:2, 0:2].copy(), "team_b": r[:, 2:4, 0:2].copy(), "ball": r[:, 4, 0:2].copy(), "owner": owner}
    if variant == "v1":
        if D < 12 or (D - 4) % 8 != 0:
            raise ValueError("v1 observations have 4 + 8 N values, got %d" % D)
        N = (D - 4) // 8
        avg = np.array(_V1_AVG)
        ball = rows[:, 0:2] * np.array(_V1_RANGE_BALL) + avg                     # envs_v1/futbol_env.py:154-180 undone
        pl = rows[:, 4:].reshape(K, 2 * N, 4)[:, :, 0:2] * np.array(_V1_RANGE_PLAYER) + avg
        return {"team_a": pl[:, :N].copy(), "team_b": pl[:, N:].copy(), "ball": ball}
    raise ValueError("variant must be 'v0' or 'v1'")


def frames(traj, scale=6, margin=5.0):
    """One PIL image per step: pitch outline, team A red, team B blue, ball green (the reference's colours)."""
    from PIL import Image, ImageDraw
    W, H = PITCH
    size = (int((W + 2 * margin) * scale), int((H + 2 * margin) * scale))

    def px(p):
        return ((p[0] + margin) * scale, (H + margin - p[1]) * scale)              # y up, like the matplotlib view

    def dot(d, p, radius, colour):
        x, y = px(p)
        d.ellipse((x - radius, y - radius, x + radius, y + radius), fill=colour)

    out = []
    for k in range(len(traj["ball"])):
        im = Image.new("RGB", size, (255, 255, 255))
        d = ImageDraw.Draw(im)
        x0, y0 = px((0.0, H))
        x1, y1 = px((W, 0.0))
        d.rectangle((x0, y0, x1, y1), outline=(0, 0, 0))
        d.line((px((W / 2, 0.0)), px((W / 2, H))), fill=(160, 160, 160))
        for p in traj["team_a"][k]:
            dot(d, p, 1.5 * scale, (220, 30, 30))
        for p in traj["team_b"][k]:
            dot(d, p, 1.5 * scale, (30, 60, 220))
        dot(d, traj["ball"][k], 1.0 * scale, (20, 160, 40))
        out.append(im)
    return out


def save_gif(path, traj, step_ms=100, **kw):
    """Writes the trajectory as an animated GIF (one frame per env step, 0.1 s of game time each)."""
    fr = frames(traj, **kw)
    if not fr:
        raise ValueError("empty trajectory")
    fr[0].save(path, save_all=True, append_images=fr[1:], duration=step_ms, loop=0)
    return len(fr)

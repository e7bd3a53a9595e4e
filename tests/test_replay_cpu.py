"""Replay export: env i cut out of a rollout buffer and drawn with PIL; pure host code."""
import numpy as np
import pytest


def test_v0_trajectory_from_a_rollout_buffer():
    from gym_futbol_b200.replay import trajectory
    K, n = 7, 3
    obs = np.zeros((K, n, 30), np.float32)
    for k in range(K):
        rows = np.zeros((5, 5), np.float32)
        rows[:, 0] = [10 + k, 20, 60, 70, 50 + k]
        rows[:, 1] = [30, 35, 40, 45, 34]
        obs[k, 1, :25] = rows.reshape(-1)
        obs[k, 1, 25 + (k % 5)] = 10.0
    obs[3, 1, 25:30] = 0.0                                  # the reference's row of zeros right after a reset
    t = trajectory(obs, env=1, variant="v0")
    assert t["team_a"].shape == (K, 2, 2) and t["team_b"].shape == (K, 2, 2) and t["ball"].shape == (K, 2)
    assert np.array_equal(t["team_a"][:, 0, 0], 10.0 + np.arange(K)) and np.array_equal(t["ball"][:, 0], 50.0 + np.arange(K))
    assert list(t["owner"]) == [0, 1, 2, 4, 4, 0, 1]
    with pytest.raises(ValueError):
        trajectory(obs[:, :, :20], variant="v0")


def test_v1_trajectory_undoes_the_normalisation():
    from gym_futbol_b200.replay import trajectory
    N, K = 5, 4
    ball = np.array([[52.5, 34.0], [0.0, 0.0], [105.0, 68.0], [26.25, 17.0]])
    players = np.random.default_rng(0).uniform([0, 0], [105, 68], size=(K, 2 * N, 2))
    obs = np.zeros((K, 1, 4 + 8 * N))
    obs[:, 0, 0:2] = (ball - [52.5, 34.0]) / [52.5, 34.0]
    obs[:, 0, 4:] = np.concatenate([(players - [52.5, 34.0]) / [55.5, 34.0], np.zeros((K, 2 * N, 2))], axis=2).reshape(K, -1)
    t = trajectory(obs, env=0, variant="v1")
    assert np.allclose(t["ball"], ball, atol=1e-12) and np.allclose(t["team_a"], players[:, :N], atol=1e-12)
    assert np.allclose(t["team_b"], players[:, N:], atol=1e-12)


def test_frames_and_gif(tmp_path):
    from PIL import Image
    from gym_futbol_b200.replay import frames, save_gif, trajectory
    K = 5
    obs = np.zeros((K, 1, 30), np.float32)
    obs[:, 0, 0] = 43.5; obs[:, 0, 1] = 39.0; obs[:, 0, 20] = np.linspace(52.5, 80.0, K); obs[:, 0, 21] = 34.0
    t = trajectory(obs, variant="v0")
    fr = frames(t, scale=4)
    assert len(fr) == K and fr[0].size == (460, 312)
    # the ball (green) moves to the right between the first and the last frame
    def green_x(im):
        a = np.asarray(im).astype(int)
        ys, xs = np.nonzero((a[:, :, 1] > 120) & (a[:, :, 0] < 80) & (a[:, :, 2] < 80))
        return xs.mean()
    assert green_x(fr[-1]) > green_x(fr[0]) + 50
    path = tmp_path / "env0.gif"
    assert save_gif(str(path), t, scale=4) == K
    with Image.open(path) as im:
        assert getattr(im, "n_frames", 1) == K

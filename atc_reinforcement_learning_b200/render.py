"""Headless renderer: AtcGym.render(mode='rgb_array') of the reference
(/root/reference/envs/atc/atc_gym.py:367-552, themes.py) as one CUDA kernel — no pyglet, no X server.  Same layout
(600 px of sector width + 10 px padding, north up), same elements and colours; text labels are not drawn."""
import ctypes as C

import torch

from . import _native as nat

SCREEN_WIDTH, PADDING = 600, 10                 # atc_gym.py:373-374


def image_size(sector, screen_width=SCREEN_WIDTH):
    """(width, height) of the reference's viewer for this sector (atc_gym.py:376-382)."""
    bx0, by0, bx1, by1 = [float(v) for v in sector.bbox]
    scale = screen_width / (bx1 - bx0)
    return screen_width + 2 * PADDING, int((by1 - by0) * scale) + 2 * PADDING


def trail_from_original_state(original_state, env_index=0):
    """Positions the reference draws as the trail (atc_gym.py:440-447): of the history before the current step, every
    5th of the last 25.  original_state: [T, N, A, 10] of a rollout (slots 0, 1 = x, y).  Returns [K, 2] float64."""
    xy = original_state[:, env_index, :, :2].to(torch.float64)              # [T, A, 2]
    n = xy.shape[0]
    idx = [i for i in range(n - 5, max(0, n - 25), -1) if i % 5 == 0]
    if not idx:
        return xy.new_zeros((0, 2))
    return xy[idx].reshape(-1, 2).contiguous()


def render_rgb(env, env_index=0, trail_xy=None, screen_width=SCREEN_WIDTH):
    """RGB image (torch uint8 [H, W, 3] on the env's device) of env `env_index`: sector + its aircraft (+ trail dots)."""
    if not 0 <= env_index < env.num_envs:
        raise IndexError("env_index out of range")
    w, h = image_size(env.sector, screen_width)
    dev = env.device
    A, NA = env.num_aircraft, env.num_envs * env.num_aircraft
    st = env.state.reshape(5, NA)
    heads = st[:2, env_index * A:(env_index + 1) * A].t().contiguous()       # [A, 2] float64
    if trail_xy is None:
        trail = torch.zeros((0, 2), dtype=torch.float64, device=dev)
    else:
        trail = torch.as_tensor(trail_xy).to(device=dev, dtype=torch.float64).reshape(-1, 2).contiguous()
    img = torch.empty((h, w, 3), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = nat.lib().atc_render(env._handle, img.data_ptr(), w, h, trail.data_ptr() if trail.numel() else None,
                                  int(trail.shape[0]), heads.data_ptr(), int(heads.shape[0]),
                                  C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        nat.check(env._handle, rc)
    return img

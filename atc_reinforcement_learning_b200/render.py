"""Headless renderer: AtcGym.render(mode='rgb_array') of the reference
(/root/reference/envs/atc/atc_gym.py:367-552, themes.py) as one CUDA kernel — no pyglet, no X server.  Same layout
(600 px of sector width + 10 px padding, north up), same elements and colours; the text labels (reward lines, aircraft
name and "FL  speed") are stamped by a second small kernel with a 5 x 7 bitmap font at the reference's anchor points."""
import ctypes as C
import math

import numpy as np
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


_LABEL_DTYPE = np.dtype([('x', np.float32), ('y', np.float32), ('bold', np.int32), ('n', np.int32),
                         ('text', 'S%d' % nat.TEXT_MAX)])


def label_list(sector, aircraft, total_reward=None, last_reward=None, screen_width=SCREEN_WIDTH):
    """The reference's labels as (x, y, bold, text) in its screen coordinates (origin bottom-left, anchor top-left):
    the reward lines (atc_gym.py:404-412) and, per aircraft (x, y, h, v[, name]), name and "FL  speed" below it at
    rot_matrix(135) . (0, 8) from the symbol (atc_gym.py:436-443)."""
    bx0, by0 = float(sector.bbox[0]), float(sector.bbox[1])
    scale = screen_width / (float(sector.bbox[2]) - bx0)
    out = []
    if total_reward is not None:
        out.append((10.0, 40.0, 0, "Total reward: %.2f" % total_reward))
    if last_reward is not None:
        out.append((10.0, 25.0, 0, "Last reward: %.2f" % last_reward))
    dx, dy = 8.0 * math.sin(math.radians(135.0)), 8.0 * math.cos(math.radians(135.0))      # model.rot_matrix(135) . (0, 8)
    for k, ac in enumerate(aircraft):
        x, y, h, v = [float(t) for t in ac[:4]]
        name = ac[4] if len(ac) > 4 else "FLT%02d" % (k + 1)                                # atc_gym.py:347
        sx, sy = (x - bx0) * scale + PADDING + dx, (y - by0) * scale + PADDING + dy
        out.append((sx, sy, 1, name))
        out.append((sx, sy - 15.0, 1, "%d  %d" % (round(h / 100), round(v / 10))))          # atc_gym.py:437-441
    return out


def stamp_labels(img, labels, device):
    """Draws (x, y, bold, text) labels into the uint8 [H, W, 3] cuda image `img` (atc_render_text)."""
    if not labels:
        return img
    arr = np.zeros(len(labels), dtype=_LABEL_DTYPE)
    for i, (x, y, bold, text) in enumerate(labels):
        t = text.encode('ascii', 'replace')[:nat.TEXT_MAX]
        arr[i] = (x, y, int(bold), len(t), t)
    d = torch.from_numpy(arr.view(np.uint8).reshape(len(labels), -1)).to(device)
    with torch.cuda.device(device):
        rc = nat.lib().atc_render_text(img.data_ptr(), int(img.shape[1]), int(img.shape[0]), d.data_ptr(), len(labels),
                                       C.c_void_p(torch.cuda.current_stream(device).cuda_stream))
    if rc != 0:
        raise nat.AtcError('atc_render_text failed (%d)' % rc)
    return img


def render_rgb(env, env_index=0, trail_xy=None, screen_width=SCREEN_WIDTH, labels=True, last_reward=None):
    """RGB image (torch uint8 [H, W, 3] on the env's device) of env `env_index`: sector + its aircraft (+ trail dots)
    (+ the reference's text labels: total reward of the running episode, `last_reward` if given, aircraft name /
    flight level / speed)."""
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
    if labels:
        ac = st[:, env_index * A:(env_index + 1) * A].t().cpu().numpy()                      # [A, 5] x, y, h, phi, v
        items = label_list(env.sector, [(r[0], r[1], r[2], r[4]) for r in ac], float(env.ep_return[env_index]),
                           last_reward, screen_width)
        stamp_labels(img, items, dev)
    return img

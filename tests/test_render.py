"""Headless renderer: layout helpers on the CPU; the CUDA image against a numpy restatement of
the picture's geometry (reference: /root/reference/envs/atc/atc_gym.py:367-552, themes.py)."""
import numpy as np
import pytest
import torch

INACTIVE, ACTIVE, LINES, PLANE = (29, 69, 76), (84, 121, 128), (69, 173, 168), (157, 224, 173)      # themes.py x 256


def _sector():
    from atc_reinforcement_learning_b200 import LOWW
    from atc_reinforcement_learning_b200.sector import CompiledSector
    return CompiledSector(LOWW(random_entrypoints=True), cell=0.25)


def test_image_size_is_the_reference_viewers():
    from atc_reinforcement_learning_b200.render import image_size
    cs = _sector()
    bx0, by0, bx1, by1 = cs.bbox
    scale = 600 / (bx1 - bx0)                                        # atc_gym.py:376-380
    assert image_size(cs) == (600 + 20, int((by1 - by0) * scale) + 20)


def test_trail_selection_follows_the_reference_loop():
    from atc_reinforcement_learning_b200.render import trail_from_original_state
    T = 47
    raw = torch.zeros(T, 3, 1, 10)
    raw[:, 1, 0, 0] = torch.arange(T, dtype=torch.float32)           # x = step index
    got = trail_from_original_state(raw, env_index=1)[:, 0].tolist()
    n = T
    exp = [float(i) for i in range(n - 5, max(0, n - 25), -1) if i % 5 == 0]       # atc_gym.py:440-447
    assert got == exp and len(exp) == 4
    assert trail_from_original_state(raw[:3], 1).shape == (0, 2)


@pytest.mark.gpu
def test_rendered_image_matches_the_geometry():
    from atc_reinforcement_learning_b200 import BatchedAtcEnv, LOWW, SimParameters
    from atc_reinforcement_learning_b200.render import image_size
    env = BatchedAtcEnv(4, 2, SimParameters(1), LOWW(random_entrypoints=True), seed=3)
    cs = env.sector
    trail = np.array([[30.0, 30.0], [31.0, 31.5]])
    img = env.render('rgb_array', env_index=2, trail_xy=trail, labels=False)
    W, H = image_size(cs)
    assert img.shape == (H, W, 3) and img.dtype == np.uint8
    scale = 600 / (cs.bbox[2] - cs.bbox[0])
    colours = {tuple(c) for c in img.reshape(-1, 3)[::7].tolist()}
    assert colours <= {INACTIVE, ACTIVE, LINES, PLANE} and {INACTIVE, ACTIVE, LINES} <= colours

    def to_screen(x, y):                                             # atc_gym.py:543-552 + padding; row 0 = north
        return (x - cs.bbox[0]) * scale + 10, (y - cs.bbox[1]) * scale + 10

    def pixel(su, sv):
        return tuple(img[H - 1 - int(sv), int(su)].tolist())

    # fill: every pixel that is not a line / aircraft pixel is ACTIVE exactly where the reference scan finds an MVA
    pu, pr = np.meshgrid(np.arange(W), np.arange(H))
    wx = cs.bbox[0] + (pu + 0.5 - 10) / scale
    wy = cs.bbox[1] + ((H - 1 - pr) + 0.5 - 10) / scale
    inside = cs.find_mva_np(wx.ravel(), wy.ravel()).reshape(H, W) >= 0
    is_fill = np.all(img == np.array(ACTIVE, np.uint8), -1) | np.all(img == np.array(INACTIVE, np.uint8), -1)
    assert is_fill.mean() > 0.9
    np.testing.assert_array_equal(np.all(img == np.array(ACTIVE, np.uint8), -1)[is_fill], inside[is_fill])
    # runway centre, the FAF symbol's top corner and a drawn dash of the approach line are line pixels
    assert pixel(*to_screen(*cs.runway[:2])) == LINES
    fu, fv = to_screen(*cs.faf)
    assert pixel(fu, fv + 6) == LINES
    (rx, ry), (ix, iy) = cs.runway[:2], cs.iaf
    t = 0.5 / 48                                                     # middle of the first (drawn) segment
    assert pixel(*to_screen(rx + t * (ix - rx), ry + t * (iy - ry))) == LINES
    t = 1.5 / 48                                                     # middle of the first gap ... unless an MVA edge runs there
    assert pixel(*to_screen(rx + t * (ix - rx), ry + t * (iy - ry))) in (ACTIVE, INACTIVE, LINES)
    # the env's two aircraft: square outline 1.8 .. 3.8 px around the position, hollow centre; trail dots are filled
    st, _ = env.get_state()
    for a in range(2):
        su, sv = to_screen(float(st[2, a, 0]), float(st[2, a, 1]))
        assert pixel(su + 2.9, sv) == PLANE and pixel(su, sv - 2.9) == PLANE
        assert pixel(su, sv) != PLANE
    for x, y in trail:
        assert pixel(*to_screen(x, y)) == PLANE
    with pytest.raises(NotImplementedError):
        env.render('human')
    with pytest.raises(IndexError):
        env.render('rgb_array', env_index=4)


@pytest.mark.gpu
def test_adaptor_renders_its_own_trail():
    from atc_reinforcement_learning_b200 import make
    env = make('AtcEnv-v0')
    env.reset()
    for _ in range(40):
        env.step(np.array([0.0, 0.5, 0.2], np.float32))
    img = env.render('rgb_array')
    plane = np.all(img == np.array(PLANE, np.uint8), -1)
    assert plane.sum() > 30                                          # symbol outline + 4 trail dots
    env.close()


FONT_8 = ['.###.', '#...#', '#...#', '.###.', '#...#', '#...#', '.###.']
FONT_COLON = ['.....', '.##..', '.##..', '.....', '.##..', '.##..', '.....']


@pytest.mark.gpu
def test_text_labels_are_stamped_at_the_reference_anchor_points():
    """rendering.py:7-23 / atc_gym.py:404-443: reward lines at (10, 40) and (10, 25), aircraft name and "FL  speed" at
    rot_matrix(135) . (0, 8) from the symbol, anchored top-left, ColorScheme.label — in a 5 x 7 bitmap font."""
    from atc_reinforcement_learning_b200 import BatchedAtcEnv, LOWW, SimParameters, make
    from atc_reinforcement_learning_b200.render import image_size, label_list, stamp_labels
    # the glyphs themselves: "8:" on a black image, top-left corner at screen (20, 50)
    H, W = 80, 100
    img = torch.zeros(H, W, 3, dtype=torch.uint8, device='cuda')
    stamp_labels(img, [(20.0, 50.0, 0, "8:"), (60.0, 30.0, 1, "1")], img.device)
    a = (img.cpu().numpy() == np.array(PLANE, np.uint8)).all(-1)
    top = H - 1 - 49                                                   # image row of screen y = 49, the glyph's top row
    got8 = [''.join('#' if a[top + r, 20 + c] else '.' for c in range(5)) for r in range(7)]
    gotc = [''.join('#' if a[top + r, 26 + c] else '.' for c in range(5)) for r in range(7)]
    assert got8 == FONT_8 and gotc == FONT_COLON
    one = a[H - 1 - 29:H - 1 - 29 + 7, 60:66]
    assert one[:, 2:4].all(axis=0).all() and int(one.sum()) > 7 + 7      # bold: the vertical stroke is two pixels wide
    assert int(a.sum()) == sum(r.count('#') for r in FONT_8 + FONT_COLON) + int(one.sum())
    # on the env: labels change pixels only inside their boxes, and every label leaves some
    env = BatchedAtcEnv(3, 2, SimParameters(1), LOWW(random_entrypoints=True), seed=5)
    plain = env.render('rgb_array', env_index=1, labels=False)
    lab = env.render('rgb_array', env_index=1, last_reward=-0.05)
    Wd, Hd = image_size(env.sector)
    st, _ = env.get_state()
    ac = st[1].cpu().numpy()
    items = label_list(env.sector, [(r[0], r[1], r[2], r[4]) for r in ac], float(env.ep_return[1]), -0.05)
    assert [t for _, _, _, t in items][:2] == ["Total reward: %.2f" % float(env.ep_return[1]), "Last reward: -0.05"]
    assert items[2][3] == "FLT01" and items[3][3] == "%d  %d" % (round(ac[0][2] / 100), round(ac[0][4] / 10))
    changed = (plain != lab).any(-1)
    allowed = np.zeros_like(changed)
    for x, y, bold, text in items:
        x0, y1 = int(np.floor(x)), int(np.floor(y))                     # top-left corner, screen coordinates
        r0, r1 = Hd - 1 - (y1 - 1), Hd - 1 - (y1 - 7)
        box = (slice(max(r0, 0), max(r1 + 1, 0)), slice(max(x0, 0), x0 + 6 * len(text) + 1))
        allowed[box] = True
        assert (lab[box] == np.array(PLANE, np.uint8)).all(-1).any(), text
    assert changed.any() and not (changed & ~allowed).any()
    # the single-env adaptor passes its last reward
    g = make('AtcEnv-v0')
    g.reset()
    g.step(np.zeros(3, np.float32))
    assert g.render('rgb_array').shape == (Hd, Wd, 3)


def test_label_list_follows_the_reference_layout():
    """atc_gym.py:404-412, 436-443 (no GPU): reward lines at (10, 40) / (10, 25), not bold; per aircraft the name at
    rot_matrix(135) . (0, 8) from the symbol's screen position and "%d  %d" (flight level, speed / 10) 15 px below, bold."""
    import math
    from atc_reinforcement_learning_b200.render import PADDING, label_list
    cs = _sector()
    bx0, by0, bx1, by1 = [float(v) for v in cs.bbox]
    scale = 600 / (bx1 - bx0)
    items = label_list(cs, [(30.0, 40.0, 15049.0, 246.0), (10.0, 51.0, 3000.0, 200.0, 'AUA12')], 12.3456, -0.05)
    assert items[0] == (10.0, 40.0, 0, 'Total reward: 12.35') and items[1] == (10.0, 25.0, 0, 'Last reward: -0.05')
    c, s_ = math.cos(math.radians(135)), math.sin(math.radians(135))          # model.rot_matrix(135) . (0, 8)
    ex, ey = (30.0 - bx0) * scale + PADDING + 8 * s_, (40.0 - by0) * scale + PADDING + 8 * c
    x, y, bold, text = items[2]
    assert (round(x, 9), round(y, 9), bold, text) == (round(ex, 9), round(ey, 9), 1, 'FLT01')
    assert items[3][:3] == (x, y - 15.0, 1) and items[3][3] == '150  25'       # round(150.49), round(24.6)
    assert items[4][3] == 'AUA12' and items[5][3] == '30  20'
    assert label_list(cs, [], None, None) == []

"""CPU: the headless software renderer (spacegym_b200/replay.py) draws the reference's scene
(gym_space/rendering.py): planet outlines, goal cross, ship disc, thrust-dependent exhaust, decaying trace."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("sg_rendering", os.path.join(ROOT, "spacegym_b200", "replay.py"))
R = importlib.util.module_from_spec(spec)
spec.loader.exec_module(R)


def px(frame, x, y):
    return frame[frame.shape[0] - 1 - int(y), int(x)].astype(int)  # screen coordinates, y up


def test_scene_geometry_matches_reference_layout():
    r = R.Renderer([0.3, 0.3], 3.0)
    assert r.world_scale == 200.0 and r.screen == 600  # MAX_SCREEN_SIZE / world_size (rendering.py:25-27)
    r.reset([[0.9, 1.0], [-0.6, -0.5]], goal_pos=[0.1, -0.9])
    f = r.render([0.0, 0.0, 0.0], [0.0, 0.0])
    assert f.shape == (600, 600, 3) and f.dtype == np.uint8
    cx, cy = (0.9 + 1.5) * 200, (1.0 + 1.5) * 200
    assert px(f, cx + 60, cy).max() < 140 and px(f, cx, cy + 60).max() < 140  # outline at radius 0.3 * 200
    assert (px(f, cx, cy) == 255).all() and (px(f, cx + 30, cy) == 255).all()  # not filled
    gx, gy = (0.1 + 1.5) * 200, (-0.9 + 1.5) * 200
    near = lambda x, y: min(px(f, x + i, y + j).max() for i in (-1, 0, 1) for j in (-1, 0, 1))  # noqa: E731
    assert near(gx + 5, gy + 5) == 0 and near(gx + 5, gy - 5) == 0 and near(gx - 9, gy + 9) == 0  # the 20 px goal cross
    assert (px(f, gx + 5, gy) == 255).all()
    # ship: white disc of radius 15 with a black outline, engine triangle ahead of it (heading 0 = +x)
    assert (px(f, 300 - 7, 300 + 3) == 255).all()
    assert px(f, 300 - 15, 300).max() < 160
    assert px(f, 300 + 20, 300).max() == 0
    assert (px(f, 300 + 31, 300) == 255).all()  # no exhaust without thrust


def test_exhaust_opacity_and_trace_decay():
    r = R.Renderer([0.2, 3.0], 6.0)
    r.reset()
    full = r.render([0.0, 0.0, 0.0], [1.0, 0.0])
    half = R.Renderer([0.2, 3.0], 6.0).render([0.0, 0.0, 0.0], [0.5, 0.0])
    x = 300 + int(15 * 2.05)
    assert px(full, x, 300).max() == 0 and 100 < px(half, x, 300).max() < 160  # opacity = thrust (rendering.py:61-62)
    for k in range(1, 40):
        f = r.render([0.05 * k, 1.0, 0.0], None)
    # trace: the newest segment is black, older ones fade by 0.85 per segment, only 30 positions are kept
    y = int((1.0 + 3.0) * 100)
    newest = px(f, (0.05 * 38.5 + 3.0) * 100, y).max()
    older = px(f, (0.05 * 28.5 + 3.0) * 100, y).max()
    gone = px(f, (0.05 * 5 + 3.0) * 100, y).max()
    assert newest < 60 < older < 255 and gone == 255


def test_png_and_ppm_writers(tmp_path):
    import struct
    import zlib
    f = R.Renderer([0.3], 3.0).render([0.0, 0.0, 1.0], [0.3, 0.0])
    p = R.write_png(str(tmp_path / "f.png"), f)
    raw = open(p, "rb").read()
    assert raw[:8] == b"\x89PNG\r\n\x1a\n"
    w, h = struct.unpack(">II", raw[16:24])
    assert (w, h) == (600, 600)
    i = raw.index(b"IDAT")
    n = struct.unpack(">I", raw[i - 4:i])[0]
    px_rows = zlib.decompress(raw[i + 4:i + 4 + n])
    assert len(px_rows) == 600 * (1 + 600 * 3)
    back = np.frombuffer(px_rows, dtype=np.uint8).reshape(600, 1 + 1800)[:, 1:].reshape(600, 600, 3)
    assert (back == f).all()
    q = R.write_ppm(str(tmp_path / "f.ppm"), f)
    assert open(q, "rb").read(15) == b"P6\n600 600\n255\n"

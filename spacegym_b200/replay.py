"""Headless replay of recorded trajectories.

The reference draws through an OpenGL window (gym_space/rendering.py:15-182, pyglet viewer of
gym <= 0.21).  A batched engine on a GPU box has no display, so the same scene is rasterised in
software from recorded state: planets as circle outlines (`_init_planets` :78-85; for Kepler and
DoNotCrash the second outline is the border), the ship as a white disc with a black outline, its engine
triangle and three exhaust flames whose opacity is the thrust (:87-119, :61-62), the decaying trace of the
last 30 positions (:160-167) and the goal cross (:141-147).  World -> screen mapping as `_world_to_screen`
(:169-170): 600 px for `world_size`, origin in the window centre, y up.

    rec = TrajectoryRecorder(env, indices=[0, 1])     # env: BatchedSpaceEnv
    obs = env.reset(); rec.start()
    for t in range(T):
        obs, rew, done, info = env.step(actions); rec.record(actions, rew, done)
    rec.save("episode.npz")
    frames = replay("episode.npz", env_slot=0)        # uint8 [T, 600, 600, 3]
    write_png("frame.png", frames[-1])

Not on the hot path: `record` copies a handful of rows device -> host per step.
"""
import collections

import numpy as np

MAX_SCREEN_SIZE = 600   # rendering.py:11
SHIP_BODY_RADIUS = 15   # rendering.py:12


class TrajectoryRecorder:
    """Per-step snapshots of a few envs of a BatchedSpaceEnv, saved as one .npz.

    Arrays (T = recorded steps + 1, n = len(indices)): ship [T, n, 6]; planets [T, n, P, 2] and goal [T, n, 2]
    (Goal tasks; they change on reset / goal hits); action [T-1, n, 2] translated to (engine in [0, 1],
    thruster in [-1, 1]) like `last_action` (spaceship_env.py:72-73); reward, done [T-1, n]; plus the scene
    constants the renderer needs (world_size, radii, task).  After an auto-reset the snapshot is the new
    episode's first state, as in gym.vector.
    """

    def __init__(self, env, indices):
        self.env = env
        self.idx = np.asarray(indices, dtype=np.int64)
        self.rows = collections.defaultdict(list)

    def _snapshot(self):
        import torch
        st = self.env.state_views()
        ii = torch.as_tensor(self.idx, device=st["ship"].device)
        torch.cuda.current_stream(st["ship"].device).synchronize()
        self.rows["ship"].append(st["ship"][:, ii].t().cpu().numpy())
        if "planets" in st:
            p = st["planets"][:, ii].t().cpu().numpy()
            self.rows["planets"].append(p.reshape(len(self.idx), -1, 2))
            self.rows["goal"].append(st["goal"][:, ii].t().cpu().numpy())

    def start(self):
        self.rows.clear()
        self._snapshot()

    def record(self, actions, reward, done):
        a = np.asarray(actions.detach().cpu().numpy() if hasattr(actions, "detach") else actions, dtype=np.float32)
        a = a.reshape(self.env.num_envs, -1)[self.idx]
        if self.env.cfg.discrete:  # spaceship_env.py:189-202
            table = np.array([[0, 0], [1, 0], [0, -1], [0, 1], [1, -1], [1, 1]], dtype=np.float32)
            a = table[a[:, 0].astype(np.int64)]
        else:  # spaceship_env.py:211-214
            a = np.stack([(a[:, 0] + 1) / 2, a[:, 1]], axis=1)
        self.rows["action"].append(a)
        self.rows["reward"].append(reward.reshape(-1)[self.idx].cpu().numpy())
        self.rows["done"].append(done.reshape(-1)[self.idx].cpu().numpy())
        self._snapshot()

    def save(self, path):
        c = self.env.cfg
        out = {k: np.stack(v) for k, v in self.rows.items()}
        out.update(task=np.int32(c.task), world_size=np.float64(c.world_size), indices=self.idx,
                   env_id=np.array(self.env.env_id))
        if c.task == 0:
            out["radii"] = np.full(c.n_planets, c.planets_radius)
        else:  # planet + border at the origin (kepler.py:204-206, do_not_crash.py:15-17)
            out["radii"] = np.array([c.kepler_planet_radius, c.kepler_border_radius])
        np.savez_compressed(path, **out)
        return path


# ---- software rasteriser ------------------------------------------------------------------------
class Canvas:
    def __init__(self, size):
        self.n = int(size)
        self.img = np.ones((self.n, self.n, 3), dtype=np.float32)  # the gym viewer clears to white

    def _blend(self, rows, cols, color, alpha):
        ok = (rows >= 0) & (rows < self.n) & (cols >= 0) & (cols < self.n)
        r, c = rows[ok], cols[ok]
        a = (alpha[ok] if isinstance(alpha, np.ndarray) else alpha)
        a = np.asarray(a, dtype=np.float32)[..., None] if np.ndim(a) else np.float32(a)
        self.img[r, c] = self.img[r, c] * (1 - a) + np.asarray(color, dtype=np.float32) * a

    def _box(self, cx, cy, r):
        x0, x1 = int(np.floor(cx - r - 1)), int(np.ceil(cx + r + 1))
        y0, y1 = int(np.floor(cy - r - 1)), int(np.ceil(cy + r + 1))
        xs, ys = np.meshgrid(np.arange(x0, x1 + 1), np.arange(y0, y1 + 1))
        return xs.ravel(), ys.ravel()

    def circle(self, cx, cy, r, color=(0, 0, 0), filled=False, alpha=1.0):
        xs, ys = self._box(cx, cy, r)
        d = np.hypot(xs + 0.5 - cx, ys + 0.5 - cy)
        cover = np.clip(r + 0.5 - d, 0, 1) if filled else np.clip(1.0 - np.abs(d - r), 0, 1)
        m = cover > 0
        self._blend(self.n - 1 - ys[m], xs[m], color, cover[m] * alpha)  # y up -> row down

    def line(self, p, q, color=(0, 0, 0), alpha=1.0):
        p, q = np.asarray(p, dtype=np.float64), np.asarray(q, dtype=np.float64)
        k = int(np.ceil(np.hypot(*(q - p)))) + 1
        t = np.linspace(0, 1, k)
        pts = p[None] * (1 - t[:, None]) + q[None] * t[:, None]
        cols, rows = np.floor(pts[:, 0]).astype(np.int64), self.n - 1 - np.floor(pts[:, 1]).astype(np.int64)
        _, first = np.unique(rows * self.n + cols, return_index=True)  # each pixel once: opacity is not accumulated
        self._blend(rows[first], cols[first], color, alpha)

    def polygon(self, pts, color=(0, 0, 0), alpha=1.0):
        pts = np.asarray(pts, dtype=np.float64)
        cx, cy = pts.mean(axis=0)
        r = np.max(np.hypot(pts[:, 0] - cx, pts[:, 1] - cy))
        xs, ys = self._box(cx, cy, r)
        x, y = xs + 0.5, ys + 0.5
        inside = np.ones(len(x), dtype=bool)
        sign = None
        for i in range(len(pts)):  # convex polygon: same side of every edge
            a, b = pts[i], pts[(i + 1) % len(pts)]
            s = (b[0] - a[0]) * (y - a[1]) - (b[1] - a[1]) * (x - a[0])
            if sign is None:
                sign = np.sign(np.sum((b[0] - a[0]) * (cy - a[1]) - (b[1] - a[1]) * (cx - a[0]))) or 1.0
            inside &= (s * sign) >= 0
        self._blend(self.n - 1 - ys[inside], xs[inside], color, alpha)

    def rgb(self):
        return (np.clip(self.img, 0, 1) * 255 + 0.5).astype(np.uint8)


def _unit(angle):
    return np.array([np.cos(angle), np.sin(angle)])  # helpers.angle_to_unit_vector


class Renderer:
    """Software twin of gym_space.rendering.Renderer: same geometry, `render` returns an rgb array."""

    def __init__(self, planet_radii, world_size, goal_pos=None, num_prev_pos_vis=30, prev_pos_color_decay=0.85):
        self.world_size = float(world_size)
        self.world_translation = np.full(2, -self.world_size / 2)
        self.world_scale = MAX_SCREEN_SIZE / self.world_size
        self.screen = int(self.world_size * self.world_scale)
        self.radii = np.asarray(planet_radii, dtype=np.float64)
        self.planet_pos = np.zeros((len(self.radii), 2))
        self.goal_pos = goal_pos
        self.prev_ship_pos = collections.deque(maxlen=num_prev_pos_vis)
        self.decay = prev_pos_color_decay

    def _world_to_screen(self, p):
        return self.world_scale * (np.asarray(p, dtype=np.float64) - self.world_translation)

    def reset(self, planet_pos=None, goal_pos=None):
        if planet_pos is not None:
            self.planet_pos = np.asarray(planet_pos, dtype=np.float64).reshape(-1, 2)
        self.goal_pos = goal_pos
        self.prev_ship_pos.clear()

    def move_goal(self, goal_pos):
        self.goal_pos = goal_pos

    def render(self, ship_world_position, action=None):
        cv = Canvas(self.screen)
        for pos, r in zip(self.planet_pos, self.radii):
            cv.circle(*self._world_to_screen(pos), r * self.world_scale)
        if self.goal_pos is not None:
            gx, gy = self._world_to_screen(self.goal_pos)
            cv.line((gx - 10, gy - 10), (gx + 10, gy + 10))
            cv.line((gx - 10, gy + 10), (gx + 10, gy - 10))
        pos = self._world_to_screen(ship_world_position[:2])
        th = float(ship_world_position[2])
        self.prev_ship_pos.append(pos)
        thrust = float(action[0]) if action is not None else 0.0
        # engine triangle and exhaust (rendering.py:87-119), in the ship frame rotated by the heading
        w = np.pi / 4
        edge = SHIP_BODY_RADIUS * 1.7
        cv.polygon([pos, pos + edge * _unit(th - w / 2), pos + edge * _unit(th + w / 2)])
        if thrust > 0:
            for fa in np.linspace(-w / 4, w / 4, 3):
                v = _unit(th + fa)
                cv.line(pos + SHIP_BODY_RADIUS * 1.9 * v, pos + SHIP_BODY_RADIUS * 2.2 * v, alpha=min(1.0, thrust))
        cv.circle(*pos, SHIP_BODY_RADIUS, color=(1, 1, 1), filled=True)
        cv.circle(*pos, SHIP_BODY_RADIUS)
        cv.circle(*pos, 1.0, color=(0.5, 0.5, 0.5), filled=True)
        opacity = 1.0
        trace = list(self.prev_ship_pos)
        for i in range(1, len(trace)):  # rendering.py:160-167
            cv.line(trace[-i], trace[-i - 1], alpha=opacity)
            opacity *= self.decay
        return cv.rgb()


def replay(path, env_slot=0, every=1):
    """Frames of one recorded env: uint8 [T', 600, 600, 3]."""
    d = np.load(path)
    ship = d["ship"][:, env_slot]
    goal_task = int(d["task"]) == 0
    r = Renderer(d["radii"], float(d["world_size"]))
    frames = []
    for t in range(0, len(ship), every):
        fresh = t == 0 or bool(d["done"][t - 1, env_slot])
        planets = d["planets"][t, env_slot] if goal_task else None
        goal = d["goal"][t, env_slot] if goal_task else None
        if fresh:
            r.reset(planets, goal)
        elif goal_task:
            r.move_goal(goal)
        action = d["action"][t - 1, env_slot] if t > 0 and not fresh else None
        frames.append(r.render(ship[t], action))
    return np.stack(frames)


def write_png(path, frame):
    """8-bit RGB PNG with the standard library only (zlib)."""
    import struct
    import zlib
    frame = np.ascontiguousarray(frame, dtype=np.uint8)
    h, w, _ = frame.shape
    raw = b"".join(b"\x00" + frame[y].tobytes() for y in range(h))

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0))
                + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))
    return path


def write_ppm(path, frame):
    with open(path, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (frame.shape[1], frame.shape[0]))
        f.write(np.ascontiguousarray(frame, dtype=np.uint8).tobytes())
    return path

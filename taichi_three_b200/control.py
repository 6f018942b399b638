"""Orbit camera control (reference tina/util/control.py:7-163): turns GUI mouse events into
Engine.set_camera(view, proj).  Works with any object exposing the small subset of the
ti.GUI interface used here; a headless stand-in only needs `.res`."""
import numpy as np

from .matrix import affine, euler_matrix, orthogonal, perspective


def _quat_from_matrix(R):
    m = np.asarray(R, dtype=float)[:3, :3]
    w = np.sqrt(max(0.0, 1 + m[0, 0] + m[1, 1] + m[2, 2])) / 2
    x = np.sqrt(max(0.0, 1 + m[0, 0] - m[1, 1] - m[2, 2])) / 2
    y = np.sqrt(max(0.0, 1 - m[0, 0] + m[1, 1] - m[2, 2])) / 2
    z = np.sqrt(max(0.0, 1 - m[0, 0] - m[1, 1] + m[2, 2])) / 2
    x, y, z = np.copysign(x, m[2, 1] - m[1, 2]), np.copysign(y, m[0, 2] - m[2, 0]), np.copysign(z, m[1, 0] - m[0, 1])
    return np.array([w, x, y, z])


def _quat_mul(a, b):
    w0, x0, y0, z0 = a
    w1, x1, y1, z1 = b
    return np.array([w0 * w1 - x0 * x1 - y0 * y1 - z0 * z1, w0 * x1 + x0 * w1 + y0 * z1 - z0 * y1,
                     w0 * y1 - x0 * z1 + y0 * w1 + z0 * x1, w0 * z1 + x0 * y1 - y0 * x1 + z0 * w1])


def _quat_matrix(q):
    w, x, y, z = q / np.linalg.norm(q)
    R = np.eye(4)
    R[:3, :3] = [[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                 [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                 [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]]
    return R


def RotationStep(R, wx, wy, wz):  # matrix.py:14-19
    q = _quat_from_matrix(R)
    q = q + _quat_mul(q, 0.5 * np.array([0, wx, wy, wz]))
    return _quat_matrix(q)


class Control:
    def __init__(self, gui, fov=60, is_ortho=False, blendish=True):
        self.gui = gui
        self.center = np.zeros(3)
        self.is_ortho = is_ortho
        self.radius = 3.0
        self.R = np.eye(4)
        self.fov = fov
        self.blendish = blendish
        self.last_mouse = None

    def init_rot(self, theta, phi):  # control.py:20-27
        self.R = euler_matrix(-(theta or 0), phi or 0, 0, 'sxyz')

    @property
    def back(self):
        return self.R[:3, :3] @ np.array([0, 0, self.radius], dtype=float)

    def get_camera(self):  # control.py:102-113
        res = self.gui.res
        aspect = res[0] / res[1]
        if self.is_ortho:
            view = np.linalg.inv(affine(self.R[:3, :3], self.center + self.back / self.radius))
            proj = orthogonal(self.radius, aspect)
        else:
            view = np.linalg.inv(affine(self.R[:3, :3], self.center + self.back))
            proj = perspective(self.fov, aspect)
        return view, proj

    def apply_camera(self, engine):  # control.py:115-119
        changed = self.process_events()
        engine.set_camera(*self.get_camera())
        return changed

    # ---- events (only when the gui object supports them) ------------------------------
    def on_orbit(self, delta):
        self.R = RotationStep(self.R, delta[1] * np.pi, -delta[0] * np.pi, 0)

    def on_pan(self, delta):
        v = self.R[:3, :3] @ np.array([-delta[0] * np.pi, -delta[1] * np.pi, 0]) * 0.5
        self.center = self.center + v * self.radius

    def on_zoom(self, delta):
        self.radius *= pow(0.89, delta)

    def process_events(self):
        gui = self.gui
        if not hasattr(gui, 'get_events'):
            return False
        changed = False
        for e in gui.get_events():
            if e.type == gui.PRESS and e.key == gui.TAB:
                self.is_ortho = not self.is_ortho
                changed = True
            elif e.type == gui.PRESS and e.key == gui.ESCAPE:
                gui.running = False
            elif e.type == gui.MOTION and e.key == gui.WHEEL:
                self.on_zoom(e.delta[1] / 120)
                changed = True
        cur = np.array(gui.get_cursor_pos())
        lmb, mmb, rmb = (gui.is_pressed(b) for b in (gui.LMB, gui.MMB, gui.RMB))
        if self.last_mouse is not None and (lmb or mmb or rmb):
            delta = cur - self.last_mouse
            if delta[0] or delta[1]:
                shift = gui.is_pressed(gui.SHIFT)
                if self.blendish and mmb:
                    (self.on_pan if shift else self.on_orbit)(delta)
                elif not self.blendish and lmb:
                    self.on_orbit(delta)
                elif not self.blendish and rmb:
                    self.on_pan(delta)
                changed = True
        self.last_mouse = cur if (lmb or mmb or rmb) else None
        return changed

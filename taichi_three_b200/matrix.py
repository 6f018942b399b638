"""Host-side camera / transform matrices (float64 numpy), behaviourally matching the
reference's tina/util/matrix.py:5-104 and the parts of the third-party `transformations`
package (absent here) that tina/util/control.py:21-27 uses for orbit cameras.
"""
import numpy as np


def identity():
    return np.eye(4)


def affine(lin, pos):
    """4x4 from a 3x3 linear part and a translation (matrix.py:8-12)."""
    m = np.eye(4)
    m[:3, :3] = np.asarray(lin, dtype=float)[:3, :3]
    m[:3, 3] = np.asarray(pos, dtype=float)[:3]
    return m


def lookat(pos=(0, 0, 0), back=(0, 0, 3), up=(0, 1, 1e-12)):
    """View matrix looking at `pos` from `pos + back` (matrix.py:22-35)."""
    pos, back, up = (np.array(v, dtype=float) for v in (pos, back, up))
    fwd = -back / np.linalg.norm(back)
    right = np.cross(fwd, up)
    right = right / np.linalg.norm(right)
    up = np.cross(right, fwd)
    return np.linalg.inv(affine(np.stack([right, up, -fwd], axis=1), pos + back))


def ortho(left=-1, right=1, bottom=-1, top=1, near=-100, far=100):
    m = np.eye(4)
    m[0, 0], m[1, 1], m[2, 2] = 2 / (right - left), 2 / (top - bottom), -2 / (far - near)
    m[0, 3] = -(right + left) / (right - left)
    m[1, 3] = -(top + bottom) / (top - bottom)
    m[2, 3] = -(far + near) / (far - near)
    return m


def frustum(left=-1, right=1, bottom=-1, top=1, near=1, far=100):
    m = np.zeros((4, 4))
    m[0, 0], m[1, 1] = 2 * near / (right - left), 2 * near / (top - bottom)
    m[0, 2], m[1, 2] = (right + left) / (right - left), (top + bottom) / (top - bottom)
    m[2, 2], m[2, 3] = -(far + near) / (far - near), -2 * far * near / (far - near)
    m[3, 2] = -1
    return m


def orthogonal(size=1, aspect=1, near=-100, far=100):
    ax, ay = size * aspect, size
    return ortho(-ax, ax, -ay, ay, near, far)


def perspective(fov=60, aspect=1, near=0.05, far=500):
    t = np.tan(np.radians(fov) / 2)
    ax, ay = t * aspect, t
    return frustum(-near * ax, near * ax, -near * ay, near * ay, near, far)


def scale(factor):
    return affine(np.eye(3) * np.array(factor), np.zeros(3))


def translate(offset):
    return affine(np.eye(3), np.array(offset) * np.ones(3))


def quaternion(q):
    """Rotation from an (x, y, z, w) quaternion as glTF stores it (matrix.py:80-92)."""
    x, y, z, w = (float(v) for v in q)
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (w * y + x * z)],
                  [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                  [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
    return affine(R, np.zeros(3))


def eularXYZ(theta):
    cx, cy, cz = np.cos(theta[0]), np.cos(theta[1]), np.cos(theta[2])
    sx, sy, sz = np.sin(theta[0]), np.sin(theta[1]), np.sin(theta[2])
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return affine(Rz @ Ry @ Rx, np.zeros(3))


def euler_matrix(ai, aj, ak, axes='sxyz'):
    """`transformations.euler_matrix` for the static-xyz convention control.py:27 uses:
    R = Rz(ak) @ Ry(aj) @ Rx(ai)."""
    if axes != 'sxyz':
        raise NotImplementedError(axes)
    return eularXYZ((ai, aj, ak))


def orbit_camera(center=(0, 0, 0), radius=3.0, theta=0.0, phi=0.0, fov=60, aspect=1.0, is_ortho=False):
    """(view, proj) of the reference's orbit Control for fixed parameters
    (control.py:24-27 init_rot, :102-113 get_camera)."""
    R = euler_matrix(-theta, phi, 0)
    center = np.asarray(center, dtype=float)
    back = R[:3, :3] @ np.array([0, 0, radius], dtype=float)
    if is_ortho:
        view = np.linalg.inv(affine(R[:3, :3], center + back / radius))
        proj = orthogonal(radius, aspect)
    else:
        view = np.linalg.inv(affine(R[:3, :3], center + back))
        proj = perspective(fov, aspect)
    return view, proj

"""Octahedral-group constants of the E2PN configuration SE3ET uses (kanchor 6 x quotient 4 = 24 rotations, 15
kernel points).  Derived here from the group itself; the reference builds the same tables through trimesh
(utils_epn/rotation.py:484-523, anchors.py:85-90) and KPConvInterSO3.init_permute_idxs_* (blocks_epn.py:228-332).
tests/ pin them against the reference's buffers and against the tables compiled into the gather kernel."""
import functools
import math

import numpy as np

KANCHOR, QUOTIENT, NUM_KP, K_REAL = 6, 4, 15, 6

VERTICES = np.array([[0, 0, 1], [1, 0, 0], [0, 1, 0], [-1, 0, 0], [0, -1, 0], [0, 0, -1]], dtype=np.float64)
FACE_NORMALS = np.array([[1, 1, 1], [-1, 1, 1], [-1, -1, 1], [1, -1, 1], [1, 1, -1], [-1, 1, -1], [-1, -1, -1],
                         [1, -1, -1]], dtype=np.float64) / math.sqrt(3.0)


def rot_z(t):
    c, s = math.cos(t), math.sin(t)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


def rot_y(t):
    c, s = math.cos(t), math.sin(t)
    return np.array([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]])


@functools.lru_cache(maxsize=None)
def tables():
    """dict: anchors (6,3,3) = section of SO(3)/C4 taking +z to each vertex (zyz convention, gamma = 0);
    quotient (4,3,3) = C4 about z; kp_unit (15,3); kidx (15,6); ridx (6,6); trace_idx_ori (24,6)."""
    # R = Rz(alpha) Ry(beta) sends e_z to the vertex; alpha = 0 at +z and pi at -z (the reference's zyz section)
    anchors = []
    for v in VERTICES:
        beta = math.acos(max(-1.0, min(1.0, v[2])))
        if abs(v[2]) < 0.5:
            alpha = math.atan2(v[1], v[0])
        else:
            alpha = 0.0 if v[2] > 0 else math.pi
        anchors.append(np.round(rot_z(alpha) @ rot_y(beta)) + 0.0)
    anchors = np.stack(anchors)
    quotient = np.stack([np.round(rot_z(k * math.pi / 2)) for k in range(4)])
    kp = np.concatenate([VERTICES, FACE_NORMALS, np.zeros((1, 3))], 0)

    def nearest(points, target):
        return int(np.argmin(((points - target) ** 2).sum(1)))

    cls = -np.ones(NUM_KP, dtype=np.int64)
    n_cls = 0
    for i in range(NUM_KP):
        if cls[i] < 0:
            for q in quotient:
                cls[nearest(kp, q @ kp[i])] = n_cls
            n_cls += 1
    assert n_cls == K_REAL
    kidx = np.zeros((NUM_KP, KANCHOR), dtype=np.int64)
    for r in range(KANCHOR):
        rotated = kp @ anchors[r].T
        for k in range(NUM_KP):
            kidx[k, r] = cls[nearest(rotated, kp[k])]
    ridx = np.zeros((KANCHOR, KANCHOR), dtype=np.int64)
    for a in range(KANCHOR):
        for r in range(KANCHOR):
            scores = [max(np.trace((anchors[a] @ q).T @ anchors[r] @ anchors[b]) for q in quotient) for b in range(KANCHOR)]
            ridx[a, r] = int(np.argmax(scores))
    return {"anchors": anchors, "quotient": quotient, "kp_unit": kp, "kidx": kidx, "ridx": ridx}

"""Seeded synthetic point-cloud pairs shaped like the reference's datasets.

3DMatch-shaped: an indoor fragment (planar patches + boxes in a 3 x 2.5 x 2.7 m room, 2 mm noise)
voxel-downsampled at 0.025 m to ~15k points; the source cloud is an overlapping fragment moved by a
random SE(3).  KITTI-shaped: a LiDAR-like ground disc + street walls / boxes within +-60 m,
voxel-downsampled at 0.3 m to ~30k points.  Pair i always uses numpy.default_rng(1000 + i).
Host-side numpy only; nothing here is on the timed path.
"""
import numpy as np


def _voxel_downsample(points, voxel, rng=None):
    keys = np.floor(points / voxel).astype(np.int64)
    keys -= keys.min(0)
    dims = keys.max(0) + 1
    flat = keys[:, 0] + dims[0] * (keys[:, 1] + dims[1] * keys[:, 2])
    _, first = np.unique(flat, return_index=True)
    return points[np.sort(first)]


def _random_rotation(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
    ])


def _patch(rng, room, density):
    """A random axis-aligned-ish planar rectangle inside the room."""
    axis = rng.integers(0, 3)
    size = rng.uniform(0.3, 0.62, size=2) * np.delete(room, axis)
    origin = rng.uniform(0, 1, size=3) * room
    n = int(size[0] * size[1] * density)
    uv = rng.uniform(0, 1, size=(n, 2)) * size
    pts = np.zeros((n, 3))
    others = [d for d in range(3) if d != axis]
    pts[:, others[0]] = np.clip(origin[others[0]] - size[0] / 2 + uv[:, 0], 0, room[others[0]])
    pts[:, others[1]] = np.clip(origin[others[1]] - size[1] / 2 + uv[:, 1], 0, room[others[1]])
    pts[:, axis] = origin[axis]
    tilt = _random_rotation(rng)
    tilt = np.eye(3) + 0.08 * (tilt - np.eye(3))  # slight tilt
    c = pts.mean(0)
    return (pts - c) @ tilt.T + c


def _box(rng, room, density):
    size = rng.uniform(0.25, 0.65, size=3)
    corner = rng.uniform(0, 1, size=3) * (room - size)
    faces = []
    for axis in range(3):
        for side in (0, 1):
            others = [d for d in range(3) if d != axis]
            n = int(size[others[0]] * size[others[1]] * density)
            p = np.zeros((n, 3))
            p[:, others[0]] = corner[others[0]] + rng.uniform(0, size[others[0]], n)
            p[:, others[1]] = corner[others[1]] + rng.uniform(0, size[others[1]], n)
            p[:, axis] = corner[axis] + side * size[axis]
            faces.append(p)
    return np.concatenate(faces, 0)


def make_3dmatch_pair(index, target_points=15000, voxel=0.025, crop=None):
    """Returns dict(ref_points (N,3) f32, src_points (M,3) f32, transform (4,4) f32 mapping src -> ref)."""
    rng = np.random.default_rng(1000 + index)
    room = np.array([3.0, 2.5, 2.7])
    density = 6000.0
    surfaces = [_patch(rng, room, density) for _ in range(rng.integers(6, 11))]
    surfaces += [_box(rng, room, density) for _ in range(rng.integers(2, 5))]
    k = len(surfaces)
    share = rng.uniform(0.3, 0.8)
    n_shared = max(2, int(round(share * k)))
    perm = rng.permutation(k)
    shared, rest = perm[:n_shared], perm[n_shared:]
    half = len(rest) // 2
    ref_ids = np.concatenate([shared, rest[:half]])
    src_ids = np.concatenate([shared, rest[half:]])

    def build(ids):
        pts = np.concatenate([surfaces[i] for i in ids], 0)
        if crop is not None:
            centre = room / 2
            pts = pts[np.all(np.abs(pts - centre) <= crop / 2, axis=1)]
        pts = pts + rng.normal(scale=0.002, size=pts.shape)
        pts = _voxel_downsample(pts, voxel)
        if len(pts) > target_points:
            pts = pts[np.sort(rng.choice(len(pts), target_points, replace=False))]
        return pts

    ref = build(ref_ids)
    src_in_ref = build(src_ids)
    rot = _random_rotation(rng)
    trans = rng.uniform(-1, 1, size=3)
    trans *= min(1.0, 1.0 / max(np.linalg.norm(trans), 1e-9)) * rng.uniform(0, 1)
    # ref = R src + t  =>  src = R^T (ref - t)
    src = (src_in_ref - trans) @ rot
    transform = np.eye(4)
    transform[:3, :3] = rot
    transform[:3, 3] = trans
    return {
        "ref_points": ref.astype(np.float32),
        "src_points": src.astype(np.float32),
        "transform": transform.astype(np.float32),
    }


def make_kitti_pair(index, target_points=30000, voxel=0.3):
    rng = np.random.default_rng(1000 + index)

    def scan(offset):
        n_ground = 160000
        r = np.abs(rng.normal(0, 22, n_ground)) + 2.0
        th = rng.uniform(0, 2 * np.pi, n_ground)
        ground = np.stack([r * np.cos(th), r * np.sin(th), -1.7 + rng.normal(0, 0.03, n_ground)], 1)
        walls = []
        for side in (-1, 1):
            n = 40000
            x = rng.uniform(-60, 60, n)
            y = side * (8 + 0.5 * np.sin(x / 7.0)) + rng.normal(0, 0.05, n)
            z = rng.uniform(-1.7, 4.0, n)
            walls.append(np.stack([x, y, z], 1))
        boxes = []
        for _ in range(12):
            c = np.array([rng.uniform(-40, 40), rng.uniform(-6, 6), -1.0])
            s = np.array([rng.uniform(1.5, 4.5), rng.uniform(1.5, 2.0), rng.uniform(1.2, 1.8)])
            p = rng.uniform(-0.5, 0.5, size=(3000, 3))
            ax = rng.integers(0, 3, 3000)
            p[np.arange(3000), ax] = np.sign(p[np.arange(3000), ax]) * 0.5
            boxes.append(c + p * s)
        pts = np.concatenate([ground] + walls + boxes, 0) - offset
        pts = pts[np.all(np.abs(pts[:, :2]) <= 60, axis=1)]
        pts = _voxel_downsample(pts, voxel)
        if len(pts) > target_points:
            pts = pts[np.sort(rng.choice(len(pts), target_points, replace=False))]
        return pts

    ref = scan(np.zeros(3))
    move = np.array([rng.uniform(5, 10), rng.uniform(-0.5, 0.5), 0.0])
    yaw = rng.uniform(-0.1, 0.1)
    rot = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1]])
    src_world = scan(np.zeros(3))
    src = (src_world - move) @ rot
    transform = np.eye(4)
    transform[:3, :3] = rot
    transform[:3, 3] = move
    return {
        "ref_points": ref.astype(np.float32),
        "src_points": src.astype(np.float32),
        "transform": transform.astype(np.float32),
    }

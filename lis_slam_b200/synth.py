"""Seeded synthetic LiDAR data: a street-canyon scene, HDL-64 / VLP-16
ray-caster, edge/surf local-map sampler, poses and initial guesses.

There is no dataset access in this environment, so every benchmark/test input is
generated here from `numpy.random.default_rng(seed)`; identical bytes feed the CPU
oracle and the GPU engine.  Shapes follow KITTI HDL-64 (64 x 1800, 10 Hz).
"""
import ctypes as C
import os
import subprocess

import numpy as np

GROUND_Z = -1.73

_HERE = os.path.dirname(os.path.abspath(__file__))
_SYNTH_LIB = None


def build_synth_lib(force=False):
    """gcc -O2 -fopenmp -shared tools/synth_raycast.c -> libsynth.so (in-tree, next to liblisreg.so)."""
    src = os.path.join(_HERE, "tools", "synth_raycast.c"); out = os.path.join(_HERE, "libsynth.so")
    if force or not os.path.exists(out) or os.path.getmtime(src) > os.path.getmtime(out):
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-ffp-contract=off", "-o", out, src, "-lm"])
    return out


def _synth_lib():
    global _SYNTH_LIB
    if _SYNTH_LIB is None:
        _SYNTH_LIB = C.CDLL(build_synth_lib())
        _SYNTH_LIB.synth_raycast.restype = None
    return _SYNTH_LIB


def euler_to_R(roll, pitch, yaw):
    """R = Rz(yaw) Ry(pitch) Rx(roll) (pcl::getTransformation convention), float64."""
    A, B = np.cos(yaw), np.sin(yaw)
    Cc, D = np.cos(pitch), np.sin(pitch)
    E, F = np.cos(roll), np.sin(roll)
    return np.array([[A * Cc, A * D * F - B * E, B * F + A * D * E],
                     [B * Cc, A * E + B * D * F, B * D * E - A * F],
                     [-D, Cc * F, Cc * E]], dtype=np.float64)


def R_to_euler(R):
    """Inverse of euler_to_R (pcl::getTranslationAndEulerAngles)."""
    return np.array([np.arctan2(R[2, 1], R[2, 2]), np.arcsin(-R[2, 0]), np.arctan2(R[1, 0], R[0, 0])])


def pose_to_T(pose6):
    p = np.asarray(pose6, dtype=np.float64)
    T = np.eye(4)
    T[:3, :3] = euler_to_R(p[0], p[1], p[2])
    T[:3, 3] = p[3:6]
    return T


def T_to_pose(T):
    return np.concatenate([R_to_euler(T[:3, :3]), T[:3, 3]]).astype(np.float32)


def pose_error(pa, pb):
    """(rotation error [rad], translation error [m]) between two pose6 vectors."""
    Ta, Tb = pose_to_T(pa), pose_to_T(pb)
    dR = Ta[:3, :3].T @ Tb[:3, :3]
    ang = np.arccos(np.clip((np.trace(dR) - 1) / 2, -1, 1))
    return float(ang), float(np.linalg.norm(Ta[:3, 3] - Tb[:3, 3]))


class Scene:
    """Ground plane + building blocks with pilasters + poles + parked boxes.

    Labels follow config/label.yaml's 0..19 index space: ground 9 (road), facade 13
    (building), pole 18, box 1 (car).
    """

    def __init__(self, seed=1001, extent=72.0):
        rng = np.random.default_rng(seed)
        self.extent = extent
        boxes, labels = [], []
        # building blocks either side of the main street (facades at |y| = 8.5)
        for side in (+1, -1):
            for x0 in (-72.0, -42.0, -12.0, 18.0, 48.0):
                for (ya, yb) in ((8.5, 28.5), (38.5, 58.5)):
                    lo, hi = (ya, yb) if side > 0 else (-yb, -ya)
                    boxes.append([x0, lo, GROUND_Z, x0 + 24.0, hi, 6.27]); labels.append(13)
            # pilasters: 3 m wide, 0.5 m deep, every 6 m on the main-street facade
            for k in range(-12, 12):
                x0 = 6.0 * k + 1.5
                inside = any(bx <= x0 and x0 + 3.0 <= bx + 24.0 for bx in (-72.0, -42.0, -12.0, 18.0, 48.0))
                if not inside:
                    continue
                lo, hi = (8.0, 8.5) if side > 0 else (-8.5, -8.0)
                boxes.append([x0, lo, GROUND_Z, x0 + 3.0, hi, 6.27]); labels.append(13)
        # parked "cars"
        for _ in range(12):
            cx, cy = rng.uniform(-60, 60), rng.uniform(2.5, 6.0) * rng.choice([-1, 1])
            boxes.append([cx - 2.0, cy - 1.0, GROUND_Z, cx + 2.0, cy + 1.0, GROUND_Z + 1.5]); labels.append(1)
        self.boxes = np.array(boxes, dtype=np.float64)
        self.box_labels = np.array(labels, dtype=np.uint16)
        # poles / trunks
        n_pole = 60
        px = rng.uniform(-60, 60, n_pole)
        py = rng.uniform(3.0, 7.5, n_pole) * rng.choice([-1, 1], n_pole)
        extra = 240  # trunks in the cross alleys / back streets
        ex = rng.uniform(-70, 70, extra)
        ey = rng.uniform(29.5, 37.5, extra) * rng.choice([-1, 1], extra)
        self.poles = np.stack([np.concatenate([px, ex]), np.concatenate([py, ey])], 1)
        self.pole_r, self.pole_h = 0.15, 4.0

    # ------------------------------------------------------------------ ray casting
    def raycast(self, origin, dirs):
        """Nearest hit along rays. Returns (range [inf = miss], label u16)."""
        o = np.asarray(origin, np.float64)
        d = np.asarray(dirs, np.float64)
        n = d.shape[0]
        best = np.full(n, np.inf)
        lab = np.zeros(n, np.uint16)
        reach = 80.0
        with np.errstate(divide="ignore", invalid="ignore"):
            # ground
            t = (GROUND_Z - o[2]) / d[:, 2]
            hx, hy = o[0] + t * d[:, 0], o[1] + t * d[:, 1]
            ok = (t > 0.5) & (np.abs(hx) < self.extent + 5) & (np.abs(hy) < self.extent + 5)
            best = np.where(ok, t, best); lab = np.where(ok, 9, lab).astype(np.uint16)
            inv = 1.0 / d
            for b, bl in zip(self.boxes, self.box_labels):
                # cull boxes farther than the sensor reach
                c = np.clip(o, b[:3], b[3:])
                if np.linalg.norm(c - o) > reach:
                    continue
                t1 = (b[:3] - o) * inv
                t2 = (b[3:] - o) * inv
                tmin = np.max(np.minimum(t1, t2), axis=1)
                tmax = np.min(np.maximum(t1, t2), axis=1)
                ok = (tmax >= tmin) & (tmin > 0.5) & (tmin < best)
                best = np.where(ok, tmin, best); lab = np.where(ok, bl, lab).astype(np.uint16)
            # vertical cylinders: only rays whose azimuth can see the pole are tested
            a = d[:, 0] ** 2 + d[:, 1] ** 2
            az = np.arctan2(d[:, 1], d[:, 0])
            for (cx, cy) in self.poles:
                ox, oy = o[0] - cx, o[1] - cy
                dist = np.hypot(ox, oy)
                if dist > reach or dist <= self.pole_r:
                    continue
                caz = np.arctan2(-oy, -ox)
                half = np.arcsin(min(1.0, self.pole_r / dist)) + 1e-3
                dif = np.abs(((az - caz) + np.pi) % (2 * np.pi) - np.pi)
                sel = np.nonzero(dif <= half)[0]
                if len(sel) == 0:
                    continue
                ds = d[sel]
                bq = ox * ds[:, 0] + oy * ds[:, 1]
                cq = ox * ox + oy * oy - self.pole_r ** 2
                disc = bq * bq - a[sel] * cq
                t = (-bq - np.sqrt(np.where(disc > 0, disc, np.nan))) / a[sel]
                z = o[2] + t * ds[:, 2]
                ok = (disc > 0) & (t > 0.5) & (t < best[sel]) & (z >= GROUND_Z) & (z <= GROUND_Z + self.pole_h)
                hit = sel[ok]
                best[hit] = t[ok]; lab[hit] = 18
        return best, lab

    def raycast_fast(self, origin, dirs):
        """Same as raycast() through the C generator (tools/synth_raycast.c, libsynth.so): ~100x faster; used for the
        long synthetic streams.  Ranges agree with raycast() to rounding (the pole test skips numpy's azimuth
        pre-cull, which only removes rays that cannot hit)."""
        lib = _synth_lib()
        o = np.ascontiguousarray(origin, np.float64); d = np.ascontiguousarray(dirs, np.float64)
        n = len(d)
        rng_out = np.empty(n, np.float64); lab = np.empty(n, np.uint16)
        boxes = np.ascontiguousarray(self.boxes, np.float64); bl = np.ascontiguousarray(self.box_labels, np.uint16)
        poles = np.ascontiguousarray(self.poles, np.float64)
        vp = C.c_void_p
        lib.synth_raycast(o.ctypes.data_as(vp), d.ctypes.data_as(vp), C.c_int64(n), boxes.ctypes.data_as(vp), bl.ctypes.data_as(vp),
                          C.c_int32(len(boxes)), poles.ctypes.data_as(vp), C.c_int32(len(poles)), C.c_double(self.pole_r),
                          C.c_double(self.pole_h), C.c_double(GROUND_Z), C.c_double(self.extent), rng_out.ctypes.data_as(vp),
                          lab.ctypes.data_as(vp))
        return rng_out, lab

    def scan(self, pose6, sensor="hdl64", seed=2000, noise=0.01, max_range=70.0, fast=False):
        """Raw sweep in firing order (column-major).  Returns dict with
        pts (N,4) f32 {x,y,z,intensity} in the SENSOR frame, ring u16, time f32, label u16."""
        if sensor == "hdl64":
            n_ring, period = 64, 0.1
            elev = np.deg2rad(np.linspace(-24.8, 2.0, n_ring))  # ring 0 = lowest beam
        elif sensor == "vlp16":
            n_ring, period = 16, 0.01
            elev = np.deg2rad(np.linspace(-15.0, 15.0, n_ring))
        else:
            raise ValueError(sensor)
        H = 1800
        rng = np.random.default_rng(seed)
        az = np.deg2rad((np.arange(H) - H // 2) * (360.0 / H))
        azg, elg = np.meshgrid(az, elev, indexing="ij")  # (H, n_ring): column-major firing order
        d_s = np.stack([np.cos(elg) * np.cos(azg), np.cos(elg) * np.sin(azg), np.sin(elg)], -1).reshape(-1, 3)
        T = pose_to_T(pose6)
        rngs, lab = (self.raycast_fast if fast else self.raycast)(T[:3, 3], d_s @ T[:3, :3].T)
        rngs = rngs + rng.normal(0.0, noise, rngs.shape)
        keep = np.isfinite(rngs) & (rngs < max_range) & (rngs > 1.0)
        ring = np.tile(np.arange(n_ring, dtype=np.uint16), H)
        col = np.repeat(np.arange(H), n_ring)
        pts = np.zeros((keep.sum(), 4), np.float32)
        pts[:, :3] = (d_s[keep] * rngs[keep, None]).astype(np.float32)
        pts[:, 3] = (lab[keep] * 10).astype(np.float32)
        return {"pts": pts, "ring": ring[keep].copy(), "time": (period * col[keep] / H).astype(np.float32),
                "label": lab[keep].copy(), "n_ring": n_ring, "H": H}

    # ------------------------------------------------------------------ map sampling
    def _edge_lines(self):
        segs = []
        for b, bl in zip(self.boxes, self.box_labels):
            x0, y0, z0, x1, y1, z1 = b
            for (x, y) in ((x0, y0), (x0, y1), (x1, y0), (x1, y1)):
                segs.append([x, y, z0, x, y, z1, 13 if bl == 13 else 1])
            if bl == 1:  # car roof outline
                for (xa, ya, xb, yb) in ((x0, y0, x1, y0), (x1, y0, x1, y1), (x1, y1, x0, y1), (x0, y1, x0, y0)):
                    segs.append([xa, ya, z1, xb, yb, z1, 1])
        for (cx, cy) in self.poles:
            segs.append([cx, cy, GROUND_Z, cx, cy, GROUND_Z + self.pole_h, 18])
        return np.array(segs)

    def sample_map(self, n_edge=40000, n_surf=160000, seed=3001, jitter=0.01):
        """Edge map (points along vertical/roof edges and pole axes) and surf map (ground,
        facades, box faces) sampled like voxel-grid output (~0.2 m / ~0.4 m spacing)."""
        rng = np.random.default_rng(seed)
        segs = self._edge_lines()
        e_pts, e_lab = [], []
        for s in segs:
            L = np.linalg.norm(s[3:6] - s[0:3])
            k = max(2, int(L / 0.07))
            t = (np.arange(k) + rng.uniform(0, 1, k)) / k
            e_pts.append(s[0:3] + t[:, None] * (s[3:6] - s[0:3])); e_lab.append(np.full(k, int(s[6]), np.uint16))
        e_pts = np.concatenate(e_pts); e_lab = np.concatenate(e_lab)
        s_pts, s_lab = [], []
        h = 0.4
        g = np.arange(-self.extent, self.extent, h)
        gx, gy = np.meshgrid(g, g, indexing="ij")
        gp = np.stack([gx.ravel(), gy.ravel(), np.full(gx.size, GROUND_Z)], 1)
        inside = np.zeros(len(gp), bool)
        for b in self.boxes:
            inside |= (gp[:, 0] > b[0]) & (gp[:, 0] < b[3]) & (gp[:, 1] > b[1]) & (gp[:, 1] < b[4])
        s_pts.append(gp[~inside]); s_lab.append(np.full((~inside).sum(), 9, np.uint16))
        for b, bl in zip(self.boxes, self.box_labels):
            x0, y0, z0, x1, y1, z1 = b
            zz = np.arange(z0 + 0.2, z1, h)
            for (axis, c) in ((0, x0), (0, x1), (1, y0), (1, y1)):
                u = np.arange((y0 if axis == 0 else x0) + 0.2, (y1 if axis == 0 else x1), h)
                uu, vv = np.meshgrid(u, zz, indexing="ij")
                if axis == 0:
                    p = np.stack([np.full(uu.size, c), uu.ravel(), vv.ravel()], 1)
                else:
                    p = np.stack([uu.ravel(), np.full(uu.size, c), vv.ravel()], 1)
                s_pts.append(p); s_lab.append(np.full(len(p), bl, np.uint16))
            if bl == 1:
                u = np.arange(x0 + 0.2, x1, h); v = np.arange(y0 + 0.2, y1, h)
                uu, vv = np.meshgrid(u, v, indexing="ij")
                p = np.stack([uu.ravel(), vv.ravel(), np.full(uu.size, z1)], 1)
                s_pts.append(p); s_lab.append(np.full(len(p), 1, np.uint16))
        s_pts = np.concatenate(s_pts); s_lab = np.concatenate(s_lab)

        def fit(p, l, n):
            p = p + rng.normal(0, jitter, p.shape)
            if len(p) >= n:
                sel = rng.choice(len(p), n, replace=False)
                sel.sort()
                return p[sel], l[sel]
            extra = rng.choice(len(p), n - len(p), replace=True)
            q = p[extra] + rng.normal(0, 0.05, (len(extra), 3))
            return np.concatenate([p, q]), np.concatenate([l, l[extra]])

        e_pts, e_lab = fit(e_pts, e_lab, n_edge)
        s_pts, s_lab = fit(s_pts, s_lab, n_surf)
        mc = np.zeros((n_edge, 4), np.float32); mc[:, :3] = e_pts; mc[:, 3] = e_lab
        ms = np.zeros((n_surf, 4), np.float32); ms[:, :3] = s_pts; ms[:, 3] = s_lab
        return {"corner": mc, "surf": ms, "corner_label": e_lab, "surf_label": s_lab}

    def sample_scan_features(self, pose6, n_corner=4000, n_surf=12000, seed=0, noise=0.02, max_range=55.0):
        """Scan-like corner/surf feature clouds in the SENSOR frame at `pose6` (quick
        stand-in for raycast + feature extraction + voxel grid, used by LM unit tests)."""
        rng = np.random.default_rng(seed)
        m = self.sample_map(n_edge=60000, n_surf=200000, seed=seed + 77, jitter=noise)
        T = pose_to_T(pose6)
        out = {}
        for key, n in (("corner", n_corner), ("surf", n_surf)):
            p = m[key][:, :3].astype(np.float64)
            lab = m[key + "_label"]
            r = np.linalg.norm(p - T[:3, 3], axis=1)
            w = np.where((r > 2.0) & (r < max_range), 1.0 / np.maximum(r, 4.0) ** 1.5, 0.0)
            sel = rng.choice(len(p), n, replace=False, p=w / w.sum())
            sel.sort()
            q = (p[sel] - T[:3, 3]) @ T[:3, :3]  # R^T (p - t)
            arr = np.zeros((n, 4), np.float32); arr[:, :3] = q; arr[:, 3] = lab[sel]
            out[key] = arr; out[key + "_label"] = lab[sel].copy()
        return out


def random_pose(rng):
    """Ground-truth pose distribution of SURVEY.md §8d (street-aligned)."""
    return np.array([rng.uniform(-0.02, 0.02), rng.uniform(-0.02, 0.02), rng.uniform(-np.pi, np.pi),
                     rng.uniform(-40, 40), rng.uniform(-2, 2), rng.uniform(-0.1, 0.1)], dtype=np.float32)


def perturb_pose(pose6, rng, rot=0.02, trans=0.3):
    """Initial guess = truth o perturbation with |dtheta| <= rot, |dt| <= trans."""
    dth = rng.standard_normal(3); dth *= rng.uniform(0, rot) / np.linalg.norm(dth)
    dt = rng.standard_normal(3); dt *= rng.uniform(0, trans) / np.linalg.norm(dt)
    T = pose_to_T(pose6) @ pose_to_T(np.concatenate([dth, dt]))
    return T_to_pose(T)

"""Host-side scene helpers for the harness: camera / projection matrices in the reference's
conventions (DirectXMath row-vector, left-handed; MultiVolumes.cpp:264-279, 328-337,
ObjectRenderer.cpp:171-186) and the synthetic inputs of SURVEY.md §8d. Pure numpy."""
import numpy as np

G_ZNEAR, G_ZFAR = 1.0, 1000.0           # SharedConsts.h:9-10
FOV_Y = np.pi / 4                       # MultiVolumes.cpp:21


def look_at_lh(eye, focus, up=(0, 1, 0)):
    eye, focus, up = (np.asarray(v, np.float64) for v in (eye, focus, up))
    z = focus - eye; z /= np.linalg.norm(z)
    x = np.cross(up, z); x /= np.linalg.norm(x)
    y = np.cross(z, x)
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2] = x, y, z
    m[3, :3] = [-x @ eye, -y @ eye, -z @ eye]
    return m


def perspective_fov_lh(fov_y, aspect, zn, zf):
    h = 1.0 / np.tan(fov_y / 2); w = h / aspect; r = zf / (zf - zn)
    return np.array([[w, 0, 0, 0], [0, h, 0, 0], [0, 0, r, 1], [0, 0, -r * zn, 0]], np.float64)


def orthographic_lh(w, h, zn, zf):
    r = 1.0 / (zf - zn)
    return np.array([[2 / w, 0, 0, 0], [0, 2 / h, 0, 0], [0, 0, r, 0], [0, 0, -r * zn, 1]], np.float64)


def default_camera(width, height, eye=(4.0, 16.0, -80.0), focus=(0.0, 0.0, 0.0)):
    """view * proj of the reference's start-up camera, as float32 row-major (row-vector convention)."""
    view = look_at_lh(eye, focus)
    proj = perspective_fov_lh(FOV_Y, width / float(height), G_ZNEAR, G_ZFAR)
    return (view @ proj).astype(np.float32), np.asarray(eye, np.float32)


def orbit_camera(width, height, frame, fps=60.0):
    """MultiVolumes.cpp:328-337 auto animation: eye = (r sin t, 6, r cos t), t = 0.5 * time."""
    t = 0.5 * (frame / fps)
    eye = (np.sin(t) * 60.0, 6.0, np.cos(t) * 60.0)
    return default_camera(width, height, eye=eye)


def shadow_view_proj(light_pt=(75.0, 75.0, -75.0), scene_size=None):
    """ObjectRenderer.cpp:177-184: ortho size = sceneSize * 1.5, z in [1, 200]."""
    size = (scene_size if scene_size is not None else 40.0) * 1.5
    view = look_at_lh(light_pt, (0, 0, 0))
    proj = orthographic_lh(size, size, 1.0, 200.0)
    return (view @ proj).astype(np.float32)


LIGHT_PT = (75.0, 75.0, -75.0)                                   # MultiVolumes.cpp:339-345
LIGHT_COLOR, LIGHT_INTENSITY = (1.0, 0.7, 0.3), 3.0 * np.pi
AMBIENT_COLOR, AMBIENT_INTENSITY = (0.4, 0.6, 1.0), 2.0 * np.pi


def procedural_sky(size=64, light_dir=LIGHT_PT):
    """Stand-in for the missing LA_Radiance.dds: zenith-blue / horizon-white / ground-brown gradient
    plus one Gaussian sun lobe toward the light (SURVEY.md §8d cfg 2). Returns (6, size, size, 3) f32."""
    l = np.asarray(light_dir, np.float64); l /= np.linalg.norm(l)
    out = np.empty((6, size, size, 3), np.float32)
    idx = (np.arange(size) + 0.5) / size * 2 - 1
    px, py = np.meshgrid(idx, -idx)      # GetLocalPos: y flipped
    for f in range(6):
        if f == 0: d = np.stack([np.ones_like(px), py, -px], -1)
        elif f == 1: d = np.stack([-np.ones_like(px), py, px], -1)
        elif f == 2: d = np.stack([px, np.ones_like(px), -py], -1)
        elif f == 3: d = np.stack([px, -np.ones_like(px), py], -1)
        elif f == 4: d = np.stack([px, py, np.ones_like(px)], -1)
        else: d = np.stack([-px, py, -np.ones_like(px)], -1)
        d = d / np.linalg.norm(d, axis=-1, keepdims=True)
        up = d[..., 1:2]
        sky = np.where(up > 0, (1 - up) * np.array([1.0, 1.0, 1.0]) + up * np.array([0.25, 0.45, 1.0]),
                       (1 + up) * np.array([1.0, 1.0, 1.0]) - up * np.array([0.35, 0.25, 0.15]))
        sun = np.exp((d @ l - 1.0) * 60.0)[..., None] * np.array([8.0, 6.0, 3.0])
        out[f] = (sky + sun).astype(np.float32)
    return out


def sphere_depth(width, height, view_proj, center=(0.0, -9.0 + 9.0, 0.0), radius=9.0):
    """Analytic stand-in for the mesh pre-pass (SURVEY.md §8d cfg 3 fallback): D32 depth of a sphere."""
    vp = view_proj.astype(np.float64)
    inv = np.linalg.inv(vp)
    xs = (np.arange(width) + 0.5) / width * 2 - 1
    ys = -((np.arange(height) + 0.5) / height * 2 - 1)
    X, Y = np.meshgrid(xs, ys)
    def unproj(z):
        p = np.stack([X, Y, np.full_like(X, z), np.ones_like(X)], -1) @ inv
        return p[..., :3] / p[..., 3:4]
    p0, p1 = unproj(0.0), unproj(1.0)
    d = p1 - p0; d /= np.linalg.norm(d, axis=-1, keepdims=True)
    oc = p0 - np.asarray(center)
    b = (oc * d).sum(-1); cc = (oc * oc).sum(-1) - radius * radius
    disc = b * b - cc
    hit = disc > 0
    t = -b - np.sqrt(np.where(hit, disc, 0))
    hit &= t > 0
    hp = p0 + d * t[..., None]
    clip = np.concatenate([hp, np.ones_like(hp[..., :1])], -1) @ vp
    z = clip[..., 2] / clip[..., 3]
    return np.where(hit, z, 1.0).astype(np.float32)


def occluder_mesh(rings=132, sectors=264):
    """Stand-in for the reference's Bin/Assets/bunny.obj (34 835 vertices / 69 666 triangles, x in [-5, 5], y in [0, 9.9],
    z in [-3.9, 3.9]; `-mesh bunny.obj 0 -9 0 1.8` in Bin/all64.bat), which is not redistributed with this package: a closed,
    lobed blob of the same extent and 69 696 triangles, so the depth / shadow producer sees the same load.
    Returns (positions (V, 3) float32, indices (3 T,) uint32) in the OBJ importer's output convention."""
    pos, idx = [], []
    for r in range(rings + 1):
        th = np.pi * r / rings
        for q in range(sectors):
            ph = 2.0 * np.pi * q / sectors
            d = np.array([np.sin(th) * np.cos(ph), np.cos(th), np.sin(th) * np.sin(ph)])
            lobes = 1.0 + 0.18 * np.sin(3.0 * ph) * np.sin(th) ** 2 + 0.12 * np.cos(2.0 * th)      # ears / haunches, kept star-shaped
            pos.append(tuple(d * lobes))
    for r in range(rings):
        for q in range(sectors):
            a, b = r * sectors + q, r * sectors + (q + 1) % sectors
            c, d = a + sectors, b + sectors
            idx += [a, c, b, b, c, d]
    p = np.asarray(pos, np.float64)
    lo, hi = p.min(0), p.max(0)
    p = (p - lo) / (hi - lo) * np.array([10.03, 9.94, 7.77]) + np.array([-5.015, -0.044, -3.887])      # the bunny's bounding box
    return p.astype(np.float32), np.asarray(idx, np.uint32)


MESH_WORLD = (1.8, (0.0, -9.0, 0.0))          # scale, position: Bin/all64.bat:1

"""Shared seeded Ken-Burns scene for the parity tests (numpy, CPU): image, disparity -> cloud via the ORACLE."""
import numpy as np

from cartoonsegmentation_b200.utils.synthetic import smooth_disparity, smooth_image
from oracle import kb_oracle as orc

FOCAL, BASELINE = 512.0, 40.0


def make_scene(h, w, seed=0, extra_points=0):
    img = smooth_image(h, w, seed=1000 + seed)
    raw = smooth_disparity(h, w, seed=2000 + seed)
    cloud = orc.disparity_to_cloud(raw, FOCAL, BASELINE)
    img_t = np.ascontiguousarray(img.transpose(2, 0, 1)[None].astype(np.float32) * np.float32(1.0 / 255.0))
    pts = cloud['points'].reshape(1, 3, -1).copy()
    data = np.concatenate([img_t.reshape(1, 3, -1), cloud['depth'].reshape(1, 1, -1)], 1)
    if extra_points:   # appended (inpainted-style) points: arbitrary positions behind / beside the grid
        rng = np.random.default_rng(seed + 7)
        z = rng.uniform(600, 5000, extra_points).astype(np.float32)
        x = (rng.uniform(-w / 2 - 20, w / 2 + 20, extra_points) * z / FOCAL).astype(np.float32)
        y = (rng.uniform(-h / 2 - 20, h / 2 + 20, extra_points) * z / FOCAL).astype(np.float32)
        pts = np.concatenate([pts, np.stack([x, y, z])[None]], 2)
        ed = rng.uniform(0, 1, (1, 3, extra_points)).astype(np.float32)
        data = np.concatenate([data, np.concatenate([ed, z[None, None]], 1)], 2)
    common = {'objDepthrange': cloud['depthrange'], 'intWidth': w, 'intHeight': h, 'fltFocal': FOCAL, 'fltBaseline': BASELINE}
    if min(h, w) <= 256:   # depthrange needs a >256 px image; small scenes get a synthetic closest point
        d = cloud['depth'][0, 0]
        iy, ix = np.unravel_index(np.argmin(d), d.shape)
        common['objDepthrange'] = (float(d.min()), float(d.max()), (int(ix), int(iy)), (0, 0))
    return dict(img=img, raw=raw, cloud=cloud, points=np.ascontiguousarray(pts), data=np.ascontiguousarray(data), common=common)

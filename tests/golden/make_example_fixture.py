"""BASELINE.json configs[0] input: the reference's example image `examples/1562990.jpg` (1500 x 750), INTER_AREA-resized so that its short side is 512
and centre-cropped to 512 x 512 -> tests/golden/example_1562990_512.jpg (JPEG quality 92).  Run in the build container
(/root/reference is not on the GPU box); bench.py's `infer_512_example` workload and its CPU arm read the fixture."""
import os

import cv2

REF = os.environ.get("CSB_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.abspath(__file__))

img = cv2.imread(os.path.join(REF, "examples", "1562990.jpg"))
h, w = img.shape[:2]
s = 512.0 / min(h, w)
nw, nh = max(512, int(round(w * s))), max(512, int(round(h * s)))
small = cv2.resize(img, (nw, nh), interpolation=cv2.INTER_AREA)
y0, x0 = (nh - 512) // 2, (nw - 512) // 2
crop = small[y0:y0 + 512, x0:x0 + 512]
out = os.path.join(ROOT, "example_1562990_512.jpg")
cv2.imwrite(out, crop, [cv2.IMWRITE_JPEG_QUALITY, 92])
print(out, crop.shape, os.path.getsize(out))

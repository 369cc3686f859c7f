"""First end-to-end check on a GPU box: demo case through the C-ABI vs the oracle."""
import hashlib, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from contour3d_b200 import api, casefile
from oracle import oracle

case = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/demo.bin"
scene, grids, fields, contours, lines = casefile.read_case(case)
ctx = api.Context()
print(api.lib().c3d_version().decode())
nz, ny, nx = fields[0].shape
fs = [ctx.field((nx, ny, nz), dtype=f.dtype).upload(f) for f in fields]
cs = [dict(field=fs[c["field"]], grids=grids, threshold=c["threshold"], color=c["color"]) for c in contours]
for it in range(3):
    t = time.time()
    img = ctx.render(scene, cs, lines)
    print("frame", it, time.time() - t, ctx.timings())
ocs = [dict(field=fields[c["field"]], grids=grids, threshold=c["threshold"], color=c["color"]) for c in contours]
t = time.time()
rgb, depth, stats = oracle.frame(scene, ocs, lines, nthreads=8)
print("oracle", time.time() - t, stats)
ref = rgb[::-1]
diff = (img != ref).any(axis=2)
print("pixels differing:", int(diff.sum()), "of", diff.size, "max abs diff", int(np.abs(img.astype(int) - ref.astype(int)).max()))
gd, gk = ctx.canvas(scene["width"], scene["height"])
print("depth equal:", bool((gd == depth).all()), "mismatch count", int((gd != depth).sum()))
# codes
codes, counts = ctx.case_codes(fs[0], contours[0]["threshold"])
ocodes, ocounts = oracle.case_codes(fields[0], grids, contours[0]["threshold"])
print("codes equal", bool((codes == ocodes).all()), "counts equal", bool((counts == ocounts).all()), "triangles", int(counts.sum()))
tri = ctx.tessellate(fs[0], grids, contours[0]["threshold"])
otri, ostats = oracle.tessellate(fields[0], grids, contours[0]["threshold"])
print("tri count", len(tri), len(otri))
if len(tri) == len(otri):
    for name in ("cube", "tri", "has_vn", "edge"):
        print(name, bool((tri[name] == otri[name]).all()))
    for name in ("v", "fn", "area", "vn"):
        a, b = tri[name], otri[name]
        same = (a == b) | (np.isnan(a) & np.isnan(b))
        print(name, "bit-equal", bool(same.all()), "max abs", float(np.nanmax(np.abs(a - b))))

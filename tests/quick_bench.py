import sys, time
sys.path.insert(0, '.')
import numpy as np
from basicrenderer_b200 import load, meshgen
lib = load(0)
for n in (330, 707, 2236):
    m = meshgen.grid(n, seed=1234)
    w = np.ones(3, np.float32)
    h = lib.upload_mesh(m.positions, m.indices, attributes=m.normals, attribute_weights=w, protect_mask=7)
    for it in range(2):
        t = time.time(); rec = lib.build_dag_resident(h, keep_indices=False); dt = time.time() - t
        print(m.name, m.triangle_count, 'tris', 'time %.3f s' % dt, '%.2f Mtris/s' % (m.triangle_count / dt / 1e6), 'levels', rec.levels, 'groups', rec.groups, 'clusters', rec.total_clusters, 'passes', rec.simplify_passes, 'rounds', rec.simplify_rounds, 'launches', rec.launches, flush=True)
    print(' tris', rec.level_triangles.tolist(), ' groups', rec.level_groups.tolist())
    lib.free_mesh(h)

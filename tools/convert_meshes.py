"""Convert the reference's dolfin-XML meshes into compact ``.npz`` fixtures.

Run once in the build container (the reference tree is not present on the GPU
boxes):  ``python tools/convert_meshes.py [/root/reference/tests/mesh]``.
Only vertex coordinates and cell connectivity are stored (float64 / int32);
facet-region files are *not* converted -- boundary parts are classified
geometrically and the marker histograms of the
facet-region files are stored as ``facet_hist`` for the count check.
"""
import gzip
import os
import re
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
from dolfin_navier_scipy_b200.fem import read_dolfin_xml  # noqa: E402

src = sys.argv[1] if len(sys.argv) > 1 else '/root/reference/tests/mesh'
dst = os.path.join(os.path.dirname(__file__), '..',
                   'dolfin_navier_scipy_b200', 'mesh')
os.makedirs(dst, exist_ok=True)
names = ['cylinder_%d.xml' % k for k in range(5)] + \
    ['karman2D-rotcyl_lvl%d.xml.gz' % k for k in range(1, 5)] + \
    ['karman2D-outlets_lvl%d.xml.gz' % k for k in range(1, 3)] + \
    ['2D-double-rotcyl_lvl%d.xml.gz' % k for k in range(1, 3)]
for n in names:
    m = read_dolfin_xml(os.path.join(src, n))
    base = n.replace('.xml.gz', '').replace('.xml', '')
    extra = {}
    fr = os.path.join(src, base + '_facet_region.xml.gz')
    if os.path.isfile(fr):
        with gzip.open(fr, 'rt') as f:
            vals = np.array([int(v) for v in
                             re.findall(r'value="(\d+)"', f.read())])
        extra['facet_hist'] = np.bincount(vals)
    np.savez_compressed(os.path.join(dst, base + '.npz'),
                        coords=m.coords, cells=m.cells, **extra)
    print(base, m.num_vertices, m.num_cells, m.num_edges,
          extra.get('facet_hist'))
import json  # noqa: E402
for j in os.listdir(src):
    if j.endswith('.json'):
        # geometry/boundary descriptions: re-serialised (sorted, compact)
        with open(os.path.join(src, j)) as f:
            geo = json.load(f)
        with open(os.path.join(dst, j.replace('_cntrlbc', '')), 'w') as g:
            json.dump(geo, g, sort_keys=True, separators=(',', ':'))

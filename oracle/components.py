"""CPU oracle of the object labelling (flood_fill_3d, /root/reference/handy_utils.py:295-480, with an untrained
in-situ model).  TEST INFRASTRUCTURE ONLY.

Parity status: pinned - tests/golden/objects.npz holds the output of the unmodified reference flood_fill_3d
(run in the build container through tests/golden/_ref_loader.py) on a random class grid.

Restatement: scipy.ndimage.label with the full 3x3x3 structuring element per class id, objects smaller than 3
voxels dropped, the rest numbered -2, -3, ... by their first voxel in scan (flat-index) order."""
import numpy as np
from scipy import ndimage


def label_objects(class_grid, null_class=133, min_voxels=3):
    grid = np.asarray(class_grid)
    out = np.full(grid.shape, -1, np.int32)
    firsts = []   # (first flat index, class, component mask)
    structure = np.ones((3, 3, 3), bool)
    for c in np.unique(grid):
        if c == -1 or c == null_class:
            continue
        lab, n = ndimage.label(grid == c, structure=structure)
        flat = lab.reshape(-1)
        for k in range(1, n + 1):
            idx = np.flatnonzero(flat == k)
            if len(idx) >= min_voxels:
                firsts.append((int(idx[0]), idx))
    firsts.sort(key=lambda t: t[0])
    for rank, (_, idx) in enumerate(firsts):
        out.reshape(-1)[idx] = -2 - rank
    return out, len(firsts)

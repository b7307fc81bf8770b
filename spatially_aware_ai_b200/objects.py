"""Object labelling of the fused class grid: the flood fill of ``flood_fill_3d``
(/root/reference/handy_utils.py:295-480) and the bookkeeping of ``add_object`` (:244-292), for the case the
reference runs first - an untrained in-situ model (`insitu_model.model_trained` false), i.e. pure geometry.

The reference's triple Python loop over every voxel is replaced by a device connected-component labelling
(csrc/saf_components.cu); the per-object dictionary is assembled on the host from its result.
"""
import ctypes

import numpy as np

NULL_CLASS = 133   # handy_utils.py:381: "null class is 133, empty voxels are -1"
MIN_VOXELS = 3     # handy_utils.py:389: "reject small objects less than 3 voxels"


def label_objects(class_grid, null_class=NULL_CLASS, min_voxels=MIN_VOXELS):
    """class_grid: CUDA tensor [nx,ny,nz] of class ids (-1 = unobserved), e.g.
    ``volume.label_argmax().view(*nvox)``.  Returns (voxel_obj_ids int32 CUDA tensor [nx,ny,nz], n_objects):
    26-connected components of equal class, objects smaller than `min_voxels` rejected, numbered -2, -3, ... in
    the reference's scan order, -1 elsewhere."""
    import torch
    from . import _lib
    if not class_grid.is_cuda:
        raise RuntimeError("label_objects runs on a CUDA (sm_100) device only; there is no CPU path")
    if class_grid.dim() != 3:
        raise ValueError("class_grid must be [nx, ny, nz]")
    dev = class_grid.device
    labels = class_grid.to(torch.int64).contiguous()
    nx, ny, nz = (int(v) for v in labels.shape)
    lib = _lib.load()
    nbytes = ctypes.c_uint64()
    _lib.check(lib.saf_label_components_workspace_bytes(labels.numel(), ctypes.byref(nbytes)),
               "saf_label_components_workspace_bytes")
    scratch = torch.empty(nbytes.value + 256, dtype=torch.uint8, device=dev)
    base = (scratch.data_ptr() + 255) // 256 * 256
    out = torch.empty((nx, ny, nz), dtype=torch.int32, device=dev)
    n_obj = ctypes.c_uint32()
    with torch.cuda.device(dev):
        _lib.check(lib.saf_label_components(labels.data_ptr(), nx, ny, nz, int(null_class), int(min_voxels), out.data_ptr(),
                                        base, nbytes.value, ctypes.byref(n_obj),
                                        torch.cuda.current_stream(dev).cuda_stream), "saf_label_components")
    del scratch
    return out, int(n_obj.value)


def build_scene_knowledge(voxel_obj_ids, class_grid, class_names, class_colors, scene_knowledge=None):
    """unique_objects / object_counts as flood_fill_3d + add_object build them for an untrained in-situ model
    (handy_utils.py:244-292, 351-480).  `voxels` lists are in ascending voxel order (the reference's are in its
    flood-fill stack order; same sets)."""
    ids = voxel_obj_ids.detach().cpu().numpy() if hasattr(voxel_obj_ids, "detach") else np.asarray(voxel_obj_ids)
    classes = class_grid.detach().cpu().numpy() if hasattr(class_grid, "detach") else np.asarray(class_grid)
    flat = ids.reshape(-1)
    sel = np.flatnonzero(flat < -1)
    order = np.argsort(-flat[sel].astype(np.int64), kind="stable")      # -2, -3, ... ; voxels ascending within
    sel = sel[order]
    keys = -flat[sel].astype(np.int64) - 2
    bounds = np.flatnonzero(np.diff(keys)) + 1
    unique_objects, object_counts = {}, {}
    for group in np.split(sel, bounds) if len(sel) else []:
        coords = np.stack(np.unravel_index(group, ids.shape), axis=1)
        class_id = int(classes.reshape(-1)[group[0]])
        class_label = class_names[class_id]
        object_counts[class_label] = object_counts.get(class_label, 0) + 1      # get_obj_counts (:483-498)
        obj_id = "%s:%d" % (class_label, object_counts[class_label])
        unique_objects[obj_id] = {
            "class_id": class_id, "class_label": class_label, "voxels": [tuple(int(c) for c in v) for v in coords],
            "object_index": int(flat[group[0]]), "gt_label": obj_id, "user_modified": False,
            "merged": "merged" in class_label, "removed": False, "color": class_colors[class_id],
        }
    if scene_knowledge is None:
        scene_knowledge = {}
    scene_knowledge["unique_objects"] = unique_objects
    scene_knowledge["object_counts"] = object_counts
    scene_knowledge["unchanged_objects"] = {}
    scene_knowledge["new_objects"] = {}
    scene_knowledge["missing_objects"] = {}
    return scene_knowledge

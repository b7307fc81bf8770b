"""ctypes binding of libsaf_b200.so (include/saf_b200.h).

There is deliberately no fallback: if the shared library is missing the import of any compute
entry point raises, and on a machine without an sm_100 GPU every call returns SAF_ERR_DEVICE /
a CUDA error which is raised as RuntimeError.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# SAF_LIB_PATH: load another build of the same ABI (A/B timing of kernel variants)
LIB_PATH = os.environ.get("SAF_LIB_PATH") or os.path.join(HERE, "libsaf_b200.so")

SAF_ABI_VERSION = 3
SAF_MAX_BATCH = 16

SAF_SEG_NONE, SAF_SEG_U8, SAF_SEG_I16, SAF_SEG_I32, SAF_SEG_I64, SAF_SEG_F32 = range(6)
SAF_RGB_NEAREST, SAF_RGB_BILINEAR = 0, 1
SAF_STAGE_TILE_SETUP, SAF_STAGE_ACCUMULATE = 1, 2     # saf_feature_accumulate_window_stages
SAF_DEPTH_F32, SAF_DEPTH_U16_MM = 0, 1
SAF_RGB_F32, SAF_RGB_U8 = 0, 1
SAF_TABLE_PATCH_GRID, SAF_TABLE_SEGMENTS = 0, 1
SAF_BLOCK_EDGE = 8
SAF_FLAG_BAD_CLASS_ID = 1
SAF_NORM_NONE, SAF_NORM_NAN_TO_NUM, SAF_NORM_CLAMP_MIN = 0, 1, 2
SAF_SCORE_DOT, SAF_SCORE_SOFTMAX100, SAF_SCORE_SURGERY = 0, 1, 2
SAF_PRECISION_FP32, SAF_PRECISION_TF32 = 0, 1
SAF_SAMPLE_TRILINEAR, SAF_SAMPLE_NEAREST = 0, 1

c_void_p, c_int32, c_int64, c_uint64, c_float = (ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64,
                                                 ctypes.c_uint64, ctypes.c_float)


class GridDesc(ctypes.Structure):
    _fields_ = [("origin", c_float * 3), ("voxel_size", c_float), ("nvox", c_int32 * 3),
                ("x_begin", c_int32), ("x_end", c_int32), ("x_span", c_int32), ("x_stride", c_int32),
                ("y_ranks", c_int32), ("y_rank", c_int32)]


class Volume(ctypes.Structure):
    _fields_ = [("tsdf", c_void_p), ("tsdf_weight", c_void_p), ("weight", c_void_p), ("rgb", c_void_p),
                ("clip_feat", c_void_p), ("labels_one_hot", c_void_p), ("feature_dim", c_int32),
                ("n_classes", c_int32)]


class Frame(ctypes.Structure):
    _fields_ = [("depth", c_void_p), ("rgb", c_void_p), ("seg", c_void_p), ("table", c_void_p),
                ("table_stride_c", c_int64), ("table_stride_r", c_int64), ("npy", c_int32), ("npx", c_int32),
                ("seg_dtype", c_int32), ("depth_dtype", c_int32), ("pose", c_float * 16), ("K", c_float * 9),
                ("rgb_dtype", c_int32), ("pose_device", c_void_p), ("K_device", c_void_p), ("table_mode", c_int32),
                ("reserved", c_int32)]


def frame_numpy_dtype():
    """numpy structured dtype with saf_frame's layout: lets callers fill an array of frames column-wise."""
    import numpy as np
    names, formats, offsets = [], [], []
    for name, ctype in Frame._fields_:
        names.append(name)
        offsets.append(getattr(Frame, name).offset)
        if name == "pose":
            formats.append(("<f4", (16,)))
        elif name == "K":
            formats.append(("<f4", (9,)))
        elif ctype is c_void_p:
            formats.append("<u8")
        elif ctype is c_int64:
            formats.append("<i8")
        else:
            formats.append("<i4")
    return np.dtype({"names": names, "formats": formats, "offsets": offsets, "itemsize": ctypes.sizeof(Frame)})


class Stats(ctypes.Structure):
    _fields_ = [("total_frames", c_uint64), ("total_valid", c_uint64), ("total_tsdf_valid", c_uint64),
                ("total_blocks", c_uint64), ("last_blocks", ctypes.c_uint32),
                ("last_valid", ctypes.c_uint32 * SAF_MAX_BATCH), ("last_tsdf_valid", ctypes.c_uint32 * SAF_MAX_BATCH),
                ("error_flags", ctypes.c_uint32), ("last_processed", ctypes.c_uint32),
                ("depth_cull_on", ctypes.c_uint32), ("total_calls", c_uint64), ("total_union", c_uint64),
                ("last_union", ctypes.c_uint32), ("reserved", ctypes.c_uint32)]


class Workspace(ctypes.Structure):
    _fields_ = [("base", c_void_p), ("bytes", c_uint64), ("max_batch", c_int32), ("reserved", c_int32),
                ("max_table_elems", c_int64)]


P = ctypes.POINTER

# name -> (restype, argtypes); every symbol include/saf_b200.h declares
SIGNATURES = {
    "saf_abi_version": (ctypes.c_int, []),
    "saf_error_string": (ctypes.c_char_p, [ctypes.c_int]),
    "saf_workspace_bytes": (ctypes.c_int, [P(GridDesc), c_int32, c_int64, P(c_uint64)]),
    "saf_workspace_init": (ctypes.c_int, [P(Workspace), P(GridDesc), c_void_p]),
    "saf_read_stats": (ctypes.c_int, [P(Workspace), P(Stats), c_void_p]),
    "saf_frustum_cull": (ctypes.c_int, [P(GridDesc), P(Frame), c_int32, c_int32, c_int32, c_float, P(Workspace),
                                        c_void_p]),
    "saf_tsdf_update": (ctypes.c_int, [P(GridDesc), P(Volume), P(Frame), c_int32, c_int32, c_int32, c_float,
                                       P(Workspace), c_void_p, c_void_p, c_void_p]),
    "saf_feature_accumulate": (ctypes.c_int, [P(GridDesc), P(Volume), P(Frame), c_int32, c_int32, c_int32, c_int32,
                                              c_int32, P(Workspace), c_void_p]),
    "saf_tsdf_update_window": (ctypes.c_int, [P(GridDesc), P(Volume), P(Frame), c_int32, c_int32, c_int32, c_float,
                                              P(Workspace), c_void_p]),
    "saf_feature_accumulate_window": (ctypes.c_int, [P(GridDesc), P(Volume), P(Frame), c_int32, c_int32, c_int32,
                                                     c_int32, P(Workspace), c_void_p]),
    "saf_feature_accumulate_window_stages": (ctypes.c_int, [P(GridDesc), P(Volume), P(Frame), c_int32, c_int32, c_int32,
                                                            c_int32, P(Workspace), c_int32, c_void_p]),
    "saf_integrate": (ctypes.c_int, [P(GridDesc), P(Volume), P(Frame), c_int32, c_int32, c_int32, c_float, c_int32,
                                     P(Workspace), c_void_p]),
    "saf_frame_reaches_slab": (ctypes.c_int, [P(GridDesc), P(c_float), P(c_float), c_int32, c_int32]),
    "saf_integrate_sequence": (ctypes.c_int, [P(GridDesc), P(Volume), P(Frame), c_int32, c_int32, c_int32, c_float,
                                              c_int32, P(Workspace), c_void_p]),
    "saf_label_argmax": (ctypes.c_int, [c_void_p, c_int64, c_int32, c_void_p, c_void_p]),
    "saf_query_scores": (ctypes.c_int, [c_void_p, c_int64, c_int32, c_int64, c_void_p, c_int32, c_int32, c_int32,
                                        c_void_p, c_int32, c_void_p, c_void_p]),
    "saf_query_topk_workspace_bytes": (ctypes.c_int, [c_int64, c_int32, c_int32, P(c_uint64)]),
    "saf_query_topk": (ctypes.c_int, [c_void_p, c_int64, c_int32, c_int64, c_void_p, c_int32, c_int32, c_int32,
                                      c_void_p, c_int32, c_int32, c_int64, c_void_p, c_void_p, c_void_p, c_uint64,
                                      c_void_p]),
    "saf_query_rows_workspace_bytes": (ctypes.c_int, [c_int32, P(c_uint64)]),
    "saf_query_row_labels": (ctypes.c_int, [c_void_p, c_int64, c_int32, c_int64, c_void_p, c_int32, c_int32, c_int32,
                                            c_int32, c_void_p, c_void_p, c_void_p, c_uint64, c_void_p]),
    "saf_query_text_presence": (ctypes.c_int, [c_void_p, c_int64, c_int32, c_int64, c_void_p, c_int32, c_int32, c_int32,
                                               c_int32, c_void_p, c_void_p, c_uint64, c_void_p]),
    "saf_label_components_workspace_bytes": (ctypes.c_int, [c_int64, P(c_uint64)]),
    "saf_label_components": (ctypes.c_int, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p,
                                            c_uint64, P(ctypes.c_uint32), c_void_p]),
    "saf_backproject_samples": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                                               c_int32, c_int32, c_float, c_void_p, c_void_p, c_void_p]),
    "saf_mesh_workspace_bytes": (ctypes.c_int, [P(GridDesc), P(c_uint64)]),
    "saf_mesh_count": (ctypes.c_int, [P(GridDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_uint64,
                                      P(c_uint64), P(c_uint64), c_void_p]),
    "saf_mesh_emit": (ctypes.c_int, [P(GridDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_uint64, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_void_p]),
    "saf_mesh_sample": (ctypes.c_int, [P(GridDesc), c_void_p, c_int64, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                                       c_void_p, c_void_p]),
}

_lib = None


def load():
    """Load libsaf_b200.so (raises if it has not been built: there is no other code path)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libsaf_b200.so is missing (%s). Build it with `python -m spatially_aware_ai_b200.build`; "
                "this package has no CPU or PyTorch fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        if lib.saf_abi_version() != SAF_ABI_VERSION:
            raise RuntimeError("libsaf_b200.so ABI version mismatch")
        _lib = lib
    return _lib


def check(rc, what):
    if rc != 0:
        msg = load().saf_error_string(rc).decode()
        raise RuntimeError("%s failed: %s (code %d)" % (what, msg, rc))

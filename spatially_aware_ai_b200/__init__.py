"""B200-native RGB-D fusion + language query: a drop-in for the hot path of cy-xu/spatially_aware_AI.

Public surface (same names as the reference's):
    ClipSeemFusion, ClipFusion          volumes with integrate() / extract_mesh() and the reference's buffers
    Clip                                run_query / clip_feature_surgery / text_inference wrapper
    extract_mesh_by_object, extract_mesh_by_id
    label_objects, build_scene_knowledge  object labelling (the flood fill of handy_utils.flood_fill_3d)
    save_state, load_state              grid files in the reference's .npy formats + full state for re-fusion
plus the functional query API (query_scores, query_topk, surgery_weights, relevance_*, segment_labels - the
scorer of eval_scannet_segmentation.segment() - and presence_scores - the one of hypersim_eval).
All compute goes through libsaf_b200.so (include/saf_b200.h); there is no CPU or PyTorch fallback.
"""
from .checkpoint import load_state, save_state  # noqa: F401
from .fusion import ClipFusion, ClipSeemFusion  # noqa: F401
from .mesh import extract_mesh_by_id, extract_mesh_by_object  # noqa: F401
from .objects import build_scene_knowledge, label_objects  # noqa: F401
from .query import (Clip, minmax_per_text, presence_scores, query_scores, query_topk, relevance_half,  # noqa: F401
                    relevance_minmax, relevance_outliers, segment_labels, surgery_weights)

__all__ = ["ClipSeemFusion", "ClipFusion", "Clip", "extract_mesh_by_object", "extract_mesh_by_id", "label_objects",
           "build_scene_knowledge", "save_state", "load_state", "query_scores",
           "query_topk", "surgery_weights", "relevance_minmax", "relevance_half", "relevance_outliers",
           "minmax_per_text", "segment_labels", "presence_scores"]

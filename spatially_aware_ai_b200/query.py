"""Natural-language query over fused features: the reference's ``Clip.run_query`` /
``Clip.clip_feature_surgery`` (/root/reference/clipfusion.py:899-934) and the query drivers'
normalisation and post-processing (clip_seem_fusion.py:507-533, query_mesh.py:24-73,
eval_scannet_segmentation.py:546-561, hypersim_eval.py:50-51,80-89), computed by the kernels in
libsaf_b200.so.  The operand is any row-major [M,C] fp32 matrix on the GPU: mesh-vertex features
as in the reference, or the voxel grid's ``clip_feat`` itself.
"""
import ctypes

import torch

from . import _lib

NORM_MODES = {None: _lib.SAF_NORM_NONE, "none": _lib.SAF_NORM_NONE, "nan_to_num": _lib.SAF_NORM_NAN_TO_NUM,
              "clamp_min": _lib.SAF_NORM_CLAMP_MIN}
SCORE_MODES = {"dot": _lib.SAF_SCORE_DOT, "softmax100": _lib.SAF_SCORE_SOFTMAX100, "surgery": _lib.SAF_SCORE_SURGERY}
PRECISIONS = {"fp32": _lib.SAF_PRECISION_FP32, "tf32": _lib.SAF_PRECISION_TF32, "3xtf32": _lib.SAF_PRECISION_3XTF32}


def _prep(feats, text):
    if not feats.is_cuda:
        raise RuntimeError("query features must be on a CUDA (sm_100) device; there is no CPU path")
    if feats.dim() != 2 or text.dim() != 2:
        raise ValueError("feats must be [M,C] and text [T,C]")
    if feats.dtype != torch.float32:
        feats = feats.to(torch.float32)
    if feats.stride(1) != 1:
        feats = feats.contiguous()
    C = feats.shape[1]
    text = text.to(device=feats.device, dtype=torch.float32)[:, :C].contiguous()
    return feats, text


def surgery_weights(row0, text):
    """w = T * softmax(2 * <row0, X_t>) (clipfusion.py:913-915).  row0: the first feature row [C]
    (already normalised), text [T,C].  A zero row gives w == 1."""
    feats, text = _prep(row0.reshape(1, -1), text)
    s0 = query_scores(feats, text, norm=None, mode="dot", precision="fp32")[0]
    prob = (s0 * 2).softmax(-1)
    return prob / prob.mean(-1, keepdim=True)


def query_scores(feats, text, norm=None, mode="dot", surgery_w=None, precision="fp32", out=None):
    """[M,T] scores of feature rows against text embeddings.

    norm: None | "nan_to_num" | "clamp_min" row normalisation fused into the kernel;
    mode: "dot" | "softmax100" | "surgery"; precision: "fp32" | "tf32" | "3xtf32"."""
    feats, text = _prep(feats, text)
    M, C = feats.shape
    T = text.shape[0]
    if out is None:
        out = torch.empty((M, T), dtype=torch.float32, device=feats.device)
    w_ptr = None
    if mode == "surgery":
        if surgery_w is None:
            raise ValueError("mode='surgery' needs surgery_w (see surgery_weights)")
        surgery_w = surgery_w.to(device=feats.device, dtype=torch.float32).contiguous()
        w_ptr = surgery_w.data_ptr()
    stream = torch.cuda.current_stream(feats.device).cuda_stream
    rc = _lib.load().saf_query_scores(feats.data_ptr(), M, C, feats.stride(0), text.data_ptr(), T, NORM_MODES[norm],
                                      SCORE_MODES[mode], w_ptr, PRECISIONS[precision], out.data_ptr(), stream)
    _lib.check(rc, "saf_query_scores")
    return out


def query_topk(feats, text, k, norm=None, mode="dot", surgery_w=None, precision="fp32", index_base=0):
    """Top-k rows per text: (scores [T,k] descending, indices [T,k] int64), ties to the lower row.
    The [M,T] score matrix is never materialised."""
    feats, text = _prep(feats, text)
    M, C = feats.shape
    T = text.shape[0]
    lib = _lib.load()
    nbytes = ctypes.c_uint64()
    _lib.check(lib.saf_query_topk_workspace_bytes(M, T, k, ctypes.byref(nbytes)), "saf_query_topk_workspace_bytes")
    ws = torch.empty(nbytes.value + 256, dtype=torch.uint8, device=feats.device)
    base = (ws.data_ptr() + 255) // 256 * 256
    out_s = torch.empty((T, k), dtype=torch.float32, device=feats.device)
    out_i = torch.empty((T, k), dtype=torch.int64, device=feats.device)
    w_ptr = None
    if mode == "surgery":
        if surgery_w is None:
            raise ValueError("mode='surgery' needs surgery_w (see surgery_weights)")
        surgery_w = surgery_w.to(device=feats.device, dtype=torch.float32).contiguous()
        w_ptr = surgery_w.data_ptr()
    stream = torch.cuda.current_stream(feats.device).cuda_stream
    rc = lib.saf_query_topk(feats.data_ptr(), M, C, feats.stride(0), text.data_ptr(), T, NORM_MODES[norm],
                            SCORE_MODES[mode], w_ptr, PRECISIONS[precision], k, index_base, out_s.data_ptr(),
                            out_i.data_ptr(), base, nbytes.value, stream)
    _lib.check(rc, "saf_query_topk")
    return out_s, out_i


# ---- post-processing the reference's drivers apply to a relevance column -----------------------

def relevance_minmax(rel):
    """clip_seem_fusion.py:527-533: subtract the mean, clip to [0,1], min-max normalise."""
    rel = rel - rel.mean()
    rel = rel.clamp(0, 1)
    return (rel - rel.min()) / (rel.max() - rel.min())


def relevance_half(rel):
    """query_mesh.py:39."""
    return ((rel - 0.5) * 2).clamp(0, 1)


def relevance_outliers(sim):
    """query_mesh.py:59-73: per-text min-max, keep entries above median + 2 sigma.  sim [M,T] -> bool [M,T]."""
    mn, mx = sim.min(dim=0, keepdim=True).values, sim.max(dim=0, keepdim=True).values
    rel = (sim - mn) / (mx - mn)
    thr = rel.median(dim=0, keepdim=True).values + 2 * rel.std(dim=0, keepdim=True)
    return rel > thr


class Clip(torch.nn.Module):
    """Mirror of the reference's Clip wrapper (clipfusion.py:766-1039).  The image/text encoders
    are third-party DNN inference (open_clip) and out of this package's scope: pass `backend`, any
    object with `encode_text(tokens)`, `tokenizer(str_list)`, `feature_dim` and optionally
    `img_inference_tiled`; with backend=None the reference's open_clip construction is attempted."""

    def __init__(self, clip_model=None, pretraining=None, backend=None):
        super().__init__()
        if backend is None:
            try:
                import open_clip
            except ImportError as e:
                raise ImportError("open_clip is not installed; pass backend= with encode_text/tokenizer") from e
            self.clip = open_clip.create_model(clip_model, pretrained=pretraining, require_pretrained=True)
            self.tokenizer = open_clip.get_tokenizer(clip_model)
            self.feature_dim = self.clip.visual.output_dim
        else:
            self.clip = backend
            self.tokenizer = backend.tokenizer
            self.feature_dim = backend.feature_dim

    def img_inference_tiled(self, rgb_imgs, patch_size, patch_stride):
        return self.clip.img_inference_tiled(rgb_imgs, patch_size, patch_stride)

    def text_inference(self, str_list):
        """clipfusion.py:892-897: unit-norm text embeddings [L, C]."""
        feats = self.clip.encode_text(self.tokenizer(str_list))
        return feats / feats.norm(dim=-1, keepdim=True)

    def run_query(self, img_feats, labels, precision="fp32"):
        """clipfusion.py:899-904: softmax(100 * img_feats @ text^T) over the labels."""
        text = self.text_inference(labels)[:, : img_feats.shape[-1]]
        lead = img_feats.shape[:-1]
        out = query_scores(img_feats.reshape(-1, img_feats.shape[-1]), text, mode="softmax100", precision=precision)
        return out.view(*lead, text.shape[0])

    @staticmethod
    def clip_feature_surgery(image_features, text_features, redundant_feats=None, t=2, precision="fp32"):
        """clipfusion.py:906-934.  image_features [b,M,C] (caller-normalised), text_features [T,C]
        -> [b,M,T].  Evaluated as a GEMM plus row epilogue instead of the reference's [b,M,T,C]
        broadcast product; `t` is unused there as well."""
        if image_features.dim() != 3:
            raise ValueError("image_features must be [b, M, C]")
        outs = []
        for b in range(image_features.shape[0]):
            F = image_features[b]
            if redundant_feats is not None:
                outs.append(query_scores(F, text_features.to(F.device) - redundant_feats.to(F.device), mode="dot",
                                         precision=precision))
            else:
                w = surgery_weights(F[0], text_features)
                outs.append(query_scores(F, text_features, mode="surgery", surgery_w=w, precision=precision))
        return torch.stack(outs, dim=0)

    def encode_text_with_prompt_ensemble(self, texts, device, prompt_templates=None):
        """clipfusion.py:936-1039: mean of unit embeddings over prompt templates, re-normalised."""
        if prompt_templates is None:
            prompt_templates = DEFAULT_PROMPT_TEMPLATES
        feats = []
        for t in texts:
            emb = self.clip.encode_text(self.tokenizer([tpl.format(t) for tpl in prompt_templates]))
            emb = emb / emb.norm(dim=-1, keepdim=True)
            emb = emb.mean(dim=0)
            feats.append(emb / emb.norm())
        return torch.stack(feats, dim=1).to(device).t()


# the reference's default ensemble is the 85-template ImageNet list (clipfusion.py:939-1025); its
# own query driver passes ["a photo of {}"] (clip_seem_fusion.py:496-505), which is the default here.
DEFAULT_PROMPT_TEMPLATES = ["a photo of {}"]

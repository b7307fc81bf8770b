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
PRECISIONS = {"fp32": _lib.SAF_PRECISION_FP32, "tf32": _lib.SAF_PRECISION_TF32}


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
    mode: "dot" | "softmax100" | "surgery"; precision: "fp32" | "tf32"."""
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
    with torch.cuda.device(feats.device):
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
    with torch.cuda.device(feats.device):
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


def minmax_per_text(sim):
    """query_mesh.py:59-61: per-text min-max normalisation of a similarity block [..., M, T] over the rows."""
    mn, mx = sim.min(dim=-2, keepdim=True).values, sim.max(dim=-2, keepdim=True).values
    return (sim - mn) / (mx - mn)


def relevance_outliers(rel):
    """query_mesh.py:63-73 for one relevance column [M] (already min-max normalised): entries above
    median + 2 sigma keep their value, the rest become 0."""
    thr = torch.median(rel) + 2 * torch.std(rel)
    return torch.where(rel > thr, rel, torch.zeros_like(rel))


def _rows_workspace(T, device):
    nbytes = ctypes.c_uint64()
    _lib.check(_lib.load().saf_query_rows_workspace_bytes(T, ctypes.byref(nbytes)), "saf_query_rows_workspace_bytes")
    ws = torch.empty(nbytes.value + 256, dtype=torch.uint8, device=device)
    return ws, (ws.data_ptr() + 255) // 256 * 256, nbytes.value


def segment_labels(feats, text, k=None, norm="clamp_min", precision="fp32", return_probs=False):
    """The scorer of segment() (eval_scannet_segmentation.py:546-561): for every feature row the texts ordered
    by descending softmax(100 * cos), as int64 [M, k].  The reference returns the full argsort (k = T, the
    default) and its evaluation reads columns 0 (top-1) and 0..4 (top-5); pass k to get only those without ever
    materialising [M, T] (M = 24 M voxels x T = 256 texts would be 24.6 GB).  norm="clamp_min" is segment()'s
    feat_norm.clamp_min_(0.1).  Ties go to the lower text index."""
    feats, text = _prep(feats, text)
    M, C = feats.shape
    T = text.shape[0]
    k = T if k is None else int(k)
    labels = torch.empty((M, k), dtype=torch.int64, device=feats.device)
    probs = torch.empty((M, k), dtype=torch.float32, device=feats.device) if return_probs else None
    ws, base, nbytes = _rows_workspace(T, feats.device)
    stream = torch.cuda.current_stream(feats.device).cuda_stream
    with torch.cuda.device(feats.device):
        rc = _lib.load().saf_query_row_labels(feats.data_ptr(), M, C, feats.stride(0), text.data_ptr(), T,
                                              NORM_MODES[norm], PRECISIONS[precision], k, labels.data_ptr(),
                                              probs.data_ptr() if return_probs else None, base, nbytes, stream)
    _lib.check(rc, "saf_query_row_labels")
    return (labels, probs) if return_probs else labels


def presence_scores(feats, background_text, target_text, norm="clamp_min", precision="fp32"):
    """hypersim_eval.py:80-89: for every target text i, max over the rows of softmax(100 * cos([background..,
    target_i]))[-1] - the number the reference compares with its 101 thresholds (`relevance.max() > thresholds`).
    background_text [nb,C] (the four "a picture of an object / things / stuff / texture" embeddings), target_text
    [L,C] -> float32 [L].  One pass over the features for all L targets instead of L passes."""
    feats, text = _prep(feats, torch.cat([background_text, target_text.to(background_text.device)], dim=0))
    M, C = feats.shape
    nb, L = background_text.shape[0], target_text.shape[0]
    out = torch.empty(L, dtype=torch.float32, device=feats.device)
    ws, base, nbytes = _rows_workspace(nb + L, feats.device)
    stream = torch.cuda.current_stream(feats.device).cuda_stream
    with torch.cuda.device(feats.device):
        rc = _lib.load().saf_query_text_presence(feats.data_ptr(), M, C, feats.stride(0), text.data_ptr(), nb, L,
                                                 NORM_MODES[norm], PRECISIONS[precision], out.data_ptr(), base, nbytes,
                                                 stream)
    _lib.check(rc, "saf_query_text_presence")
    return out


# CLIP's image normalisation constants (clipfusion.py:773-780)
CLIP_CHANNEL_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_CHANNEL_STD = (0.26862954, 0.26130258, 0.27577711)


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
        self.channel_mean = torch.nn.Parameter(torch.tensor(CLIP_CHANNEL_MEAN)[None, :, None, None], requires_grad=False)
        self.channel_std = torch.nn.Parameter(torch.tensor(CLIP_CHANNEL_STD)[None, :, None, None], requires_grad=False)

    def _model_device(self):
        params = list(self.clip.parameters()) if hasattr(self.clip, "parameters") else []
        return params[0].device if params else getattr(self.clip, "device", torch.device("cpu"))

    def normalize_img(self, rgb_img_0_1):
        """clipfusion.py:783-784."""
        return (rgb_img_0_1 - self.channel_mean.to(rgb_img_0_1.device)) / self.channel_std.to(rgb_img_0_1.device)

    def unnormalize_img(self, rgb_img_normed):
        """clipfusion.py:786-787."""
        return rgb_img_normed * self.channel_std.to(rgb_img_normed.device) + self.channel_mean.to(rgb_img_normed.device)

    def get_patches(self, rgb_imgs, patch_size, patch_stride):
        """clipfusion.py:789-806: [B,3,H,W] -> [B, npy, npx, 3, patch, patch] overlapping tiles."""
        batch_size, _, imheight, imwidth = rgb_imgs.shape
        assert (imheight - patch_size) % patch_stride == 0
        assert (imwidth - patch_size) % patch_stride == 0
        npatches_x = 1 + (imwidth - patch_size) // patch_stride
        npatches_y = 1 + (imheight - patch_size) // patch_stride
        # two unfolds are strided views of the image (no copy until the reshape below)
        tiles = rgb_imgs.unfold(2, patch_size, patch_stride).unfold(3, patch_size, patch_stride)
        assert tiles.shape[2] == npatches_y and tiles.shape[3] == npatches_x
        return tiles.permute(0, 2, 3, 1, 4, 5)

    def img_inference_tiled(self, rgb_imgs, patch_size, patch_stride):
        """clipfusion.py:808-839: normalise, cut into tiles, resize every tile to 224 x 224 (bilinear,
        align_corners=False), encode in batches of 8, return the feature image [B, C, npy, npx] as the permuted
        view of [B, npy, npx, C] (the memory layout the fusion kernels read without repacking).  A backend that
        brings its own img_inference_tiled is used instead."""
        if hasattr(self.clip, "img_inference_tiled"):
            return self.clip.img_inference_tiled(rgb_imgs, patch_size, patch_stride)
        rgb_imgs = self.normalize_img(rgb_imgs)
        patches = self.get_patches(rgb_imgs, patch_size, patch_stride)
        batch_size, npatches_y, npatches_x = patches.shape[:3]
        patches = patches.reshape(batch_size * npatches_y * npatches_x, 3, patch_size, patch_size)
        patches = torch.nn.functional.interpolate(patches, size=(224, 224), mode="bilinear", align_corners=False)
        clip_feats = torch.empty(len(patches), self.feature_dim, device=rgb_imgs.device)
        max_patch_batch_size = 8
        for start in range(0, len(patches), max_patch_batch_size):
            stop = min(len(patches), start + max_patch_batch_size)
            clip_feats[start:stop] = self.clip.encode_image(patches[start:stop])
        return clip_feats.view(batch_size, npatches_y, npatches_x, self.feature_dim).permute(0, 3, 1, 2)

    def text_inference(self, str_list):
        """clipfusion.py:891-897: unit-norm text embeddings [L, C] (tokens moved to the model's device)."""
        tokens = self.tokenizer(str_list)
        if hasattr(tokens, "to"):
            tokens = tokens.to(self._model_device())
        feats = self.clip.encode_text(tokens)
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
        model_device = self._model_device()
        for t in texts:
            tokens = self.tokenizer([tpl.format(t) for tpl in prompt_templates])
            if hasattr(tokens, "to"):
                tokens = tokens.to(model_device)
            emb = self.clip.encode_text(tokens)
            emb = emb / emb.norm(dim=-1, keepdim=True)
            emb = emb.mean(dim=0)
            feats.append(emb / emb.norm())
        return torch.stack(feats, dim=1).to(device).t()


from .prompt_templates import DEFAULT_PROMPT_TEMPLATES  # noqa: E402  (the reference's 85-template default)

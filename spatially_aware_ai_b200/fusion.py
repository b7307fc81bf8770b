"""Drop-in fusion volumes: same constructor / integrate / buffer surface as the reference's
``ClipSeemFusion`` (/root/reference/clip_seem_fusion.py:611-888) and ``ClipFusion``
(/root/reference/clipfusion.py:575-763), with the per-frame work done by the sm_100a kernels in
libsaf_b200.so.  PyTorch is only the owner of device memory and streams here.

Additions to the reference surface: ``integrate_sequence`` (the frame loop as one call, fused 16 frames at a time
on the device), ``label_argmax``, ``stats`` / ``check_errors``; ``extract_mesh`` runs on the device.

Differences a caller can observe:
  * tensors must live on a CUDA device (sm_100); there is no CPU path - it raises instead.
  * ``xyz_world`` is not stored (12 B/voxel re-read every frame in the reference); voxel centres
    are recomputed in-kernel with the reference's exact roundings.  The attribute still exists
    as a lazily computed property.
  * optional ``x_begin`` / ``x_end`` keyword arguments hold only an x-slab of the grid
    (multi-GPU partitioning); buffers then cover that slab, coordinates stay global.  ``x_span`` /
    ``x_stride`` make the slab block-cyclic (stripes of x_span planes every x_stride planes); ``y_ranks`` /
    ``y_rank`` select the sheared block-column layout (column (bx, by) of 8 x 8 x nz voxels on rank
    (bx + by) % y_ranks), which gives every rank the same share of every view AND of every axis-aligned wall.
  * class ids outside [0, n_classes) cannot raise inside a kernel; they set a sticky flag that
    ``check_errors()`` (and ``stats()``) turn into the RuntimeError torch's one_hot would raise.
"""
import ctypes

import numpy as np
import torch

from . import _lib

_FRAME_DTYPE = _lib.frame_numpy_dtype()
_SEG_DTYPES = {torch.uint8: _lib.SAF_SEG_U8, torch.int16: _lib.SAF_SEG_I16, torch.int32: _lib.SAF_SEG_I32,
               torch.int64: _lib.SAF_SEG_I64, torch.float32: _lib.SAF_SEG_F32}


def _as_int_list(nvox):
    if isinstance(nvox, torch.Tensor):
        return [int(v) for v in nvox.tolist()]
    return [int(v) for v in np.asarray(nvox).tolist()]


class _FusionVolume(torch.nn.Module):
    """State layout of clip_seem_fusion.py:640-672 plus the kernel plumbing shared by both classes."""

    _with_labels = True
    _rgb_mode = _lib.SAF_RGB_BILINEAR
    # "patch_grid": the reference's feature source (bilinear sample of the tiled-patch CLIP image).
    # "segment_table": one embedding per class id, a voxel takes the row of its nearest-sampled id
    # (BASELINE.json north_star / SURVEY.md 7.2; include/saf_b200.h SAF_TABLE_SEGMENTS).  Set on the instance.
    feature_source = "patch_grid"

    def _init_volume(self, origin, voxel_size, nvox, trunc, feature_dim, x_begin=0, x_end=None, x_span=0, x_stride=0,
                     y_ranks=0, y_rank=0):
        self.origin = origin
        self.voxel_size = voxel_size
        self.nvox = nvox
        self.trunc = trunc
        self.n_clip_feats = int(feature_dim)
        dims = _as_int_list(nvox)
        self.x_begin = int(x_begin)
        self.x_end = dims[0] if x_end is None else int(x_end)
        if not (0 <= self.x_begin < self.x_end <= dims[0]):
            raise ValueError("x-slab [%d,%d) outside the grid" % (self.x_begin, self.x_end))
        # block-cyclic slabs: stripes of x_span planes every x_stride planes (saf_grid_desc)
        self.x_span, self.x_stride = int(x_span), int(x_stride)
        if self.x_span and (self.x_span % _lib.SAF_BLOCK_EDGE or self.x_stride % _lib.SAF_BLOCK_EDGE or
                            self.x_stride < self.x_span or self.x_span < 0):
            raise ValueError("x_span / x_stride must be multiples of %d with x_stride >= x_span" % _lib.SAF_BLOCK_EDGE)
        # sheared block columns: 8 x 8 x nz columns, column (bx, by) on rank (bx + by) % y_ranks (saf_grid_desc)
        self.y_ranks, self.y_rank = int(y_ranks), int(y_rank)
        if self.y_ranks > 1 and (self.x_span or self.x_begin or self.x_end != dims[0] or not 0 <= self.y_rank < self.y_ranks):
            raise ValueError("the sheared block-column layout holds the whole x range and needs 0 <= y_rank < y_ranks")
        self._dims = dims
        n = len(self.global_x_planes()) * self.ny_local * dims[2]
        self.register_buffer("tsdf", torch.zeros(n, dtype=torch.float32))
        self.register_buffer("rgb", torch.zeros((n, 3), dtype=torch.float32))
        self.register_buffer("clip_feat", torch.zeros((n, self.n_clip_feats), dtype=torch.float32))
        self.register_buffer("weight", torch.zeros(n, dtype=torch.int32))
        self.register_buffer("tsdf_weight", torch.zeros(n, dtype=torch.int32))
        if self._with_labels:
            # 133 + 10 spare classes, clip_seem_fusion.py:653-659
            self.n_classes = 133 + 10
            self.register_buffer("labels_one_hot", torch.zeros((n, self.n_classes), dtype=torch.int32))
        self._ws_tensor = None
        self._ws = None
        self._ws_key = None
        self._grid = None

    # -- geometry ------------------------------------------------------------------------------

    def _grid_desc(self):
        if self._grid is None:
            g = _lib.GridDesc()
            origin = self.origin.detach().cpu().to(torch.float32).tolist() if isinstance(self.origin, torch.Tensor) \
                else np.asarray(self.origin, dtype=np.float32).tolist()
            g.origin[:] = origin
            g.voxel_size = float(self.voxel_size)
            g.nvox[:] = self._dims
            g.x_begin, g.x_end = self.x_begin, self.x_end
            g.x_span, g.x_stride = self.x_span, self.x_stride
            g.y_ranks, g.y_rank = (self.y_ranks, self.y_rank) if self.y_ranks > 1 else (0, 0)
            self._grid = g
        return self._grid

    @property
    def ny_local(self):
        """Rows of the slab's y axis: the grid's ny, or the sheared layout's 8 * ceil(n_y_blocks / y_ranks)."""
        if self.y_ranks <= 1:
            return self._dims[1]
        e = _lib.SAF_BLOCK_EDGE
        nby = (self._dims[1] + e - 1) // e
        return e * ((nby + self.y_ranks - 1) // self.y_ranks)

    def global_rows(self, device=None):
        """Global flat voxel index (x*ny + y)*nz + z of every row of the slab's buffers, int64 [n]; -1 for the
        sheared layout's padding rows (local columns beyond the grid)."""
        nx, ny, nz = self._dims
        e = _lib.SAF_BLOCK_EDGE
        xs = torch.as_tensor(self.global_x_planes(), dtype=torch.int64, device=device)
        ly = torch.arange(self.ny_local, dtype=torch.int64, device=device)
        if self.y_ranks > 1:
            lx = torch.arange(len(xs), dtype=torch.int64, device=device)
            shift = (self.y_rank - lx // e) % self.y_ranks
            gy = ((ly[None, :] // e) * self.y_ranks + shift[:, None]) * e + ly[None, :] % e       # [nx, ny_local]
        else:
            gy = ly[None, :].expand(len(xs), -1)
        rows = (xs[:, None] * ny + gy)[:, :, None] * nz + torch.arange(nz, dtype=torch.int64, device=device)
        rows = torch.where((gy < ny)[:, :, None], rows, torch.full_like(rows, -1))
        return rows.reshape(-1)

    def global_x_planes(self):
        """Global x index of every slab-local plane (a contiguous range, or the rank's block-cyclic stripes)."""
        if not self.x_span:
            return list(range(self.x_begin, self.x_end))
        out = []
        for start in range(self.x_begin, self.x_end, self.x_stride):
            out.extend(range(start, min(start + self.x_span, self.x_end)))
        return out

    @property
    def xyz_world(self):
        """World coordinates of the slab's voxel centres, computed like clip_seem_fusion.py:664-669."""
        dev = self.tsdf.device
        if self.y_ranks > 1:
            rows = self.global_rows(dev).clamp(min=0)     # padding rows report voxel 0's centre
            nx, ny, nz = self._dims
            xyz_idx = torch.stack((rows // (ny * nz), rows // nz % ny, rows % nz), dim=-1)
        else:
            x = torch.as_tensor(self.global_x_planes(), device=dev)
            y = torch.arange(self._dims[1], device=dev)
            z = torch.arange(self._dims[2], device=dev)
            xx, yy, zz = torch.meshgrid(x, y, z, indexing="ij")
            xyz_idx = torch.stack((xx, yy, zz), dim=-1).view(-1, 3)
        origin = self.origin if isinstance(self.origin, torch.Tensor) else torch.as_tensor(self.origin)
        return xyz_idx * self.voxel_size + origin.to(dev)

    # -- kernel plumbing -----------------------------------------------------------------------

    def _volume_desc(self):
        v = _lib.Volume()
        v.tsdf = self.tsdf.data_ptr()
        v.tsdf_weight = self.tsdf_weight.data_ptr()
        v.weight = self.weight.data_ptr()
        v.rgb = self.rgb.data_ptr()
        v.clip_feat = self.clip_feat.data_ptr()
        v.labels_one_hot = self.labels_one_hot.data_ptr() if self._with_labels else None
        v.feature_dim = self.n_clip_feats
        v.n_classes = self.n_classes if self._with_labels else 0
        return v

    def _workspace(self, batch, table_elems):
        dev = self.tsdf.device
        key = self._ws_key
        if key is not None and key[0] == dev and key[1] >= batch and key[2] >= table_elems:
            return self._ws
        lib = _lib.load()
        old_stats = self.stats(check=False) if self._ws is not None else None
        max_batch = max(batch, key[1] if key else 1)
        max_table = max(table_elems, key[2] if key else 0)
        nbytes = ctypes.c_uint64()
        _lib.check(lib.saf_workspace_bytes(ctypes.byref(self._grid_desc()), max_batch, max_table,
                                           ctypes.byref(nbytes)), "saf_workspace_bytes")
        self._ws_tensor = torch.empty(nbytes.value + 256, dtype=torch.uint8, device=dev)
        base = (self._ws_tensor.data_ptr() + 255) // 256 * 256
        ws = _lib.Workspace()
        ws.base, ws.bytes, ws.max_batch, ws.max_table_elems = base, nbytes.value, max_batch, max_table
        with torch.cuda.device(dev):     # the library launches on the CURRENT device (stream 0 is valid on all)
            _lib.check(lib.saf_workspace_init(ctypes.byref(ws), ctypes.byref(self._grid_desc()),
                                              torch.cuda.current_stream(dev).cuda_stream), "saf_workspace_init")
        self._ws, self._ws_key = ws, (dev, max_batch, max_table)
        self._stats_carry = old_stats
        return ws

    def _make_frames(self, depth_imgs, rgb_imgs, poses, K, clip_feat_img, seg_maps):
        B, H, W, _ = rgb_imgs.shape
        dev = self.tsdf.device
        table_mode = _lib.SAF_TABLE_PATCH_GRID
        if self.feature_source == "segment_table":
            # [B, n_segments, C] rows -> the [B, C, 1, n_segments] view the kernels index as a one-row patch grid
            if clip_feat_img.dim() != 3:
                raise RuntimeError("segment-table mode takes per-frame tables [B, n_segments, C]")
            if seg_maps is None:
                raise RuntimeError("segment-table mode needs the class maps")
            clip_feat_img = clip_feat_img.permute(0, 2, 1)[:, :, None, :]
            table_mode = _lib.SAF_TABLE_SEGMENTS
        for name, t in (("depth_imgs", depth_imgs), ("rgb_imgs", rgb_imgs), ("clip feature image", clip_feat_img)):
            if t.device != dev:
                raise RuntimeError("%s is on %s but the volume is on %s" % (name, t.device, dev))
        # sensor formats stay as they are (converted in-kernel with the dataset classes' roundings,
        # clipfusion.py:185-188): uint16 millimetres / uint8; anything else is handed over as fp32
        depth = depth_imgs.contiguous() if depth_imgs.dtype == torch.uint16 else depth_imgs.to(torch.float32).contiguous()
        rgb = rgb_imgs.contiguous() if rgb_imgs.dtype == torch.uint8 else rgb_imgs.to(torch.float32).contiguous()
        depth_dtype = _lib.SAF_DEPTH_U16_MM if depth.dtype == torch.uint16 else _lib.SAF_DEPTH_F32
        rgb_dtype = _lib.SAF_RGB_U8 if rgb.dtype == torch.uint8 else _lib.SAF_RGB_F32
        table = clip_feat_img[:, : self.n_clip_feats]
        if table.dtype != torch.float32:
            table = table.to(torch.float32)
        if table.shape[1] != self.n_clip_feats:
            raise RuntimeError("feature image has %d channels, volume needs %d" % (table.shape[1], self.n_clip_feats))
        _, _, npy, npx = table.shape
        sb, sc, sy, sx = table.stride()
        if npy > 1 and npx > 1 and sy != sx * npx:
            table = table.contiguous()
            sb, sc, sy, sx = table.stride()
        sr = sx if npx > 1 else (sy if npy > 1 else max(1, self.n_clip_feats if sc == 1 else 1))
        # the B descriptors are filled column-wise (a Python loop over frames costs more than a window's kernels)
        arr = np.zeros(B, dtype=_FRAME_DTYPE)
        steps = np.arange(B, dtype=np.uint64)
        arr["depth"] = depth.data_ptr() + steps * np.uint64(depth.stride(0) * depth.element_size())
        arr["rgb"] = rgb.data_ptr() + steps * np.uint64(rgb.stride(0) * rgb.element_size())
        arr["table"] = table.data_ptr() + steps * np.uint64(sb * table.element_size())
        arr["depth_dtype"], arr["rgb_dtype"] = depth_dtype, rgb_dtype
        arr["table_stride_c"], arr["table_stride_r"] = sc, sr
        arr["table_mode"] = table_mode
        arr["npy"], arr["npx"] = npy, npx
        keep = [depth, rgb, table, arr]
        if seg_maps is not None:
            if isinstance(seg_maps, torch.Tensor):       # stacked [B,H,W]
                segs = seg_maps if seg_maps.dtype in _SEG_DTYPES else seg_maps.to(torch.int64)
                segs = segs.contiguous()
                if segs.device != dev:
                    raise RuntimeError("segmentation maps are on %s but the volume is on %s" % (segs.device, dev))
                if tuple(segs.shape) != (B, H, W):
                    raise RuntimeError("segmentation maps shape %s != %s" % (tuple(segs.shape), (B, H, W)))
                keep.append(segs)
                arr["seg"] = segs.data_ptr() + steps * np.uint64(segs.stride(0) * segs.element_size())
                arr["seg_dtype"] = _SEG_DTYPES[segs.dtype]
            else:
                for b in range(B):
                    seg = seg_maps[b]
                    if seg.device != dev:
                        raise RuntimeError("segmentation map is on %s but the volume is on %s" % (seg.device, dev))
                    if seg.dtype not in _SEG_DTYPES:
                        seg = seg.to(torch.int64)
                    seg = seg.contiguous()
                    if tuple(seg.shape) != (H, W):
                        raise RuntimeError("segmentation map shape %s != image %s" % (tuple(seg.shape), (H, W)))
                    keep.append(seg)
                    arr["seg"][b] = seg.data_ptr()
                    arr["seg_dtype"][b] = _SEG_DTYPES[seg.dtype]
        if poses.is_cuda:
            poses_dev = poses.to(torch.float32).contiguous()
            K_dev = K.to(device=dev, dtype=torch.float32).contiguous()
            keep += [poses_dev, K_dev]
            arr["pose_device"] = poses_dev.data_ptr() + steps * np.uint64(64)
            arr["K_device"] = K_dev.data_ptr() + steps * np.uint64(36)
        else:
            arr["pose"] = poses.detach().to(torch.float32).contiguous().view(B, 16).numpy()
            arr["K"] = K.detach().cpu().to(torch.float32).contiguous().view(B, 9).numpy()
        frames = arr.ctypes.data_as(ctypes.POINTER(_lib.Frame))
        return frames, keep, (B, H, W, npy * npx * self.n_clip_feats)

    def _integrate_frames(self, depth_imgs, rgb_imgs, poses, K, clip_feat_img, seg_maps, sequence=False):
        if not self.tsdf.is_cuda:
            raise RuntimeError("spatially_aware_ai_b200 volumes run on a CUDA (sm_100) device only; "
                               "call .to('cuda') - there is no CPU path")
        frames, keep, (B, H, W, table_elems) = self._make_frames(depth_imgs, rgb_imgs, poses, K, clip_feat_img, seg_maps)
        vol = self._volume_desc()
        stream = torch.cuda.current_stream(self.tsdf.device).cuda_stream
        with torch.cuda.device(self.tsdf.device):
            self._launch_integrate(frames, vol, B, H, W, table_elems, stream, sequence)
        del keep

    def _launch_integrate(self, frames, vol, B, H, W, table_elems, stream, sequence):
        if sequence:
            # B successive single-frame calls; a max_batch = SAF_MAX_BATCH workspace lets the library fuse them 16 at a time
            ws = self._workspace(_lib.SAF_MAX_BATCH, table_elems)
            rc = _lib.load().saf_integrate_sequence(ctypes.byref(self._grid_desc()), ctypes.byref(vol), frames, B, H, W,
                                                    float(self.trunc), self._rgb_mode, ctypes.byref(ws), stream)
            _lib.check(rc, "saf_integrate_sequence")
        else:
            if B > _lib.SAF_MAX_BATCH:
                raise RuntimeError("batch of %d frames exceeds SAF_MAX_BATCH=%d" % (B, _lib.SAF_MAX_BATCH))
            ws = self._workspace(B, table_elems)
            rc = _lib.load().saf_integrate(ctypes.byref(self._grid_desc()), ctypes.byref(vol), frames, B, H, W,
                                           float(self.trunc), self._rgb_mode, ctypes.byref(ws), stream)
            _lib.check(rc, "saf_integrate")

    @staticmethod
    def _producer_inputs(depth_imgs, rgb_imgs):
        """The CLIP / segmentation producers take what the reference hands them: fp32 metres and fp32 [0,1]."""
        rgb_f = rgb_imgs.float() / 255 if rgb_imgs.dtype == torch.uint8 else rgb_imgs
        depth_f = depth_imgs.float() / 1000 if depth_imgs.dtype == torch.uint16 else depth_imgs
        return depth_f, rgb_f

    def integrate_sequence(self, depth_imgs, rgb_imgs, poses, K, clip_feat_img=None, seg_maps=None):
        """The reference's frame loop (clip_seem_fusion.py:305-313) as one call: same result as
        ``for i in range(F): self.integrate(depth_imgs[i:i+1], rgb_imgs[i:i+1], poses[i:i+1], K[i:i+1])``,
        but consecutive frames are fused 16 at a time on the device (each voxel's state is read and written
        once per window instead of once per frame).  The producers are called once with all F frames, unless
        their outputs are passed in: ``clip_feat_img`` [F,C,npy,npx] and (ClipSeemFusion) ``seg_maps``, a list
        of F class maps [H,W].

        ``depth_imgs`` may be uint16 millimetres and ``rgb_imgs`` uint8 - the formats the reference's dataset
        classes read from disk before ``float() / 1000`` and ``float() / 255`` (clipfusion.py:185-188); the
        kernels convert with those exact roundings, so the result equals feeding the converted fp32 tensors.
        (The producers are then handed ``rgb_imgs.float() / 255``.)"""
        if clip_feat_img is None:
            clip_feat_img, seg_auto = self._run_producers(*self._producer_inputs(depth_imgs, rgb_imgs), K)
            if seg_maps is None:
                seg_maps = seg_auto
        if not self._with_labels:
            seg_maps = None
        self._integrate_frames(depth_imgs, rgb_imgs, poses, K, clip_feat_img, seg_maps, sequence=True)

    # -- bookkeeping ---------------------------------------------------------------------------

    def stats(self, check=True):
        """Counters kept by the kernels: frames, sum of valid / tsdf_valid voxels, visible blocks."""
        if self._ws is None:
            return dict(total_frames=0, total_valid=0, total_tsdf_valid=0, total_blocks=0, total_calls=0, last_blocks=0,
                        last_valid=[], last_tsdf_valid=[], error_flags=0, total_union=0, last_union=0)
        st = _lib.Stats()
        stream = torch.cuda.current_stream(self.tsdf.device).cuda_stream
        with torch.cuda.device(self.tsdf.device):
            _lib.check(_lib.load().saf_read_stats(ctypes.byref(self._ws), ctypes.byref(st), stream), "saf_read_stats")
        out = dict(total_frames=st.total_frames, total_valid=st.total_valid, total_tsdf_valid=st.total_tsdf_valid,
                   total_blocks=st.total_blocks, last_blocks=st.last_blocks, last_valid=list(st.last_valid),
                   last_tsdf_valid=list(st.last_tsdf_valid), error_flags=st.error_flags,
                   last_processed=st.last_processed, depth_cull_on=st.depth_cull_on, total_calls=st.total_calls,
                   total_union=st.total_union, last_union=st.last_union)
        carry = getattr(self, "_stats_carry", None)
        if carry:
            for k in ("total_frames", "total_valid", "total_tsdf_valid", "total_blocks", "total_calls", "total_union"):
                out[k] += carry[k]
            out["error_flags"] |= carry["error_flags"]
        if check and out["error_flags"] & _lib.SAF_FLAG_BAD_CLASS_ID:
            raise RuntimeError("Class values must be smaller than num_classes.")
        return out

    def check_errors(self):
        self.stats(check=True)

    def label_argmax(self):
        """argmax_with_check_2d_efficient (clip_seem_fusion.py:315-325): int64 [N], -1 where unobserved."""
        out = torch.empty(self.labels_one_hot.shape[0], dtype=torch.int64, device=self.tsdf.device)
        stream = torch.cuda.current_stream(self.tsdf.device).cuda_stream
        with torch.cuda.device(self.tsdf.device):
            _lib.check(_lib.load().saf_label_argmax(self.labels_one_hot.data_ptr(), out.numel(), self.n_classes,
                                                    out.data_ptr(), stream), "saf_label_argmax")
        return out


class ClipSeemFusion(_FusionVolume):
    """clip_seem_fusion.py:611-888."""

    _with_labels = True
    _rgb_mode = _lib.SAF_RGB_BILINEAR

    def __init__(self, origin, voxel_size, nvox, trunc, scale_patches_by_depth, clip_patch_size, clip_patch_stride,
                 clip_model, seg_model, x_begin=0, x_end=None, x_span=0, x_stride=0, y_ranks=0, y_rank=0):
        super().__init__()
        self.clip = clip_model
        self.clip_patch_size = clip_patch_size
        self.clip_patch_stride = clip_patch_stride
        self.scale_patches_by_depth = scale_patches_by_depth
        self.segmentation_model = seg_model
        self._init_volume(origin, voxel_size, nvox, trunc, self.clip.feature_dim, x_begin, x_end, x_span, x_stride,
                          y_ranks, y_rank)
        self.debug_counter = 0

    def _run_producers(self, depth_imgs, rgb_imgs, K):
        """clip_seem_fusion.py:683-695, 753-757: tiled-patch CLIP feature image and one class map per frame."""
        batch_size = rgb_imgs.shape[0]
        rgb_chw = rgb_imgs.permute(0, 3, 1, 2)
        if self.feature_source == "segment_table":
            seg_maps = [self.segmentation_model.run_on_image(rgb_chw[i]) for i in range(batch_size)]
            return self.clip.segment_features(rgb_chw, seg_maps), seg_maps     # [B, n_segments, C]
        if self.scale_patches_by_depth:
            clip_feat_img = self.clip.img_inference_tiled_depthscaled(rgb_chw, depth_imgs, K,
                                                                      patch_stride=self.clip_patch_stride)
        else:
            clip_feat_img = self.clip.img_inference_tiled(rgb_chw, patch_size=self.clip_patch_size,
                                                          patch_stride=self.clip_patch_stride)
        seg_maps = [self.segmentation_model.run_on_image(rgb_chw[i]) for i in range(batch_size)]
        return clip_feat_img, seg_maps

    def integrate(self, depth_imgs, rgb_imgs, poses, K):
        """clip_seem_fusion.py:676-822.  depth [B,H,W], rgb [B,H,W,3] in [0,1], poses [B,4,4], K [B,3,3]."""
        clip_feat_img, seg_maps = self._run_producers(*self._producer_inputs(depth_imgs, rgb_imgs), K)
        self._integrate_frames(depth_imgs, rgb_imgs, poses, K, clip_feat_img, seg_maps)

    def extract_mesh(self, halo=None, return_edge_ids=False):
        """clip_seem_fusion.py:824-888 on the device.  halo / return_edge_ids: x-slabs only, see
        mesh.extract_mesh_seem."""
        from .mesh import extract_mesh_seem
        return extract_mesh_seem(self, halo, return_edge_ids)


class ClipFusion(_FusionVolume):
    """clipfusion.py:575-763.  `clip_model` may be an open_clip model name (as in the reference) or an
    already constructed object with feature_dim / img_inference_tiled (what the tests inject)."""

    _with_labels = False
    _rgb_mode = _lib.SAF_RGB_NEAREST

    def __init__(self, origin, voxel_size, nvox, trunc, scale_patches_by_depth, clip_model, clip_pretraining,
                 clip_patch_size, clip_patch_stride, x_begin=0, x_end=None, x_span=0, x_stride=0, y_ranks=0, y_rank=0):
        super().__init__()
        if isinstance(clip_model, str):
            from .query import Clip
            self.clip = Clip(clip_model, clip_pretraining)
            self.clip.requires_grad_(False)
            self.clip.eval()
        else:
            self.clip = clip_model
        self.clip_patch_size = clip_patch_size
        self.clip_patch_stride = clip_patch_stride
        self.scale_patches_by_depth = scale_patches_by_depth
        self._init_volume(origin, voxel_size, nvox, trunc, self.clip.feature_dim, x_begin, x_end, x_span, x_stride,
                          y_ranks, y_rank)

    def _run_producers(self, depth_imgs, rgb_imgs, K):
        """clipfusion.py:634-646."""
        rgb_chw = rgb_imgs.permute(0, 3, 1, 2)
        if self.scale_patches_by_depth:
            clip_feat_img = self.clip.img_inference_tiled_depthscaled(rgb_chw, depth_imgs, K,
                                                                      patch_stride=self.clip_patch_stride)
        else:
            clip_feat_img = self.clip.img_inference_tiled(rgb_chw, patch_size=self.clip_patch_size,
                                                          patch_stride=self.clip_patch_stride)
        return clip_feat_img, None

    def integrate(self, depth_imgs, rgb_imgs, poses, K):
        """clipfusion.py:627-721."""
        clip_feat_img, _ = self._run_producers(*self._producer_inputs(depth_imgs, rgb_imgs), K)
        self._integrate_frames(depth_imgs, rgb_imgs, poses, K, clip_feat_img, None)

    def extract_mesh(self, halo=None, return_edge_ids=False):
        """clipfusion.py:723-763 on the device.  halo / return_edge_ids: x-slabs only, see
        mesh.extract_mesh_fusion."""
        from .mesh import extract_mesh_fusion
        return extract_mesh_fusion(self, halo, return_edge_ids)

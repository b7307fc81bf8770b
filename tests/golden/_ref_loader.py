"""Import the UNMODIFIED reference modules from /root/reference (this container only).

Used only by tests/golden/make_golden.py to generate the committed golden vectors; nothing
that runs on the GPU box imports this file (the reference checkout does not travel).

The reference's module-top imports pull in packages that are absent here and that the
fusion/query hot path never touches; they are replaced by empty stub
modules before import.  torch, numpy, cv2, pandas, yaml, tqdm are the real packages.
"""
import importlib
import sys
import types

REFERENCE_ROOT = "/root/reference"


def _stub(name, **attrs):
    mod = sys.modules.get(name)
    if mod is None:
        mod = types.ModuleType(name)
        mod.__path__ = []  # behave as a package so submodule imports resolve
        sys.modules[name] = mod
    for key, val in attrs.items():
        setattr(mod, key, val)
    return mod


def _try_real(name):
    try:
        importlib.import_module(name)
        return True
    except Exception:
        return False


def install_stubs():
    class _Anything:
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return _Anything()

        def __getattr__(self, item):
            return _Anything()

    def _fn(*a, **k):
        return _Anything()

    for name in ("h5py", "open_clip", "trimesh", "open3d", "pretty_errors"):
        if not _try_real(name):
            _stub(name)
    if not _try_real("matplotlib.pyplot"):
        _stub("matplotlib")
        _stub("matplotlib.pyplot")
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if not _try_real("skimage.measure"):
        _stub("skimage")
        _stub("skimage.measure")
        sys.modules["skimage"].measure = sys.modules["skimage.measure"]
    if not _try_real("vedo"):
        _stub("vedo", __all__=[])
    _stub("dgcnn")
    _stub("dgcnn.main_cls", InSituLearning=_Anything)
    _stub("dgcnn.data", InSituVoxelData=_Anything)
    _stub("detectron2")
    _stub("detectron2.config", get_cfg=_fn)
    _stub("detectron2.projects")
    _stub("detectron2.projects.deeplab", add_deeplab_config=_fn)
    _stub("detectron2.utils")
    _stub("detectron2.utils.visualizer", ColorMode=_Anything, Visualizer=_Anything,
          _PanopticPrediction=_Anything)
    _stub("detectron2.modeling", build_model=_fn)
    _stub("detectron2.data", MetadataCatalog=_Anything())
    _stub("detectron2.data.transforms")
    _stub("detectron2.checkpoint", DetectionCheckpointer=_Anything)
    _stub("kmax")
    _stub("kmax.kmax_deeplab", add_kmax_deeplab_config=_fn)
    _stub("kmax.constants",
          COCO_PANOPTIC_CLASSES=["class%d" % i for i in range(133)],
          COCO_PANOPTIC_COLORS=[[i, i, i] for i in range(133)])


def load_reference():
    """Returns (clipfusion, clip_seem_fusion, handy_utils) reference modules."""
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    clipfusion = importlib.import_module("clipfusion")
    handy_utils = importlib.import_module("handy_utils")
    clip_seem_fusion = importlib.import_module("clip_seem_fusion")
    return clipfusion, clip_seem_fusion, handy_utils

"""Import the UNMODIFIED reference modules (clipfusion, clip_seem_fusion, handy_utils).

Two users: tests/golden/make_golden.py (build container, from /root/reference) to generate the committed golden
vectors, and bench.py's `--impl reference` arm (GPU box, from baseline/_ref/, where tools/install_reference.py put
unmodified copies - the reference checkout itself does not travel).  Nothing in the product package imports this.

The reference's module-top imports pull in packages that are absent here and that the
fusion/query hot path never touches; they are replaced by empty stub
modules before import.  torch, numpy, cv2, pandas, yaml, tqdm are the real packages.
"""
import importlib
import sys
import types

import os

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOTS = ("/root/reference", os.path.join(_HERE, "_ref"))


def reference_root():
    for root in REFERENCE_ROOTS:
        if os.path.exists(os.path.join(root, "clip_seem_fusion.py")):
            return root
    return None


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


def load_reference(root=None):
    """Returns (clipfusion, clip_seem_fusion, handy_utils) reference modules."""
    root = root or reference_root()
    if root is None:
        raise RuntimeError("no reference checkout: neither /root/reference nor baseline/_ref holds the modules "
                           "(run tools/install_reference.py in the build container)")
    install_stubs()
    if root not in sys.path:
        sys.path.insert(0, root)
    clipfusion = importlib.import_module("clipfusion")
    handy_utils = importlib.import_module("handy_utils")
    clip_seem_fusion = importlib.import_module("clip_seem_fusion")
    return clipfusion, clip_seem_fusion, handy_utils

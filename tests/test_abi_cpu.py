"""CPU: the C-ABI library builds, loads and exports every symbol include/saf_b200.h declares;
host-side argument checking works without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from spatially_aware_ai_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build_library()
    return _lib.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "saf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(saf_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    names = declared_symbols()
    assert len(names) >= 14
    for name in names:
        assert hasattr(lib, name), name
    assert set(names) == set(_lib.SIGNATURES)


def test_struct_sizes_match_header(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "saf_b200.h"\nint main(void){printf("%zu %zu %zu %zu %zu\\n",'
                   'sizeof(saf_frame),sizeof(saf_stats),sizeof(saf_grid_desc),sizeof(saf_volume),sizeof(saf_workspace));return 0;}\n')
    exe = tmp_path / "sz"
    import subprocess
    subprocess.run(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert sizes == [ctypes.sizeof(_lib.Frame), ctypes.sizeof(_lib.Stats), ctypes.sizeof(_lib.GridDesc),
                     ctypes.sizeof(_lib.Volume), ctypes.sizeof(_lib.Workspace)]


def test_version_and_error_strings(lib):
    assert lib.saf_abi_version() == _lib.SAF_ABI_VERSION
    assert lib.saf_error_string(0) == b"ok"
    assert b"sm_100" in lib.saf_error_string(-9)
    assert b"null" in lib.saf_error_string(-1)


def test_workspace_bytes_host_logic(lib):
    g = _lib.GridDesc()
    g.origin[:] = [0, 0, 0]
    g.voxel_size = 0.02
    g.nvox[:] = [304, 304, 154]
    g.x_begin, g.x_end = 0, 304
    n = ctypes.c_uint64()
    assert lib.saf_workspace_bytes(ctypes.byref(g), 1, 35 * 768, ctypes.byref(n)) == 0
    nblocks = 38 * 38 * 20
    # header + block bookkeeping + one 16-byte list entry per voxel of every 8^3 block + one packed table
    # (two scratch slots: K1/K2 of the next frame overlap K3 of the current one)
    lists = 2 * 16 * nblocks * 512
    assert n.value >= 512 + lists + 2 * 4 * 35 * 768
    assert n.value < 512 + lists + 2 * 4 * 35 * 768 + 64 * nblocks
    n2 = ctypes.c_uint64()
    g.x_begin, g.x_end = 76, 152   # a quarter slab
    assert lib.saf_workspace_bytes(ctypes.byref(g), 1, 35 * 768, ctypes.byref(n2)) == 0
    assert n2.value < n.value / 3.7
    # argument errors
    assert lib.saf_workspace_bytes(ctypes.byref(g), 0, 0, ctypes.byref(n)) == -2
    assert lib.saf_workspace_bytes(ctypes.byref(g), 17, 0, ctypes.byref(n)) == -2
    g.x_end = 400
    assert lib.saf_workspace_bytes(ctypes.byref(g), 1, 0, ctypes.byref(n)) == -3
    assert lib.saf_workspace_bytes(None, 1, 0, ctypes.byref(n)) == -1


def test_topk_workspace_bytes(lib):
    n = ctypes.c_uint64()
    assert lib.saf_query_topk_workspace_bytes(24_000_000, 256, 100, ctypes.byref(n)) == 0
    assert n.value < 2 << 30
    assert lib.saf_query_topk_workspace_bytes(10, 0, 5, ctypes.byref(n)) == -5


def test_volume_refuses_cpu():
    import torch
    import spatially_aware_ai_b200 as saf
    from tests.helpers import FakeClip, FakeSeg
    vol = saf.ClipSeemFusion(torch.zeros(3), 0.1, torch.tensor([8, 8, 8]), 0.2, False, 0, 0, FakeClip(4), FakeSeg())
    assert vol.tsdf.shape == (512,) and vol.labels_one_hot.shape == (512, 143) and vol.clip_feat.shape == (512, 4)
    assert vol.weight.dtype == torch.int32 and vol.tsdf_weight.dtype == torch.int32
    ref_xyz = (torch.stack(torch.meshgrid(torch.arange(8), torch.arange(8), torch.arange(8), indexing="ij"), -1)
               .view(-1, 3) * 0.1 + torch.zeros(3))
    assert torch.equal(vol.xyz_world, ref_xyz)
    vol.clip.next_table = torch.zeros(1, 4, 1, 1)
    vol.segmentation_model.queue = [torch.zeros(4, 4, dtype=torch.int64)]
    with pytest.raises(RuntimeError, match="no CPU path"):
        vol.integrate(torch.zeros(1, 4, 4), torch.zeros(1, 4, 4, 3), torch.eye(4)[None], torch.eye(3)[None])


def test_extract_mesh_by_object_matches_reference_golden():
    import spatially_aware_ai_b200 as saf
    from tests import helpers as Hh
    g = Hh.load_golden("query")
    v, f, c, _ = saf.extract_mesh_by_object(g["mesh_verts"], g["mesh_faces"], g["mesh_colors"], g["mesh_vidx"], 1)
    assert np.array_equal(v, g["obj1_verts"]) and np.array_equal(f, g["obj1_faces"]) and np.array_equal(c, g["obj1_colors"])
    v, f, c, _ = saf.extract_mesh_by_object(g["mesh_verts"], g["mesh_faces"], g["mesh_colors"], g["mesh_vidx"], 99)
    assert len(v) == 0 and len(f) == 0 and len(c) == 0


def test_mesh_workspace_bytes_host_logic(lib):
    g = _lib.GridDesc()
    g.voxel_size = 0.02
    g.nvox[:] = [304, 304, 154]
    g.x_begin, g.x_end = 0, 304
    n = ctypes.c_uint64()
    assert lib.saf_mesh_workspace_bytes(ctypes.byref(g), ctypes.byref(n)) == 0
    voxels = 304 * 304 * 154
    plane = 304 * 154                       # room for the successor's first plane (slab meshes with seam cells)
    assert 12 * (voxels + plane) <= n.value < 12 * (voxels + plane) + 8 * ((voxels + plane) // 256 + 2) + 4096
    g.x_begin, g.x_end = 10, 5
    assert lib.saf_mesh_workspace_bytes(ctypes.byref(g), ctypes.byref(n)) == -3
    assert lib.saf_mesh_workspace_bytes(None, ctypes.byref(n)) == -1


def test_mc_table_header_is_current():
    """csrc/saf_mc_tables.h is generated by tools/gen_mc_tables.py; the committed copy must match the generator
    (the CPU oracle derives its table from the same generator at import time)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_mc_tables", os.path.join(ROOT, "tools", "gen_mc_tables.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "t.h")
        gen.write_header(path)
        assert open(path).read() == open(os.path.join(ROOT, "spatially_aware_ai_b200", "csrc", "saf_mc_tables.h")).read()


def test_build_scene_knowledge_matches_reference_golden():
    """Host-side bookkeeping of flood_fill_3d / add_object (handy_utils.py:244-292, 351-480) from the labelled grid
    the reference itself produced: same object ids, classes, indices, sizes and counts, in the same order."""
    import spatially_aware_ai_b200 as saf
    from tests import helpers as Hh
    g = Hh.load_golden("objects")
    names = [str(s) for s in g["class_names"]]
    know = saf.build_scene_knowledge(g["voxel_obj_ids"], g["class_grid"], names, [[i, i, i] for i in range(len(names))])
    uo = know["unique_objects"]
    assert list(uo.keys()) == [str(s) for s in g["obj_ids"]]
    assert [uo[k]["class_id"] for k in uo] == g["obj_class_id"].tolist()
    assert [uo[k]["object_index"] for k in uo] == g["obj_index"].tolist()
    assert [len(uo[k]["voxels"]) for k in uo] == g["obj_size"].tolist()
    assert all(uo[k]["color"] == [uo[k]["class_id"]] * 3 and uo[k]["gt_label"] == k and not uo[k]["merged"] for k in uo)
    assert know["object_counts"] == dict(zip([str(s) for s in g["count_keys"]], g["count_vals"].tolist()))
    # every listed voxel carries the object's index in the grid
    first = next(iter(uo.values()))
    assert all(g["voxel_obj_ids"][v] == first["object_index"] for v in first["voxels"])
    assert set(know) >= {"unique_objects", "object_counts", "unchanged_objects", "new_objects", "missing_objects"}


def test_ctypes_signatures_have_the_header_arity():
    """Every prototype of include/saf_b200.h takes as many parameters as its ctypes signature in _lib.py."""
    text = open(os.path.join(ROOT, "include", "saf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = re.findall(r"\b(saf_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S)
    assert len(protos) >= 28
    for name, params in protos:
        params = " ".join(params.split())
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(_lib.SIGNATURES[name][1]), (name, n, len(_lib.SIGNATURES[name][1]))


def test_window_stage_argument_checks(lib):
    """saf_feature_accumulate_window_stages refuses stage masks it does not know before touching the device."""
    g, v, ws, f = _lib.GridDesc(), _lib.Volume(), _lib.Workspace(), (_lib.Frame * 2)()
    for stages in (0, 4, -1):
        rc = lib.saf_feature_accumulate_window_stages(ctypes.byref(g), ctypes.byref(v), f, 2, 48, 64, _lib.SAF_RGB_BILINEAR,
                                                      ctypes.byref(ws), stages, None)
        assert rc == -8, (stages, rc)      # SAF_ERR_UNSUPPORTED

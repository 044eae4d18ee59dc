"""Repository contract: the product never touches the oracle, the C-ABI library loads on a
CPU-only machine and exports exactly what include/voxblox_b200.h declares."""
import ctypes
import os
import re
import subprocess

import pytest

import voxblox_b200 as vb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "voxblox_b200.h")).read()
    return sorted(set(re.findall(r"VBX_API\s+[\w\s\*]+?\b(vbx_\w+)\s*\(", text)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = vb.load_library()  # no CUDA call happens at load time
    declared = _header_functions()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/voxblox_b200.h but not exported"
    nm = subprocess.check_output(["nm", "-D", "--defined-only", vb.LIB_PATH], text=True)
    exported = sorted(set(re.findall(r" T (vbx_\w+)", nm)))
    assert exported == declared, (set(exported) ^ set(declared))
    assert sorted(vb.api.EXPORTS) == declared
    assert b"sm_100a" in lib.vbx_version()


def test_engine_is_built_for_sm_100a_only():
    out = subprocess.run(["cuobjdump", "--list-elf", vb.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_(\d+a?)", out.stdout))
    assert archs == {"100a"}, archs


def test_product_never_references_the_oracle():
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "voxblox_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp")) or f == "Makefile":
                text = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"\boracle\b|vbx_oracle|libvbx_ref|pyoracle|/root/reference", text):
                    bad.append(os.path.join(base, f))
    assert not bad, bad
    for f in ("include/voxblox_b200.h",):
        assert "pyoracle" not in open(os.path.join(ROOT, f)).read()


def test_dev_scripts_outside_the_test_tree_do_not_touch_the_oracle():
    """Only tests/, smoke() and bench.py's CPU-baseline legs may load the checkers: the measurement / probe scripts
    under scripts/ drive the product alone (the ones that compare against the oracle live under tests/tools/)."""
    bad = []
    for f in os.listdir(os.path.join(ROOT, "scripts")):
        if f.endswith((".py", ".sh")):
            text = open(os.path.join(ROOT, "scripts", f), errors="ignore").read()
            if re.search(r"from oracle|import oracle|pyoracle|libvbx_ref|libvbx_oracle", text):
                bad.append(f)
    assert not bad, bad


def test_no_cpu_fallback_without_a_device():
    """On a machine without CUDA the engine must refuse loudly, never compute on the host."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    layer = vb.Layer(0.1, 16)
    with pytest.raises(vb.VoxbloxError):
        vb.TsdfIntegratorFactory.create("merged", vb.TsdfIntegratorConfig(), layer)


def test_operator_interface_mirrors_reference_errors():
    layer = vb.Layer(0.1, 16)
    with pytest.raises(vb.VoxbloxError, match="Unknown TSDF integrator type"):
        vb.TsdfIntegratorFactory.create("octomap", vb.TsdfIntegratorConfig(), layer)  # cc:22
    with pytest.raises(vb.VoxbloxError, match="Unknown TSDF integrator type"):
        vb.TsdfIntegratorFactory.create(7, vb.TsdfIntegratorConfig(), layer)  # cc:41
    with pytest.raises(vb.VoxbloxError):
        vb.TsdfIntegratorFactory.create("merged", vb.TsdfIntegratorConfig(), None)  # CHECK_NOTNULL cc:29
    with pytest.raises(vb.VoxbloxError, match="integration order mode"):
        vb.TsdfIntegratorConfig(integration_order_mode="random")  # integrator_utils.cc:12
    with pytest.raises(vb.VoxbloxError):
        vb.Layer(0.0, 16)  # CHECK_GT(voxel_size_, 0.0f), core/layer.h:38
    cfg = vb.TsdfIntegratorConfig()
    assert abs(cfg.default_truncation_distance - 0.1) < 1e-7 and cfg.max_weight == 10000.0
    assert cfg.max_ray_length_m == 5.0 and abs(cfg.min_ray_length_m - 0.1) < 1e-7
    assert cfg.start_voxel_subsampling_factor == 2.0 and cfg.max_consecutive_ray_collisions == 2
    e = vb.EsdfIntegratorConfig()
    assert e.max_distance_m == 2.0 and abs(e.min_diff_m - 0.001) < 1e-9 and e.num_buckets == 20
    assert vb.TSDF_DTYPE.itemsize == 12 and vb.ESDF_DTYPE.itemsize == 20


def test_mesh_host_mirror_matches_reference_defaults_and_layout():
    """MeshIntegratorConfig defaults (mesh/mesh_integrator.h:49-52), the C struct layout of
    vbx_mesh_config, and MeshLayer::allocateMeshPtrByIndex's origin = index * block_size
    (mesh/mesh_layer.h:108-121)."""
    import ctypes as C

    import numpy as np

    cfg = vb.MeshIntegratorConfig()
    assert cfg.use_color == 1 and abs(cfg.min_weight - 1e-4) < 1e-10
    assert C.sizeof(vb.MeshIntegratorConfig) == 8
    text = open(os.path.join(ROOT, "include", "voxblox_b200.h")).read()
    body = re.search(r"typedef struct vbx_mesh_config \{(.*?)\} vbx_mesh_config;", text, re.S).group(1)
    assert re.findall(r"\b(int32_t|float)\s+(\w+);", body) == [("int32_t", "use_color"), ("float", "min_weight")]
    ml = vb.MeshLayer(0.8)
    m = ml.allocateMeshPtrByIndex((1, -2, 3))
    assert m is ml.allocateMeshPtrByIndex(np.array([1, -2, 3]))          # same mesh on the second call
    assert np.allclose(m.origin, np.float32(0.8) * np.array([1, -2, 3], np.float32)) and m.size() == 0 and not m.updated
    assert ml.getNumberOfAllocatedMeshes() == 1 and ml.getMeshPtrByIndex((0, 0, 0)) is None
    ml.allocateMeshPtrByIndex((0, 0, 0)).updated = True
    assert ml.getAllAllocatedMeshes().tolist() == [[0, 0, 0], [1, -2, 3]]
    assert ml.getAllUpdatedMeshes().tolist() == [[0, 0, 0]]
    with pytest.raises(vb.VoxbloxError):
        vb.MeshIntegrator(cfg, None, ml)     # CHECK_NOTNULL(sdf_layer), mesh_integrator.h:89

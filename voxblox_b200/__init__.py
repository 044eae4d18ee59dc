"""voxblox_b200: B200-native TSDF / ESDF integration behind voxblox's integrator API.

Only the hot path lives here: csrc/ (hand-written sm_100a CUDA + the C-ABI of
include/voxblox_b200.h), api.py (host-side mirror of the reference's operator
interface) and scenes.py (synthetic scan generators for tests and bench).
"""
from .api import (EngineOptions, ICP, ICPConfig, Mesh, MeshIntegrator, MeshIntegratorConfig, MeshLayer, EsdfIntegrator, EsdfIntegratorConfig, FastTsdfIntegrator, Layer,
                  MergedTsdfIntegrator, SimpleTsdfIntegrator, TsdfIntegratorBase,
                  TsdfIntegratorConfig, TsdfIntegratorFactory, TsdfIntegratorType, VoxbloxError,
                  ESDF_DTYPE, TSDF_DTYPE, LIB_PATH, load_library)

__all__ = ["EngineOptions", "ICP", "ICPConfig", "Mesh", "MeshIntegrator", "MeshIntegratorConfig", "MeshLayer", "EsdfIntegrator", "EsdfIntegratorConfig", "FastTsdfIntegrator", "Layer",
           "MergedTsdfIntegrator", "SimpleTsdfIntegrator", "TsdfIntegratorBase",
           "TsdfIntegratorConfig", "TsdfIntegratorFactory", "TsdfIntegratorType", "VoxbloxError",
           "ESDF_DTYPE", "TSDF_DTYPE", "LIB_PATH", "load_library"]

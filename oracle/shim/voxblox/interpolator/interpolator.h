// TEST INFRASTRUCTURE ONLY (oracle/shim): voxblox/mesh/mesh_integrator.h includes the trilinear
// interpolator header but uses nothing from it; the real header needs Eigen features the shim does
// not provide (dynamic-size products), so the oracle build sees this empty stand-in instead.
#ifndef VOXBLOX_INTERPOLATOR_INTERPOLATOR_H_
#define VOXBLOX_INTERPOLATOR_INTERPOLATOR_H_
#endif

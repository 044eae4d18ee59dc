// ESDF wavefront on the device (EsdfIntegrator, voxblox/src/integrator/esdf_integrator.cc).
// Placeholder until the TSDF path is parity-green on hardware.
#include "vbx_engine.h"

namespace vbx {

int esdf_create(vbx_ctx* c, const vbx_esdf_config*) {
  return fail(c, VBX_E_STATE, "ESDF device path not built yet");
}
int esdf_update(vbx_ctx* c, int, int) { return fail(c, VBX_E_STATE, "ESDF device path not built yet"); }
int esdf_destroy(vbx_ctx*) { return VBX_OK; }

}  // namespace vbx

// rv_abi.cu -- context management and error reporting of the C ABI (include/repas_vision.h).
#include <new>

#include "rv_common.cuh"

extern "C" {

int rv_abi_version(void) { return RV_ABI_VERSION; }
int rv_sizeof_cam(void) { return (int)sizeof(RvCam); }
int rv_sizeof_deproject_params(void) { return (int)sizeof(RvDeprojectParams); }

const char *rv_build_info(void) {
  return "repas_vision_b200 sm_100a (compute_100a) --fmad=false; built " __DATE__ " " __TIME__;
}

const char *rv_status_string(int status) {
  switch (status) {
    case RV_OK: return "ok";
    case RV_EINVAL: return "invalid argument";
    case RV_ECAPACITY: return "output capacity exceeded";
    case RV_ECUDA: return "CUDA error";
    case RV_EALIGN: return "misaligned pointer";
    case RV_EWORKSPACE: return "workspace too small";
    default: return "unknown status";
  }
}

int rv_create(int device, rv_ctx **out_ctx) {
  if (!out_ctx) return RV_EINVAL;
  *out_ctx = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return RV_ECUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return RV_ECUDA;
  if (prop.major < 10) return RV_ECUDA;  // sm_100a code only: no fallback path exists
  rv_ctx *c = new (std::nothrow) rv_ctx();
  if (!c) return RV_ECUDA;
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->launches = 0;
  c->err[0] = 0;
  *out_ctx = c;
  return RV_OK;
}

int rv_destroy(rv_ctx *ctx) {
  delete ctx;
  return RV_OK;
}

const char *rv_last_error(const rv_ctx *ctx) { return ctx ? ctx->err : "null context"; }
int rv_sm_count(const rv_ctx *ctx) { return ctx ? ctx->sm_count : 0; }
int64_t rv_launch_count(const rv_ctx *ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"

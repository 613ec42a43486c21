// Handle management for the gccnmf_b200 C ABI.  There is no CPU fallback: without a CUDA device
// gccnmf_create fails with GCCNMF_ERR_NO_DEVICE.
#include <cstdlib>

#include "common.cuh"

namespace {
thread_local std::string g_create_error = "no error";
}

extern "C" {

int gccnmf_abi_version(void) { return GCCNMF_ABI_VERSION; }

int gccnmf_create(gccnmf_handle** out, int device) {
  if (!out) return GCCNMF_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  int count = 0;
  cudaError_t err = cudaGetDeviceCount(&count);
  if (err != cudaSuccess || count == 0) {
    g_create_error = std::string("gccnmf_create: no CUDA device (") + cudaGetErrorString(err) + "); there is no CPU fallback";
    return GCCNMF_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= count) {
    g_create_error = "gccnmf_create: device index out of range";
    return GCCNMF_ERR_INVALID_ARGUMENT;
  }
  err = cudaSetDevice(device);
  cudaDeviceProp prop;
  if (err == cudaSuccess) err = cudaGetDeviceProperties(&prop, device);
  if (err != cudaSuccess) {
    g_create_error = std::string("gccnmf_create: ") + cudaGetErrorString(err);
    return GCCNMF_ERR_CUDA;
  }
  if (prop.major != 10) {
    g_create_error = "gccnmf_create: this library is built for sm_100a (B200) only, found compute capability " +
                     std::to_string(prop.major) + "." + std::to_string(prop.minor);
    return GCCNMF_ERR_UNSUPPORTED;
  }
  gccnmf_handle* h = new gccnmf_handle();
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  h->last_error = "no error";
  const char* path = getenv("GCCNMF_NMF_PATH");
  h->force_simt_nmf = path && strcmp(path, "simt") == 0;
  const char* pdl = getenv("GCCNMF_NMF_PDL");
  h->nmf_pdl = !(pdl && strcmp(pdl, "0") == 0);
  const char* split2 = getenv("GCCNMF_WH_SPLIT2");
  if (split2) h->wh_split2 = atoi(split2);
  const char* light = getenv("GCCNMF_MC_LIGHT_SIGNAL");
  if (light) h->mc_light_signal = atoi(light);
  const char* pers = getenv("GCCNMF_ARGMAX_PERSISTENT");
  if (pers) h->argmax_persistent = strcmp(pers, "0") != 0;
  *out = h;
  return GCCNMF_OK;
}

int gccnmf_destroy(gccnmf_handle* h) {
  if (!h) return GCCNMF_OK;
  for (int i = 0; i < gccnmf_handle::kMaxPlans; ++i) {
    if (h->plan_tw64[i]) cudaFree(h->plan_tw64[i]);
    if (h->plan_tw32[i]) cudaFree(h->plan_tw32[i]);
  }
  gccnmf_tmap_cache_free(h);
  delete h;
  return GCCNMF_OK;
}

const char* gccnmf_last_error(const gccnmf_handle* h) { return h ? h->last_error.c_str() : g_create_error.c_str(); }

const char* gccnmf_status_string(int status) {
  switch (status) {
    case GCCNMF_OK: return "ok";
    case GCCNMF_ERR_INVALID_ARGUMENT: return "invalid argument";
    case GCCNMF_ERR_CUDA: return "CUDA error";
    case GCCNMF_ERR_WORKSPACE: return "workspace missing or too small";
    case GCCNMF_ERR_UNSUPPORTED: return "unsupported shape or device";
    case GCCNMF_ERR_NO_DEVICE: return "no CUDA device (no CPU fallback)";
    default: return "unknown status";
  }
}

int64_t gccnmf_launch_count(const gccnmf_handle* h) { return h ? h->launches : 0; }

int gccnmf_set_option(gccnmf_handle* h, const char* name, int value) {
  if (!h || !name) return GCCNMF_ERR_INVALID_ARGUMENT;
  if (strcmp(name, "force_simt_nmf") == 0) { h->force_simt_nmf = value != 0; return GCCNMF_OK; }
  if (strcmp(name, "nmf_pdl") == 0) { h->nmf_pdl = value != 0; return GCCNMF_OK; }
  if (strcmp(name, "wh_tile") == 0) { h->wh_tile = value; return GCCNMF_OK; }
  if (strcmp(name, "argmax_refine_shared") == 0) { h->argmax_refine_shared = value != 0; return GCCNMF_OK; }
  if (strcmp(name, "gemm_pair") == 0) { h->gemm_pair = value; return GCCNMF_OK; }
  if (strcmp(name, "gemm_streaming") == 0) { h->gemm_streaming = value; return GCCNMF_OK; }
  if (strcmp(name, "argmax_persistent") == 0) { h->argmax_persistent = value != 0; return GCCNMF_OK; }
  if (strcmp(name, "wh_split2") == 0) { h->wh_split2 = value; return GCCNMF_OK; }
  if (strcmp(name, "w_cluster_reduce") == 0) { h->w_cluster_reduce = value != 0; return GCCNMF_OK; }
  if (strcmp(name, "pull_force_pack") == 0) { h->pull_force_pack = value; return GCCNMF_OK; }
  if (strcmp(name, "mc_light_signal") == 0) { h->mc_light_signal = value; return GCCNMF_OK; }
  if (strcmp(name, "l2_persist") == 0) { h->l2_persist = value; return GCCNMF_OK; }
  if (strcmp(name, "gemm_preload") == 0) { h->gemm_preload = value; return GCCNMF_OK; }
  if (strcmp(name, "gemm_cluster") == 0) { h->gemm_cluster = value; return GCCNMF_OK; }
  return gccnmf_fail(h, GCCNMF_ERR_INVALID_ARGUMENT, "unknown option '%s'", name);
}

}  // extern "C"

// Handle lifecycle for libnerf_b200 (include/nerf_b200.h).
#include "nb_common.cuh"
#include <new>

extern "C" int nb_abi_version(void) { return NB_ABI_VERSION; }

extern "C" int nb_create(nb_handle_t* out, int device, unsigned flags) {
  (void)flags;
  if (!out) return NB_ERR_INVALID;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return NB_ERR_CUDA;
  nb_handle_s* h = new (std::nothrow) nb_handle_s();
  if (!h) return NB_ERR_INVALID;
  memset(h, 0, sizeof(*h));
  h->device = device;
  cudaDeviceProp p;
  if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&p, device) != cudaSuccess) {
    delete h;
    return NB_ERR_CUDA;
  }
  h->sm_count = p.multiProcessorCount;
  h->cc_major = p.major;
  h->cc_minor = p.minor;
  if (p.major != 10) {  // the only SASS in this library is sm_100a
    delete h;
    return NB_ERR_UNSUPPORTED;
  }
  *out = h;
  return NB_OK;
}

extern "C" int nb_destroy(nb_handle_t h) {
  if (!h) return NB_ERR_INVALID;
  delete h;
  return NB_OK;
}

extern "C" const char* nb_last_error(nb_handle_t h) { return h ? h->err : "null handle"; }

extern "C" int nb_device_info(nb_handle_t h, int32_t info[4]) {
  if (!h || !info) return NB_ERR_INVALID;
  info[0] = h->sm_count;
  info[1] = h->cc_major;
  info[2] = h->cc_minor;
  info[3] = (int32_t)(h->launches & 0x7fffffff);
  return NB_OK;
}

extern "C" int64_t nb_launch_count(nb_handle_t h) { return h ? h->launches : -1; }

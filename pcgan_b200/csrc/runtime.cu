// Error plumbing and device queries shared by every entry point of the C ABI.
#include <stdarg.h>
#include <stdlib.h>
#include <mutex>

#include "common.cuh"

namespace pcgan {

static thread_local std::string g_last_error;

void set_last_error(const std::string& s) { g_last_error = s; }

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    cached[dev] = n;
  }
  return cached[dev];
}

bool sync_launches() {
  static int cached = -1;
  if (cached < 0) {
    const char* e = getenv("PCGAN_SYNC");
    cached = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return cached == 1;
}

bool pdl_enabled() {
  static int cached = -1;
  if (cached < 0) {
    const char* e = getenv("PCGAN_PDL");
    cached = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return cached == 1;
}

}  // namespace pcgan

extern "C" int pcgan_abi_version(void) { return PCGAN_ABI_VERSION; }
extern "C" const char* pcgan_last_error(void) { return pcgan::g_last_error.c_str(); }

// sizeof() of every ABI struct, so a binding can verify its own mirror of the layout.
extern "C" int64_t pcgan_sizeof(const char* name) {
  const std::string s(name ? name : "");
  if (s == "pcgan_tmap") return sizeof(pcgan_tmap);
  if (s == "pcgan_comp") return sizeof(pcgan_comp);
  if (s == "pcgan_igemm_desc") return sizeof(pcgan_igemm_desc);
  if (s == "pcgan_pack_args") return sizeof(pcgan_pack_args);
  if (s == "pcgan_unpack_args") return sizeof(pcgan_unpack_args);
  if (s == "pcgan_norm_finalize_args") return sizeof(pcgan_norm_finalize_args);
  if (s == "pcgan_norm_apply_args") return sizeof(pcgan_norm_apply_args);
  if (s == "pcgan_fold_args") return sizeof(pcgan_fold_args);
  if (s == "pcgan_norm_bwd_args") return sizeof(pcgan_norm_bwd_args);
  if (s == "pcgan_maxpool_args") return sizeof(pcgan_maxpool_args);
  if (s == "pcgan_loss_args") return sizeof(pcgan_loss_args);
  if (s == "pcgan_batch_item") return sizeof(pcgan_batch_item);
  if (s == "pcgan_running_item") return sizeof(pcgan_running_item);
  if (s == "pcgan_adam_item") return sizeof(pcgan_adam_item);
  if (s == "pcgan_image_item") return sizeof(pcgan_image_item);
  if (s == "pcgan_augment_args") return sizeof(pcgan_augment_args);
  return -1;
}

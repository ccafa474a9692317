// Library-wide C ABI helpers (error text, version).
#include <string.h>

#include "common.cuh"

namespace se3et {
static thread_local char g_last_error[512] = "";

void set_last_error(const char* what, cudaError_t err) {
  snprintf(g_last_error, sizeof(g_last_error), "%s: %s (%s)", what, cudaGetErrorName(err), cudaGetErrorString(err));
}
}  // namespace se3et

extern "C" const char* se3et_last_error(void) { return se3et::g_last_error; }
extern "C" int se3et_version(void) { return 100; }

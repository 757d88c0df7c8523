// Library-level entry points: version, error text.
#include <math.h>
#include <stdarg.h>

#include "common.cuh"

namespace fod {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

float iou_threshold_as_float(double thr) {
  float f = static_cast<float>(thr);
  if (!(static_cast<double>(f) > thr)) f = nextafterf(f, INFINITY);
  return f;
}

}  // namespace fod

extern "C" int fod_version(void) { return 100; }

extern "C" int fod_last_error(char* buf, size_t len) {
  if (!buf || len == 0) return FOD_ERR_BAD_ARG;
  strncpy(buf, fod::g_error, len - 1);
  buf[len - 1] = '\0';
  return FOD_OK;
}

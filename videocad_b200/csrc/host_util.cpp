#include "host_util.h"
#include <string.h>

namespace vck {

static thread_local char g_err[1024] = {0};

int set_error(const char* msg) {
  strncpy(g_err, msg ? msg : "unknown error", sizeof(g_err) - 1);
  g_err[sizeof(g_err) - 1] = 0;
  return 1;
}

const char* last_error() { return g_err; }

}  // namespace vck

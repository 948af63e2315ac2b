// Library-wide runtime helpers: error string, device properties.
#include "common.cuh"
#include <string.h>
#include <stdlib.h>

namespace brtpe {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static int g_pdl = -1;
bool pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("BRTPE_PDL");
    g_pdl = (e && atoi(e) != 0) ? 1 : 0;      // opt-in: measured neutral to -1.5 % (profiles/r02_chain.md)
  }
  return g_pdl != 0;
}
void pdl_set(bool on) { g_pdl = on ? 1 : 0; }

int num_sms() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      cached = 148;
  }
  return cached;
}

}  // namespace brtpe

extern "C" const char* brtpe_last_error(void) { return brtpe::g_err; }
extern "C" int brtpe_version(void) { return 100; }

// ABI bookkeeping: version, thread-local error string, launch counter.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace cmmvae {
static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
static std::atomic<int> g_sm_budget{kNumSMs};
int sm_budget() { return g_sm_budget.load(std::memory_order_relaxed); }
static std::atomic<int> g_pdl{1};
bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("CMMVAE_PDL");
    return !(e && strcmp(e, "0") == 0);
  }();
  return on && g_pdl.load(std::memory_order_relaxed) != 0;
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace cmmvae

extern "C" int cmmvae_abi_version(void) { return CMMVAE_ABI_VERSION; }
extern "C" const char* cmmvae_last_error(void) { return cmmvae::g_err; }
extern "C" long long cmmvae_launch_count(void) { return cmmvae::g_launches.load(); }
extern "C" int cmmvae_set_sm_budget(int sms) {
  if (sms < 1 || sms > cmmvae::kNumSMs) {
    cmmvae::set_error("set_sm_budget: %d outside [1, %d]", sms, cmmvae::kNumSMs);
    return -1;
  }
  cmmvae::g_sm_budget.store(sms);
  return 0;
}

extern "C" int cmmvae_set_pdl(int on) {
  cmmvae::g_pdl.store(on ? 1 : 0);
  return 0;
}

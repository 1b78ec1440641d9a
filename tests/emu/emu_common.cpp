// Definitions the library's translation units expect from dkt_api.cu / dkt_solve.cu, for the emulated build (tests only).
#include "dkt_internal.h"
#include "cuda_emu.h"

#include <string>

namespace dkt
{
std::string g_emu_err;
uint64_t g_launches = 0;
void set_error(const std::string &msg) { g_emu_err = msg; }
}  // namespace dkt
extern "C" const char *emu_last_error() { return dkt::g_emu_err.c_str(); }
extern "C" void emu_set_order(int order) { emu::state().order = order; }

// Library-wide state of liblvcb200.so.
#include "common.cuh"

namespace lvcb200 {
thread_local char g_last_error[512] = "";
std::atomic<long long> g_launch_count{0};
}  // namespace lvcb200

extern "C" int lvcb200_abi_version(void) { return LVCB200_ABI_VERSION; }
extern "C" const char* lvcb200_last_error(void) { return lvcb200::g_last_error; }
extern "C" int64_t lvcb200_launch_count(void) { return (int64_t)lvcb200::g_launch_count.load(); }

// TEST INFRASTRUCTURE ONLY -- stubs for the timing-baseline build of the reference (stock src/random.cpp in use):
// the Philox replay hooks of ref_shim.cpp do nothing here.
#include <cstdint>
extern "C" void ref_shim_seed(uint64_t, uint64_t) {}
extern "C" uint64_t ref_shim_pos(void) { return 0; }

// TEST TOOL: the product's kernels + engine + C ABI compiled against cuda_emul.h so the whole prove / verify pipeline
// can be run on CPU threads in this GPU-less container (tests/test_emul_vs_oracle.py).  Exports the same rofl_* symbols
// as librofl_b200.so but is never shipped, loaded or referenced by the product package.
#define ROFL_EMUL 1
#define KG_ALL 1
#include "cuda_emul.h"
#include "../../rofl-project-code_b200/csrc/capi.cuh"
#pragma GCC visibility push(default)
extern "C" int rofl_ctx_create(rofl_ctx **out, int device) {
    rofl_ctx *c = new rofl_ctx(); c->e.device = device; c->e.host_threads = 4; c->e.rt_bits = 8;   // small tables: the CPU emulation builds them thread by thread
    engine_init(c->e); *out = c; return 0;
}
extern "C" void rofl_ctx_destroy(rofl_ctx *c) { if (!c) return; engine_destroy(c->e); delete c; }
extern "C" void rofl_prof_enable(int) {}
extern "C" void rofl_prof_reset(void) {}
extern "C" double rofl_prof_ms(int) { return 0; }
extern "C" long rofl_prof_launches(int) { return 0; }
extern "C" double rofl_prof_work(int) { return 0; }
extern "C" double rofl_probe_imad_wide(rofl_ctx *) { return 0; }
extern "C" void *rofl_ctx_stream(rofl_ctx *) { return nullptr; }
#pragma GCC visibility pop

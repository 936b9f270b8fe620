#include "ctx.h"
namespace mdbn {
bool tensor_supported(const mdbn_ctx*, const mdbn_cd_args&) { return false; }
int tensor_cd_step(mdbn_ctx*, const mdbn_cd_args&, cudaStream_t) { set_error("tensor path not built"); return 4; }
}

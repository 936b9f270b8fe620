#include "ctx.h"
namespace mdbn {
bool skinny_supported(const mdbn_ctx*, const mdbn_cd_args&) { return false; }
int skinny_cd_step(mdbn_ctx*, const mdbn_cd_args&, cudaStream_t) { set_error("skinny path not built"); return 4; }
}

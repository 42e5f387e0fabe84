// smallk_b200 — solver-level entry points used by the C ABI (internal header).
#pragma once

#include "context.h"

namespace smk {

void solver_alloc(smk_ctx* c);
void solver_init(smk_ctx* c);
void solver_step(smk_ctx* c);
int solver_progress(smk_ctx* c, double* metric);
int solver_normalize(smk_ctx* c);
int solver_fail_iter(smk_ctx* c);
void solver_product(smk_ctx* c, int which);

} // namespace smk

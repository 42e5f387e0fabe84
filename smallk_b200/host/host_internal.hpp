// smallk_b200 host — shared between the translation units of libsmallk_host.so (not installed).
#pragma once
#include "../../include/smallk_b200.h"
#include "nmf.hpp"

smk_ctx* NmfContext();                 // the context NmfInitialize created, or nullptr
smk_nmf_options NmfToAbi(const NmfOptions& o);
Result NmfFromAbi(int rc);
void NmfSetLastError(const char* msg);
void HierReleaseWorkers();             // clust.cpp: the score workers' library contexts (kept from one Clust call to the next)

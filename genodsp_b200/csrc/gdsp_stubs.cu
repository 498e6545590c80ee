// temporary: entry points still being written
#include "gdsp_common.cuh"
extern "C" int gdsp_select_ranks (gdsp_ctx*, const gdsp_layout*, const double*, uint32_t, double, double,
                                  const uint64_t*, int, double*, uint64_t*)
	{ gdsp_set_error ("gdsp_select_ranks: not built yet"); return GDSP_ERR_ARG; }
extern "C" int gdsp_sort_genome (gdsp_ctx*, const gdsp_layout*, double*, double*, uint64_t)
	{ gdsp_set_error ("gdsp_sort_genome: not built yet"); return GDSP_ERR_ARG; }
extern "C" size_t gdsp_clump_work_bytes (uint64_t) { return 0; }
extern "C" int gdsp_clump (gdsp_ctx*, const gdsp_layout*, double*, uint64_t, void*, double, uint32_t, double, int, double, double)
	{ gdsp_set_error ("gdsp_clump: not built yet"); return GDSP_ERR_ARG; }

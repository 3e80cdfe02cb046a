// ALS / DMRG sweeps on device-resident tensor trains (placeholder translation unit: filled in below this commit).
#include "xb_internal.cuh"

using namespace xb;

extern "C" {

xb_status xb_als_default_options(xb_als_options* opt, uint32_t sites, int assume_spd) {
	return guard([&] {
		XB_REQUIRE(opt, "null");
		XB_REQUIRE(sites > 0, "sites must be positive");          // als.h:141
		opt->sites = sites;
		opt->assume_spd = assume_spd;
		opt->num_half_sweeps = 0;
		opt->convergence_epsilon = 1e-6;                           // als.h:137
		opt->preserve_core_position = 1;                           // als.h:138
		opt->local_tolerance = 0.0;
		opt->local_max_iterations = 0;
	});
}

xb_status xb_als_solve(const xb_tt*, xb_tt*, const xb_tt*, const xb_als_options*, double*, size_t*) {
	return guard([&] { throw Error(XB_ERR_UNSUPPORTED, "xb_als_solve: not built yet"); });
}

} // extern "C"

// placeholder until the device-side L-BFGS lands
#pragma once
#include "host_common.cuh"
struct ist_lbfgs { int dummy; };
extern "C" {
int ist_lbfgs_create(ist_lbfgs**, ist_plan*, int, int, int, float, double, double) { return ist::fail(IST_ERR_STATE, "ist_lbfgs: not built yet"); }
int ist_lbfgs_destroy(ist_lbfgs*) { return IST_OK; }
int ist_lbfgs_step(ist_lbfgs*, float*, int*, float*, void*) { return ist::fail(IST_ERR_STATE, "ist_lbfgs: not built yet"); }
int ist_lbfgs_last_losses(ist_lbfgs*, float*) { return ist::fail(IST_ERR_STATE, "ist_lbfgs: not built yet"); }
}

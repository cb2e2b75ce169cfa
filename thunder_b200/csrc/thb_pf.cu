// placeholder - replaced by the device particle filter
#include "thb_context.h"
namespace thb { void pf_free(thb_ctx*) {} }
extern "C" {
int thb_pf_load(thb_ctx* c, int, const thb_pf_params*, const double*, const double*, const double*, const double*) { return thb::set_error(c, THB_E_STATE, "pf: not built"); }
int thb_pf_get(thb_ctx* c, double*, double*, double*, double*, double*) { return thb::set_error(c, THB_E_STATE, "pf: not built"); }
int thb_pf_set(thb_ctx* c, const double*, const double*, const double*, const double*, const double*) { return thb::set_error(c, THB_E_STATE, "pf: not built"); }
int thb_expectation(thb_ctx* c, int*) { return thb::set_error(c, THB_E_STATE, "pf: not built"); }
int thb_reconstruct_insert(thb_ctx* c, int, int, const double*) { return thb::set_error(c, THB_E_STATE, "pf: not built"); }
int thb_pf_op(thb_ctx* c, int, double, const float*, const float*) { return thb::set_error(c, THB_E_STATE, "pf: not built"); }
}

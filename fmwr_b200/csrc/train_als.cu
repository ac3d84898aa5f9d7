// placeholder: filled in below in this round
#include "common.cuh"
namespace fmwr {
void train_als_mcmc(fmwr_ctx*, fmwr_model*, fmwr_data*, const fmwr_solver_cfg*, fmwr_trace*)
{
  throw Error(FMWR_ERR_UNSUPPORTED, "ALS/MCMC not built yet");
}
}

#include "mtg_common.cuh"
extern "C" int mtg_loglike_grad_dev(const mtg_model_t *, int, int, int, int, const double *, const double *,
                                    const double *, const double *, const double *, double *, double *,
                                    int32_t *, uint32_t, void *)
{ mtg_set_error("mtg_loglike_grad_dev: not built yet"); return MTG_EUNSUP; }

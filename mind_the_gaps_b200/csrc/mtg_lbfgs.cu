#include "mtg_common.cuh"
extern "C" int mtg_lbfgs_dev(const mtg_model_t *, int, int, int, int, const double *, const double *,
                             const double *, const double *, const double *, const mtg_lbfgs_opts_t *,
                             double *, double *, int32_t *, int32_t *, uint32_t, void *)
{ mtg_set_error("mtg_lbfgs_dev: not built yet"); return MTG_EUNSUP; }

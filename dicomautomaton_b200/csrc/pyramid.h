// dcma_b200 -- multi-resolution pyramid (extension, see pyramid.cu)
#pragma once
#include <array>

#include "engine.h"

namespace dcma {
void coarsen_geom(const dcma_b200_geom &fine, dcma_b200_geom *coarse, int halved[3]);
void restrict_image_host(const dcma_b200_geom &fine, const float *image, dcma_b200_geom *coarse, float *out, int device);
void prolong_field_host(const dcma_b200_geom &fine, const double *coarse_field, double *out, int device);
void pyramid_register(const dcma_b200_geom &geom, const float *fixed, const float *moving, const dcma_b200_demons_params &params,
                      double *deformation_out, double *mse_history, dcma_b200_run_stats *stats);
}  // namespace dcma

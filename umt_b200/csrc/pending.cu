// Entry points whose kernels are not written yet fail loudly (never a CPU fallback).
#include "umt_internal.h"
#define NOT_YET(ctx, what) do { if (!(ctx)) return UMT_ERR_ARG; UMT_FAIL(ctx, UMT_ERR_STATE, what " is not implemented yet"); } while (0)

extern "C" int umt_gta_set_opacity(umt_ctx *ctx, const double *, const double *, const double *) { NOT_YET(ctx, "GTA"); }
extern "C" int umt_gta_sweep(umt_ctx *ctx, const double *, const double *, double *, double *, int) { NOT_YET(ctx, "GTA"); }

// Entry points whose kernels are not written yet fail loudly (never a CPU fallback).
#include "umt_internal.h"
#define NOT_YET(ctx, what) do { if (!(ctx)) return UMT_ERR_ARG; UMT_FAIL(ctx, UMT_ERR_STATE, what " is not implemented yet"); } while (0)

int umt_launch_sweeprz(umt_ctx *ctx, int) { NOT_YET(ctx, "RZ sweep"); }
int umt_exchange_begin_pass(umt_ctx *ctx) { NOT_YET(ctx, "psib exchange"); }
int umt_exchange_test_convergence(umt_ctx *ctx, double, int *) { NOT_YET(ctx, "psib exchange"); }
extern "C" int umt_add_shared_boundary(umt_ctx *ctx, int, int, int) { NOT_YET(ctx, "umt_add_shared_boundary"); }
extern "C" int umt_nccl_unique_id(unsigned char *) { return UMT_ERR_NCCL; }
extern "C" int umt_set_comm(umt_ctx *ctx, int, int, const unsigned char *) { NOT_YET(ctx, "umt_set_comm"); }
extern "C" int umt_gta_set_opacity(umt_ctx *ctx, const double *, const double *, const double *) { NOT_YET(ctx, "GTA"); }
extern "C" int umt_gta_sweep(umt_ctx *ctx, const double *, const double *, double *, double *, int) { NOT_YET(ctx, "GTA"); }

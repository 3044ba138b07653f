#include "reach_common.cuh"
using namespace reach;
struct reach_glgm06_plan { reach_ctx* ctx; };
#define NOTYET(ctx) set_error((ctx), REACH_EUNSUPPORTED, "GLGM06 device path not built yet")
extern "C" int reach_glgm06_plan_create(reach_ctx* ctx, int, int64_t, int, int, int, int, reach_glgm06_plan**) { return NOTYET(ctx); }
extern "C" int reach_glgm06_plan_upload(reach_glgm06_plan* p, const double*, const double*, const double*, const double*, const double*) { return NOTYET(p ? p->ctx : nullptr); }
extern "C" int reach_glgm06_plan_run(reach_glgm06_plan* p) { return NOTYET(p ? p->ctx : nullptr); }
extern "C" int reach_glgm06_plan_sync(reach_glgm06_plan* p) { return NOTYET(p ? p->ctx : nullptr); }
extern "C" int reach_glgm06_plan_ngens(reach_glgm06_plan* p, int, int64_t, int*) { return NOTYET(p ? p->ctx : nullptr); }
extern "C" int reach_glgm06_plan_fetch(reach_glgm06_plan* p, int, int64_t, double*, double*, int32_t*) { return NOTYET(p ? p->ctx : nullptr); }
extern "C" int reach_glgm06_plan_support(reach_glgm06_plan* p, const double*, double*) { return NOTYET(p ? p->ctx : nullptr); }
extern "C" int reach_glgm06_plan_stats(reach_glgm06_plan* p, double*, int64_t*, int64_t*) { return NOTYET(p ? p->ctx : nullptr); }
extern "C" void reach_glgm06_plan_destroy(reach_glgm06_plan*) {}
extern "C" int reach_glgm06(reach_ctx* ctx, int, int64_t, int, int, const double*, const double*, const double*, int, const double*, const double*, int, double*, double*, int, int32_t*) { return NOTYET(ctx); }
extern "C" int reach_reduce_order_gir05(reach_ctx* ctx, int, int, int, const double*, double*, int*, int32_t*) { return NOTYET(ctx); }

// lucifit_fit_inst.cu -- compiled once per (model, n_lines): -DLF_INST_MODEL=m -DLF_INST_L=l.
// Defines  lf_launcher_m<m>_l<l>(bucket)  returning the launcher of the fit kernel for that group-count bucket.
#include "lucifit_fit_kernel.cuh"
#include "lucifit_instances.h"

#define LF_CAT2(a, b, c, d) a##b##c##d
#define LF_CAT(a, b, c, d) LF_CAT2(a, b, c, d)

template <int L_, int G_>
static lf_launch_fn lf_pick()
{
    if constexpr (L_ == LF_INST_L) return &lf_launch<LF_INST_MODEL, L_, G_>;
    else return nullptr;
}

extern "C" lf_launch_fn LF_CAT(lf_launcher_m, LF_INST_MODEL, _l, LF_INST_L)(int bucket)
{
#define X(L_, G_) if (L_ == LF_INST_L && G_ == bucket) return lf_pick<L_, G_>();
    LF_INSTANCES(X)
#undef X
    return nullptr;
}

// tcgen05 / TMEM backward of one coupling layer (gwtf_tc_bwd.cuh).
#include "gwtf_host.h"
#include "gwtf_bwd.cuh"

namespace gwtf {

int launch_bwd_layer_tc(const BwdArgs& a, int phase, cudaStream_t st) {
    return launch_bwd_layer_mma(a, phase, st);      // placeholder while the tcgen05 kernels are brought up
}

}  // namespace gwtf

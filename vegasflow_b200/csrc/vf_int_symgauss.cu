// Fused-kernel instantiations for the symgauss integrand (examples/simgauss_tf.py:22-32).
#include "vf_event.cuh"
namespace vf {
VF_INSTANTIATE_GENERIC_INTEGRAND(SymGauss)
}

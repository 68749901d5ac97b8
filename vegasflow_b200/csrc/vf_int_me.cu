// Fused-kernel instantiations for the LO matrix elements
// (examples/drellyan_lo_tf.py, examples/singletop_lo_tf.py).
#include "vf_event.cuh"
namespace vf {
VF_INSTANTIATE_FIXED_INTEGRAND(DrellYanLO)
VF_INSTANTIATE_FIXED_INTEGRAND(SingleTopLO)
}

// Fused-kernel instantiations for the product integrand (README.md:63-68).
#include "vf_event.cuh"
namespace vf {
VF_INSTANTIATE_GENERIC_INTEGRAND(Product)
}

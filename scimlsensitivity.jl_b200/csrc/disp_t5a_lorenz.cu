// disp_t5a_lorenz.cu -- adaptive Tsit5 kernels instantiated for the Lorenz family
#include "disp_t5a.inc"
namespace b200adj {
template int launch_t5a_fwd<Lorenz>(Handle*, const T5aArgs&);
template int launch_t5a_rev<Lorenz>(Handle*, const T5aArgs&);
}

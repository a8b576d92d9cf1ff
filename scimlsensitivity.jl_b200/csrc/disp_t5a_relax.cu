// disp_t5a_relax.cu -- adaptive Tsit5 kernels instantiated for the Relax family (parameter-dependent event condition)
#include "disp_t5a.inc"
namespace b200adj {
template int launch_t5a_fwd<Relax>(Handle*, const T5aArgs&);
template int launch_t5a_rev<Relax>(Handle*, const T5aArgs&);
}

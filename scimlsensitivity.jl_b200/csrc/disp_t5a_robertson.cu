// disp_t5a_robertson.cu -- adaptive Tsit5 kernels instantiated for the Robertson family
#include "disp_t5a.inc"
namespace b200adj {
template int launch_t5a_fwd<Robertson>(Handle*, const T5aArgs&);
template int launch_t5a_rev<Robertson>(Handle*, const T5aArgs&);
}

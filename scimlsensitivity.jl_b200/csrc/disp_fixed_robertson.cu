// disp_fixed_robertson.cu -- fixed-step Tsit5 kernels instantiated for the Robertson family
#include "disp_fixed.inc"
namespace b200adj {
template int launch_fwd<Robertson>(Handle*, const OdeFwdArgs&);
template int launch_rev<Robertson>(Handle*, const OdeRevArgs&);
}

// disp_fixed_lorenz.cu -- fixed-step Tsit5 kernels instantiated for the Lorenz family
#include "disp_fixed.inc"
namespace b200adj {
template int launch_fwd<Lorenz>(Handle*, const OdeFwdArgs&);
template int launch_rev<Lorenz>(Handle*, const OdeRevArgs&);
template int launch_fwd_f32<Lorenz>(Handle*, const OdeFwdArgsT<float>&);
template int launch_rev_f32<Lorenz>(Handle*, const OdeRevArgsT<float>&);
}

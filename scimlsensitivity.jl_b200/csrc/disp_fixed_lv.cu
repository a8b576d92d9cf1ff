// disp_fixed_lv.cu -- fixed-step Tsit5 kernels instantiated for the LotkaVolterra family
#include "disp_fixed.inc"
namespace b200adj {
template int launch_fwd<LotkaVolterra>(Handle*, const OdeFwdArgs&);
template int launch_rev<LotkaVolterra>(Handle*, const OdeRevArgs&);
template int launch_fwd_f32<LotkaVolterra>(Handle*, const OdeFwdArgsT<float>&);
template int launch_rev_f32<LotkaVolterra>(Handle*, const OdeRevArgsT<float>&);
}

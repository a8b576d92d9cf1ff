// disp_t5a_lv.cu -- adaptive Tsit5 kernels instantiated for the LotkaVolterra family
#include "disp_t5a.inc"
namespace b200adj {
template int launch_t5a_fwd<LotkaVolterra>(Handle*, const T5aArgs&);
template int launch_t5a_rev<LotkaVolterra>(Handle*, const T5aArgs&);
}

// disp_t5a_ball.cu -- adaptive Tsit5 kernels instantiated for the BouncingBall family (state-dependent events)
#include "disp_t5a.inc"
namespace b200adj {
template int launch_t5a_fwd<BouncingBall>(Handle*, const T5aArgs&);
template int launch_t5a_rev<BouncingBall>(Handle*, const T5aArgs&);
}
